# where does K2's time go at C4?  (a) relation without edges = ptr reads + zero writes only, (b) a real relation
import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import mpgnn_b200
from mpgnn_b200 import _lib
lib = _lib.load()
n, e, r, f = 10_000_000, 200_000_000, 64, 128
gen = torch.Generator(device="cuda").manual_seed(0)
ei = torch.randint(0, n, (2, e), device="cuda", generator=gen); et = torch.randint(0, r, (e,), device="cuda", generator=gen)
g = mpgnn_b200.RelationGraph(ei, et, n, r + 1); del ei, et
x = torch.randn(n, f, device="cuda", generator=gen); out = torch.empty_like(x)
def run(rel, transpose=0, mean=1):
    _lib.check(lib.mpgnn_spmm(g.handle, rel, transpose, mean, _lib.ptr(x), f, f, None, 0, _lib.ptr(out), f, _lib.current_stream()))
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
print("empty relation (ptr + zero rows): %.3f ms" % t(lambda: run(r)))
print("relation 3 (3.1M edges) fwd:      %.3f ms" % t(lambda: run(3)))
print("relation 3 transposed, no init:   %.3f ms" % t(lambda: run(3, 1, 0)))

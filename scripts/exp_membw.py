# plain torch memory-bandwidth reference points for the K2 discussion (experiment helper)
import torch
n, f = 10_000_000, 128
x = torch.randn(n, f, device="cuda"); y = torch.empty_like(x)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
gb = n * f * 4 / 1e9
ms = t(lambda: y.zero_());            print("write-only  %.2f GB: %.3f ms  %.2f TB/s" % (gb, ms, gb / ms))
ms = t(lambda: y.copy_(x));           print("copy      2x%.2f GB: %.3f ms  %.2f TB/s" % (gb, ms, 2 * gb / ms))
ms = t(lambda: x.sum());              print("read-only   %.2f GB: %.3f ms  %.2f TB/s" % (gb, ms, gb / ms))
ms = t(lambda: torch.add(x, 1.0, out=y)); print("read+write 2x%.2f GB: %.3f ms  %.2f TB/s" % (gb, ms, 2 * gb / ms))

import torch, sys
sys.path.insert(0, '.')
import mpgnn_b200
DEV='cuda'
def run(n, f_in, f_out, relu):
    g = torch.Generator().manual_seed(n)
    ei = torch.randint(0, n, (2, 6*n), generator=g); et = torch.randint(0, 3, (6*n,), generator=g)
    x = torch.randn(n, f_in, generator=g).to(DEV)
    gy = torch.randn(n, f_out, generator=g).to(DEV)
    conv = mpgnn_b200.CustomRGCNConv(f_in, f_out, 1, flow='target_to_source', device=DEV)
    graph = mpgnn_b200.RelationGraph(ei, et, n, 3, device=DEV)
    res = {}
    for prec in ('fp32', 'tf32x3', 'tf32x3'):
        xi = x.clone().requires_grad_(True)
        y = conv.hop(1, xi, graph, relu=relu, precision=prec)
        y.backward(gy)
        if prec == 'fp32':
            res['a'] = xi.grad.clone(); continue
        a, b = res['a'], xi.grad
        d = (a-b).abs()
        bad = d > 1e-4 * a.abs().max()
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print(f'GX n={n} f_in={f_in} f_out={f_out} relu={relu} maxerr={d.max().item():.3e} amax={a.abs().max().item():.3f} bad_elems={int(bad.sum())} bad_rows={rows.numel()} bad_cols={cols.numel()}')
        if rows.numel():
            tiles = torch.unique(rows // 128)
            print('  bad tiles (first 40):', tiles[:40].tolist(), ' count', tiles.numel())
            print('  rows within tile (first 40):', (rows % 128)[:40].tolist())
            print('  cols (first 40):', cols[:40].tolist())
            r0 = rows[0].item()
            print('  row', r0, 'a', a[r0][:8].tolist(), 'b', b[r0][:8].tolist())
for shape in [(40000,128,128,False),(40000,128,128,True),(5000,128,128,True),(40000,64,64,True),(19001,64,128,True)]:
    run(*shape)

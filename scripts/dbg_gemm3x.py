import sys, torch
sys.path.insert(0, '.')
from linear_operator_b200 import _kernels
dev = 'cuda:0'
torch.manual_seed(0)
def run(M, N, K, ta, tb, splits=1, kind='rand'):
    if kind == 'ones':
        A = torch.ones(1, K, M, device=dev) if ta else torch.ones(1, M, K, device=dev)
        B = torch.ones(1, N, K, device=dev) if tb else torch.ones(1, K, N, device=dev)
    elif kind == 'rowid':  # A[m,k] = m+1 (const over k), B[k,n] = 1 -> D[m,n] = K*(m+1)
        Am = (torch.arange(M, device=dev).float() + 1)[:, None].expand(M, K)
        A = (Am.t() if ta else Am).contiguous()[None]
        B = torch.ones(1, N, K, device=dev) if tb else torch.ones(1, K, N, device=dev)
    elif kind == 'colid':  # B[k,n] = n+1
        A = torch.ones(1, K, M, device=dev) if ta else torch.ones(1, M, K, device=dev)
        Bm = (torch.arange(N, device=dev).float() + 1)[None, :].expand(K, N)
        B = (Bm.t() if tb else Bm).contiguous()[None]
    elif kind == 'kid':  # A[m,k] = 1, B[k,n] = delta(k, n)  -> D = 1 for n<K
        A = torch.ones(1, K, M, device=dev) if ta else torch.ones(1, M, K, device=dev)
        Bm = torch.eye(K, N, device=dev)
        B = (Bm.t() if tb else Bm).contiguous()[None]
    else:
        A = torch.randn(1, K, M, device=dev) if ta else torch.randn(1, M, K, device=dev)
        B = torch.randn(1, N, K, device=dev) if tb else torch.randn(1, K, N, device=dev)
    D = _kernels.gemm3x(A, B, trans_a=ta, trans_b=tb, splits=splits)
    torch.cuda.synchronize()
    Ad = (A.mT if ta else A).double(); Bd = (B.mT if tb else B).double()
    W = Ad @ Bd
    err = ((D.double() - W).norm() / W.norm()).item()
    print(f"M{M} N{N} K{K} ta={int(ta)} tb={int(tb)} splits={splits} {kind:6s} relerr {err:.3e}  D[0,:2,:4]={D[0,:2,:4].flatten().tolist()}  want={W[0,:2,:4].flatten().tolist()}", flush=True)
for ta in (False, True):
    for tb in (True, False):
        for kind in ('ones', 'rowid', 'colid', 'kid', 'rand'):
            try:
                run(128, 128, 32, ta, tb, 1, kind)
            except Exception as e:
                print('EXC', ta, tb, kind, repr(e)[:200]); raise
run(128, 128, 8, False, True)
run(128, 128, 64, False, True)
run(128, 128, 256, False, True)
run(256, 256, 256, False, True)
run(200, 136, 1000, False, True)
run(200, 136, 1000, False, True, splits=4)
run(200, 136, 1000, False, False, splits=4)
run(200, 136, 1000, True, False, splits=4)
run(200, 136, 1000, True, True, splits=4)

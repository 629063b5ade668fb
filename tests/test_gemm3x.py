"""lob_gemm3x (csrc/gemm3x.cu): the tcgen05 3xTF32 batched GEMM against an fp64 product of the same fp32 inputs, for
every operand major-ness, ragged tiles, split-K, grouped operand batches and the fused epilogue."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from linear_operator_b200 import _kernels  # noqa: E402
from test_gpu_parity import DEV, check, npy  # noqa: E402

# 3xTF32: the dropped lo*lo term and the tf32 rounding of the lo parts are ~2^-22 relative per product; the fp32
# accumulator in TMEM adds ~1e-8 per accumulated k step relative to the running magnitude (measured 8e-6 at K = 1000
# on random data without split-K)
TOL = 2e-5


def rnd(*shape, gen):
    return torch.randn(*shape, device=DEV, generator=gen)


@pytest.mark.parametrize("trans_a", [False, True])
@pytest.mark.parametrize("trans_b", [False, True])
@pytest.mark.parametrize("shape", [(3, 200, 136, 1000), (1, 128, 128, 32), (2, 100, 100, 4100), (1, 512, 256, 20000)])
def test_gemm3x_all_layouts(trans_a, trans_b, shape):
    nb, M, N, K = shape
    gen = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = rnd(nb, K, M, gen=gen) if trans_a else rnd(nb, M, K, gen=gen)
    B = rnd(nb, N, K, gen=gen) if trans_b else rnd(nb, K, N, gen=gen)
    D = _kernels.gemm3x(A, B, trans_a=trans_a, trans_b=trans_b)
    assert D is not None and D.shape == (nb, M, N)
    Ad = (A.mT if trans_a else A).double()
    Bd = (B.mT if trans_b else B).double()
    check(npy(D), npy(Ad @ Bd), TOL)


def test_gemm3x_split_k_double_output_and_determinism():
    gen = torch.Generator(device=DEV).manual_seed(5)
    L = rnd(4, 5000, 100, gen=gen)  # Gram matrix L^T L: both operands MN-major, long contraction
    G = _kernels.gemm3x(L, L, trans_a=True, splits=8, out_dtype=torch.float64)
    assert G is not None and G.dtype == torch.float64
    check(npy(G), npy(L.double().mT @ L.double()), TOL)
    G2 = _kernels.gemm3x(L, L, trans_a=True, splits=8, out_dtype=torch.float64)
    assert torch.equal(G, G2)


def test_gemm3x_epilogue_and_grouped_batches():
    """x = (R - w U^T) / sigma with one U shared by the whole batch (b_div = batch) and per-row factors."""
    gen = torch.Generator(device=DEV).manual_seed(6)
    Bsz, Nl, r = 130, 3000, 64
    U = rnd(Nl, r, gen=gen) / 8
    w = rnd(Bsz, r, gen=gen)
    R = rnd(Bsz, Nl, gen=gen)
    sig = 0.5 + torch.rand(Bsz, device=DEV, generator=gen)
    out = _kernels.gemm3x(w.unsqueeze(0), U.unsqueeze(0), trans_b=True, row_alpha=(-1.0 / sig).unsqueeze(0),
                          E=R.unsqueeze(0), row_beta=(1.0 / sig).unsqueeze(0))
    want = (R.double() - w.double() @ U.double().mT) / sig.double().unsqueeze(-1)
    check(npy(out[0]), npy(want), TOL)
    # the same product with swapped roles and a transposed store: x^T = U w^T, written as (B, N)
    out_t = _kernels.gemm3x(U.unsqueeze(0), w.unsqueeze(0), trans_b=True, row_alpha=(-1.0 / sig).unsqueeze(0),
                            E=R.unsqueeze(0), row_beta=(1.0 / sig).unsqueeze(0), store_transposed=True)
    assert out_t.shape == (1, Bsz, Nl)
    check(npy(out_t[0]), npy(want), TOL)
    Dt = _kernels.gemm3x(R.unsqueeze(0), U.unsqueeze(0), splits=4, store_transposed=True)  # split-K + transposed store
    check(npy(Dt[0]), npy((R.double() @ U.double()).mT), TOL)
    # grouped operand batches: A[b // 3] B[b]
    A = rnd(2, 96, 40, gen=gen)
    Bm = rnd(6, 40, 72, gen=gen)
    D = _kernels.gemm3x(A, Bm, a_div=3)
    want = A.double().repeat_interleave(3, dim=0) @ Bm.double()
    check(npy(D), npy(want), TOL)


def test_gemm3x_rejects_unaligned_layouts():
    gen = torch.Generator(device=DEV).manual_seed(7)
    A = rnd(1, 64, 33, gen=gen)  # leading dimension 33: not a multiple of 4 elements
    B = rnd(1, 33, 64, gen=gen)
    assert _kernels.gemm3x(A, B) is None

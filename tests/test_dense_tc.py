"""The tcgen05 split-TF32 dense operator matmuls (csrc/dense_stream2p.cu: CTA-pair version of the A-stationary streaming
kernel, the default for long contractions; csrc/dense_stream2.cu: its single-CTA form, serves short contractions and
calls without a workspace; csrc/dense_stream.cu: its X-stationary predecessor, serves 48 < C <= 64; csrc/dense_tc.cu: first-generation
kernel, kept as the workspace-free fallback and for short contractions) against an fp64 product of the same fp32
inputs and against the CUDA-core kernel.  fp32 inputs; the bar is fp32 accuracy (error << 1e-4, north-star parity
tolerance)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from linear_operator_b200 import _kernels, _lib  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(params=["stream2p", "stream2", "stream", "tc"])
def impl(request):
    """Pins the fp32 tensor-core kernel through the library's explicit test hook (lob_debug_pin_dense_impl)."""
    _lib.pin_dense_impl(request.param)
    yield request.param
    _lib.pin_dense_impl(None)


def _case(B, N, C, with_diag, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    A = torch.randn(B, N, N, device=DEV, generator=g) / N**0.5
    X = torch.randn(B, N, C, device=DEV, generator=g)
    d = (0.1 + torch.rand(B, N, device=DEV, generator=g)) if with_diag else None
    ref = A.double() @ X.double()
    if with_diag:
        ref = ref + d.double().unsqueeze(-1) * X.double()
    return A, X, d, ref


@pytest.mark.parametrize(
    "B,N,C,with_diag",
    [(2, 384, 33, True), (1, 260, 1, False), (2, 1000, 17, True), (2, 512, 48, True), (3, 2052, 33, True),
     (1, 128, 8, False), (2, 5000, 33, True), (2, 776, 64, True), (40, 300, 33, True), (2, 640, 36, True),
     (1, 300, 18, True), (2, 600, 35, False)],
)
def test_dense_tc_matches_fp64(B, N, C, with_diag, impl):
    A, X, d, ref = _case(B, N, C, with_diag, 100 + N + C)
    Y, dots, n_parts = _kernels.dense_matmul(A, X, d=d, want_dots=True)
    torch.cuda.synchronize()
    scale = ref.abs().max()
    err = ((Y.double() - ref).abs().max() / scale).item()
    print(f"B={B} N={N} C={C}: tc err {err:.3e}")
    # the tensor core truncates (does not round) when it adds into the fp32 accumulator: a bias that grows linearly in
    # the number of k steps (measured ~3e-9 * K); still >6x below the 1e-4 parity bar at N = 5000
    assert err < 3e-6 + 3.5e-9 * N, f"tensor-core matmul error {err}"
    dots_ref = (X.double() * ref).sum(-2)
    derr = ((dots.sum(1) - dots_ref).abs().max() / dots_ref.abs().max()).item()
    assert derr < 3e-6 + 3.5e-9 * N, f"fused <x,y> partials error {derr}"
    # same call through the CUDA-core kernel
    _lib.pin_dense_impl("simt")
    try:
        Y2 = _kernels.dense_matmul(A, X, d=d)
    finally:
        _lib.pin_dense_impl(impl)
    err2 = ((Y2.double() - ref).abs().max() / scale).item()
    print(f"   simt err {err2:.3e}")
    assert err2 < 3e-6
    assert ((Y - Y2).abs().max() / scale).item() < 3e-6 + 3.5e-9 * N


def test_dense_tc_plain_tf32_would_fail(impl):
    """Sanity of the bar itself: a single-pass TF32 product is ~1e-3 off, the 3xTF32 kernel is not."""
    A, X, d, ref = _case(1, 1024, 33, False, 7)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        tf32 = A @ X
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    e_tf32 = ((tf32.double() - ref).abs().max() / ref.abs().max()).item()
    Y = _kernels.dense_matmul(A, X)
    e_ours = ((Y.double() - ref).abs().max() / ref.abs().max()).item()
    assert e_ours < 3e-6 + 3.5e-9 * 1024 and e_ours < e_tf32 / 20


def test_dense_tc_back_to_back_launches_are_deterministic(impl):
    """Stress: 12 launches of a 48 x 5000 x 5000 problem queued without host synchronisation must be bit-identical and
    agree with the CUDA-core kernel.  Regression test for pipeline hazards of the TMEM-operand hand-off (tcgen05.st ->
    mbarrier -> tcgen05.mma): a fast-issue variant of the first-generation kernel corrupted single rows once the tensor
    pipe was backed up; the shipped kernels (elect.sync'd issue on a converged warp, wait::st directly behind the
    store) must stay bit-reproducible under load."""
    import os

    B, N, C = 48, 5000, 33
    g = torch.Generator(device=DEV).manual_seed(5)
    A = torch.randn(B, N, N, device=DEV, generator=g) / N**0.5
    X = torch.randn(B, N, C, device=DEV, generator=g)
    _lib.pin_dense_impl("simt")
    try:
        ref = _kernels.dense_matmul(A, X)
    finally:
        _lib.pin_dense_impl(None)
    outs = [_kernels.dense_matmul(A, X) for _ in range(12)]
    torch.cuda.synchronize()
    scale = ref.abs().max()
    for Y in outs:
        assert torch.equal(Y, outs[0])
        assert int((((Y - ref).abs() / scale) > 1e-4).sum()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("constant_diag", [True, False])
def test_dense_matmul_generalised_epilogue(dtype, constant_diag, impl):
    """Y = alpha[b] (A X) + d (.) E and the partial sums of E * Y -- the fused precondition_closure form
    (A = Q, X = Q^T r, E = r).  fp32 goes through the tensor-core kernel (K = 100), fp64 through the CUDA-core one."""
    B, N, k, C = 3, 1000, 100, 33
    g = torch.Generator(device=DEV).manual_seed(3)
    Q = torch.randn(B, N, k, device=DEV, generator=g, dtype=dtype) / N**0.5
    t = torch.randn(B, k, C, device=DEV, generator=g, dtype=dtype)
    E = torch.randn(B, N, C, device=DEV, generator=g, dtype=dtype)
    alpha = -(0.5 + torch.rand(B, device=DEV, generator=g, dtype=dtype))
    d = (0.5 + torch.rand(B, 1 if constant_diag else N, device=DEV, generator=g, dtype=dtype))
    Y, dots, _n = _kernels.dense_matmul(Q, t, d=d, want_dots=True, E=E, alpha=alpha)
    ref = alpha.double().view(B, 1, 1) * (Q.double() @ t.double()) + d.double().expand(B, N).unsqueeze(-1) * E.double()
    tol = 1e-5 if dtype == torch.float32 else 1e-13
    assert ((Y.double() - ref).abs().max() / ref.abs().max()).item() < tol
    dref = (E.double() * ref).sum(-2)
    assert ((dots.sum(1) - dref).abs().max() / dref.abs().max()).item() < tol
    # skinny output (M < 128): t = Q^T r through the same kernel
    Qt = Q.mT.contiguous()
    T2 = _kernels.dense_matmul(Qt, E)
    ref2 = Qt.double() @ E.double()
    assert ((T2.double() - ref2).abs().max() / ref2.abs().max()).item() < tol


@pytest.mark.parametrize("B,N,C", [(1, 512, 17), (3, 500, 33), (2, 100, 5), (1, 1000, 64)])
def test_dense_matmul_small_fp64_row_kernel(B, N, C):
    """Small fp64 problems (BASELINE config 1) take the row-per-warp kernel: no fused <X, Y> partial sums (linear_cg adds
    its own dot pass), fused + d (.) X, operator shared across the batch."""
    g = torch.Generator(device=DEV).manual_seed(B + N + C)
    A = torch.randn(B, N, N, device=DEV, generator=g, dtype=torch.float64) / N**0.5
    X = torch.randn(B, N, C, device=DEV, generator=g, dtype=torch.float64)
    d = 0.1 + torch.rand(B, N, device=DEV, generator=g, dtype=torch.float64)
    Y, dots, n_parts = _kernels.dense_matmul(A, X, d=d, want_dots=True)
    assert dots is None and n_parts == 0
    ref = A @ X + d.unsqueeze(-1) * X
    assert ((Y - ref).abs().max() / ref.abs().max()).item() < 1e-13
    Y1 = _kernels.dense_matmul(A[:1], X)  # one operator for every batch element, no diagonal
    assert ((Y1 - A[:1] @ X).abs().max() / ref.abs().max()).item() < 1e-13


@pytest.mark.parametrize("kernel", ["stream2p", "stream2"])
def test_dense_stream_accumulation_bias_on_positive_data(kernel):
    """Worst case for the tensor core's truncating fp32 accumulate: all-positive operands, no cancellation.  The bias is
    linear in the number of indices ONE accumulator sees (measured 1.5e-8 per index): 7.5e-5 at K = 5000 for the
    single-CTA kernel, 2.5e-5 -- one 2048-index segment -- for the pair kernel, which drains its accumulator into a
    round-to-nearest running sum every 64 k blocks.  Both are pinned below the 1e-4 parity bar."""
    _lib.pin_dense_impl(kernel)
    try:
        g = torch.Generator(device=DEV).manual_seed(11)
        A = torch.rand(1, 512, 5000, device=DEV, generator=g)
        X = torch.rand(1, 5000, 16, device=DEV, generator=g)
        Y = _kernels.dense_matmul(A, X)
    finally:
        _lib.pin_dense_impl(None)
    ref = A.double() @ X.double()
    rel = ((Y.double() - ref).abs() / ref.abs()).max().item()
    print(f"all-positive K=5000: rel err {rel:.3e}")
    assert rel < (4e-5 if kernel == "stream2p" else 1e-4)
    assert (Y.double() <= ref * (1 + 1e-6)).all()  # truncation only ever loses magnitude


def test_dense_matmul_operator_shared_across_batch(impl):
    """One operator (batch 1, also as a strided view with lda > K) applied to a batch of right-hand sides: the TMA
    descriptor pins the batch coordinate to 0 and keeps the caller's leading dimension."""
    g = torch.Generator(device=DEV).manual_seed(9)
    N, C, B = 768, 33, 5
    big = torch.randn(1, N, N + 64, device=DEV, generator=g) / N**0.5
    for A in (big[..., :N].contiguous(), big[..., :N]):
        X = torch.randn(B, N, C, device=DEV, generator=g)
        d = 0.1 + torch.rand(B, N, device=DEV, generator=g)
        Y, dots, _ = _kernels.dense_matmul(A, X, d=d, want_dots=True)
        ref = A.double() @ X.double() + d.double().unsqueeze(-1) * X.double()
        assert ((Y.double() - ref).abs().max() / ref.abs().max()).item() < 3e-6 + 3.5e-9 * N
        dref = (X.double() * ref).sum(-2)
        assert ((dots.sum(1) - dref).abs().max() / dref.abs().max()).item() < 3e-6 + 3.5e-9 * N


def test_dense_matmul_c_abi_without_workspace():
    """A C-ABI caller that passes no scratch (ws = NULL) still gets a tensor-core product: the streaming kernel then
    produces the split right-hand side inside the kernel (workspace-free mode)."""
    from linear_operator_b200 import _lib
    from linear_operator_b200._lib import check, dt, ptr, stream

    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(12)
    B, N, C = 3, 1000, 33
    A = torch.randn(B, N, N, device=DEV, generator=g) / N**0.5
    X = torch.randn(B, N, C, device=DEV, generator=g)
    d = 0.1 + torch.rand(B, N, device=DEV, generator=g)
    Y = torch.empty(B, N, C, device=DEV)
    n_parts = int(lib.lob_dense_matmul_parts(N))
    dots = torch.empty(B, n_parts, C, dtype=torch.float64, device=DEV)
    before = _lib.launch_count()
    check(lib.lob_dense_matmul(dt(X), B, N, N, C, ptr(A), N, N * N, ptr(X), ptr(Y), ptr(d), N, 1, ptr(dots), None, 0,
                               stream(X)), "lob_dense_matmul")
    torch.cuda.synchronize()
    assert _lib.launch_count() - before == 1  # one kernel, no split pass
    ref = A.double() @ X.double() + d.double().unsqueeze(-1) * X.double()
    assert ((Y.double() - ref).abs().max() / ref.abs().max()).item() < 3e-6 + 3.5e-9 * N
    dref = (X.double() * ref).sum(-2)
    assert ((dots.sum(1) - dref).abs().max() / dref.abs().max()).item() < 3e-6 + 3.5e-9 * N

"""CPU-only tests: C-ABI surface, host-side dispatch / shape / error logic, settings, the patchable solver symbol.
No compute is requested from the library here (there is no GPU in the build container)."""
import ctypes
import os
import re
from unittest import mock

import pytest
import torch

import linear_operator_b200 as lo
from linear_operator_b200 import _lib, settings
from linear_operator_b200.operators import (
    AddedDiagLinearOperator,
    ConstantDiagLinearOperator,
    DenseLinearOperator,
    DiagLinearOperator,
    IdentityLinearOperator,
    KroneckerProductLinearOperator,
    LowRankRootAddedDiagLinearOperator,
    LowRankRootLinearOperator,
    PsdSumLinearOperator,
    RootLinearOperator,
    ToeplitzLinearOperator,
)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lob_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lob_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_header_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lob_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms  # the ctypes table binds exactly the header's surface
    assert _lib.load().lob_version() >= 100


def test_argument_errors_come_back_through_the_abi():
    lib = _lib.load()
    p = _lib.CgParams(0, 10, 1, 0, 0, 0, 10, 10, 0, 1.0, 1e-10, 1e-10)
    assert lib.lob_cg_workspace_bytes(ctypes.byref(p)) == 0
    rc = lib.lob_dense_matmul(0, 0, 4, 4, 1, None, 4, 16, None, None, None, 0, 0, None, None, 0, None)
    assert rc == -1 and b"positive" in lib.lob_last_error()
    rc = lib.lob_tn_matmul(7, 0, 1, 4, 2, 2, None, 0, None, 0, None, None, None)
    assert rc == -1


def test_new_entry_points_validate_arguments_without_a_gpu():
    lib = _lib.load()
    # lanczos: bad mode / iteration index / NULL pointers come back as argument errors, nothing is launched
    assert lib.lob_lanczos_step(0, 5, 1, 8, 1, 4, 1, None, None, None, None, 1e-5, None) == -1
    assert b"mode" in lib.lob_last_error()
    assert lib.lob_lanczos_step(0, 0, 1, 8, 1, 4, 2, None, None, None, None, 1e-5, None) == -1  # mode 0 needs k = 0
    assert lib.lob_lanczos_step(0, 1, 1, 8, 1, 4, 1, None, None, None, None, 1e-5, None) == -1
    assert b"NULL" in lib.lob_last_error()
    assert lib.lob_lanczos_init(0, 0, 8, 1, None, None, None) == -1
    # streaming matmul scratch: (B, 2 * round_up(C, 8), K) floats for fp32, none for fp64 or C > 64
    assert lib.lob_dense_matmul_workspace_bytes(0, 1024, 5000, 5000, 33) == 1024 * 80 * 5000 * 4  # pair kernel: 2 x (2 g + 8) rows, g = 16
    # operand rows of the pair kernel per column count: 2 x (2 g + 8) when C = 2 g + e with e <= 4, else 4 x ceil8(C / 2)
    for C, rows in ((1, 32), (16, 32), (17, 48), (20, 48), (21, 64), (32, 64), (33, 80), (36, 80), (37, 96), (48, 96)):
        assert lib.lob_dense_matmul_workspace_bytes(0, 2, 1000, 1000, C) >= 2 * rows * 1000 * 4, C
    assert lib.lob_dense_matmul_workspace_bytes(1, 1024, 5000, 5000, 33) == 0
    assert lib.lob_dense_matmul_workspace_bytes(0, 2, 100, 100, 65) == 0
    # Toeplitz column-pair path: sizes and NULL pointers are rejected before anything is launched
    assert lib.lob_toeplitz_colmax(0, 0, 8, 3, None, None, None) == -1
    assert lib.lob_toeplitz_pack(0, 1, 8, 3, 4, None, None, None, None) == -1  # L < N
    assert b"bad sizes" in lib.lob_last_error()
    assert lib.lob_toeplitz_pack(0, 1, 8, 3, 16, None, None, None, None) == -1
    assert b"NULL" in lib.lob_last_error()
    assert lib.lob_toeplitz_mulr(0, 1, 2, 15, None, 0, None, None) == -1  # odd transform length
    assert lib.lob_toeplitz_unpack(0, 1, 8, 3, 16, None, 1.0, None, None, None, 0, 0, None, None, None) == -1
    assert lib.lob_toeplitz_unpack_parts(0, 1000, 33) == 8 and lib.lob_toeplitz_unpad_parts(0, 1000, 33) == 8
    assert lib.lob_toeplitz_unpack_parts(0, 1000, 100) == 32 and lib.lob_toeplitz_unpack_parts(1, 1000, 300) == 0


def test_lanczos_host_checks_and_no_cpu_fallback():
    from linear_operator_b200.utils.lanczos import lanczos_tridiag

    with pytest.raises(RuntimeError, match="matmul_closure should be a function callable object"):
        lanczos_tridiag(torch.eye(3), 2, dtype=torch.float32, device="cpu", matrix_shape=(3, 3))
    A = torch.eye(8)
    with pytest.raises(_lib.LobError, match="CUDA tensors only"):
        lanczos_tridiag(lambda v: A @ v, 4, dtype=torch.float32, device=torch.device("cpu"), matrix_shape=A.shape,
                        init_vecs=torch.ones(8, 1))


def test_no_cpu_fallback():
    op = DenseLinearOperator(torch.eye(8)).add_jitter(0.5)
    with settings.max_cholesky_size(0):
        with pytest.raises(_lib.LobError, match="CUDA tensors only"):
            op.inv_quad_logdet(torch.ones(8, 1), logdet=True)
    with pytest.raises(_lib.LobError):
        op @ torch.ones(8, 2)


def test_settings_defaults_and_context_managers():
    assert settings.cg_tolerance.value() == 1
    assert settings.max_cg_iterations.value() == 1000
    assert settings.max_lanczos_quadrature_iterations.value() == 20
    assert settings.max_cholesky_size.value() == 800
    assert settings.max_preconditioner_size.value() == 15
    assert settings.min_preconditioning_size.value() == 2000
    assert settings.preconditioner_tolerance.value() == 1e-3
    assert settings.num_trace_samples.value() == 10
    with settings.cg_tolerance(1e-3), settings.max_preconditioner_size(100):
        assert settings.cg_tolerance.value() == 1e-3 and settings.max_preconditioner_size.value() == 100
        with settings.cg_tolerance(0.5):
            assert settings.cg_tolerance.value() == 0.5
        assert settings.cg_tolerance.value() == 1e-3
    assert settings.cg_tolerance.value() == 1
    assert settings.fast_computations.log_prob.on() and settings.skip_logdet_forward.off()
    with settings.fast_computations(log_prob=False):
        assert settings.fast_computations.log_prob.off() and settings.fast_computations.solves.on()
    assert settings.cholesky_jitter.value(torch.float64) == 1e-8


def test_dispatch_table():
    """SURVEY.md section 3.5 (rows that stay inside the built scope)."""
    k = torch.eye(6)
    assert type(DenseLinearOperator(k).add_jitter(0.1)) is AddedDiagLinearOperator
    assert type(DenseLinearOperator(k) + DiagLinearOperator(torch.ones(6))) is AddedDiagLinearOperator
    assert type(DenseLinearOperator(k).add_jitter(0.1)._diag_tensor) is ConstantDiagLinearOperator
    t = ToeplitzLinearOperator(torch.arange(6.0))
    assert type(t.add_jitter(0.1)) is ToeplitzLinearOperator  # column[0] += jitter
    assert type(t + DiagLinearOperator(torch.ones(6))) is AddedDiagLinearOperator
    lr = LowRankRootLinearOperator(torch.ones(6, 2))
    assert type(lr + DiagLinearOperator(torch.ones(6))) is LowRankRootAddedDiagLinearOperator
    assert type(lr.add_jitter(0.1)) is LowRankRootAddedDiagLinearOperator
    assert type(AddedDiagLinearOperator(RootLinearOperator(torch.ones(6, 2)), DiagLinearOperator(torch.ones(6)))) \
        is AddedDiagLinearOperator
    kron = KroneckerProductLinearOperator(torch.eye(2), torch.eye(3))
    assert kron.shape == torch.Size([6, 6])
    assert type(AddedDiagLinearOperator(kron, DiagLinearOperator(torch.ones(6)))) is AddedDiagLinearOperator
    with pytest.raises(RuntimeError):
        AddedDiagLinearOperator(DenseLinearOperator(k), DenseLinearOperator(k))
    with pytest.raises(RuntimeError):
        LowRankRootAddedDiagLinearOperator(DenseLinearOperator(k), DiagLinearOperator(torch.ones(6)))


def test_representation_round_trip_and_shapes():
    a = torch.randn(2, 3, 5, 5)
    op = AddedDiagLinearOperator(DenseLinearOperator(a), DiagLinearOperator(torch.rand(2, 3, 5)))
    rep = op.representation()
    assert len(rep) == 2 and rep[0] is a
    rebuilt = op.representation_tree()(*rep)
    assert type(rebuilt) is AddedDiagLinearOperator and rebuilt.shape == torch.Size([2, 3, 5, 5])
    assert op.batch_shape == torch.Size([2, 3]) and op.matrix_shape == torch.Size([5, 5]) and op.is_square
    assert op.dtype == torch.float32 and op.device.type == "cpu" and op.dim() == 4 and op.size(-1) == 5
    ident = IdentityLinearOperator(5, batch_shape=torch.Size([2]), dtype=torch.float64)
    assert ident.shape == torch.Size([2, 5, 5]) and ident.representation() == ()
    assert type(ident.representation_tree()()) is IdentityLinearOperator
    psd = PsdSumLinearOperator(RootLinearOperator(torch.ones(5, 2)), DiagLinearOperator(torch.ones(5)))
    assert type(psd.representation_tree()(*psd.representation())) is PsdSumLinearOperator
    c = ConstantDiagLinearOperator(torch.tensor([2.0]), diag_shape=4)
    assert c._diag.stride(-1) == 0 and c.shape == torch.Size([4, 4])  # expanded, never materialised
    assert torch.equal(op.detach().representation()[0], a)


def test_shape_errors_match_reference_messages():
    op = DenseLinearOperator(torch.eye(900)).add_jitter(0.5)
    with pytest.raises(RuntimeError, match="cannot be multiplied"):
        op.inv_quad_logdet(torch.ones(899), logdet=True)
    with pytest.raises(RuntimeError, match="same number of dimensions"):
        op.inv_quad_logdet(torch.ones(2, 900, 1), logdet=True)
    with pytest.raises(RuntimeError, match="Either `inv_quad_rhs` or `logdet`"):
        op.inv_quad_logdet(None, logdet=False)
    rect = DenseLinearOperator(torch.ones(900, 901))
    with pytest.raises(RuntimeError, match="square"):
        rect.inv_quad_logdet(torch.ones(901, 1), logdet=True)
    with pytest.raises(RuntimeError, match="square"):
        rect.solve(torch.ones(900, 1))


def test_linear_cg_is_a_patchable_module_attribute():
    """linear_operator/test/linear_operator_test_case.py:555-556 wraps utils.linear_cg in a MagicMock to see whether CG
    ran; LinearOperator._solve must therefore look the symbol up on the module at call time."""
    op = DenseLinearOperator(torch.eye(12))
    rhs = torch.ones(12, 3)
    with mock.patch("linear_operator_b200.utils.linear_cg", return_value=torch.zeros(12, 3)) as spy:
        out = op._solve(rhs, None)
    assert spy.called and out.shape == (12, 3)
    kwargs = spy.call_args.kwargs
    assert kwargs["max_iter"] == 1000 and kwargs["max_tridiag_iter"] == 20 and kwargs["n_tridiag"] == 0


def test_preconditioner_gating_and_torch_function():
    small = DenseLinearOperator(torch.eye(100)).add_jitter(0.5)
    assert small._preconditioner() == (None, None, None)  # N < min_preconditioning_size
    with settings.max_preconditioner_size(0), settings.min_preconditioning_size(10):
        assert small._preconditioner() == (None, None, None)
    with pytest.raises(NotImplementedError, match="is not implemented"):
        torch.trace(small)
    assert torch.equal(torch.diagonal(DiagLinearOperator(torch.arange(4.0)), dim1=-2, dim2=-1), torch.arange(4.0))
    # no tensor requires grad: nothing to differentiate (reference _linear_operator.py:377-378)
    assert small._bilinear_derivative(None, None) == (None, None)

    class NoDerivative(lo.LinearOperator):
        def __init__(self, t):
            super().__init__(t)

    with pytest.raises(NotImplementedError, match="_bilinear_derivative"):
        NoDerivative(torch.eye(3, requires_grad=True))._bilinear_derivative(None, None)
    assert lo.to_dense(torch.eye(2)).shape == (2, 2) and type(lo.to_linear_operator(torch.eye(2))) is DenseLinearOperator


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the unmodified reference from baseline/_ref on the host cores; the oracle port when
    that install is absent) keeps the driver's contract: exactly one JSON line on stdout with the shared metric / config
    keys, impl = reference, a cpu_baseline and a zero-copy e2e object.  Under torchrun (WORLD_SIZE > 1) rank 0 alone
    prints, for the repo arm's global batch, with all host threads although torchrun exports OMP_NUM_THREADS=1."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run(
        [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
         "--cpu-sample-batch", "1", "--n", "300", "--batch", "4"],
        capture_output=True, text=True, timeout=300, check=True,
    ).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "calls/s" and d["value"] > 0
    have_ref = os.path.isdir(os.path.join(root, "baseline", "_ref", "linear_operator"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["config"]["global_batch"] == 4
    # torchrun environment: rank 1 is silent, rank 0 reports the global batch of the repo arm and its real thread count
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--cpu-sample-batch", "1", "--n", "300", "--batch", "4"]
    env = dict(os.environ, WORLD_SIZE="2", RANK="1", LOCAL_RANK="1", OMP_NUM_THREADS="1")
    assert subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=True, env=env).stdout.strip() == ""
    env["RANK"] = env["LOCAL_RANK"] = "0"
    d2 = json.loads(subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=True, env=env).stdout)
    assert d2["config"]["global_batch"] == 8 and d2["n_gpus"] == 2
    ncpu = len(os.sched_getaffinity(0))
    assert d2["cpu_baseline"]["cores"] == ncpu or ncpu == 1

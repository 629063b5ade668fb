"""Lanczos helpers of the Krylov path (reference: utils/lanczos.py).

``lanczos_tridiag_to_diag`` (reference :167-189) runs on the device through ``lob_tridiag_eigh_slq`` -- the reference
moves every T < 32 problem to CPU LAPACK (:179-180).
"""
from __future__ import annotations

from .. import _kernels, settings


def lanczos_tridiag_to_diag(t_mat):
    """``t_mat``: (num_init_vecs, *batch, k, k) tridiagonal.  Returns eigenvalues (num_init_vecs, *batch, k), ascending,
    negative ones replaced by 1, and eigenvectors (num_init_vecs, *batch, k, k) whose columns for negative eigenvalues
    are zeroed (reference :184-187)."""
    if settings.verbose_linalg.on():
        settings.verbose_linalg.logger.debug(f"Running symeig on a matrix of size {t_mat.shape}.")
    out = _kernels.tridiag_eigh_slq(t_mat, t_mat.shape[-1], want_evals=True, want_evecs=True, want_logdet=False)
    return out["evals"], out["evecs"]

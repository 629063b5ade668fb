"""PivotedCholesky Function (reference: functions/_pivoted_cholesky.py:13-105).  The batched pivoted Cholesky
runs in ``csrc/pivchol.cu``.  Dense / Kronecker / Toeplitz operators hand the kernels a device functor as row source, so
the reference's per-step ``__getitem__`` gathers and host synchronisations (:57-98) disappear; every other operator
(Root, Sum, user classes) takes the reference's generic route -- one ``_get_indices`` row gather per step on
device-resident pivot indices -- through the same kernels (``_kernels.pivoted_cholesky_rows``)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from .. import settings


class PivotedCholesky(Function):
    @staticmethod
    def forward(ctx, representation_tree, max_iter, error_tol, *matrix_args):
        matrix = representation_tree(*matrix_args)
        if error_tol is None:
            error_tol = settings.preconditioner_tolerance.value()
        if settings.verbose_linalg.on():
            settings.verbose_linalg.logger.debug(
                f"Running Pivoted Cholesky on a {matrix.shape} RHS for {max_iter} iterations."
            )
        L, perm = matrix._pivoted_cholesky(max_iter, error_tol)
        ctx.representation_tree = representation_tree
        ctx.mark_non_differentiable(perm)
        ctx.save_for_backward(perm, L, *matrix_args)
        return L, perm

    @staticmethod
    def backward(ctx, grad_output, _):
        """Reference :106-150.  The reference rebuilds the factor as ``Krows chol(Krows[:m])^-T`` from the selected
        rows ``Krows = K[pi, pi[:m]]`` and lets autograd differentiate it.  Here the adjoint of that expression is
        written out (triangular inverse in ``lob_tri_inverse``, the tall products in the skinny-matmul kernels); only
        the row gather itself -- ``_get_indices``, pure indexing -- is differentiated by autograd, which is what makes
        the backward generic over operator classes exactly like the reference's.

        With G = grad_L[pi] (pivoted row order), C = L[pi[:m]] (lower triangular), L2 = L[pi[m:]], G1 / G2 alike:
            gK2 = G2 C^-1,   gC = tril(G1 - C^-T (G2^T L2)),
            gKmm = sym(C^-T Phi(C^T gC) C^-1)   (Phi: lower triangle, diagonal halved)."""
        from .. import _kernels

        perm, L, *_matrix_args = ctx.saved_tensors
        if not any(ctx.needs_input_grad[3:]):
            return tuple([None] * (3 + len(_matrix_args)))
        m = L.size(-1)
        n = L.size(-2)
        short = perm[..., :m]
        with torch.no_grad():
            idx = perm.unsqueeze(-1).expand(*perm.shape, m)
            G = torch.gather(grad_output, -2, idx)  # rows in pivoted order
            Lp = torch.gather(L, -2, idx)
            C = Lp[..., :m, :].contiguous()
            Cinv = _kernels.tri_inverse(C)
            G1 = G[..., :m, :]
            if n > m:
                G2, L2 = G[..., m:, :].contiguous(), Lp[..., m:, :].contiguous()
                gK2 = _kernels.matmul_nn(G2, Cinv)  # (N-m, m)
                T = _kernels.tn_matmul(G2, L2)  # G2^T L2, (m, m)
                gC = torch.tril(G1 - _kernels.tn_matmul(Cinv, T))
            else:
                gK2 = G[..., m:, :]
                gC = torch.tril(G1)
            P = torch.tril(_kernels.tn_matmul(C, gC))  # Phi(C^T gC)
            P.diagonal(dim1=-2, dim2=-1).mul_(0.5)
            S = _kernels.matmul_nn(_kernels.tn_matmul(Cinv, P), Cinv)  # C^-T P C^-1
            gKmm = 0.5 * (S + S.mT)
            gKrows = torch.cat([gKmm, gK2], dim=-2)  # (*b, N, m): gradient w.r.t. K[pi, pi[:m]]

        with torch.enable_grad():
            matrix_args = []
            for a in _matrix_args:
                if a.dtype in (torch.float, torch.double, torch.half):
                    a = a.detach().requires_grad_(True)
                matrix_args.append(a)
            matrix = ctx.representation_tree(*matrix_args)
            # Krows = apply_permutation(matrix, full_permutation, short_permutation)  (utils/permutation.py:9-88)
            batch_shape = perm.shape[:-1]
            full_shape = (*batch_shape, n, m)
            batch_idx = []
            for i, sz in enumerate(batch_shape):
                shape = [1] * len(full_shape)
                shape[i] = sz
                batch_idx.append(torch.arange(sz, device=perm.device).reshape(shape).expand(full_shape))
            rows = perm.unsqueeze(-1).expand(full_shape)
            cols = short.unsqueeze(-2).expand(full_shape)
            Krows = matrix._get_indices(rows, cols, *batch_idx)
            leaves = [a for a in matrix_args if a.requires_grad]
            grads = list(torch.autograd.grad(Krows, leaves, grad_outputs=gKrows, allow_unused=True))
        out = []
        for a in matrix_args:
            out.append(grads.pop(0) if a.requires_grad else None)
        return tuple([None, None, None] + out)

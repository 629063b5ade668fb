/*
 * lob_b200.h -- C ABI of liblob_b200.so: sm_100a kernels for the batched Krylov hot path of
 * cornellius-gp/linear_operator (mBCG + stochastic Lanczos quadrature + pivoted-Cholesky preconditioner +
 * structured matmuls).  The reference is pure Python on PyTorch and has no FFI of its own; each entry point below
 * names the reference code (path:line relative to /root/reference/linear_operator) whose arithmetic it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions (every call):
 *   - all pointers are DEVICE pointers unless the parameter name starts with host_; the caller owns every buffer
 *     (inputs, outputs, workspaces); the library never allocates device memory and keeps no pointer after return;
 *   - tensors are dense row-major; vectors are (B, N, C) with C fastest, B = product of the batch dimensions;
 *   - dtype: LOB_F32 or LOB_F64; `stream` is a cudaStream_t passed as void*;
 *   - calls are stream-ordered and never synchronise the host unless the name ends in _sync;
 *   - return value: LOB_OK or a negative error code, message via lob_last_error() (thread local).
 *     Numerical conditions (NaNs, no convergence) are NOT errors: they are reported through device status words
 *     which the Python host turns into the reference's RuntimeError / NumericalWarning.
 */
#ifndef LOB_B200_H
#define LOB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOB_F32 0
#define LOB_F64 1

#define LOB_OK 0
#define LOB_ERR_ARG (-1)
#define LOB_ERR_CUDA (-2)
#define LOB_ERR_UNSUPPORTED (-3)

int lob_version(void);
const char* lob_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
int64_t lob_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Modified batched CG (utils/linear_cg.py:98-359).  The host drives the iteration (the operator's _matmul and the
 * preconditioner are closures on the reference's API), every per-iteration vector/scalar update is fused here.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t B, N, C;          /* batch, operator size, right-hand-side columns                         */
  int32_t dtype;            /* LOB_F32 / LOB_F64                                                     */
  int32_t n_tridiag;        /* S: tridiagonals are recovered for the first S columns (0 = none)      */
  int32_t n_tridiag_iter;   /* T = min(max_tridiag_iter, N)            (linear_cg.py:171)            */
  int32_t max_iter;         /* settings.max_cg_iterations               (linear_cg.py:139-140)       */
  int32_t n_iter;           /* iterations the loop may run              (linear_cg.py:170)           */
  int32_t has_precond;      /* 1: z = M^-1 r is supplied by the caller each iteration                */
  double tolerance;         /* settings.cg_tolerance                    (linear_cg.py:150-151)       */
  double eps;               /* 1e-10 safe-division threshold            (linear_cg.py:103,172)       */
  double stop_updating_after; /* 1e-10 per-column freeze threshold      (linear_cg.py:104,205,300)   */
} lob_cg_params;

/* control words kept in the workspace and copied out by lob_cg_poll_sync */
typedef struct {
  int32_t stop;               /* loop has ended (tolerance reached or NaN)            */
  int32_t tolerance_reached;  /* linear_cg.py:307                                      */
  int32_t iterations;         /* completed iterations (k+1)                            */
  int32_t update_tridiag;     /* linear_cg.py:238,326-327                              */
  int32_t last_tridiag_iter;  /* linear_cg.py:239,329                                  */
  int32_t nan_detected;       /* linear_cg.py:199-200                                  */
  int32_t all_converged_at_start; /* linear_cg.py:207                                  */
  int32_t reserved;
  double residual_norm_mean;  /* for the NumericalWarning text (linear_cg.py:337-347)  */
} lob_cg_status;

size_t lob_cg_workspace_bytes(const lob_cg_params* p);

/* linear_cg.py:177-183: column norms of rhs (zero columns flagged, norm := 1), rhs_n = rhs / norm,
 * x = x0 / norm (x0 may be NULL -> x = 0).  Also resets the control words and zero-fills t_mat (may be NULL,
 * shape (S, B, T, T)). */
int lob_cg_setup(const lob_cg_params* p, void* ws, const void* rhs, const void* x0, void* rhs_n, void* x,
                 void* t_mat, void* stream);

/* linear_cg.py:186-208: r = rhs_n - ax0 (ax0 = A x0, may be NULL when x0 was NULL: r = rhs_n), NaN check,
 * residual norms, has_converged. */
int lob_cg_residual_init(const lob_cg_params* p, void* ws, const void* rhs_n, const void* ax0, void* r, void* stream);

/* linear_cg.py:213-215: rz = sum z*r, p = z.  z may alias r (no preconditioner).  rz_partials: optional
 * (B, n_rz_parts, C) double partial sums of r*z from a preconditioner with a fused epilogue (lob_dense_matmul_ex). */
int lob_cg_direction_init(const lob_cg_params* p, void* ws, const void* r, const void* z, void* pvec,
                          const double* rz_partials, int32_t n_rz_parts, void* stream);

/* linear_cg.py:250-264 (+ :31 x update): alpha = rz / <p,Ap> with the eps rule and the converged mask,
 * r -= alpha Ap, x += alpha p; also accumulates <r,r> for the residual norm.
 * pap_partials: optional (B, n_parts, C) partial sums of p*Ap in double precision written by a fused matmul
 * (lob_dense_matmul); NULL -> computed here. */
int lob_cg_step_xr(const lob_cg_params* p, void* ws, int32_t k, const void* ap, const void* pvec, void* x, void* r,
                   const double* pap_partials, int32_t n_parts, void* stream);

/* linear_cg.py:31-46 + :298-332: beta = <r,z>_new / <r,z>_old (eps rule), p = z + beta p, residual norm / converged
 * flags, stop test, tridiagonal update.  z NULL -> z = r.  t_mat (S,B,T,T) may be NULL when n_tridiag == 0.
 * rz_partials as in lob_cg_direction_init (NULL: <r,z> is reduced here). */
int lob_cg_step_p(const lob_cg_params* p, void* ws, int32_t k, const void* z, const void* r, void* pvec, void* t_mat,
                  const double* rz_partials, int32_t n_rz_parts, void* stream);

/* copies the control words to host memory and synchronises the stream */
int lob_cg_poll_sync(const lob_cg_params* p, void* ws, lob_cg_status* host_status, void* stream);

/* linear_cg.py:335: x *= rhs_norm */
int lob_cg_finish(const lob_cg_params* p, void* ws, void* x, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dense operator matmul: Y = A X (+ d (.) X), A (B|1, M, K) row-major with leading dimension lda,
 * X (B, K, C), Y (B, M, C).   dense_linear_operator.py:60-64 + added_diag_linear_operator.py:72-76.
 * a_batch_stride / d_batch_stride in elements (0 broadcasts over the batch); d may be NULL; d_stride is the element
 * stride inside one diagonal (0 for a constant diagonal stored as one value per batch element).
 * If dots != NULL (requires M == K) it receives (B, lob_dense_matmul_parts(M), C) double partial sums of X*Y over
 * row tiles: the <p, A p> reduction of linear_cg.py:250-251 fused into the matmul epilogue.
 * ---------------------------------------------------------------------------------------------------------- */
int32_t lob_dense_matmul_parts(int64_t M);
/* Scratch for the streaming tensor-core kernel (dense_stream.cu): the tf32 split of X^T, (B, R(C), K) floats.
 * 0 when the shape / dtype is served by a kernel that needs none.  ws may be NULL (or too small): the call then runs
 * on the slower workspace-free CUDA kernels. */
size_t lob_dense_matmul_workspace_bytes(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C);
int lob_dense_matmul(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C, const void* A, int64_t lda,
                     int64_t a_batch_stride, const void* X, void* Y, const void* d, int64_t d_batch_stride,
                     int64_t d_stride, double* dots, void* ws, size_t ws_bytes, void* stream);

/* Test hook: pins the fp32 dense-matmul implementation (0 automatic dispatch, 1 dense_stream2, 2 dense_stream,
 * 5 dense_stream2p (CTA pairs),
 * 3 dense_tc, 4 CUDA cores) so the tests can compare the kernels with each other on the same call.  Every choice
 * computes the same product to fp32 accuracy; nothing in the release library changes numerics through the environment
 * (the harness-only experiment switches of the streaming kernels exist only when compiled with -DLOB_DIAG). */
int lob_debug_pin_dense_impl(int32_t impl);

/* Generalised epilogue:  Y = alpha[b] * (A X) + d (.) E,  dots = per-row-tile partial sums of E * Y.
 * E (B, M, C) may be NULL (then E = X, which needs M == K); alpha (one value per batch element, stride
 * alpha_batch_stride) may be NULL (= 1).  With A = Q, X = Q^T r, E = r, alpha = -1/s, d = 1/s this is the whole
 * precondition_closure z = (r - Q Q^T r)/s of added_diag_linear_operator.py:135-140 with <r, z> (linear_cg.py:35-36)
 * coming out of the epilogue. */
int lob_dense_matmul_ex(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C, const void* A, int64_t lda,
                        int64_t a_batch_stride, const void* X, void* Y, const void* E, const void* alpha,
                        int64_t alpha_batch_stride, const void* d, int64_t d_batch_stride, int64_t d_stride,
                        double* dots, void* ws, size_t ws_bytes, void* stream);

/* Out (B, I, J) = P^T Q with P (B, N, I), Q (B, N, J) row-major: reductions over the long dimension N
 * (Q^T r of added_diag_linear_operator.py:137, L^T L of the preconditioner build, U^T (D^-1 b) of
 * low_rank_root_added_diag_linear_operator.py:77).  Accumulates in double; deterministic two-stage reduction.
 * ws must hold lob_tn_matmul_workspace_bytes(). out_dtype selects the type of Out (LOB_F64 allowed for f32 in). */
size_t lob_tn_matmul_workspace_bytes(int64_t B, int64_t N, int64_t I, int64_t J);
int lob_tn_matmul(int32_t dtype, int32_t out_dtype, int64_t B, int64_t N, int64_t I, int64_t J, const void* P,
                  int64_t p_batch_stride, const void* Q, int64_t q_batch_stride, void* out, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Pivoted-Cholesky preconditioner of AddedDiagLinearOperator
 * ---------------------------------------------------------------------------------------------------------- */
/* functions/_pivoted_cholesky.py:13-105 for a dense operator A (B, N, N): greedy diagonal-pivoted partial Cholesky.
 * Lt (B, rank, N) zero-filled on entry receives row m = step m (the reference's internal layout, :36-42);
 * perm (B, N) int64; m_out (device int32) = number of steps actually taken (common to the batch, :57).
 * ws: lob_pivchol_workspace_bytes().  Row sources other than dense: see the kron / toeplitz variants. */
size_t lob_pivchol_workspace_bytes(int64_t B, int64_t N, int32_t rank);
int lob_pivchol_dense(int32_t dtype, int64_t B, int64_t N, int32_t rank, double error_tol, const void* A, int64_t lda,
                      int64_t a_batch_stride, void* Lt, int64_t* perm, int32_t* m_out, void* ws, void* stream);
/* kronecker_product_linear_operator.py:198-216 row source: up to 4 factors, factor f is (B|1, n_f, n_f) */
int lob_pivchol_kron(int32_t dtype, int64_t B, int32_t n_factors, const int64_t* host_sizes,
                     const void* const* host_factors, const int64_t* host_batch_strides, int32_t rank,
                     double error_tol, void* Lt, int64_t* perm, int32_t* m_out, void* ws, void* stream);
/* toeplitz_linear_operator.py:38-40 row source: T[i,j] = col[|i-j|], col (B|1, N) */
int lob_pivchol_toeplitz(int32_t dtype, int64_t B, int64_t N, const void* col, int64_t col_batch_stride, int32_t rank,
                         double error_tol, void* Lt, int64_t* perm, int32_t* m_out, void* ws, void* stream);

/* Generic operators (RootLinearOperator, SumLinearOperator, user-defined classes): the same pivoted Cholesky with the
 * pivot row of every step supplied by the caller, who evaluates the operator's own `_get_indices`
 * (operators/_linear_operator.py:412-461 through utils/permutation.py:76-87) on the device-resident pivot indices:
 *   lob_pivchol_rows_begin (diag (B,N) = operator._approx_diagonal(), _pivoted_cholesky.py:26-31)
 *   for m in 0..rank-1:  lob_pivchol_rows_pivot  -> pivot_rows (B) int64 = pi_m (:61-70), sqrt on the diagonal (:73-74)
 *                        rows (B,N) = op[b, pi_m[b], :]                         (host: `_get_indices`, :79)
 *                        lob_pivchol_rows_update                                (:80-98)
 *   lob_pivchol_rows_status -> m_out = steps taken, active_out = loop still running (optional early-exit poll)
 * Lt (B, rank, N) zero-filled on entry, perm (B, N) int64, ws = lob_pivchol_workspace_bytes(); the stop rule (:57)
 * lives on the device: after it fires, pivot/update launches are no-ops. */
int lob_pivchol_rows_begin(int32_t dtype, int64_t B, int64_t N, int32_t rank, const void* diag, int64_t* perm,
                           void* ws, void* stream);
int lob_pivchol_rows_pivot(int32_t dtype, int64_t B, int64_t N, int32_t rank, int32_t m, double error_tol, void* Lt,
                           int64_t* perm, int64_t* pivot_rows, void* ws, void* stream);
int lob_pivchol_rows_update(int32_t dtype, int64_t B, int64_t N, int32_t rank, int32_t m, const void* rows, void* Lt,
                            void* ws, void* stream);
int lob_pivchol_rows_status(int32_t dtype, int64_t B, int64_t N, int32_t rank, int32_t* m_out, int32_t* active_out,
                            void* ws, void* stream);

/* (B, R, N) -> (B, N, m) keeping the first m rows: the `L[..., :m, :].mT.contiguous()` of _pivoted_cholesky.py:104 */
int lob_transpose_rows(int32_t dtype, int64_t B, int64_t R, int64_t N, int64_t m, const void* Lt, void* L,
                       void* stream);

/* added_diag_linear_operator.py:144-184 restated through the Gram matrix (R^T R = L^T D^-1 L + I, resp. L^T L + s I):
 * given G (B, k, k) in double (from lob_tn_matmul), adds the diagonal term, factors it in-SM and returns
 * Rinv (B, k, k) in `dtype` (upper triangular inverse: Q = L Rinv) and logdet_r (B) = 2 sum log R_ii.
 * The diagonal added to G is sigma2[b * sigma2_stride] when sigma2 != NULL, else add_identity_scale.
 * info (B) != 0 flags a non-positive pivot.  ws: B*k*k doubles.  k <= 160. */
int lob_precond_factor(int32_t dtype, int64_t B, int32_t k, const double* G, double add_identity_scale,
                       const void* sigma2, int64_t sigma2_stride, void* rinv, void* logdet_r, int32_t* info, void* ws,
                       void* stream);

/* fused elementwise pieces of precondition_closure (added_diag_linear_operator.py:135-140):
 * z = (r - w) * inv_sigma2[b]  (constant diagonal)   or   z = r / d - w   (general diagonal), w = Q (Q^T r). */
int lob_precond_combine(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* r, const void* w, const void* d,
                        int64_t d_batch_stride, int64_t d_stride, int32_t constant_diag, void* z, void* stream);

/* rows scaled: out[b,n,:] = in[b,n,:] * f(d[b,n]),  mode 0: *d, 1: /d, 2: *sqrt(d), 3: /sqrt(d)
 * (D^-1/2 L of added_diag_linear_operator.py:176-178, D^-1 b of low_rank_root_added_diag_linear_operator.py:77) */
int lob_scale_rows(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* in, const void* d,
                   int64_t d_batch_stride, int64_t d_stride, int32_t mode, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Probe vectors, column reductions
 * ---------------------------------------------------------------------------------------------------------- */
/* functions/_inv_quad_logdet.py:107-110 + psd_sum_linear_operator.py:15-18 + diag_linear_operator.py:273-277:
 * z (B,N,S) = z_root (B,N,S, may be NULL) + sqrt(d[b,n]) * eps_diag (S,B,N) (d NULL -> 1); then column norms and
 * normalisation.  probes (B,N,S) and norms (B,1,S) are outputs.  ws: lob_colred_workspace_bytes(B,N,S). */
size_t lob_colred_workspace_bytes(int64_t B, int64_t N, int64_t C);
int lob_probe_assemble(int32_t dtype, int64_t B, int64_t N, int64_t S, const void* z_root, const void* eps_diag,
                       const void* d, int64_t d_batch_stride, int64_t d_stride, void* probes, void* norms, void* ws,
                       void* stream);

/* out[b, j] = sum_n U[b, n, u_off + j] * V[b, n, v_off + j], j < R  (inv_quad = (solves[..., S:] * rhs).sum(-2),
 * functions/_inv_quad_logdet.py:151-153).  U is (B,N,Cu), V is (B,N,Cv). */
int lob_col_dots(int32_t dtype, int64_t B, int64_t N, int64_t R, const void* U, int64_t Cu, int64_t u_off,
                 const void* V, int64_t Cv, int64_t v_off, void* out, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Stochastic Lanczos quadrature (utils/lanczos.py:167-189 + utils/stochastic_lq.py:45-82)
 * ---------------------------------------------------------------------------------------------------------- */
/* t_mat (S, B, T, T) symmetric tridiagonal.  evals (S,B,T) ascending with negative eigenvalues replaced by 1 and
 * their eigenvector columns zeroed; evecs (S,B,T,T) may be NULL (first rows only are needed for the quadrature);
 * logdet (B) = (n / S) sum_j sum_i V_j[0,i]^2 log(lambda_ji) may be NULL.  All arithmetic in double on device
 * (the reference ships T<32 problems to CPU LAPACK, lanczos.py:179-180).  ws: lob_tridiag_workspace_bytes(). */
size_t lob_tridiag_workspace_bytes(int64_t S, int64_t B, int32_t T, int32_t want_evecs);
int lob_tridiag_eigh_slq(int32_t dtype, int64_t S, int64_t B, int32_t T, int64_t n, const void* t_mat, void* evals,
                         void* evecs, void* logdet, int32_t* info, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Structured matmuls
 * ---------------------------------------------------------------------------------------------------------- */
/* kronecker_product_linear_operator.py:34-45: one mode product  Y[b, q, i, c] = sum_j K[b, i, j] X[b, j, q, c]
 * for X viewed as (B, n, Q, C) and Y as (B, Q, n, C): the bmm and the transposing copy of the reference fused. */
int lob_kron_mode_matmul(int32_t dtype, int64_t B, int64_t n, int64_t Q, int64_t C, const void* K,
                         int64_t k_batch_stride, const void* X, void* Y, void* stream);

/* utils/toeplitz.py:131-149 with a length-L (>= 2N-1) real circulant embedding; the FFTs themselves are cuFFT.
 * pad: xt (B, C, L) = [X^T, 0];  embed: c (B, L) = [col, 0.., rev(col[1:])];
 * mul: Fy = Fx * Fc (complex, (B,C,L/2+1) x (B,L/2+1));  unpad: Y (B,N,C) = scale * yt[..., :N]^T (+ d (.) X). */
int lob_toeplitz_pad(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* X, void* xt, void* stream);
int lob_toeplitz_embed(int32_t dtype, int64_t B, int64_t N, int64_t L, const void* col, int64_t col_batch_stride,
                       void* c, void* stream);
int lob_toeplitz_mul(int32_t dtype, int64_t B, int64_t C, int64_t H, const void* fc, int64_t fc_batch_stride, void* fx,
                     void* stream);
int lob_toeplitz_unpad(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* yt, double scale,
                       const void* X, const void* d, int64_t d_batch_stride, int64_t d_stride, void* Y, double* dots,
                       void* stream);
/* dots (optional, both unpad and unpack): (B, parts, C) partial sums of X * Y per row block and column -- linear_cg's
 * <p, A p> (linear_cg.py:250-251) out of the last pass of the product; parts = lob_toeplitz_unpad_parts() (0: this
 * column count has no fused form) resp. lob_toeplitz_unpack_parts() (0: more columns than one shared-memory tile
 * holds -- pack / unpack then refuse the call and the caller splits the column block). */
int32_t lob_toeplitz_unpad_parts(int32_t dtype, int64_t N, int64_t C);
int32_t lob_toeplitz_unpack_parts(int32_t dtype, int64_t N, int64_t C);
/* The same product through complex FFTs of column PAIRS (the symmetric embedding has a real spectrum, so two real
 * columns ride one C2C transform: no real-to-complex pre/post-processing passes, transposes folded into pack/unpack).
 * colmax: maxbits (B, C) = bit patterns of max_n |X| (uint32 / uint64);  pack: zt (B, ceil(C/2), L) complex =
 * transposed column pairs, each column scaled by an exact power of two to [1, 2), zero above N;  mulr: zt *= fr
 * (fr (B|1, L/2+1) real spectrum, indexed min(k, L-k));  unpack: Y (B,N,C) = scale / s_c * zt[..., :N] (+ d (.) X). */
int lob_toeplitz_colmax(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* X, void* maxbits, void* stream);
int lob_toeplitz_pack(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* X, const void* maxbits,
                      void* zt, void* stream);
int lob_toeplitz_mulr(int32_t dtype, int64_t B, int64_t P, int64_t L, const void* fr, int64_t fr_batch_stride, void* zt,
                      void* stream);
int lob_toeplitz_unpack(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* zt, double scale,
                        const void* maxbits, const void* X, const void* d, int64_t d_batch_stride, int64_t d_stride,
                        void* Y, double* dots, void* stream);

/* generic batched small-K product  Y (B, M, C) = A (B|1, M, K) X (B, K, C): Q t, L eps, U w
 * (added_diag_linear_operator.py:137, _linear_operator.py:2784-2791, low_rank_root_added_diag_...py:83).
 * Same kernel as lob_dense_matmul without the fused epilogues; beta_y: Y = A X + beta_y * Y. */
int lob_matmul_nn(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C, const void* A, int64_t lda,
                  int64_t a_batch_stride, const void* X, int64_t x_batch_stride, void* Y, double beta_y, void* stream);

/* Woodbury pieces of low_rank_root_added_diag_linear_operator.py:36-47,62-101: in-SM Cholesky solve of the k x k
 * capacitance system for C right-hand sides: W (B,k,C) <- cap^-1 W, cap = I + G (G (B|1,k,k) double);
 * logdet_cap (B) = 2 sum log diag chol(cap). */
size_t lob_cap_solve_workspace_bytes(int64_t B, int32_t k, int64_t C);
int lob_cap_solve(int32_t dtype, int64_t B, int32_t k, int64_t C, const double* G, int64_t g_batch_stride, void* W,
                  void* logdet_cap, int32_t* info, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Lanczos tridiagonalisation with full re-orthogonalisation (utils/lanczos.py:9-164), one fused kernel per iteration.
 * q_mat (t_cap, B, N, C) and t_mat (t_cap, t_cap, B, C) are the reference's own work arrays (zero-initialised by the
 * caller); w (B, N, C) = A q_k is the closure product of the current iteration.
 *   lob_lanczos_init : q_mat[0] = init / ||init||                                            (lanczos.py:83-84)
 *   lob_lanczos_step : mode 0 (k = 0): alpha_0, beta_0, q_1                                  (:86-100)
 *                      mode 1 (k >= 1): alpha_k and, unless k is the last iteration, r, one Gram-Schmidt pass against
 *                                       q_0..q_k, beta_k, q_{k+1}                             (:103-131, :153)
 *                      mode 2: one more re-orthogonalisation of q_{k+1}                       (:141-147)
 * flags (2 x int32, device): [0] some <q_j, q_{k+1}> > tol (signed, as the reference :133,:147) -- rewritten by every
 * mode 1 / 2 launch; [1] some |beta_k| > 1e-6 (:150) -- rewritten by mode 0 / 1.  The host reads them to take the
 * reference's own control decisions.
 * ---------------------------------------------------------------------------------------------------------- */
int lob_lanczos_init(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* init, void* q0, void* stream);
int lob_lanczos_step(int32_t dtype, int32_t mode, int64_t B, int64_t N, int64_t C, int32_t t_cap, int32_t k,
                     const void* w, void* q_mat, void* t_mat, int32_t* flags, double tol, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Batched fp32-accurate GEMM on tcgen05 (3xTF32, fp32 accumulation in TMEM, TMA-fed):
 *     D[b] (M x N) = alpha * A[b / a_batch_div] B[b / b_batch_div]   (+ row scalings / E, see below)
 * for the products of the path that are plain GEMMs: the tall products of the low-rank Woodbury solve with the batch
 * as the row dimension (low_rank_root_added_diag_linear_operator.py:62-87), the Kronecker mode products
 * (kronecker_product_linear_operator.py:34-45), Q = L R^-1 (added_diag_linear_operator.py:164-166).
 * A is (M x K): a_mn == 0: element (m,k) at A[m*lda + k] ("K-major"), a_mn == 1: A[k*lda + m] ("MN-major").
 * B is (K x N): b_mn == 1: element (k,n) at B[k*ldb + n] (row-major K x N), b_mn == 0: B[n*ldb + k].
 * Epilogue: D[m][n] = ra(m) * acc + rb(m) * E[m][n] with ra(m) = row_alpha ? row_alpha[b*stride + m] : alpha and
 * rb(m) = row_beta ? row_beta[b*stride + m] : 1 (E NULL: no second term).  d_trans != 0 stores the TRANSPOSE: element
 * (m,n) at D[n*ldd + m], E read likewise, and the two factor vectors are then indexed by n -- pick the operand roles so
 * that the memory-contiguous index of the result is m (the tensor core's lane index): every store is a full line.  Split-K (requested_splits: 0 = automatic,
 * lob_gemm3x_splits() tells what will be used) needs ws = lob_gemm3x_workspace_bytes(); d_dtype LOB_F64 is available on
 * the split-K path (the reduction runs in double).  fp32 operands only; leading dimensions and batch strides must be
 * multiples of 4 elements and bases 16-byte aligned (TMA) -- otherwise LOB_ERR_UNSUPPORTED and the caller uses
 * lob_matmul_nn / lob_tn_matmul.
 * ---------------------------------------------------------------------------------------------------------- */
int32_t lob_gemm3x_splits(int64_t batch, int64_t M, int64_t N, int64_t K, int32_t requested_splits);
size_t lob_gemm3x_workspace_bytes(int64_t batch, int64_t M, int64_t N, int64_t K, int32_t requested_splits);
int lob_gemm3x(int64_t batch, int64_t M, int64_t N, int64_t K, const void* A, int32_t a_mn, int64_t lda,
               int64_t a_batch_stride, int64_t a_batch_div, const void* B, int32_t b_mn, int64_t ldb,
               int64_t b_batch_stride, int64_t b_batch_div, void* D, int32_t d_dtype, int64_t ldd,
               int64_t d_batch_stride, double alpha, const void* row_alpha, int64_t row_alpha_batch_stride,
               const void* E, int64_t lde, int64_t e_batch_stride, const void* row_beta, int64_t row_beta_batch_stride,
               int32_t d_trans, int32_t requested_splits, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Backward pass (functions/_inv_quad_logdet.py:163-226, _solve.py:70-131, _inv_quad.py:63-93,
 * _pivoted_cholesky.py:106-150)
 * ---------------------------------------------------------------------------------------------------------- */
/* DenseLinearOperator._bilinear_derivative (dense_linear_operator.py:69-71) with the per-column scalings of the
 * callers folded in:  out[b,i,j] (+)= sum_c w[b,c] * left[b,i,c] * right[b,j,c];  left (B,N,C), right (B,M,C),
 * w (B,C) or NULL (= 1), out (B,N,M); accumulate != 0 adds to out. */
int lob_bilinear_dense(int32_t dtype, int64_t B, int64_t N, int64_t M, int64_t C, const void* left, const void* right,
                       const void* w, void* out, int32_t accumulate, void* stream);
/* DiagLinearOperator._bilinear_derivative (diag_linear_operator.py:37-45): out[b,n] = sum_c w[b,c] left[b,n,c] right[b,n,c] */
int lob_bilinear_diag(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* left, const void* right,
                      const void* w, void* out, void* stream);
/* out (B,k,k) = inverse of the lower-triangular k x k blocks Cm[b] (leading dimension ldc, batch stride in elements):
 * the triangular-solve adjoint of PivotedCholesky.backward (_pivoted_cholesky.py:128-137). */
int lob_tri_inverse(int32_t dtype, int64_t B, int32_t k, const void* Cm, int64_t ldc, int64_t c_batch_stride, void* out,
                    void* stream);
/* sym_toeplitz_derivative_quadratic_form (utils/toeplitz.py:164-204) in the frequency domain: fu, fv (B,C,H) complex
 * spectra (cuFFT R2C of the zero-padded left / right vectors), out (B,H) complex = sum_c w[b,c] 2 Re(conj(fu) fv);
 * finish: out (B,N) = scale * y[b, :N] with element 0 halved (y (B,L) = C2R of the spectrum). */
int lob_toeplitz_cross_spectrum(int32_t dtype, int64_t B, int64_t C, int64_t H, const void* fu, const void* fv,
                                const void* w, void* out, void* stream);
int lob_toeplitz_deriv_finish(int32_t dtype, int64_t B, int64_t N, int64_t L, const void* y, double scale, void* out,
                              void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Shifted MINRES (utils/minres.py:10-282), the solver of contour-integral quadrature (utils/contour_integral_quad.py).
 * Vectors are (B, N, C), per-column scalars (B, C), per-shift scalars (Q, B, C), shifts (Q, B), search directions and
 * solutions (Q, B, N, C).  One iteration = operator closure + 2 x lob_col_dots + these three launches:
 *   lob_minres_z        prod <- prod - alpha z_prev1 - beta_prev z_prev2                         (minres.py:137)
 *   lob_minres_scalars  beta_curr = max(sqrt(beta_sq), eps) and the Givens / QR recurrences per shift (:141-142,:231-245)
 *   lob_minres_update   z /= beta_curr, q /= beta_curr (skipped when q == z), search_curr, solution += ... (:146-147,:246-251)
 * ---------------------------------------------------------------------------------------------------------- */
int lob_minres_z(int32_t dtype, int64_t B, int64_t N, int64_t C, void* prod, const void* z1, const void* z2,
                 const void* alpha, const void* beta_prev, void* stream);
int lob_minres_scalars(int32_t dtype, int64_t Q, int64_t B, int64_t C, const void* shifts, const void* alpha,
                       const void* beta_prev, const void* beta_sq, void* beta_curr, const void* cos_prev2,
                       const void* sin_prev2, const void* cos_prev1, const void* sin_prev1, void* cos_curr, void* sin_curr,
                       void* scale_prev, void* scale_curr, void* sub_diag, void* subsub_diag, void* diag, double eps,
                       void* stream);
int lob_minres_update(int32_t dtype, int64_t Q, int64_t B, int64_t N, int64_t C, void* z, void* q, const void* beta_curr,
                      const void* q_prev1, const void* search_prev1, const void* search_prev2, void* search_curr,
                      void* solution, const void* sub_diag, const void* subsub_diag, const void* diag,
                      const void* scale_prev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LOB_B200_H */

// minres.cu -- fused pieces of shifted MINRES (reference: utils/minres.py:10-282), the solver behind contour-integral
// quadrature (utils/contour_integral_quad.py): (K * value + shift_q I) x_q = b for Q shifts at once.
//
// The reference spends ~45 ATen launches per iteration on (B, N, C) and (Q, B, N, C) tensors.  Here an iteration is the
// operator closure, two deterministic column reductions (lob_col_dots) and three launches:
//   k_minres_z        z = prod - alpha z_prev1 - beta_prev z_prev2                                  (:137)
//   k_minres_scalars  beta = max(sqrt(<z, q>), eps) and, per shift, the Givens rotation / QR update of the
//                     tridiagonal: sub-sub-diagonal, sub-diagonal, diagonal terms, next cos / sin, scale (:231-266)
//   k_minres_update   z /= beta, q /= beta;  search = (q_prev1 - sub search_prev1 - subsub search_prev2) / diag;
//                     solution += search * scale                                                    (:146-147,:268-276)
// All per-column scalars live in (Q, B, C) / (B, C) device arrays; the host only rotates pointers and, every 10th
// iteration, reads the one convergence number the reference's stop rule needs (:177-182).
#include "common.cuh"

namespace lob {

template <typename T>
__global__ void __launch_bounds__(256)
k_minres_z(int64_t total, int64_t N, int64_t C, T* __restrict__ prod, const T* __restrict__ z1, const T* __restrict__ z2,
           const T* __restrict__ alpha, const T* __restrict__ beta_prev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t c = i % C, b = i / (N * C);
  const int64_t s = b * C + c;
  // prod.addcmul_(alpha, z1, value=-1).addcmul_(beta_prev, z2, value=-1): two roundings, in this order
  T v = prod[i] - alpha[s] * z1[i];
  v = v - beta_prev[s] * z2[i];
  prod[i] = v;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_minres_scalars(int64_t Q, int64_t BC, int64_t C, const T* __restrict__ shifts /* (Q, B) */,
                 const T* __restrict__ alpha, const T* __restrict__ beta_prev, const T* __restrict__ bsq,
                 T* __restrict__ beta_curr, const T* __restrict__ cos2, const T* __restrict__ sin2,
                 const T* __restrict__ cos1, const T* __restrict__ sin1, T* __restrict__ cos_c, T* __restrict__ sin_c,
                 T* __restrict__ scale_prev, T* __restrict__ scale_curr, T* __restrict__ sub, T* __restrict__ subsub,
                 T* __restrict__ diag, double eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q * BC) return;
  const int64_t q = i / BC, s = i - q * BC, b = s / C;
  T be = (T)sqrt((double)bsq[s]);  // beta_curr.sqrt_().clamp_min_(eps)  (:141-142)
  if (!(be >= (T)eps)) be = (be != be) ? be : (T)eps;
  if (q == 0) beta_curr[s] = be;
  const T bp = beta_prev[s];
  const T ss = sin2[i] * bp;                 // subsub_diag_term (:231)
  T sd = cos2[i] * bp;                       // sub_diag_term    (:232)
  const T ash = alpha[s] + shifts[q * (BC / C) + b];  // (:234)
  T dg = ash * cos1[i] - sin1[i] * sd;       // (:236)
  sd = sd * cos1[i] + sin1[i] * ash;         // (:237)
  const T radius = (T)sqrt((double)(dg * dg + be * be));  // (:239)
  const T cc = dg / radius, sc = be / radius;  // (:240-241)
  dg = dg * cc + sc * be;                    // (:242)
  cos_c[i] = cc;
  sin_c[i] = sc;
  scale_curr[i] = -(scale_prev[i] * sc);     // (:244)
  scale_prev[i] = scale_prev[i] * cc;        // (:245)
  sub[i] = sd;
  subsub[i] = ss;
  diag[i] = dg;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_minres_update(int64_t Q, int64_t B, int64_t N, int64_t C, T* __restrict__ z, T* __restrict__ qv,
                const T* __restrict__ beta_curr, const T* __restrict__ q1, const T* __restrict__ search1,
                const T* __restrict__ search2, T* __restrict__ search_c, T* __restrict__ solution,
                const T* __restrict__ sub, const T* __restrict__ subsub, const T* __restrict__ diag,
                const T* __restrict__ scale_prev, int normalise_q) {
  const int64_t per = B * N * C;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per) return;
  const int64_t c = i % C, b = i / (N * C);
  const int64_t s = b * C + c;
  const T be = beta_curr[s];
  z[i] = z[i] / be;                                   // (:146)
  if (normalise_q) qv[i] = qv[i] / be;                // (:147); without a preconditioner q aliases z
  const T qp = q1[i];
  for (int64_t q = 0; q < Q; ++q) {
    const int64_t si = q * (B * C) + s, vi = q * per + i;
    T sc = qp - sub[si] * search1[vi];                // (:246)
    sc = sc - subsub[si] * search2[vi];               // (:247)
    sc = sc / diag[si];                               // (:248)
    search_c[vi] = sc;
    solution[vi] = solution[vi] + sc * scale_prev[si];  // (:250-251)
  }
}

}  // namespace lob

using namespace lob;

extern "C" int lob_minres_z(int32_t dtype, int64_t B, int64_t N, int64_t C, void* prod, const void* z1, const void* z2,
                            const void* alpha, const void* beta_prev, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0, "lob_minres_z: sizes must be positive");
  LOB_REQUIRE(prod && z1 && z2 && alpha && beta_prev, "lob_minres_z: NULL pointer");
  const int64_t total = B * N * C;
  LOB_DISPATCH_DTYPE(dtype, {
    k_minres_z<scalar_t><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        total, N, C, (scalar_t*)prod, (const scalar_t*)z1, (const scalar_t*)z2, (const scalar_t*)alpha,
        (const scalar_t*)beta_prev);
  });
  return check_launch("k_minres_z");
}

extern "C" int lob_minres_scalars(int32_t dtype, int64_t Q, int64_t B, int64_t C, const void* shifts, const void* alpha,
                                  const void* beta_prev, const void* beta_sq, void* beta_curr, const void* cos_prev2,
                                  const void* sin_prev2, const void* cos_prev1, const void* sin_prev1, void* cos_curr,
                                  void* sin_curr, void* scale_prev, void* scale_curr, void* sub_diag, void* subsub_diag,
                                  void* diag, double eps, void* stream) {
  LOB_REQUIRE(Q > 0 && B > 0 && C > 0, "lob_minres_scalars: sizes must be positive");
  LOB_REQUIRE(shifts && alpha && beta_prev && beta_sq && beta_curr && cos_prev2 && sin_prev2 && cos_prev1 && sin_prev1 &&
                  cos_curr && sin_curr && scale_prev && scale_curr && sub_diag && subsub_diag && diag,
              "lob_minres_scalars: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    k_minres_scalars<scalar_t><<<(unsigned)cdiv(Q * B * C, 256), 256, 0, (cudaStream_t)stream>>>(
        Q, B * C, C, (const scalar_t*)shifts, (const scalar_t*)alpha, (const scalar_t*)beta_prev,
        (const scalar_t*)beta_sq, (scalar_t*)beta_curr, (const scalar_t*)cos_prev2, (const scalar_t*)sin_prev2,
        (const scalar_t*)cos_prev1, (const scalar_t*)sin_prev1, (scalar_t*)cos_curr, (scalar_t*)sin_curr,
        (scalar_t*)scale_prev, (scalar_t*)scale_curr, (scalar_t*)sub_diag, (scalar_t*)subsub_diag, (scalar_t*)diag, eps);
  });
  return check_launch("k_minres_scalars");
}

extern "C" int lob_minres_update(int32_t dtype, int64_t Q, int64_t B, int64_t N, int64_t C, void* z, void* q,
                                 const void* beta_curr, const void* q_prev1, const void* search_prev1,
                                 const void* search_prev2, void* search_curr, void* solution, const void* sub_diag,
                                 const void* subsub_diag, const void* diag, const void* scale_prev, void* stream) {
  LOB_REQUIRE(Q > 0 && B > 0 && N > 0 && C > 0, "lob_minres_update: sizes must be positive");
  LOB_REQUIRE(z && q && beta_curr && q_prev1 && search_prev1 && search_prev2 && search_curr && solution && sub_diag &&
                  subsub_diag && diag && scale_prev,
              "lob_minres_update: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    k_minres_update<scalar_t><<<(unsigned)cdiv(B * N * C, 256), 256, 0, (cudaStream_t)stream>>>(
        Q, B, N, C, (scalar_t*)z, (scalar_t*)q, (const scalar_t*)beta_curr, (const scalar_t*)q_prev1,
        (const scalar_t*)search_prev1, (const scalar_t*)search_prev2, (scalar_t*)search_curr, (scalar_t*)solution,
        (const scalar_t*)sub_diag, (const scalar_t*)subsub_diag, (const scalar_t*)diag, (const scalar_t*)scale_prev,
        q == z ? 0 : 1);
  });
  return check_launch("k_minres_update");
}

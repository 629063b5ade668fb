// tridiag.cu -- eigendecomposition of the Lanczos tridiagonals + stochastic Lanczos quadrature on the device.
// Reference: utils/lanczos.py:167-189 (eigh, negative-eigenvalue masking) and utils/stochastic_lq.py:45-82
// (logdet ~ (N/S) sum_probes sum_i V[0,i]^2 log lambda_i).  The reference ships every T<32 problem to CPU LAPACK
// and loops over probes in Python; here one thread owns one tridiagonal (S*B of them, e.g. 32768 for BASELINE
// config 2), runs the implicit-shift QL iteration in double precision and keeps only the first eigenvector row
// unless the caller asks for the full eigenvector matrices.  Per-thread arrays live in a workspace laid out with
// the matrix index fastest, so every access of a warp is coalesced.  Small tridiagonals (T <= 48) keep d, e and the
// first eigenvector row in shared memory instead (same layout, thread index fastest): the QL iteration is one long
// dependent chain, and with a handful of matrices (16 at BASELINE config 1) every workspace access was an exposed L2
// round trip -- 256 us for sixteen 20 x 20 problems.
#include "common.cuh"

namespace lob {

__device__ __forceinline__ double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// d, e, z: element i of matrix `mat` lives at [i * nmat + mat]
template <typename T>
__global__ void k_tridiag_eig(int64_t nmat, int Tn, const T* __restrict__ t_mat, T* __restrict__ evals,
                              T* __restrict__ evecs, double* __restrict__ quad, int32_t* __restrict__ info,
                              double* __restrict__ wd, double* __restrict__ we, double* __restrict__ wz0,
                              double* __restrict__ wz, int use_smem) {
  extern __shared__ double tri_smem[];  // use_smem: [3][Tn][blockDim.x]
  const int64_t mat = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= nmat) return;
  const T* t = t_mat + mat * Tn * Tn;
  const int64_t st_ = use_smem ? (int64_t)blockDim.x : nmat;  // distance between consecutive elements of one matrix
  const int64_t ix_ = use_smem ? (int64_t)threadIdx.x : mat;
  double* pd = use_smem ? tri_smem : wd;
  double* pe = use_smem ? tri_smem + (size_t)Tn * blockDim.x : we;
  double* pz0 = use_smem ? tri_smem + 2 * (size_t)Tn * blockDim.x : wz0;
#define D(i) pd[(int64_t)(i) * st_ + ix_]
#define E(i) pe[(int64_t)(i) * st_ + ix_]
#define Z0(i) pz0[(int64_t)(i) * st_ + ix_]
#define Z(r, c) wz[((int64_t)(r) * Tn + (c)) * nmat + mat]
  const int n = Tn;
  for (int i = 0; i < n; ++i) {
    D(i) = (double)t[i * n + i];
    E(i) = (i + 1 < n) ? (double)t[(i + 1) * n + i] : 0.0;
    Z0(i) = (i == 0) ? 1.0 : 0.0;
  }
  if (wz)
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) Z(r, c) = (r == c) ? 1.0 : 0.0;
  bool failed = false;
  for (int l = 0; l < n && !failed; ++l) {
    int iter = 0;
    int m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(D(m)) + fabs(D(m + 1));
        if (fabs(E(m)) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 80) {
          failed = true;
          break;
        }
        double g = (D(l + 1) - D(l)) / (2.0 * E(l));
        double r = hypot(g, 1.0);
        g = D(m) - D(l) + E(l) / (g + sign_of(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * E(i);
          const double b = c * E(i);
          r = hypot(f, g);
          E(i + 1) = r;
          if (r == 0.0) {
            D(i + 1) -= p;
            E(m) = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = D(i + 1) - p;
          r = (D(i) - g) * s + 2.0 * c * b;
          p = s * r;
          D(i + 1) = g + p;
          g = c * r - b;
          // rotate eigenvector rows
          f = Z0(i + 1);
          Z0(i + 1) = s * Z0(i) + c * f;
          Z0(i) = c * Z0(i) - s * f;
          if (wz) {
            for (int k = 0; k < n; ++k) {
              const double fk = Z(k, i + 1);
              Z(k, i + 1) = s * Z(k, i) + c * fk;
              Z(k, i) = c * Z(k, i) - s * fk;
            }
          }
        }
        if (r == 0.0 && i >= l) continue;
        D(l) -= p;
        E(l) = g;
        E(m) = 0.0;
      }
    } while (m != l);
  }
  // ascending order (selection sort; eigh returns ascending eigenvalues)
  for (int i = 0; i < n - 1; ++i) {
    int k = i;
    double p = D(i);
    for (int j = i + 1; j < n; ++j)
      if (D(j) < p) {
        k = j;
        p = D(j);
      }
    if (k != i) {
      D(k) = D(i);
      D(i) = p;
      const double z = Z0(i);
      Z0(i) = Z0(k);
      Z0(k) = z;
      if (wz)
        for (int r = 0; r < n; ++r) {
          const double zz = Z(r, i);
          Z(r, i) = Z(r, k);
          Z(r, k) = zz;
        }
    }
  }
  // masking (lanczos.py:184-187) + quadrature (stochastic_lq.py:74-80)
  double q = 0.0;
  for (int i = 0; i < n; ++i) {
    double lam = D(i);
    const bool keep = lam >= 0.0;
    const double z0 = keep ? Z0(i) : 0.0;
    if (!keep) lam = 1.0;
    if (evals) evals[mat * n + i] = (T)lam;
    if (evecs)
      for (int r = 0; r < n; ++r) evecs[(mat * n + r) * n + i] = keep ? (T)Z(r, i) : (T)0;
    q += z0 * z0 * log(lam);
  }
  if (failed) q = NAN;
  quad[mat] = q;
  if (failed) atomicOr(info, 1);
#undef D
#undef E
#undef Z0
#undef Z
}

template <typename T>
__global__ void k_slq_reduce(int64_t S, int64_t B, double scale, const double* __restrict__ quad,
                             T* __restrict__ logdet) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int64_t j = 0; j < S; ++j) s += scale * quad[j * B + b];
  logdet[b] = (T)s;
}

}  // namespace lob

using namespace lob;

extern "C" size_t lob_tridiag_workspace_bytes(int64_t S, int64_t B, int32_t T, int32_t want_evecs) {
  if (S <= 0 || B <= 0 || T <= 0) return 0;
  const size_t nmat = (size_t)S * B;
  size_t n = nmat * (1 + 3 * (size_t)T);
  if (want_evecs) n += nmat * (size_t)T * T;
  return n * sizeof(double) + 256;
}

extern "C" int lob_tridiag_eigh_slq(int32_t dtype, int64_t S, int64_t B, int32_t T, int64_t n, const void* t_mat,
                                    void* evals, void* evecs, void* logdet, int32_t* info, void* ws, void* stream) {
  LOB_REQUIRE(S > 0 && B > 0 && T > 0, "lob_tridiag_eigh_slq: sizes must be positive");
  LOB_REQUIRE(t_mat && info && ws, "lob_tridiag_eigh_slq: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nmat = S * B;
  double* quad = (double*)ws;
  double* wd = quad + nmat;
  double* we = wd + nmat * T;
  double* wz0 = we + nmat * T;
  double* wz = evecs ? wz0 + nmat * T : nullptr;
  LOB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  LOB_DISPATCH_DTYPE(dtype, {
    const int use_smem = T <= 48;
    const int threads = use_smem ? 64 : 128;
    const size_t smem = use_smem ? sizeof(double) * 3 * (size_t)T * threads : 0;
    if (smem > 48 * 1024)
      LOB_CUDA(cudaFuncSetAttribute(k_tridiag_eig<scalar_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tridiag_eig<scalar_t><<<(unsigned)cdiv(nmat, threads), threads, smem, st>>>(
        nmat, T, (const scalar_t*)t_mat, (scalar_t*)evals, (scalar_t*)evecs, quad, info, wd, we, wz0, wz, use_smem);
    LOB_TRY(check_launch("k_tridiag_eig"));
    if (logdet) {
      k_slq_reduce<scalar_t><<<(unsigned)cdiv(B, 256), 256, 0, st>>>(S, B, (double)n / (double)S, quad,
                                                                     (scalar_t*)logdet);
      LOB_TRY(check_launch("k_slq_reduce"));
    }
  });
  return LOB_OK;
}

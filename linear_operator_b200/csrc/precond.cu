// precond.cu -- pivoted-Cholesky preconditioner of AddedDiagLinearOperator after the factor L is known
// (reference: operators/added_diag_linear_operator.py:95-184).
//
// The reference builds a thin QR of the stacked (N+k) x k matrix [L; sqrt(s) I] (resp. [D^-1/2 L; I]) and applies
// M^-1 v = (v - Q1 Q1^T v)/s.  Q1 Q1^T is the orthogonal projector onto range(L) in that metric, so it is fixed by
// R^T R = L^T L + s I alone:  Q1 = L R^-1.  We therefore accumulate the k x k Gram matrix in double on the device
// (lob_tn_matmul), factor it here with one CTA per batch element, return R^-1 and 2 sum log R_ii, and form Q1 with
// one skinny matmul.  Doing the k x k part in double keeps Q1 orthonormal to fp32 round-off (plain fp32
// CholeskyQR would lose cond(G) * eps).
#include "common.cuh"

namespace lob {

template <typename T>
__global__ void __launch_bounds__(256)
k_precond_factor(int k, const double* __restrict__ G, double add_scale, const T* __restrict__ sigma2,
                 int64_t sigma2_stride, T* __restrict__ rinv, T* __restrict__ logdet_r, int32_t* __restrict__ info,
                 double* __restrict__ xws) {
  extern __shared__ double a[];  // k x (k+1), lower Cholesky factor built in place
  __shared__ double scratch[32];
  __shared__ int bad;
  const int64_t b = blockIdx.x;
  const int ld = k + 1;
  const int tid = threadIdx.x;
  const double add = sigma2 ? (double)sigma2[b * sigma2_stride] : add_scale;
  const double* g = G + b * k * k;
  if (tid == 0) bad = 0;
  for (int e = tid; e < k * k; e += blockDim.x) {
    const int i = e / k, j = e % k;
    a[i * ld + j] = g[e] + (i == j ? add : 0.0);
  }
  __syncthreads();
  // right-looking Cholesky, lower triangle
  for (int j = 0; j < k; ++j) {
    if (tid == 0) {
      const double p = a[j * ld + j];
      if (!(p > 0.0)) bad = 1;
      a[j * ld + j] = sqrt(p);
    }
    __syncthreads();
    const double djj = a[j * ld + j];
    for (int i = j + 1 + tid; i < k; i += blockDim.x) a[i * ld + j] /= djj;
    __syncthreads();
    // trailing update: a[i][m] -= a[i][j] * a[m][j] for j < m <= i
    const int rem = k - j - 1;
    for (int e = tid; e < rem * rem; e += blockDim.x) {
      const int i = j + 1 + e / rem, m = j + 1 + e % rem;
      if (m <= i) a[i * ld + m] -= a[i * ld + j] * a[m * ld + j];
    }
    __syncthreads();
  }
  // logdet = 2 sum log diag
  double acc = 0.0;
  for (int i = tid; i < k; i += blockDim.x) acc += log(a[i * ld + i]);
  const double ls = block_sum(acc, scratch);
  if (tid == 0) {
    logdet_r[b] = (T)(2.0 * ls);
    info[b] = bad;
  }
  // inverse of the lower factor, column j by thread j: X = Lc^-1 ; Rinv = X^T (upper)
  double* x = xws + b * k * k;  // x[j*k + i] = X[i][j]
  for (int j = tid; j < k; j += blockDim.x) {
    for (int i = 0; i < j; ++i) x[j * k + i] = 0.0;
    for (int i = j; i < k; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int m = j; m < i; ++m) s -= a[i * ld + m] * x[j * k + m];
      x[j * k + i] = s / a[i * ld + i];
    }
    for (int i = 0; i < k; ++i) rinv[b * k * k + j * k + i] = (T)x[j * k + i];
  }
}

// z = (r - w) / s[b]   or   z = r / d - w
template <typename T>
__global__ void k_precond_combine(int64_t N, int64_t C, const T* __restrict__ r, const T* __restrict__ w,
                                  const T* __restrict__ d, int64_t d_bs, int64_t d_st, int constant, T* __restrict__ z,
                                  int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / (N * C);
  const int64_t n = (idx / C) % N;
  if (constant) {
    const T inv = (T)1 / d[b * d_bs];
    z[idx] = inv * (r[idx] - w[idx]);
  } else {
    z[idx] = r[idx] / d[b * d_bs + n * d_st] - w[idx];
  }
}

template <typename T>
__global__ void k_scale_rows(int64_t N, int64_t C, const T* __restrict__ in, const T* __restrict__ d, int64_t d_bs,
                             int64_t d_st, int mode, T* __restrict__ out, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / (N * C);
  const int64_t n = (idx / C) % N;
  const T dv = d[b * d_bs + n * d_st];
  T v = in[idx];
  switch (mode) {
    case 0: v = v * dv; break;
    case 1: v = v / dv; break;
    case 2: v = v * (T)sqrt((double)dv); break;
    default: v = v / (T)sqrt((double)dv); break;
  }
  out[idx] = v;
}

// (B, R, N) -> (B, N, m): out[b, n, j] = in[b, j, n], j < m
template <typename T>
__global__ void k_transpose_rows(int64_t R, int64_t N, int64_t m, const T* __restrict__ in, T* __restrict__ out) {
  __shared__ T tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t n0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int jj = ty; jj < 32; jj += 8) {
    const int64_t j = j0 + jj, n = n0 + tx;
    tile[jj][tx] = (j < m && n < N) ? in[(b * R + j) * N + n] : (T)0;
  }
  __syncthreads();
  for (int nn = ty; nn < 32; nn += 8) {
    const int64_t n = n0 + nn, j = j0 + tx;
    if (n < N && j < m) out[(b * N + n) * m + j] = tile[tx][nn];
  }
}

}  // namespace lob

using namespace lob;

extern "C" int lob_precond_factor(int32_t dtype, int64_t B, int32_t k, const double* G, double add_identity_scale,
                                  const void* sigma2, int64_t sigma2_stride, void* rinv, void* logdet_r, int32_t* info,
                                  void* ws, void* stream) {
  LOB_REQUIRE(B > 0 && k > 0, "lob_precond_factor: sizes must be positive");
  LOB_REQUIRE(k <= 160, "lob_precond_factor: rank > 160 not supported");
  LOB_REQUIRE(G && rinv && logdet_r && info && ws, "lob_precond_factor: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = sizeof(double) * k * (k + 1);
  LOB_DISPATCH_DTYPE(dtype, {
    auto kern = k_precond_factor<scalar_t>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)B, 256, smem, st>>>(k, G, add_identity_scale, (const scalar_t*)sigma2, sigma2_stride,
                                         (scalar_t*)rinv, (scalar_t*)logdet_r, info, (double*)ws);
    return check_launch("k_precond_factor");
  });
}

extern "C" int lob_precond_combine(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* r, const void* w,
                                   const void* d, int64_t d_batch_stride, int64_t d_stride, int32_t constant_diag,
                                   void* z, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0, "lob_precond_combine: sizes must be positive");
  LOB_REQUIRE(r && w && d && z, "lob_precond_combine: NULL pointer");
  const int64_t total = B * N * C;
  LOB_DISPATCH_DTYPE(dtype, {
    k_precond_combine<scalar_t><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        N, C, (const scalar_t*)r, (const scalar_t*)w, (const scalar_t*)d, d_batch_stride, d_stride, constant_diag,
        (scalar_t*)z, total);
    return check_launch("k_precond_combine");
  });
}

extern "C" int lob_scale_rows(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* in, const void* d,
                              int64_t d_batch_stride, int64_t d_stride, int32_t mode, void* out, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0, "lob_scale_rows: sizes must be positive");
  LOB_REQUIRE(in && d && out, "lob_scale_rows: NULL pointer");
  LOB_REQUIRE(mode >= 0 && mode <= 3, "lob_scale_rows: bad mode");
  const int64_t total = B * N * C;
  LOB_DISPATCH_DTYPE(dtype, {
    k_scale_rows<scalar_t><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        N, C, (const scalar_t*)in, (const scalar_t*)d, d_batch_stride, d_stride, mode, (scalar_t*)out, total);
    return check_launch("k_scale_rows");
  });
}

extern "C" int lob_transpose_rows(int32_t dtype, int64_t B, int64_t R, int64_t N, int64_t m, const void* Lt, void* L,
                                  void* stream) {
  LOB_REQUIRE(B > 0 && R > 0 && N > 0 && m > 0 && m <= R, "lob_transpose_rows: bad sizes");
  LOB_REQUIRE(B <= 65535, "lob_transpose_rows: flattened batch > 65535 not supported");
  LOB_REQUIRE(Lt && L, "lob_transpose_rows: NULL pointer");
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(m, 32), (unsigned)B);
  LOB_DISPATCH_DTYPE(dtype, {
    k_transpose_rows<scalar_t><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(R, N, m, (const scalar_t*)Lt,
                                                                              (scalar_t*)L);
    return check_launch("k_transpose_rows");
  });
}

// matmul_simt.cu -- register-tiled CUDA-core matmuls used on the Krylov path where the contraction is either fp64
// (no tcgen05 kind exists for fp64) or small (preconditioner products with K = rank).  The fp32 dense operator
// matmul A (N x N) X (N x C) of BASELINE config 2 has its own tensor-core kernel (dense_tc.cu); this file is its
// fp64 / unaligned-shape counterpart and the shared building block for Q^T r, Q t, L eps, L^T L, U^T b, U w.
//
//   k_matmul_nn : Y (B,M,C)  = A (B|1,M,K) X (B,K,C) [+ d (.) X] [+ beta Y], optional fused <X,Y> column partials
//   k_matmul_tn : Out (B,I,J) = P^T Q,  P (B,N,I), Q (B,N,J)     (reduction over the long dimension, split-N)
//
// Both use a 128 x (8*RN) output tile per CTA, 256 threads as 32 (row groups of 4) x 8 (column groups of RN),
// operands staged through shared memory in k-major order so the inner product reads one 128-bit word of A and RN
// words of B per 4*RN FMAs.
#include <stdlib.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include <atomic>

#include "common.cuh"
#include "simt_tile.cuh"

namespace lob {

// ----------------------------------------------------------------------------------------------------------
template <typename T, int RN>
__global__ void __launch_bounds__(256)
k_matmul_nn(int64_t M, int64_t K, int64_t C, const T* __restrict__ A, int64_t lda, int64_t a_bs,
            const T* __restrict__ X, int64_t x_bs, T* __restrict__ Y, const T* __restrict__ dg, int64_t d_bs,
            int64_t d_st, double* __restrict__ dots, int n_parts, T beta_y, const T* __restrict__ E,
            const T* __restrict__ alpha, int64_t alpha_bs) {
  constexpr int TK = TileK<T>::value;
  constexpr int CP = 8 * RN;
  __shared__ __align__(16) T As[TK * LDA_S];
  __shared__ __align__(16) T Bs[TK * CP];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int64_t b = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int64_t c0 = (int64_t)blockIdx.z * CP;
  const T* Ab = A + b * a_bs;
  const T* Xb = X + b * x_bs;
  T* Yb = Y + b * M * C;
  const int ncol = (int)min((int64_t)CP, C - c0);

  T acc[4][RN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = (T)0;

  const bool vec_ok = (sizeof(T) == 4) && ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(Ab) & 15) == 0);

  for (int64_t k0 = 0; k0 < K; k0 += TK) {
    // ---- A tile: TM rows x TK columns, stored k-major ----
    if (sizeof(T) == 4 && vec_ok && k0 + TK <= K) {
      // TK/4 float4 per row; thread -> (row, quad)
      constexpr int QPR = TK / 4;
#pragma unroll
      for (int it = 0; it < (TM * QPR) / 256; ++it) {
        const int e = tid + it * 256;
        const int row = e / QPR, q = e % QPR;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + row < M) v = __ldg(reinterpret_cast<const float4*>(Ab + (m0 + row) * lda + k0 + q * 4));
        As[(q * 4 + 0) * LDA_S + row] = v.x;
        As[(q * 4 + 1) * LDA_S + row] = v.y;
        As[(q * 4 + 2) * LDA_S + row] = v.z;
        As[(q * 4 + 3) * LDA_S + row] = v.w;
      }
    } else {
      for (int e = tid; e < TM * TK; e += 256) {
        const int row = e / TK, kk = e % TK;
        T v = (T)0;
        if (m0 + row < M && k0 + kk < K) v = __ldg(Ab + (m0 + row) * lda + k0 + kk);
        As[kk * LDA_S + row] = v;
      }
    }
    // ---- B tile: TK rows x CP columns (zero padded) ----
    for (int e = tid; e < TK * CP; e += 256) {
      const int kk = e / CP, cc = e % CP;
      T v = (T)0;
      if (k0 + kk < K && cc < ncol) v = __ldg(Xb + (k0 + kk) * C + c0 + cc);
      Bs[e] = v;
    }
    __syncthreads();
    tile_fma<T, T, RN, TK>(As, Bs, CP, ty, tx, acc);
    __syncthreads();
  }

  // ---- epilogue ----
  double dsum[RN];
#pragma unroll
  for (int j = 0; j < RN; ++j) dsum[j] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= M) continue;
    T dv = (T)0;
    if (dg) dv = dg[b * d_bs + row * d_st];
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int cc = tx * RN + j;
      if (cc >= ncol) continue;
      const int64_t idx = row * C + c0 + cc;
      T y = acc[i][j];
      if (alpha) y *= alpha[b * alpha_bs];
      T xv = (T)0;
      if (dg || dots) xv = E ? E[b * M * C + idx] : Xb[idx];  // E == NULL requires M == K
      if (dg) y += dv * xv;
      if (beta_y != (T)0) y += beta_y * Yb[idx];
      Yb[idx] = y;
      dsum[j] += (double)xv * (double)y;
    }
  }
  if (dots) {
    double* red = reinterpret_cast<double*>(As);  // 32 x CP doubles <= sizeof(As)
    static_assert(sizeof(T) * TK * LDA_S >= sizeof(double) * 32 * 8 * RN, "reduction scratch");
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RN; ++j) red[ty * CP + tx * RN + j] = dsum[j];
    __syncthreads();
    if (tid < ncol) {
      double s = 0.0;
      for (int r = 0; r < 32; ++r) s += red[r * CP + tid];
      dots[(b * n_parts + blockIdx.x) * C + c0 + tid] = s;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------
// Small fp64 problems (BASELINE config 1: N = 512, 17 columns, batch 1): the tiled kernel above gives such a product 4 CTAs
// and 32 single-buffered k tiles, ~100 us of exposed latency per product and 78 % of the whole call.  Here a warp owns
// one output row: lanes split the contraction (coalesced reads of the A row, X through L1), eight columns at a time in
// registers, one shuffle reduction per column.  No fused <X, Y> partial sums (linear_cg adds its own dot pass).
// ----------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_matmul_rows(int64_t M, int64_t K, int64_t C, const T* __restrict__ A, int64_t lda, int64_t a_bs, const T* __restrict__ X,
              int64_t x_bs, T* __restrict__ Y, const T* __restrict__ dg, int64_t d_bs, int64_t d_st) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = blockIdx.y;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= M) return;
  const T* Ar = A + b * a_bs + row * lda;
  const T* Xb = X + b * x_bs;
  T* Yr = Y + (b * M + row) * C;
  for (int64_t c0 = 0; c0 < C; c0 += 8) {
    const int nc = (int)min((int64_t)8, C - c0);
    T acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = (T)0;
    for (int64_t k = lane; k < K; k += 32) {
      const T a = Ar[k];
      const T* xk = Xb + k * C + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < nc) acc[j] += a * xk[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      T v = acc[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[j] = v;
    }
    if (lane < nc) {
      T y = acc[0];
#pragma unroll
      for (int j = 1; j < 8; ++j)
        if (lane == j) y = acc[j];
      if (dg) y += dg[b * d_bs + row * d_st] * Xb[row * C + c0 + lane];  // square operator: E = X
      Yr[c0 + lane] = y;
    }
  }
}

template <typename T>
static bool rows_kernel_applies(int64_t B, int64_t M, int64_t K, int64_t C) {
  return sizeof(T) == 8 && B * M <= 8192 && K <= 8192 && C <= 64;
}

// ----------------------------------------------------------------------------------------------------------
// Out partial (b, split, I, J) = sum_{n in split} P[n, i] Q[n, j]
template <typename T, typename ACC, int RN>
__global__ void __launch_bounds__(256)
k_matmul_tn(int64_t N, int64_t I, int64_t J, const T* __restrict__ P, int64_t p_bs, const T* __restrict__ Q,
            int64_t q_bs, ACC* __restrict__ partial, int nsplit, int64_t rows_per_split, int nblk_i, int nblk_j) {
  constexpr int TK = TileK<T>::value;
  constexpr int CP = 8 * RN;
  __shared__ __align__(16) T As[TK * LDA_S];
  __shared__ __align__(16) T Bs[TK * CP];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int64_t b = blockIdx.y;
  const int split = blockIdx.x;
  const int bi = blockIdx.z / nblk_j, bj = blockIdx.z % nblk_j;
  const int64_t i0 = (int64_t)bi * TM, j0 = (int64_t)bj * CP;
  const int ni = (int)min((int64_t)TM, I - i0), nj = (int)min((int64_t)CP, J - j0);
  const T* Pb = P + b * p_bs;
  const T* Qb = Q + b * q_bs;
  const int64_t n_begin = (int64_t)split * rows_per_split;
  const int64_t n_end = min(n_begin + rows_per_split, N);

  ACC acc[4][RN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = (ACC)0;

  for (int64_t n0 = n_begin; n0 < n_end; n0 += TK) {
    for (int e = tid; e < TK * TM; e += 256) {
      const int kk = e / TM, ii = e % TM;
      T v = (T)0;
      if (n0 + kk < n_end && ii < ni) v = __ldg(Pb + (n0 + kk) * I + i0 + ii);
      As[kk * LDA_S + ii] = v;
    }
    for (int e = tid; e < TK * CP; e += 256) {
      const int kk = e / CP, jj = e % CP;
      T v = (T)0;
      if (n0 + kk < n_end && jj < nj) v = __ldg(Qb + (n0 + kk) * J + j0 + jj);
      Bs[e] = v;
    }
    __syncthreads();
    tile_fma<T, ACC, RN, TK>(As, Bs, CP, ty, tx, acc);
    __syncthreads();
  }
  ACC* out = partial + ((b * nsplit + split) * I) * J;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = ty * 4 + i;
    if (ii >= ni) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int jj = tx * RN + j;
      if (jj < nj) out[(i0 + ii) * J + j0 + jj] = acc[i][j];
    }
  }
}

template <typename ACC, typename OUT>
__global__ void k_reduce_splits(int64_t total_per_batch, int nsplit, const ACC* __restrict__ partial,
                                OUT* __restrict__ out, int64_t B) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * total_per_batch) return;
  const int64_t b = idx / total_per_batch, e = idx % total_per_batch;
  const ACC* p = partial + (b * nsplit) * total_per_batch + e;
  double s = 0.0;
  for (int i = 0; i < nsplit; ++i) s += (double)p[(int64_t)i * total_per_batch];
  out[idx] = (OUT)s;
}

template <typename T>
static int launch_nn(int64_t B, int64_t M, int64_t K, int64_t C, const T* A, int64_t lda, int64_t a_bs, const T* X,
                     int64_t x_bs, T* Y, const T* d, int64_t d_bs, int64_t d_st, double* dots, double beta_y,
                     cudaStream_t st, const T* E = nullptr, const T* alpha = nullptr, int64_t alpha_bs = 0) {
  if (!dots && !E && !alpha && beta_y == 0.0 && rows_kernel_applies<T>(B, M, K, C)) {
    k_matmul_rows<T><<<dim3((unsigned)cdiv(M, 8), (unsigned)B), 256, 0, st>>>(M, K, C, A, lda, a_bs, X, x_bs, Y, d, d_bs,
                                                                             d_st);
    return check_launch("k_matmul_rows");
  }
  const int rn = pick_rn(C);
  const int cp = 8 * rn;
  dim3 grid((unsigned)cdiv(M, TM), (unsigned)B, (unsigned)cdiv(C, cp));
  const int n_parts = (int)cdiv(M, TM);
#define LOB_NN_CASE(R)                                                                                          \
  case R:                                                                                                       \
    k_matmul_nn<T, R><<<grid, 256, 0, st>>>(M, K, C, A, lda, a_bs, X, x_bs, Y, d, d_bs, d_st, dots, n_parts,      \
                                            (T)beta_y, E, alpha, alpha_bs);                                     \
    break;
  switch (rn) {
    LOB_NN_CASE(1) LOB_NN_CASE(2) LOB_NN_CASE(3) LOB_NN_CASE(4) LOB_NN_CASE(5) LOB_NN_CASE(6) LOB_NN_CASE(7)
    LOB_NN_CASE(8)
  }
#undef LOB_NN_CASE
  return check_launch("k_matmul_nn");
}

// tn_skinny.cu: register-tiled kernel for the fp32 Q^T r shape (I <= 128, J <= 48)
bool tn_skinny_applicable(int64_t N, int64_t I, int64_t J, const void* P, int64_t p_bs);
int tn_skinny_nsplit(int64_t B, int64_t N);
int launch_tn_skinny_f32(int64_t B, int64_t N, int64_t I, int64_t J, const float* P, int64_t p_bs, const float* Q,
                         int64_t q_bs, float* partial, int nsplit, cudaStream_t st);

struct TnPlan {
  int rn, nblk_i, nblk_j, nsplit;
  int64_t rows_per_split;
};
static TnPlan tn_plan(int64_t B, int64_t N, int64_t I, int64_t J) {
  TnPlan p;
  p.rn = pick_rn(J);
  p.nblk_i = (int)cdiv(I, TM);
  p.nblk_j = (int)cdiv(J, 8 * p.rn);
  int64_t ctas = B * p.nblk_i * p.nblk_j;
  int64_t ns = cdiv((int64_t)kNumSMs * 4, ctas);
  int64_t maxs = cdiv(N, 64);
  if (ns > maxs) ns = maxs;
  if (ns > 64) ns = 64;
  if (ns < 1) ns = 1;
  p.rows_per_split = cdiv(N, ns);
  p.nsplit = (int)cdiv(N, p.rows_per_split);
  return p;
}

template <typename T, typename ACC, typename OUT>
static int launch_tn(int64_t B, int64_t N, int64_t I, int64_t J, const T* P, int64_t p_bs, const T* Q, int64_t q_bs,
                     OUT* out, void* ws, cudaStream_t st) {
  TnPlan p = tn_plan(B, N, I, J);
  ACC* partial = (ACC*)ws;
  if constexpr (std::is_same<T, float>::value && std::is_same<ACC, float>::value) {
    if (tn_skinny_applicable(N, I, J, P, p_bs)) {
      // the workspace is sized for max(generic plan, skinny plan) splits (lob_tn_matmul_workspace_bytes)
      const int ns = tn_skinny_nsplit(B, N);
      int s = launch_tn_skinny_f32(B, N, I, J, P, p_bs, Q, q_bs, partial, ns, st);
      if (s == LOB_OK) {
        const int64_t per = I * J;
        k_reduce_splits<ACC, OUT><<<(unsigned)cdiv(B * per, 256), 256, 0, st>>>(per, ns, partial, out, B);
        return check_launch("k_reduce_splits");
      }
      if (s != LOB_ERR_UNSUPPORTED) return s;
    }
  }
  dim3 grid((unsigned)p.nsplit, (unsigned)B, (unsigned)(p.nblk_i * p.nblk_j));
#define LOB_TN_CASE(R)                                                                                              \
  case R:                                                                                                           \
    k_matmul_tn<T, ACC, R><<<grid, 256, 0, st>>>(N, I, J, P, p_bs, Q, q_bs, partial, p.nsplit, p.rows_per_split,     \
                                                 p.nblk_i, p.nblk_j);                                               \
    break;
  switch (p.rn) {
    LOB_TN_CASE(1) LOB_TN_CASE(2) LOB_TN_CASE(3) LOB_TN_CASE(4) LOB_TN_CASE(5) LOB_TN_CASE(6) LOB_TN_CASE(7)
    LOB_TN_CASE(8)
  }
#undef LOB_TN_CASE
  LOB_TRY(check_launch("k_matmul_tn"));
  const int64_t per = I * J;
  k_reduce_splits<ACC, OUT><<<(unsigned)cdiv(B * per, 256), 256, 0, st>>>(per, p.nsplit, partial, out, B);
  return check_launch("k_reduce_splits");
}

}  // namespace lob

using namespace lob;

extern "C" int32_t lob_dense_matmul_parts(int64_t M) { return (int32_t)cdiv(M, TM); }

extern "C" int lob_matmul_nn(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C, const void* A, int64_t lda,
                             int64_t a_batch_stride, const void* X, int64_t x_batch_stride, void* Y, double beta_y,
                             void* stream) {
  LOB_REQUIRE(B > 0 && M > 0 && K > 0 && C > 0, "lob_matmul_nn: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_matmul_nn: flattened batch > 65535 not supported");
  LOB_REQUIRE(A && X && Y, "lob_matmul_nn: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    return launch_nn<scalar_t>(B, M, K, C, (const scalar_t*)A, lda, a_batch_stride, (const scalar_t*)X,
                               x_batch_stride, (scalar_t*)Y, nullptr, 0, 0, nullptr, beta_y, (cudaStream_t)stream);
  });
}

namespace lob {
int dense_matmul_tc_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                        const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs, const float* d,
                        int64_t d_bs, int64_t d_st, double* dots, cudaStream_t st);
size_t dense_stream_workspace_bytes(int64_t B, int64_t K, int64_t C);
int dense_matmul_stream_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                            const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                            const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                            cudaStream_t st);

size_t dense_stream2_workspace_bytes(int64_t B, int64_t K, int64_t C);
int dense_matmul_stream2_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                             const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                             const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                             cudaStream_t st);

size_t dense_stream2p_workspace_bytes(int64_t B, int64_t K, int64_t C);
int dense_matmul_stream2p_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                              const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                              const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                              cudaStream_t st);

// fp32 dispatch: streaming tcgen05 kernels (need the workspace; the CTA-pair kernel dense_stream2p.cu and
// dense_stream2.cu for C <= 48, dense_stream.cu up to C = 64) -> first-generation tcgen05 kernel -> CUDA cores.
// lob_debug_pin_dense_impl() pins one of them (A/B comparisons in the tests; every choice computes the same product).
static std::atomic<int> g_dense_pin{0};  // 0 auto, 1 stream2, 2 stream, 3 tc, 4 simt, 5 stream2p
static int dense_f32_tensor_paths(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                                  const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                                  const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                                  cudaStream_t st) {
  const int pin = g_dense_pin.load(std::memory_order_relaxed);
  const bool impl = pin != 0;
  if (pin == 4) return LOB_ERR_UNSUPPORTED;
  // dense_stream2 serves long contractions (streaming regime) and short ones (K <= 256: the preconditioner product
  // Q t with the fused z, <r,z> epilogue -- 1.62 ms at config 2 since its epilogue reduces the fp64 partials with a
  // transpose-reduce, against 2.25 ms for the first-generation kernel); the window in between stays on the older kernel.
  // long contractions: the CTA-pair kernel (half the X-operand traffic through each SM's shared memory)
  if (pin == 5 || (!impl && K >= 512)) {
    int s = dense_matmul_stream2p_f32(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, ws,
                                      ws_bytes, st);
    if (s != LOB_ERR_UNSUPPORTED) return s;
  }
  const bool pinned_stream2 = pin == 1;
  if (pinned_stream2 || (!impl && (K >= 512 || (K <= 256 && a_bs != 0)))) {
    int s = dense_matmul_stream2_f32(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, ws,
                                     ws_bytes, st);
    if (s != LOB_ERR_UNSUPPORTED) return s;
  }
  const bool pinned_stream = pin == 2;
  if (pinned_stream || (!impl && K >= 512)) {
    int s = dense_matmul_stream_f32(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, ws,
                                    ws_bytes, st);
    if (s != LOB_ERR_UNSUPPORTED) return s;
  }
  return dense_matmul_tc_f32(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, st);
}
}

extern "C" int lob_debug_pin_dense_impl(int32_t impl) {
  LOB_REQUIRE(impl >= 0 && impl <= 5, "lob_debug_pin_dense_impl: 0 auto, 1 stream2, 2 stream, 3 tc, 4 simt, 5 stream2p");
  lob::g_dense_pin.store(impl);
  return LOB_OK;
}

extern "C" size_t lob_dense_matmul_workspace_bytes(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C) {
  (void)M;
  if (dtype != LOB_F32) return 0;
  const size_t w1 = lob::dense_stream_workspace_bytes(B, K, C), w2 = lob::dense_stream2_workspace_bytes(B, K, C);
  const size_t w3 = lob::dense_stream2p_workspace_bytes(B, K, C);
  return std::max(w1, std::max(w2, w3));
}

extern "C" int lob_dense_matmul_ex(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C, const void* A,
                                   int64_t lda, int64_t a_batch_stride, const void* X, void* Y, const void* E,
                                   const void* alpha, int64_t alpha_batch_stride, const void* d,
                                   int64_t d_batch_stride, int64_t d_stride, double* dots, void* ws, size_t ws_bytes,
                                   void* stream) {
  LOB_REQUIRE(B > 0 && M > 0 && K > 0 && C > 0, "lob_dense_matmul_ex: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_dense_matmul_ex: flattened batch > 65535 not supported");
  LOB_REQUIRE(A && X && Y, "lob_dense_matmul_ex: NULL pointer");
  LOB_REQUIRE((!d && !dots) || E || M == K, "lob_dense_matmul_ex: diagonal / dots need E or a square operator");
  if (dtype == LOB_F32) {
    int s = dense_f32_tensor_paths(B, M, K, C, (const float*)A, lda, a_batch_stride, (const float*)X, (float*)Y,
                                   (const float*)E, (const float*)alpha, alpha_batch_stride, (const float*)d,
                                   d_batch_stride, d_stride, dots, ws, ws_bytes, (cudaStream_t)stream);
    if (s != LOB_ERR_UNSUPPORTED) return s;
  }
  LOB_DISPATCH_DTYPE(dtype, {
    return launch_nn<scalar_t>(B, M, K, C, (const scalar_t*)A, lda, a_batch_stride, (const scalar_t*)X, K * C,
                               (scalar_t*)Y, (const scalar_t*)d, d_batch_stride, d_stride, dots, 0.0,
                               (cudaStream_t)stream, (const scalar_t*)E, (const scalar_t*)alpha, alpha_batch_stride);
  });
}

extern "C" int lob_dense_matmul(int32_t dtype, int64_t B, int64_t M, int64_t K, int64_t C, const void* A, int64_t lda,
                                int64_t a_batch_stride, const void* X, void* Y, const void* d, int64_t d_batch_stride,
                                int64_t d_stride, double* dots, void* ws, size_t ws_bytes, void* stream) {
  LOB_REQUIRE(B > 0 && M > 0 && K > 0 && C > 0, "lob_dense_matmul: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_dense_matmul: flattened batch > 65535 not supported");
  LOB_REQUIRE(A && X && Y, "lob_dense_matmul: NULL pointer");
  LOB_REQUIRE((!d && !dots) || M == K, "lob_dense_matmul: fused diagonal / dots need a square operator");
  if (dtype == LOB_F32) {
    // tensor-core paths (dense_stream.cu, dense_tc.cu); shapes they do not cover fall through to the CUDA-core kernel
    int s = dense_f32_tensor_paths(B, M, K, C, (const float*)A, lda, a_batch_stride, (const float*)X, (float*)Y,
                                   nullptr, nullptr, 0, (const float*)d, d_batch_stride, d_stride, dots, ws, ws_bytes,
                                   (cudaStream_t)stream);
    if (s != LOB_ERR_UNSUPPORTED) return s;
  }
  LOB_DISPATCH_DTYPE(dtype, {
    return launch_nn<scalar_t>(B, M, K, C, (const scalar_t*)A, lda, a_batch_stride, (const scalar_t*)X, K * C,
                               (scalar_t*)Y, (const scalar_t*)d, d_batch_stride, d_stride, dots, 0.0,
                               (cudaStream_t)stream);
  });
}

extern "C" size_t lob_tn_matmul_workspace_bytes(int64_t B, int64_t N, int64_t I, int64_t J) {
  if (B <= 0 || N <= 0 || I <= 0 || J <= 0) return 0;
  TnPlan p = tn_plan(B, N, I, J);
  const int ns = std::max(p.nsplit, tn_skinny_nsplit(B, N));
  return (size_t)B * ns * I * J * sizeof(double);
}

extern "C" int lob_tn_matmul(int32_t dtype, int32_t out_dtype, int64_t B, int64_t N, int64_t I, int64_t J,
                             const void* P, int64_t p_batch_stride, const void* Q, int64_t q_batch_stride, void* out,
                             void* ws, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && I > 0 && J > 0, "lob_tn_matmul: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_tn_matmul: flattened batch > 65535 not supported");
  LOB_REQUIRE(P && Q && out && ws, "lob_tn_matmul: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LOB_F32 && out_dtype == LOB_F32)
    return launch_tn<float, float, float>(B, N, I, J, (const float*)P, p_batch_stride, (const float*)Q,
                                          q_batch_stride, (float*)out, ws, st);
  if (dtype == LOB_F32 && out_dtype == LOB_F64)
    return launch_tn<float, double, double>(B, N, I, J, (const float*)P, p_batch_stride, (const float*)Q,
                                            q_batch_stride, (double*)out, ws, st);
  if (dtype == LOB_F64 && out_dtype == LOB_F64)
    return launch_tn<double, double, double>(B, N, I, J, (const double*)P, p_batch_stride, (const double*)Q,
                                             q_batch_stride, (double*)out, ws, st);
  return fail(LOB_ERR_ARG, "lob_tn_matmul: unsupported dtype combination");
}

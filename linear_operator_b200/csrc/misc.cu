// misc.cu -- probe-vector assembly and column reductions around the solve
// (reference: functions/_inv_quad_logdet.py:107-110,151-153; psd_sum_linear_operator.py:15-18;
//  diag_linear_operator.py:273-277; identity_linear_operator.py:262-266).
#include "common.cuh"

namespace lob {

// z[b,n,s] = z_root[b,n,s] + sqrt(d[b,n]) * eps[s,b,n]; per-tile column sums of squares.
// 32 x 32 (n x s) tiles through shared memory so both the (S,B,N) read and the (B,N,S) write are coalesced.
constexpr int PROBE_TILES = 32;  // 32-row tiles per CTA of k_probe_combine

template <typename T>
__global__ void __launch_bounds__(256)
k_probe_combine(int64_t B, int64_t N, int64_t S, const T* __restrict__ z_root, const T* __restrict__ eps,
                const T* __restrict__ d, int64_t d_bs, int64_t d_st, T* __restrict__ z, double* __restrict__ parts,
                int ntn) {
  __shared__ T tile[32][33];
  __shared__ double red[8][32];
  const int64_t b = blockIdx.z;
  const int64_t s0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t s = s0 + tx;
  double acc = 0.0;
  // PROBE_TILES 32-row tiles per CTA: one partial sum per column and CTA (a partial per tile made the norm kernel add
  // N / 32 numbers per column one after the other: 7 of the 7.4 ms of this entry point at N = 10^6)
  for (int t = 0; t < PROBE_TILES; ++t) {
    const int64_t n0 = ((int64_t)blockIdx.x * PROBE_TILES + t) * 32;
    if (n0 >= N) break;
    for (int sl = ty; sl < 32; sl += 8) {
      const int64_t ss = s0 + sl, n = n0 + tx;
      tile[sl][tx] = (ss < S && n < N) ? eps[(ss * B + b) * N + n] : (T)0;
    }
    __syncthreads();
    for (int nl = ty; nl < 32; nl += 8) {
      const int64_t n = n0 + nl;
      if (s < S && n < N) {
        T v = tile[tx][nl];
        if (d) v = v * (T)sqrt((double)d[b * d_bs + n * d_st]);
        const int64_t idx = (b * N + n) * S + s;
        if (z_root) v = z_root[idx] + v;
        z[idx] = v;
        acc += (double)v * (double)v;
      }
    }
    __syncthreads();
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && s < S) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    parts[(b * ntn + blockIdx.x) * S + s] = t;
  }
}

// norms[b, s] = sqrt(sum_k parts[b, k, s]); block (32, 8), grid (ceil(S / 32), B): eight partial sums per column, fixed order
template <typename T>
__global__ void k_probe_norms(int64_t S, int ntn, const double* __restrict__ parts, T* __restrict__ norms) {
  __shared__ double red[8][32];
  const int64_t b = blockIdx.y;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t s = (int64_t)blockIdx.x * 32 + tx;
  double t = 0.0;
  if (s < S)
    for (int k = ty; k < ntn; k += 8) t += parts[(b * ntn + k) * S + s];
  red[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && s < S) {
    double a = 0.0;
    for (int i = 0; i < 8; ++i) a += red[i][tx];
    norms[b * S + s] = (T)sqrt(a);
  }
}

template <typename T>
__global__ void k_probe_normalize(int64_t N, int64_t S, T* __restrict__ z, const T* __restrict__ norms, int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / (N * S), s = idx % S;
  z[idx] = z[idx] / norms[b * S + s];
}

// parts[b, chunk, j] = sum_{n in chunk} U[b,n,uo+j] * V[b,n,vo+j]
template <typename T>
__global__ void __launch_bounds__(256)
k_col_dots(int64_t N, int64_t R, const T* __restrict__ U, int64_t Cu, int64_t uo, const T* __restrict__ V, int64_t Cv,
           int64_t vo, double* __restrict__ parts, int nchunks, int64_t rpc) {
  extern __shared__ double red[];
  const int64_t b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y, CX = blockDim.x, RY = blockDim.y;
  const int64_t r0 = (int64_t)chunk * rpc, r1 = min(r0 + rpc, N);
  for (int64_t j0 = 0; j0 < R; j0 += CX) {
    const int64_t j = j0 + tx;
    double acc = 0.0;
    if (j < R)
      for (int64_t n = r0 + ty; n < r1; n += RY)
        acc += (double)U[(b * N + n) * Cu + uo + j] * (double)V[(b * N + n) * Cv + vo + j];
    __syncthreads();
    red[ty * CX + tx] = acc;
    __syncthreads();
    if (ty == 0 && j < R) {
      double t = 0.0;
      for (int i = 0; i < RY; ++i) t += red[i * CX + tx];
      parts[(b * nchunks + chunk) * R + j] = t;
    }
  }
}

template <typename T>
__global__ void k_col_dots_finish(int64_t BR, int64_t R, int nchunks, const double* __restrict__ parts,
                                  T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BR) return;
  const int64_t b = i / R, j = i % R;
  double t = 0.0;
  for (int k = 0; k < nchunks; ++k) t += parts[(b * nchunks + k) * R + j];
  out[i] = (T)t;
}

struct ColPlan {
  int nchunks;
  int64_t rpc;
  int cx, ry;
};
static ColPlan col_plan(int64_t B, int64_t N, int64_t C) {
  ColPlan p;
  p.cx = (int)(C < 64 ? C : 64);
  p.ry = 256 / p.cx;
  int64_t nch = cdiv((int64_t)kNumSMs * 8, B);
  int64_t maxch = cdiv(N, (int64_t)p.ry * 4);
  if (nch > maxch) nch = maxch;
  if (nch > 64) nch = 64;
  if (nch < 1) nch = 1;
  p.rpc = cdiv(N, nch);
  p.nchunks = (int)cdiv(N, p.rpc);
  return p;
}

}  // namespace lob

using namespace lob;

extern "C" size_t lob_colred_workspace_bytes(int64_t B, int64_t N, int64_t C) {
  if (B <= 0 || N <= 0 || C <= 0) return 0;
  const size_t a = (size_t)B * cdiv(N, 32) * C * sizeof(double);
  const size_t b = (size_t)B * 64 * C * sizeof(double);
  return (a > b ? a : b) + 256;
}

extern "C" int lob_probe_assemble(int32_t dtype, int64_t B, int64_t N, int64_t S, const void* z_root,
                                  const void* eps_diag, const void* d, int64_t d_batch_stride, int64_t d_stride,
                                  void* probes, void* norms, void* ws, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && S > 0, "lob_probe_assemble: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_probe_assemble: flattened batch > 65535 not supported");
  LOB_REQUIRE(eps_diag && probes && norms && ws, "lob_probe_assemble: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int ntn = (int)cdiv(N, 32 * PROBE_TILES);
  dim3 grid((unsigned)ntn, (unsigned)cdiv(S, 32), (unsigned)B);
  const int64_t total = B * N * S;
  LOB_DISPATCH_DTYPE(dtype, {
    k_probe_combine<scalar_t><<<grid, dim3(32, 8), 0, st>>>(B, N, S, (const scalar_t*)z_root,
                                                            (const scalar_t*)eps_diag, (const scalar_t*)d,
                                                            d_batch_stride, d_stride, (scalar_t*)probes, (double*)ws,
                                                            ntn);
    LOB_TRY(check_launch("k_probe_combine"));
    k_probe_norms<scalar_t><<<dim3((unsigned)cdiv(S, 32), (unsigned)B), dim3(32, 8), 0, st>>>(
        S, ntn, (const double*)ws, (scalar_t*)norms);
    LOB_TRY(check_launch("k_probe_norms"));
    k_probe_normalize<scalar_t><<<(unsigned)cdiv(total, 256), 256, 0, st>>>(N, S, (scalar_t*)probes,
                                                                           (const scalar_t*)norms, total);
    LOB_TRY(check_launch("k_probe_normalize"));
  });
  return LOB_OK;
}

extern "C" int lob_col_dots(int32_t dtype, int64_t B, int64_t N, int64_t R, const void* U, int64_t Cu, int64_t u_off,
                            const void* V, int64_t Cv, int64_t v_off, void* out, void* ws, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && R > 0, "lob_col_dots: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_col_dots: flattened batch > 65535 not supported");
  LOB_REQUIRE(U && V && out && ws, "lob_col_dots: NULL pointer");
  LOB_REQUIRE(u_off + R <= Cu && v_off + R <= Cv, "lob_col_dots: column window out of range");
  cudaStream_t st = (cudaStream_t)stream;
  ColPlan p = col_plan(B, N, R);
  dim3 grid((unsigned)p.nchunks, (unsigned)B), block((unsigned)p.cx, (unsigned)p.ry);
  LOB_DISPATCH_DTYPE(dtype, {
    k_col_dots<scalar_t><<<grid, block, sizeof(double) * p.cx * p.ry, st>>>(
        N, R, (const scalar_t*)U, Cu, u_off, (const scalar_t*)V, Cv, v_off, (double*)ws, p.nchunks, p.rpc);
    LOB_TRY(check_launch("k_col_dots"));
    k_col_dots_finish<scalar_t><<<(unsigned)cdiv(B * R, 256), 256, 0, st>>>(B * R, R, p.nchunks, (const double*)ws,
                                                                           (scalar_t*)out);
    LOB_TRY(check_launch("k_col_dots_finish"));
  });
  return LOB_OK;
}

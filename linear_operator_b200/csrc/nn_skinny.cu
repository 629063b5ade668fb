// nn_skinny.cu -- Y (B, M, C) = alpha[b] (A X) + d (.) E  with a SHORT contraction (K <= 160) and few columns (C <= 40),
// plus the per-128-row partial sums of E * Y: the second product of the pivoted-Cholesky preconditioner apply,
//   z = (r - Q (Q^T r)) / sigma^2  with A = Q (N x rank), X = Q^T r (rank x C), E = r
// (operators/added_diag_linear_operator.py:135-140) and <r, z> of linear_cg (utils/linear_cg.py:35-36) in one pass.
//
// The tensor-core kernels are per-tile-overhead bound on this shape (4 k-blocks per 256-row tile; 2.25 - 2.5 ms at
// config 2 for 3.4 GB of traffic).  Here the whole K extent of 128 operator rows is ONE shared-memory tile (51 KB for
// K = 100, landed with 16-byte cp.async), X (K x C, 13 KB) sits beside it, and every thread owns a 4 x 8 output block
// (rows l, l+32, l+64, l+96 of the tile and one column block: one warp per column block, 160 threads): per 4-wide k
// chunk it reads four conflict-free LDS.128 of A (row stride padded to an odd number of 16-byte words) and eight
// LDS.128 of X for 128 FFMA.  (A first version with one row per thread was shared-memory bound: ncu showed 27 LSU
// wavefronts per 32 FFMA -- a warp-uniform LDS.128 still costs ~3 wavefronts.)  Two CTAs per SM overlap load and
// compute.  The fused epilogue is the same contract as the tensor-core kernels (alpha, d (.) E, fixed-order double
// partials per 128 rows); E comes in and Y goes out through a shared-memory tile so global accesses stay coalesced.
#include "common.cuh"

namespace lob {

constexpr int NNS_ROWS = 128;
constexpr int NNS_CB = 8;          // columns per thread
constexpr int NNS_MAX_K = 160;
constexpr int NNS_MAX_C = 40;

__device__ __forceinline__ void nns_cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// blockDim.x = 32 * ncb (ncb = ceil(C / 8)); thread -> (column block cb = warp, rows lane + 32 i, i < 4)
__global__ void __launch_bounds__(160, 2)
k_nn_skinny(int64_t M, int K, int C, const float* __restrict__ A, int64_t lda, int64_t a_bs,
            const float* __restrict__ X, float* __restrict__ Y, const float* __restrict__ E,
            const float* __restrict__ alpha, int64_t alpha_bs, const float* __restrict__ dg, int64_t d_bs, int64_t d_st,
            double* __restrict__ dots, int n_parts, int lds_a /*padded row stride of the A tile, floats*/) {
  extern __shared__ __align__(16) float nns_smem[];
  const int ncb = blockDim.x / 32;
  const int LDX = ncb * NNS_CB;                 // padded column count of the X tile
  float* As = nns_smem;                         // [128][lds_a]
  float* Xs = nns_smem + NNS_ROWS * lds_a;      // [K][LDX]
  double* red = reinterpret_cast<double*>(Xs + (size_t)K * LDX);  // alignment spacer (8-byte aligned region)
  const int ldE = C | 1;                                          // odd row stride: conflict-free column walks
  float* Es = reinterpret_cast<float*>(red + 4 * LDX);            // [128][ldE]  E tile in, Y tile out

  const int tid = threadIdx.x;
  const int lane = tid & 31, cb = tid >> 5;
  const int64_t b = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * NNS_ROWS;
  const int rows_valid = (int)min((int64_t)NNS_ROWS, M - m0);
  const float* Ab = A + b * a_bs + m0 * lda;
  const float* Xb = X + b * (int64_t)K * C;

  // ---- stage the operator rows (cp.async, zero-fill past M) and X (zero-padded columns) ----
  const int kch = K / 4;
  const uint32_t as_base = (uint32_t)__cvta_generic_to_shared(As);
  for (int e = tid; e < NNS_ROWS * kch; e += blockDim.x) {
    const int r = e / kch, ch = e - r * kch;
    const bool ok = r < rows_valid;
    nns_cp_async_16(as_base + (uint32_t)((r * lds_a + ch * 4) * 4), Ab + (ok ? (int64_t)r * lda : 0) + ch * 4,
                    ok ? 16u : 0u);
  }
  // E rows of this tile are one contiguous run of rows_valid * C floats: coalesced 4-byte cp.async into the padded tile
  const bool need_e = (dg != nullptr) || (dots != nullptr);
  const float* Et = E + (b * M + m0) * C;
  if (need_e) {
    const uint32_t es_base = (uint32_t)__cvta_generic_to_shared(Es);
    for (int e = tid; e < rows_valid * C; e += blockDim.x) {
      const int r = e / C, c = e - r * C;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(es_base + (uint32_t)((r * ldE + c) * 4)),
                   "l"(Et + e)
                   : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int e = tid; e < K * LDX; e += blockDim.x) {
    const int k = e / LDX, c = e - k * LDX;
    Xs[e] = (c < C) ? __ldg(Xb + k * C + c) : 0.f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- 4 x 8 outputs per thread ----
  float acc[4][NNS_CB];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NNS_CB; ++j) acc[i][j] = 0.f;
  const float* arow = As + lane * lds_a;
  const float* xcol = Xs + cb * NNS_CB;
#pragma unroll 2
  for (int k4 = 0; k4 < kch; ++k4) {
    float a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a4 = *reinterpret_cast<const float4*>(arow + i * 32 * lds_a + k4 * 4);
      a[i][0] = a4.x; a[i][1] = a4.y; a[i][2] = a4.z; a[i][3] = a4.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 x0 = *reinterpret_cast<const float4*>(xcol + (k4 * 4 + q) * LDX);
      const float4 x1 = *reinterpret_cast<const float4*>(xcol + (k4 * 4 + q) * LDX + 4);
      const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NNS_CB; ++j) acc[i][j] = fmaf(a[i][q], x[j], acc[i][j]);
    }
  }

  // ---- epilogue: alpha, + d (.) E, <E, Y> partials; Y goes out through the E tile (coalesced stores) ----
  const float alpha_b = alpha ? alpha[b * alpha_bs] : 1.0f;
  double pd[NNS_CB];
#pragma unroll
  for (int j = 0; j < NNS_CB; ++j) pd[j] = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = lane + 32 * i;
    const bool rok = row < rows_valid;
    const float dv = (dg && rok) ? __ldg(dg + b * d_bs + (m0 + row) * d_st) : 0.f;
#pragma unroll
    for (int j = 0; j < NNS_CB; ++j) {
      const int c = cb * NNS_CB + j;
      const bool ok = rok && c < C;
      const float e = (need_e && ok) ? Es[row * ldE + c] : 0.f;
      const float y = fmaf(dv, e, acc[i][j] * alpha_b);
      if (ok) Es[row * ldE + c] = y;  // each element is read and rewritten by its own thread only
      pd[j] += ok ? (double)e * (double)y : 0.0;
    }
  }
  __syncthreads();
  {
    float* Yt = Y + (b * M + m0) * C;
    for (int e = tid; e < rows_valid * C; e += blockDim.x) {
      const int r = e / C, c = e - r * C;
      Yt[e] = Es[r * ldE + c];
    }
  }
  if (dots) {
    // this warp holds all 128 rows of its columns: fixed-order butterfly, lane 0 writes the 128-row partial
#pragma unroll
    for (int j = 0; j < NNS_CB; ++j) {
      double v = pd[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int c = cb * NNS_CB + j;
      if (lane == 0 && c < C && blockIdx.x < n_parts) dots[(b * n_parts + blockIdx.x) * C + c] = v;
    }
  }
  (void)red;
}

bool nn_skinny_applicable(int64_t M, int64_t K, int64_t C, const void* A, int64_t lda, int64_t a_bs) {
  return K >= 4 && K <= NNS_MAX_K && (K % 4) == 0 && C >= 1 && C <= NNS_MAX_C && (lda % 4) == 0 && (a_bs % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(A) & 15) == 0 && M >= 1;
}

int launch_nn_skinny_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                         const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs, const float* d,
                         int64_t d_bs, int64_t d_st, double* dots, cudaStream_t st) {
  const int ncb = (int)cdiv(C, NNS_CB);
  const int kch = (int)(K / 4);
  const int lds_a = 4 * (kch | 1);  // odd number of 16-byte words per row: conflict-free LDS.128 down a column of rows
  const int LDX = ncb * NNS_CB;
  const size_t smem = ((size_t)NNS_ROWS * lds_a + (size_t)K * LDX + (size_t)NNS_ROWS * ((int)C | 1)) * sizeof(float) +
                      (size_t)4 * LDX * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    LOB_CUDA(cudaFuncSetAttribute(k_nn_skinny, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    // two CTAs per SM (one loads while the other computes) need the full shared-memory carve-out
    LOB_CUDA(cudaFuncSetAttribute(k_nn_skinny, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  if (smem > 160 * 1024) return LOB_ERR_UNSUPPORTED;
  dim3 grid((unsigned)cdiv(M, NNS_ROWS), (unsigned)B);
  k_nn_skinny<<<grid, 32 * ncb, smem, st>>>(M, (int)K, (int)C, A, lda, a_bs, X, Y, E ? E : X, alpha, alpha_bs, d,
                                                   d_bs, d_st, dots, (int)cdiv(M, 128), lds_a);
  return check_launch("k_nn_skinny");
}

}  // namespace lob

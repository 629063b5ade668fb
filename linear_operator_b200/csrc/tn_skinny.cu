// tn_skinny.cu -- Out (B, I, J) = P^T Q for tall-skinny fp32 operands P (B, N, I), Q (B, N, J) with I <= 128, J <= 48:
// the Q^T r product of the pivoted-Cholesky preconditioner apply (operators/added_diag_linear_operator.py:137; I = rank,
// J = probe + rhs columns, N = operator size), once per CG iteration.
//
// This contraction stays on the CUDA cores on purpose: the tensor core truncates (does not round) when it adds into its
// fp32 accumulator, a systematic bias ~1.5e-8 * N that the preconditioner amplifies by lambda_max / sigma^2 into the
// log-determinant; FFMA accumulation rounds to nearest.  The generic kernel in matmul_simt.cu reached only
// 8.7 TFLOP/s on this shape (3.8 ms at config 2); this one is register-tiled for it:
//   * one CTA = one (batch element, row split): 8 x 4 outputs per thread, threads laid out (I/8) x (J/4)  (13 x 9 = 117
//     of 128 threads for I = 100, J = 33), every thread walks all rows of the split;
//   * operands staged with cp.async (16-byte chunks of P rows, 4-byte elements of the odd-length Q rows, zero-fill past
//     the split) into a double-buffered shared-memory tile of 32 rows;
//   * per row and thread: 3 LDS.128 (two 4-wide chunks of the P row -- chunk t and chunk t + half, so consecutive
//     threads read consecutive 16-byte words -- and one chunk of the Q row) for 32 FFMA;
//   * narrow P (few 8 x 4 output tiles, e.g. 18 for the reference's default rank 15 padded to 16): G row groups of
//     threads share the tiles -- group g takes rows g, g + G, ... of every shared-memory tile -- and are added in a
//     fixed order at the end, so a CTA still has >= 4 busy warps;
//   * per-split partial sums are combined in double by k_reduce_splits (matmul_simt.cu): deterministic.
// k_tn_skinny_bulk (round 2) is the same computation fed by bulk copies: a 32-row tile of either operand is one
// contiguous run of global memory, so ONE cp.async.bulk per operand and tile (completion on an mbarrier, three stages)
// replaces the 800 16-byte + 1056 4-byte cp.async of the first version, whose per-element cost in the LSU bounded the
// kernel at 1.9 TB/s (1.3 TB/s for a rank-16 factor: almost all of its traffic is the 33-float rows copied 4 bytes at
// a time).  The tiles keep their global pitch (I resp. J floats), so the Q row is read with scalar loads (J = 33 is
// odd) and the thread tile grows to 16 x 4 for wide factors to stay under the shared-memory pipe.
#include "common.cuh"

namespace lob {

constexpr int TNS_TK = 32;       // rows per shared-memory tile
constexpr int TNS_MAX_I = 128;
constexpr int TNS_MAX_J = 48;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// LDP = 8 * half (padded I), LDQ = 4 * nj (padded J); blockDim.x = round_up(half * nj, 32)
__global__ void __launch_bounds__(256)
k_tn_skinny(int64_t N, int I, int J, const float* __restrict__ P, int64_t p_bs, const float* __restrict__ Q,
            int64_t q_bs, float* __restrict__ partial, int nsplit, int64_t rows_per_split, int half, int nj, int G) {
  extern __shared__ __align__(16) float tns_smem[];
  const int LDP = 8 * half, LDQ = 4 * nj;
  float* Ps = tns_smem;                       // [2][TK][LDP]
  float* Qs = tns_smem + 2 * TNS_TK * LDP;    // [2][TK][LDQ]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int64_t b = blockIdx.y;
  const int split = blockIdx.x;
  const int64_t n_begin = (int64_t)split * rows_per_split;
  const int64_t n_end = min(n_begin + rows_per_split, N);
  const float* Pb = P + b * p_bs;
  const float* Qb = Q + b * q_bs;

  // padding columns are never written by cp.async: zero both stages once
  for (int e = tid; e < 2 * TNS_TK * (LDP + LDQ); e += nthr) tns_smem[e] = 0.f;
  __syncthreads();

  const int pchunks = (I + 3) / 4;           // 16-byte chunks per P row (I % 4 == 0 is required by the launcher)
  const uint32_t ps_base = (uint32_t)__cvta_generic_to_shared(Ps);
  const uint32_t qs_base = (uint32_t)__cvta_generic_to_shared(Qs);
  auto load_tile = [&](int stage, int64_t n0) {
    const int total_p = TNS_TK * pchunks;
    for (int e = tid; e < total_p; e += nthr) {
      const int kk = e / pchunks, ch = e - kk * pchunks;
      const bool ok = n0 + kk < n_end;
      const float* src = Pb + (ok ? (n0 + kk) : n_begin) * I + ch * 4;
      cp_async_16(ps_base + (uint32_t)(((stage * TNS_TK + kk) * LDP + ch * 4) * 4), src, ok ? 16u : 0u);
    }
    const int total_q = TNS_TK * J;
    for (int e = tid; e < total_q; e += nthr) {
      const int kk = e / J, jj = e - kk * J;
      const bool ok = n0 + kk < n_end;
      const float* src = Qb + (ok ? (n0 + kk) : n_begin) * J + jj;
      cp_async_4(qs_base + (uint32_t)(((stage * TNS_TK + kk) * LDQ + jj) * 4), src, ok ? 4u : 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int ntile_out = half * nj;            // 8 x 4 output tiles
  const int tt = tid % ntile_out, g = tid / ntile_out;
  const int ti = tt % half, tj = tt / half;
  const bool active = g < G;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t ntile = (n_end - n_begin + TNS_TK - 1) / TNS_TK;
  if (ntile > 0) load_tile(0, n_begin);
  for (int64_t t = 0; t < ntile; ++t) {
    const int stage = (int)(t & 1);
    if (t + 1 < ntile) {
      load_tile(stage ^ 1, n_begin + (t + 1) * TNS_TK);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (active) {
      const float* prow = Ps + stage * TNS_TK * LDP + ti * 4;
      const float* qrow = Qs + stage * TNS_TK * LDQ + tj * 4;
#pragma unroll 8
      for (int kk = g; kk < TNS_TK; kk += G) {
        const float4 a0 = *reinterpret_cast<const float4*>(prow + kk * LDP);
        const float4 a1 = *reinterpret_cast<const float4*>(prow + kk * LDP + half * 4);
        const float4 q4 = *reinterpret_cast<const float4*>(qrow + kk * LDQ);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], q[j], acc[i][j]);
      }
    }
    __syncthreads();
  }

  // row groups: g = G - 1 hands its sums to g = G - 2, ... down to g = 0 (fixed order), through the free tile buffers
  for (int gg = G - 1; gg >= 1; --gg) {
    float* xch = tns_smem + tt * 32;
    if (active && g == gg) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) xch[i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (active && g == gg - 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += xch[i * 4 + j];
    }
    __syncthreads();
  }
  if (active && g == 0) {
    float* out = partial + ((b * nsplit + split) * I) * J;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ii = (i < 4) ? (ti * 4 + i) : ((half + ti) * 4 + i - 4);
      if (ii >= I) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = tj * 4 + j;
        if (jj < J) out[ii * J + jj] = acc[i][j];
      }
    }
  }
}

__device__ __forceinline__ void tns_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

constexpr int TNB_STAGES = 3;

// NA = 16-byte chunks of the P row per thread (thread tile 4 NA x 4): chunks t, t + G, ..., t + (NA - 1) G.
// RG = row groups (threads beyond the G * nj output tiles take rows g, g + RG, ... of every tile).
// rows_per_split is a multiple of TNS_TK; q_bs % 4 == 0 and 16-byte aligned bases (launcher).
template <int NA>
__global__ void __launch_bounds__(256)
k_tn_skinny_bulk(int64_t N, int I, int J, const float* __restrict__ P, int64_t p_bs, const float* __restrict__ Q,
                 int64_t q_bs, float* __restrict__ partial, int nsplit, int64_t rows_per_split, int G, int nj, int RG) {
  extern __shared__ __align__(128) unsigned char tnb_raw[];
  // stage layout: [P tile TK x I | slack | Q tile TK x J | slack], sizes rounded up to 128 bytes
  const int p_floats = (TNS_TK * I + 4 * NA * G + 31) / 32 * 32;
  const int q_floats = (TNS_TK * J + 4 + 31) / 32 * 32;
  float* stage0 = reinterpret_cast<float*>(tnb_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + (size_t)TNB_STAGES * (p_floats + q_floats));
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int64_t b = blockIdx.y;
  const int split = blockIdx.x;
  const int64_t n_begin = (int64_t)split * rows_per_split;
  const int64_t n_end = min(n_begin + rows_per_split, N);
  const float* Pb = P + b * p_bs;
  const float* Qb = Q + b * q_bs;
  const int64_t ntile = n_end > n_begin ? (n_end - n_begin + TNS_TK - 1) / TNS_TK : 0;

  if (tid == 0) {
    for (int i = 0; i < TNB_STAGES; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // full tiles come by bulk copy; a ragged last tile (fewer than TK rows: sizes no longer multiples of 16 bytes) is
  // loaded by the threads themselves when its turn comes
  auto issue = [&](int64_t t) {  // thread 0 only
    const int64_t n0 = n_begin + t * TNS_TK;
    if (n_end - n0 < TNS_TK) return;
    const int s = (int)(t % TNB_STAGES);
    float* pt = stage0 + (size_t)s * (p_floats + q_floats);
    float* qt = pt + p_floats;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&full[s]);
    const uint32_t pbytes = (uint32_t)(TNS_TK * I * 4), qbytes = (uint32_t)(TNS_TK * J * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(pbytes + qbytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(pt)), "l"(Pb + n0 * I), "r"(pbytes), "r"(bar)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(qt)), "l"(Qb + n0 * J), "r"(qbytes), "r"(bar)
                 : "memory");
  };
  if (tid == 0)
    for (int64_t t = 0; t < TNB_STAGES - 1 && t < ntile; ++t) issue(t);

  const int ntile_out = G * nj;
  const int tt = tid % ntile_out, g = tid / ntile_out;
  const int ti = tt % G, tj = tt / G;
  const bool active = g < RG;
  float acc[4 * NA][4];
#pragma unroll
  for (int i = 0; i < 4 * NA; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t t = 0; t < ntile; ++t) {
    const int s = (int)(t % TNB_STAGES);
    float* pt = stage0 + (size_t)s * (p_floats + q_floats);
    float* qt = pt + p_floats;
    const int64_t n0 = n_begin + t * TNS_TK;
    const int rows = (int)min((int64_t)TNS_TK, n_end - n0);
    if (rows == TNS_TK) {
      tns_mbar_wait((uint32_t)__cvta_generic_to_shared(&full[s]), (uint32_t)((t / TNB_STAGES) & 1));
    } else {
      const float* ps = Pb + n0 * I;
      const float* qs = Qb + n0 * J;
      for (int e = tid; e < rows * I; e += nthr) pt[e] = ps[e];
      for (int e = tid; e < rows * J; e += nthr) qt[e] = qs[e];
      __syncthreads();
    }
    if (active) {
      const float* prow = pt + ti * 4;
      const float* qrow = qt + tj * 4;
#pragma unroll 4
      for (int kk = g; kk < rows; kk += RG) {
        float a[4 * NA];
#pragma unroll
        for (int c = 0; c < NA; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(prow + kk * I + c * G * 4);
          a[4 * c] = v.x, a[4 * c + 1] = v.y, a[4 * c + 2] = v.z, a[4 * c + 3] = v.w;
        }
        // columns past J / chunks past I read the neighbouring row or the slack: their accumulators are never stored
        const float q[4] = {qrow[kk * J], qrow[kk * J + 1], qrow[kk * J + 2], qrow[kk * J + 3]};
#pragma unroll
        for (int i = 0; i < 4 * NA; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], q[j], acc[i][j]);
      }
    }
    __syncthreads();  // every thread is done with stage s: it can take tile t + STAGES - 1 ... which is stage (t - 1) % S
    if (tid == 0 && t + TNB_STAGES - 1 < ntile) issue(t + TNB_STAGES - 1);
  }

  // row groups: g = RG - 1 hands its sums to g = RG - 2, ... down to g = 0 (fixed order), through stage 0
  for (int gg = RG - 1; gg >= 1; --gg) {
    float* xch = stage0 + tt * (16 * NA);
    if (active && g == gg) {
#pragma unroll
      for (int i = 0; i < 4 * NA; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) xch[i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (active && g == gg - 1) {
#pragma unroll
      for (int i = 0; i < 4 * NA; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += xch[i * 4 + j];
    }
    __syncthreads();
  }
  if (active && g == 0) {
    float* out = partial + ((b * nsplit + split) * I) * J;
#pragma unroll
    for (int i = 0; i < 4 * NA; ++i) {
      const int ii = ((i >> 2) * G + ti) * 4 + (i & 3);
      if (ii >= I) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = tj * 4 + j;
        if (jj < J) out[ii * J + jj] = acc[i][j];
      }
    }
  }
}

bool tn_skinny_applicable(int64_t N, int64_t I, int64_t J, const void* P, int64_t p_bs) {
  return I >= 8 && I <= TNS_MAX_I && (I % 4) == 0 && J >= 1 && J <= TNS_MAX_J && N >= 256 && (p_bs % 4) == 0 &&
         (reinterpret_cast<uintptr_t>(P) & 15) == 0;
}

// row splits: enough CTAs for ~5 full waves of the resident-CTA capacity, at least 128 rows each
int tn_skinny_nsplit(int64_t B, int64_t N) {
  const int64_t slots = (int64_t)kNumSMs * 6;
  int64_t ns = cdiv(5 * slots, B);
  const int64_t maxs = std::max<int64_t>(1, N / 128);
  if (ns > maxs) ns = maxs;
  if (ns > 512) ns = 512;  // small batches of long vectors need the CTAs for bytes in flight (64 left the GPU 3/4 idle)
  if (ns < 1) ns = 1;
  const int64_t rps = cdiv(N, ns);
  return (int)cdiv(N, rps);
}

template <int NA>
static int launch_tn_skinny_bulk(int64_t B, int64_t N, int64_t I, int64_t J, const float* P, int64_t p_bs,
                                 const float* Q, int64_t q_bs, float* partial, int nsplit, cudaStream_t st) {
  const int pchunks = (int)((I + 3) / 4);
  const int G = (pchunks + NA - 1) / NA;
  const int nj = (int)((J + 3) / 4);
  const int tiles = G * nj;
  int RG = 1;  // row groups (a power of two): about 128 threads per CTA measured best (63 tiles: 2, 18 tiles: 4)
  while (RG < 8 && tiles * RG * 2 <= 128) RG *= 2;
  const int nthr = (int)align_up((size_t)tiles * RG, 32);
  if (nthr > 256) return LOB_ERR_UNSUPPORTED;
  const int64_t rows_per_split = align_up((size_t)cdiv(N, nsplit), TNS_TK);
  const int p_floats = (int)((TNS_TK * I + 4 * NA * G + 31) / 32 * 32);
  const int q_floats = (int)((TNS_TK * J + 4 + 31) / 32 * 32);
  const size_t smem = (size_t)TNB_STAGES * (p_floats + q_floats) * sizeof(float) + TNB_STAGES * sizeof(uint64_t);
  if (smem > 96 * 1024 || (size_t)tiles * 16 * NA > (size_t)(p_floats + q_floats)) return LOB_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    LOB_CUDA(cudaFuncSetAttribute(k_tn_skinny_bulk<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    LOB_CUDA(cudaFuncSetAttribute(k_tn_skinny_bulk<NA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  dim3 grid((unsigned)nsplit, (unsigned)B);
  k_tn_skinny_bulk<NA><<<grid, nthr, smem, st>>>(N, (int)I, (int)J, P, p_bs, Q, q_bs, partial, nsplit, rows_per_split,
                                                  G, nj, RG);
  return check_launch("k_tn_skinny_bulk");
}

int launch_tn_skinny_f32(int64_t B, int64_t N, int64_t I, int64_t J, const float* P, int64_t p_bs, const float* Q,
                         int64_t q_bs, float* partial, int nsplit, cudaStream_t st) {
  // bulk-copy fed version: needs 16-byte addressable tiles of Q as well (tile = 32 rows: 128 J bytes)
  if ((q_bs % 4) == 0 && (reinterpret_cast<uintptr_t>(Q) & 15) == 0) {
    const int s = I >= 64 ? launch_tn_skinny_bulk<4>(B, N, I, J, P, p_bs, Q, q_bs, partial, nsplit, st)
                          : launch_tn_skinny_bulk<2>(B, N, I, J, P, p_bs, Q, q_bs, partial, nsplit, st);
    if (s != LOB_ERR_UNSUPPORTED) return s;
  }
  const int pchunks = (int)((I + 3) / 4);
  const int half = (pchunks + 1) / 2;
  const int nj = (int)((J + 3) / 4);
  const int tiles = half * nj;
  int G = 1;  // row groups (a power of two: divides the 32-row tile)
  while (tiles < 64 && G < 8 && tiles * G * 2 <= 256) G *= 2;
  const int nthr = (int)align_up((size_t)tiles * G, 32);
  if (nthr > 256) return LOB_ERR_UNSUPPORTED;
  const int64_t rows_per_split = cdiv(N, nsplit);
  const size_t smem = (size_t)2 * TNS_TK * (8 * half + 4 * nj) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    LOB_CUDA(cudaFuncSetAttribute(k_tn_skinny, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    // six 36 KB CTAs per SM need the full shared-memory carve-out
    LOB_CUDA(cudaFuncSetAttribute(k_tn_skinny, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    attr_set = true;
  }
  dim3 grid((unsigned)nsplit, (unsigned)B);
  k_tn_skinny<<<grid, nthr, smem, st>>>(N, (int)I, (int)J, P, p_bs, Q, q_bs, partial, nsplit, rows_per_split, half, nj, G);
  return check_launch("k_tn_skinny");
}

}  // namespace lob

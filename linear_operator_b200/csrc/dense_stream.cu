// dense_stream.cu -- fp32 dense operator matmul  Y = alpha (A X) + d (.) E  as a persistent, HBM-streaming tcgen05 kernel.
// Reference arithmetic: operators/dense_linear_operator.py:60-64 (torch.matmul -> cuBLAS SGEMM),
// operators/added_diag_linear_operator.py:72-76 (addcmul), utils/linear_cg.py:250-251 (sum(p * Ap)),
// operators/added_diag_linear_operator.py:135-140 (z = (r - Q Q^T r) / s through the generalised epilogue).
//
// The operator A (B, M, K) fp32 is read from HBM exactly once (102 GB per CG iteration at BASELINE config 2); the
// right-hand side is skinny (C <= 64 columns), so the whole design is about keeping the TMA stream of A saturated
// while the tensor cores do fp32-accurate math behind it.
//
// Formulation ("X-stationary"): the tensor core computes D = Xop * A_tile^T with
//   * the A tile (256 operator rows x BK columns, exactly as TMA lands it, K-major, hardware swizzle) as the
//     UMMA *B* operand (N = 256) -- the raw fp32 bits are consumed as tf32, i.e. the hardware drops the 13 low
//     mantissa bits, which yields A_hi for free;
//   * A_lo = A - tf32(A) produced by 8 converter warps as a same-offset elementwise pass smem -> smem (no layout
//     knowledge, no transposes, no TMEM stores) into a second ring, used by a second UMMA with the same Xop;
//   * Xop (the UMMA A operand, M = 128 rows, K-major) = rows of [X_hi ; X_lo] (tf32 split of X^T, see k_split_x),
//     pre-split once per launch into a small workspace (B, R, Kp) and TMA-loaded next to the A tile.
// so D[row(c, hi)] + D[row(c, lo)] = sum_k (x_hi + x_lo)(a_hi + a_lo): all four products of the split, accumulated in
// fp32 in TMEM.  Rows of Xop are ordered so that the hi and lo rows of one column live in the same 32-lane TMEM
// quarter; the epilogue adds them with one warp shuffle.
//
// Persistent CTA (1 per SM, 512 threads), static tile schedule  tile = (batch element, 256-row block):
//   warp 0      TMA producer: A tile + Xop tile per k-block into one ring (evict-first / evict-last L2 hints)
//   warp 1      MMA issuer (elect.sync'd lane): 2 UMMAs (raw, lo) per 8-wide k step, tcgen05.commit frees the stages
//   warp 2      TMEM allocation (2 x 256 columns: double-buffered accumulator)
//   warps 4-7   epilogue: tcgen05.ld, hi + lo, alpha, + d (.) E, <E, Y> partials in double, direct global stores;
//               overlaps with the main loop of the next tile through the second accumulator
//   warps 8-15  converters: A_lo ring
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05_util.cuh"

namespace lob {

constexpr int DS_ROWS = 256;     // operator rows per tile == UMMA N
constexpr int DS_THREADS = 512;
constexpr int DS_CONV_THREADS = 256;



// Row r of the X operand <-> (column c, hi/lo part).  Rows come in 32-lane groups (one TMEM lane quarter each):
// a full group holds 16 columns (16 hi rows then 16 lo rows); a trailing group with <= 8 columns is 8 + 8 rows wide.
struct XopLayout {
  int C, full_groups, crem, w_last, R;
  __host__ __device__ explicit XopLayout(int c) : C(c) {
    full_groups = C >> 4;
    crem = C & 15;
    w_last = (crem == 0) ? 0 : (crem <= 8 ? 8 : 16);
    R = 32 * full_groups + 2 * w_last;
  }
  __host__ __device__ int group_width(int q) const { return q < full_groups ? 16 : (q == full_groups ? w_last : 0); }
  __host__ __device__ int group_cols(int q) const { return q < full_groups ? 16 : (q == full_groups ? crem : 0); }
};

struct DsParams {
  float* Y;
  const float* E;      // (B, M, C): multiplied by the diagonal term and dotted with Y (never NULL when dg or dots)
  const float* alpha;  // per-batch scale of the product (NULL: 1)
  int64_t alpha_bs;
  const float* dg;
  int64_t d_bs, d_st;
  double* dots;        // (B, n_parts, C) partial sums over 128-row blocks
  int64_t M, K, C;
  int n_parts;
  int xbytes;          // bytes of one X-operand tile (R * BK * 4 rounded up to 1 KB)
  int SA, SL;          // stages of the (A + Xop) ring and of the A_lo ring
  int MT;              // 256-row tiles per batch element
  int64_t ntiles;
  int a_shared;        // operator broadcast over the batch
  int lo_mode;         // 0: hardware truncates fp32 -> tf32 (lo = a - trunc(a)); 1: rna model; 2: no correction
  int dbg;             // bottleneck experiments (harness only): 1 skip lo MMA, 2 skip conversion, 4 skip all MMAs
};

template <int BK>
__global__ void __launch_bounds__(DS_THREADS, 1)
k_dense_stream(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX, DsParams p) {
  using namespace ds;
  constexpr int A_STAGE = DS_ROWS * BK * 4;
  constexpr int NV = A_STAGE / 16 / DS_CONV_THREADS;  // float4 per converter thread per k-block
  constexpr uint32_t IDESC = make_idesc_tf32(128, DS_ROWS);
  constexpr int MAX_ST = 12;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = A_STAGE + p.xbytes;
  unsigned char* sRing = smem;                         // SA x [A tile | Xop tile]
  unsigned char* sLo = smem + p.SA * stage_bytes;      // SL x A_lo tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLo + p.SL * A_STAGE);
  uint64_t* full = bars;                  // [MAX_ST] TMA -> converters, MMA
  uint64_t* empty = full + MAX_ST;        // [MAX_ST] MMA (commit) -> TMA
  uint64_t* lo_full = empty + MAX_ST;     // [MAX_ST] converters -> MMA
  uint64_t* lo_empty = lo_full + MAX_ST;  // [MAX_ST] MMA (commit) -> converters
  uint64_t* acc_full = lo_empty + MAX_ST; // [2] MMA (commit) -> epilogue
  uint64_t* acc_empty = acc_full + 2;     // [2] epilogue -> MMA
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (int)((p.K + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int i = 0; i < p.SA; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    for (int i = 0; i < p.SL; ++i) {
      mbar_init(smem_u32(&lo_full[i]), DS_CONV_THREADS / 32);
      mbar_init(smem_u32(&lo_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_holder))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint64_t pol_stream, pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      if (p.dbg & 16) {
        asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
        pol_keep = pol_stream;
      }
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = (int)(tile / p.MT);
        const int m0 = (int)(tile - (int64_t)b * p.MT) * DS_ROWS;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t dst = smem_u32(sRing + s * stage_bytes);
          const uint32_t bar = smem_u32(&full[s]);
          if (p.dbg & 8) {
            mbar_arrive_expect_tx(bar, (uint32_t)A_STAGE);
          } else {
            mbar_arrive_expect_tx(bar, (uint32_t)(A_STAGE + p.xbytes));
            tma_load_3d(dst + A_STAGE, &tmX, bar, kb * BK, 0, b, pol_keep);
          }
          tma_load_3d(dst, &tmA, bar, kb * BK, m0, p.a_shared ? 0 : b, pol_stream);
          if (++s == p.SA) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int s = 0, sl = 0;
    uint32_t ph = 0, phl = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(smem_u32(&acc_empty[buf]), ((it >> 1) & 1) ^ 1);
      const uint32_t d_addr = tmem_base + buf * DS_ROWS;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        mbar_wait(smem_u32(&lo_full[sl]), phl);
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(sRing + s * stage_bytes);
          const uint64_t bdesc = make_kmajor_desc<BK>(a_addr);
          const uint64_t xdesc = make_kmajor_desc<BK>(a_addr + A_STAGE);
          const uint64_t ldesc = make_kmajor_desc<BK>(smem_u32(sLo + sl * A_STAGE));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 8 tf32 = 32 bytes inside the swizzle span
            if (!(p.dbg & 4)) umma_tf32_ss(d_addr, xdesc + adv, bdesc + adv, IDESC, (kb | k) ? 1u : 0u);
            if (!(p.dbg & 5)) umma_tf32_ss(d_addr, xdesc + adv, ldesc + adv, IDESC, 1u);
          }
          umma_commit(smem_u32(&empty[s]));
          umma_commit(smem_u32(&lo_empty[sl]));
          if (kb == nkb - 1) umma_commit(smem_u32(&acc_full[buf]));
        }
        __syncwarp();
        if (++s == p.SA) { s = 0; ph ^= 1; }
        if (++sl == p.SL) { sl = 0; phl ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter of this warp
    const int C = (int)p.C;
    const XopLayout lay(C);
    const int wq = max(lay.group_width(q), 1);
    const bool active = lane < lay.group_cols(q);
    const int c = q * 16 + lane;
    const bool need_e = (p.dg != nullptr) || (p.dots != nullptr);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      const int64_t b = tile / p.MT;
      const int mt = (int)(tile - b * p.MT);
      const int64_t m0 = (int64_t)mt * DS_ROWS;
      const float alpha_b = p.alpha ? p.alpha[b * p.alpha_bs] : 1.0f;
      const float* Eb = p.E + b * p.M * C;
      float* Yb = p.Y + b * p.M * C;
      const float* dgb = p.dg ? p.dg + b * p.d_bs : nullptr;
      mbar_wait(smem_u32(&acc_full[buf]), (it >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * DS_ROWS;
      double part0 = 0.0, part1 = 0.0;
#pragma unroll 1
      for (int chunk = 0; chunk < DS_ROWS / 32; ++chunk) {
        const int64_t n0 = m0 + chunk * 32;
        if (n0 >= p.M) break;
        // E and the diagonal of this chunk's 32 rows: all loads issued back to back, before the TMEM read
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int64_t row = n0 + i;
          e[i] = (need_e && active && row < p.M) ? __ldg(Eb + row * C + c) : 0.f;
        }
        float dvl = 0.f;
        if (dgb && n0 + lane < p.M) dvl = __ldg(dgb + (n0 + lane) * p.d_st);
        uint32_t r[32];
        DS_LD32(taddr + chunk * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = __uint_as_float(r[i]);
          const float sum = v + __shfl_down_sync(0xffffffffu, v, wq);
          const float dv = __shfl_sync(0xffffffffu, dvl, i);
          const int64_t row = n0 + i;
          if (active && row < p.M) {
            const float y = fmaf(dv, e[i], sum * alpha_b);
            Yb[row * C + c] = y;
            acc += (double)e[i] * (double)y;
          }
        }
        if (chunk < 4) part0 += acc; else part1 += acc;
      }
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
      if (p.dots && active) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int64_t pi = (int64_t)mt * 2 + h;
          if (pi < p.n_parts) p.dots[(b * p.n_parts + pi) * C + c] = h ? part1 : part0;
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== converters: A_lo = A - tf32(A), same offsets, smem -> smem =====================
    const int ct = threadIdx.x - (DS_THREADS - DS_CONV_THREADS);
    const uint32_t radd = (p.lo_mode == 1) ? 0x1000u : 0u;
    const float lscale = (p.lo_mode == 2) ? 0.f : 1.f;
    int s = 0, sl = 0;
    uint32_t ph = 0, phl = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        mbar_wait(smem_u32(&lo_empty[sl]), phl ^ 1);
        const uint4* src = reinterpret_cast<const uint4*>(sRing + s * stage_bytes) + ct;
        uint4* dst = reinterpret_cast<uint4*>(sLo + sl * A_STAGE) + ct;
        uint4 v[NV];
        if (!(p.dbg & 2)) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = src[i * DS_CONV_THREADS];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint4 o;
          o.x = __float_as_uint((__uint_as_float(v[i].x) - __uint_as_float((v[i].x + radd) & 0xFFFFE000u)) * lscale);
          o.y = __float_as_uint((__uint_as_float(v[i].y) - __uint_as_float((v[i].y + radd) & 0xFFFFE000u)) * lscale);
          o.z = __float_as_uint((__uint_as_float(v[i].z) - __uint_as_float((v[i].z + radd) & 0xFFFFE000u)) * lscale);
          o.w = __float_as_uint((__uint_as_float(v[i].w) - __uint_as_float((v[i].w + radd) & 0xFFFFE000u)) * lscale);
          dst[i * DS_CONV_THREADS] = o;
        }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&lo_full[sl]));
        if (++s == p.SA) { s = 0; ph ^= 1; }
        if (++sl == p.SL) { sl = 0; phl ^= 1; }
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Xs (B, R, Kp) = tf32 split of X^T in the row order of XopLayout; k >= K and padding rows are written as zeros.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SPLIT_KT = 64;

__device__ __forceinline__ uint32_t tf32_rna_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__global__ void __launch_bounds__(256)
k_split_x(const float* __restrict__ X, float* __restrict__ Xs, int64_t K, int64_t Kp, int C, int R) {
  extern __shared__ float xs[];  // [SPLIT_KT][C | 1]
  const int ldx = C | 1;
  const int64_t b = blockIdx.y;
  const int64_t k0 = (int64_t)blockIdx.x * SPLIT_KT;
  const int kvalid = (int)max((int64_t)0, min((int64_t)SPLIT_KT, K - k0));
  const float* src = X + (b * K + k0) * C;
  for (int e = threadIdx.x; e < kvalid * C; e += blockDim.x) xs[(e / C) * ldx + (e % C)] = src[e];
  __syncthreads();
  const XopLayout lay(C);
  const int kk = threadIdx.x & (SPLIT_KT - 1);
  const int kw = (int)min((int64_t)SPLIT_KT, Kp - k0);
  for (int r = threadIdx.x / SPLIT_KT; r < R; r += blockDim.x / SPLIT_KT) {
    const int q = r >> 5, j = r & 31;
    const int w = lay.group_width(q);
    const int part = j / w;
    const int c = q * 16 + (j - part * w);
    float out = 0.f;
    if (c < C && kk < kvalid) {
      const float x = xs[kk * ldx + c];
      const float hi = __uint_as_float(tf32_rna_bits(x));
      out = part == 0 ? hi : __uint_as_float(tf32_rna_bits(x - hi));
    }
    if (kk < kw) Xs[(b * R + r) * Kp + k0 + kk] = out;
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled_ds)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled_ds ds_encode_fn() {
  static PFN_encodeTiled_ds fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled_ds)ptr;
  }
  return fn;
}

struct DsConfig {
  int bk;       // 16 (SWIZZLE_64B) or 32 (SWIZZLE_128B)
  int sa, sl;   // ring depths (0 = pick the deepest that fits)
  int lo_mode;
  int grid;     // 0 = one CTA per SM
  int dbg;
};

static DsConfig ds_default_config() {
  static DsConfig cfg = [] {
    DsConfig c{16, 0, 0, 0, 0, 0};
#ifdef LOB_DIAG  // tuning / bottleneck-experiment knobs exist only in the harness build
    if (const char* e = getenv("LOB_DS_BK")) c.bk = atoi(e);
    if (const char* e = getenv("LOB_DS_SA")) c.sa = atoi(e);
    if (const char* e = getenv("LOB_DS_SL")) c.sl = atoi(e);
    if (const char* e = getenv("LOB_DS_LO")) c.lo_mode = atoi(e);
    if (const char* e = getenv("LOB_DS_GRID")) c.grid = atoi(e);
#endif
    return c;
  }();
  return cfg;
}

constexpr size_t DS_SMEM_MAX = 232448;  // 227 KB opt-in limit per CTA
constexpr size_t DS_SMEM_FIXED = 1024 /*alignment*/ + 512 /*barriers*/;

size_t dense_stream_workspace_bytes(int64_t B, int64_t K, int64_t C) {
  if (B <= 0 || K <= 0 || C <= 0 || C > 64) return 0;
  const XopLayout lay((int)C);
  const int64_t Kp = (K + 3) / 4 * 4;
  return (size_t)B * lay.R * Kp * sizeof(float);
}

// returns LOB_ERR_UNSUPPORTED when the shape does not qualify (caller falls back to another CUDA kernel)
int dense_matmul_stream_f32_cfg(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                                const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                                const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                                cudaStream_t st, DsConfig cfg) {
  if (C > 64 || C < 1) return LOB_ERR_UNSUPPORTED;
  if ((lda % 4) != 0 || (a_bs % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0) return LOB_ERR_UNSUPPORTED;
  if (M >= (1LL << 31) || K >= (1LL << 31) || B >= (1LL << 31)) return LOB_ERR_UNSUPPORTED;
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 15) != 0 || ws_bytes < dense_stream_workspace_bytes(B, K, C))
    return LOB_ERR_UNSUPPORTED;
  if ((d || dots) && !E && M != K) return LOB_ERR_UNSUPPORTED;
  PFN_encodeTiled_ds enc = ds_encode_fn();
  if (!enc) return LOB_ERR_UNSUPPORTED;
  const int BK = (cfg.bk == 32) ? 32 : 16;
  const XopLayout lay((int)C);
  const int R = lay.R;
  const int64_t Kp = (K + 3) / 4 * 4;
  const bool shared = (a_bs == 0);
  const CUtensorMapSwizzle swz = (BK == 32) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  CUtensorMap tmA, tmX;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)(shared ? 1 : B)};
    cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)(shared ? (cuuint64_t)M * lda : a_bs) * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, DS_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     (cfg.dbg & 32) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                    : ((cfg.dbg & 64) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)R, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)R * Kp * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)R, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ws, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  }

  // 1) split X^T into the tf32 operand rows
  {
    dim3 grid((unsigned)cdiv(Kp, SPLIT_KT), (unsigned)B);
    const size_t sm = (size_t)SPLIT_KT * ((int)C | 1) * sizeof(float);
    if (!(cfg.dbg & 128)) {
      k_split_x<<<grid, 256, sm, st>>>(X, (float*)ws, K, Kp, (int)C, R);
      LOB_TRY(check_launch("k_split_x"));
    }
  }

  // 2) the streaming matmul
  const int a_stage = DS_ROWS * BK * 4;
  const int xbytes = (int)align_up((size_t)R * BK * 4, 1024);
  const int stage = a_stage + xbytes;
  int sl = cfg.sl > 0 ? cfg.sl : (BK == 32 ? 2 : 3);
  int sa = cfg.sa;
  if (sa <= 0) sa = (int)((DS_SMEM_MAX - DS_SMEM_FIXED - (size_t)sl * a_stage) / stage);
  if (sa > 12) sa = 12;
  if (sl > 12) sl = 12;
  if (sa < 2 || sl < 1) return LOB_ERR_UNSUPPORTED;
  // the M = 128 operand read of the last ring stage runs past its R rows: keep it inside the allocation
  const size_t tail_pad = (size_t)(128 - R) * BK * 4 > (size_t)sl * a_stage ? (size_t)(128 - R) * BK * 4 : 0;
  const size_t smem = DS_SMEM_FIXED + (size_t)sa * stage + (size_t)sl * a_stage + tail_pad;
  if (smem > DS_SMEM_MAX) return LOB_ERR_UNSUPPORTED;

  DsParams p;
  p.Y = Y;
  p.E = E ? E : X;
  p.alpha = alpha;
  p.alpha_bs = alpha_bs;
  p.dg = d;
  p.d_bs = d_bs;
  p.d_st = d_st;
  p.dots = dots;
  p.M = M;
  p.K = K;
  p.C = C;
  p.n_parts = (int)cdiv(M, 128);
  p.xbytes = xbytes;
  p.SA = sa;
  p.SL = sl;
  p.MT = (int)cdiv(M, DS_ROWS);
  p.ntiles = B * p.MT;
  p.a_shared = shared ? 1 : 0;
  p.lo_mode = cfg.lo_mode;
  p.dbg = cfg.dbg;
  const int64_t grid = std::min<int64_t>(p.ntiles, cfg.grid > 0 ? cfg.grid : kNumSMs);
  if (BK == 32) {
    auto kern = k_dense_stream<32>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DS_SMEM_MAX));
    kern<<<(unsigned)grid, DS_THREADS, smem, st>>>(tmA, tmX, p);
  } else {
    auto kern = k_dense_stream<16>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DS_SMEM_MAX));
    kern<<<(unsigned)grid, DS_THREADS, smem, st>>>(tmA, tmX, p);
  }
  return check_launch("k_dense_stream");
}

int dense_matmul_stream_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                            const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                            const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
  return dense_matmul_stream_f32_cfg(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, ws,
                                     ws_bytes, st, ds_default_config());
}

}  // namespace lob

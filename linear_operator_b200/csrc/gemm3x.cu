// gemm3x.cu -- batched fp32-accurate GEMM on the 5th-generation tensor cores:  D = alpha * A B (+ epilogue), 3xTF32.
//
// Serves the products of the path that are plain GEMMs rather than operator streams:
//   * the two tall products of the low-rank Woodbury solve (low_rank_root_added_diag_linear_operator.py:62-87) with the
//     batch as the GEMM row dimension:  W = R U  (split-K over N = 10^7)  and  x = (R - w U^T) / sigma  (fused epilogue);
//   * the mode products of KroneckerProductLinearOperator._matmul (kronecker_product_linear_operator.py:34-45);
//   * Q = L R^-1 of the preconditioner build (added_diag_linear_operator.py:164-166).
//
// Operands are fp32 in global memory in either major-ness (UMMA "K-major": the contraction index is contiguous;
// "MN-major": the row / column index is contiguous -- a row-major (K, N) matrix is an MN-major B operand), so no
// caller ever transposes.  Persistent CTAs (one per SM) walk a static schedule of work items, each one 128 x 128 tile
// of D for one batch element over one K range; the accumulators are double-buffered in TMEM (2 x (main + correction) =
// 512 columns), so the epilogue of one work item overlaps the main loop of the next -- what makes short contractions
// (Kronecker modes, K = 100; w U^T, K = 256) run at the same rate as long ones:
//   warp 0       TMA producer: A and B tiles (128 x 32 fp32, SWIZZLE_128B) into a 3-stage ring
//   warps 8-15   converters: lo = x - trunc_tf32(x) for both tiles, same offsets, smem -> smem (the tensor core reads the
//                raw fp32 words as tf32, i.e. truncated: hi costs nothing)
//   warp 1       MMA issuer: per 8-wide k step  D += A_hi B_hi,  Dc += A_hi B_lo + A_lo B_hi   (two fp32 accumulators in
//                TMEM: the main one takes a single truncating add per k step, the corrections are added in the epilogue)
//   warps 4-7    epilogue: tcgen05.ld, row scalings, + E, vectorised stores; split-K tiles store raw partial sums that
//                k_g3_reduce adds in a fixed order (double accumulation, deterministic)
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "tcgen05_util.cuh"

namespace lob {

constexpr int G3_BM = 128, G3_BN = 128, G3_BK = 32;
constexpr int G3_THREADS = 512, G3_CONV_THREADS = 256;
constexpr int G3_TILE_BYTES = 128 * G3_BK * 4;      // 16 KB, either major-ness
constexpr int G3_STAGE_BYTES = 4 * G3_TILE_BYTES;   // A | B | A_lo | B_lo
constexpr int G3_MAX_STAGES = 3, G3_MAX_EPI = 6;
constexpr int G3_EPI_BYTES = 128 * 32 * 4;          // one epilogue chunk: 128 rows x 32 columns, 16 KB
// operand ring + epilogue chunk ring: 3 x 64 KB + 2 x 16 KB (long contractions) or 2 x 64 KB + 6 x 16 KB (short ones)
constexpr size_t G3_SMEM = 1024 + 3 * (size_t)G3_STAGE_BYTES + 2 * (size_t)G3_EPI_BYTES + 512;

struct G3Params {
  float* D;
  int64_t ldd, d_bs;
  const float* E;
  int64_t lde, e_bs;
  const float* row_alpha;  // out = row_alpha[batch * ra_bs + m] * acc   (NULL: alpha)
  int64_t ra_bs;
  const float* row_beta;   //     + row_beta[batch * rb_bs + m] * E[m][n] (NULL: 1)
  int64_t rb_bs;
  float alpha;
  float* partial;  // (batch, splits, M, N) raw partial sums when splits > 1
  int64_t M, N, K;
  int tiles_m, tiles_n;
  int64_t nbatch;
  int64_t total;  // work items = batch * splits * tiles_m * tiles_n
  int64_t jobs;   // column jobs = batch * splits * tiles_n
  int sched_flat;
  int splits;
  int64_t kps;           // K range per split (multiple of G3_BK)
  int64_t a_div, b_div;  // operand batch index = batch / div (shared / grouped operands)
  int a_mn, b_mn;
  int stages, epi_slots; // ring depths (3, 2) or (2, 6)
  int tma_epi;           // epilogue through shared memory: E chunks TMA-loaded by warp 3, results TMA-stored (needs
                         // 16-byte aligned D / E rows); 0: direct loads / stores from the epilogue threads
  int d_trans;           // element (m, n) is stored at D[n * ldd + m] (E likewise; the row factors then index n):
                         // lanes = m write consecutive addresses -- the layout to pick when m is the contiguous index
};

// MN-major operand tile: 4 regions of [32 k rows][32 floats = 128 B].  For 32-bit operands the tensor core accepts ONE
// MN-major shared-memory layout: 128-byte rows swizzled in 32-byte chunks (UMMA layout type 1, "SWIZZLE_128B_BASE32B";
// TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), whose atom is 4 k rows x 128 B.  Regions (32 further rows / columns of
// the operand) are 4096 B apart (leading byte offset), 4-row k atoms 512 B apart (stride byte offset).
__device__ __forceinline__ uint64_t g3_mn_desc(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  desc |= (uint64_t)(4096 >> 4) << 16;
  desc |= (uint64_t)(512 >> 4) << 32;
  desc |= (uint64_t)1 << 46;
  desc |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return desc;
}

__device__ __forceinline__ uint64_t g3_desc(uint32_t tile_addr, int mn, int kstep) {
  if (mn) return g3_mn_desc(tile_addr + kstep * 1024);
  return ds::make_kmajor_desc<32>(tile_addr) + (uint64_t)((kstep * 32) >> 4);
}

// one unit of work: a 128 x 128 tile of D for one batch element over one K range
struct G3Work {
  int64_t batch, m0, n0, k_begin;
  int split, nkb;
};

// Static schedule.  A "column job" J is one (batch element, K range, 128-column block) of D; CTA c owns the jobs
// J = c, c + G, c + 2G, ... and walks the row blocks of each job in order.  Consecutive work items of a CTA therefore
// share their B tile (L2 hit), and the CTAs running side by side write neighbouring column blocks of the SAME rows --
// with row strides of tens of MB (x = (R - w U^T)/sigma is 512 x 10^7) that locality is worth 30 % (measured: 68 vs
// 95 ms for the opposite order).  With many row blocks per job (sched_flat) the items are dealt out one by one instead,
// which balances better.
__device__ __forceinline__ bool g3_decode(const G3Params& p, int64_t it, G3Work& o) {
  int64_t J, tm;
  if (p.sched_flat) {
    const int64_t w = blockIdx.x + it * (int64_t)gridDim.x;
    if (w >= p.total) return false;
    tm = w % p.tiles_m;
    J = w / p.tiles_m;
  } else {
    J = blockIdx.x + (it / p.tiles_m) * (int64_t)gridDim.x;
    if (J >= p.jobs) return false;
    tm = it % p.tiles_m;
  }
  const int64_t tn = J % p.tiles_n, rest = J / p.tiles_n;
  o.split = (int)(rest % p.splits);
  o.batch = rest / p.splits;
  o.m0 = tm * G3_BM;
  o.n0 = tn * G3_BN;
  o.k_begin = (int64_t)o.split * p.kps;
  const int64_t k_end = min(p.K, o.k_begin + p.kps);
  o.nkb = (int)((k_end - o.k_begin + G3_BK - 1) / G3_BK);
  return true;
}

__global__ void __launch_bounds__(G3_THREADS, 1)
k_gemm3x(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
         const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmE, G3Params p) {
  using namespace ds;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int G3_STAGES = p.stages;
  unsigned char* epi = smem + G3_STAGES * G3_STAGE_BYTES;  // epi_slots x 16 KB, 1024-byte aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + p.epi_slots * G3_EPI_BYTES);
  uint64_t* full = bars;                       // TMA -> converters, MMA
  uint64_t* empty = full + G3_MAX_STAGES;      // MMA (commit) -> TMA
  uint64_t* lo_full = empty + G3_MAX_STAGES;   // converters -> MMA
  uint64_t* acc_full = lo_full + G3_MAX_STAGES;  // [2] MMA (commit) -> epilogue
  uint64_t* acc_empty = acc_full + 2;            // [2] epilogue -> MMA
  uint64_t* e_full = acc_empty + 2;              // [epi_slots] E loader (warp 3) -> epilogue
  uint64_t* e_empty = e_full + G3_MAX_EPI;       // [epi_slots] epilogue (after its TMA store has read the slot) -> loader
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(e_empty + G3_MAX_EPI);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int i = 0; i < G3_STAGES; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
      mbar_init(smem_u32(&lo_full[i]), G3_CONV_THREADS / 32);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
    }
    for (int i = 0; i < p.epi_slots; ++i) {
      mbar_init(smem_u32(&e_full[i]), 1);
      mbar_init(smem_u32(&e_empty[i]), 4);
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
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
      int s = 0;
      uint32_t ph = 0;
      G3Work wk;
      for (int64_t wi = 0; g3_decode(p, wi, wk); ++wi) {
        const int ab = (int)(wk.batch / p.a_div), bb = (int)(wk.batch / p.b_div);
        for (int kb = 0; kb < wk.nkb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t bar = smem_u32(&full[s]);
          mbar_arrive_expect_tx(bar, 2 * G3_TILE_BYTES);
          const uint32_t dst = smem_u32(smem + s * G3_STAGE_BYTES);
          const int k0 = (int)(wk.k_begin + (int64_t)kb * G3_BK);
          if (!p.a_mn) {
            tma_load_3d(dst, &tmA, bar, k0, (int)wk.m0, ab, pol);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_3d(dst + j * 4096, &tmA, bar, (int)wk.m0 + 32 * j, k0, ab, pol);
          }
          if (!p.b_mn) {
            tma_load_3d(dst + G3_TILE_BYTES, &tmB, bar, k0, (int)wk.n0, bb, pol);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              tma_load_3d(dst + G3_TILE_BYTES + j * 4096, &tmB, bar, (int)wk.n0 + 32 * j, k0, bb, pol);
          }
          if (++s == G3_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_tf32(G3_BM, G3_BN) | ((uint32_t)(p.a_mn ? 1 : 0) << 15) |
                           ((uint32_t)(p.b_mn ? 1 : 0) << 16);
    int s = 0;
    uint32_t ph = 0, it = 0;
    G3Work wk;
    for (; g3_decode(p, it, wk); ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(smem_u32(&acc_empty[buf]), ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator pair
      tc_fence_after();
      const uint32_t d_main = tmem_base + buf * (2 * G3_BN), d_corr = d_main + G3_BN;
      for (int kb = 0; kb < wk.nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        mbar_wait(smem_u32(&lo_full[s]), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * G3_STAGE_BYTES);
          const uint32_t b_addr = a_addr + G3_TILE_BYTES;
          const uint32_t al_addr = a_addr + 2 * G3_TILE_BYTES, bl_addr = a_addr + 3 * G3_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < G3_BK / 8; ++k) {
            const uint64_t a_hi = g3_desc(a_addr, p.a_mn, k), b_hi = g3_desc(b_addr, p.b_mn, k);
            const uint64_t a_lo = g3_desc(al_addr, p.a_mn, k), b_lo = g3_desc(bl_addr, p.b_mn, k);
            // the small correction terms go to their own accumulator: the main accumulator then takes one truncating
            // fp32 add per k step instead of three, and the corrections are added once, in the epilogue
            umma_tf32_ss(d_main, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
            umma_tf32_ss(d_corr, a_hi, b_lo, idesc, (kb | k) ? 1u : 0u);
            umma_tf32_ss(d_corr, a_lo, b_hi, idesc, 1u);
          }
          umma_commit(smem_u32(&empty[s]));
          if (kb == wk.nkb - 1) umma_commit(smem_u32(&acc_full[buf]));
        }
        __syncwarp();
        if (++s == G3_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== E loader: feeds the epilogue chunk ring (tma_epi) =====================
    if (p.tma_epi && elect_one()) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      const bool load_e = p.E != nullptr && p.splits == 1;
      uint32_t gch = 0;
      int slot = 0;
      G3Work wk;
      for (int64_t wi = 0; g3_decode(p, wi, wk); ++wi) {
        for (int chunk = 0; chunk < G3_BN / 32; ++chunk) {
          const int64_t nb = wk.n0 + chunk * 32;
          if (nb >= p.N) break;
          mbar_wait(smem_u32(&e_empty[slot]), ((gch / p.epi_slots) & 1) ^ 1);
          const uint32_t bar = smem_u32(&e_full[slot]);
          if (load_e) {
            mbar_arrive_expect_tx(bar, G3_EPI_BYTES);
            tma_load_3d(smem_u32(epi + slot * G3_EPI_BYTES), &tmE, bar, (int)nb, (int)wk.m0, (int)wk.batch, pol);
          } else {
            mbar_arrive(bar);
          }
          ++gch;
          if (++slot == p.epi_slots) slot = 0;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: thread = output row; overlaps the main loop of the next work item =============
    const int q = warp & 3;
    uint32_t it = 0, gch = 0;
    int eslot = 0;
    G3Work wk;
    for (; g3_decode(p, it, wk); ++it) {
      const uint32_t buf = it & 1;
      const int64_t m = wk.m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      float* dst;
      float ra = p.alpha, rb = 1.f;
      const float* erow = nullptr;
      if (p.splits > 1) {
        dst = p.partial + ((wk.batch * p.splits + wk.split) * p.M + (row_ok ? m : 0)) * p.N;
        ra = 1.f;
      } else {
        dst = p.D + wk.batch * p.d_bs + (row_ok ? m : 0) * p.ldd;
        if (row_ok && !p.d_trans) {
          if (p.row_alpha) ra = p.row_alpha[wk.batch * p.ra_bs + m];
          if (p.E) {
            erow = p.E + wk.batch * p.e_bs + m * p.lde;
            if (p.row_beta) rb = p.row_beta[wk.batch * p.rb_bs + m];
          }
        }
      }
      const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                          (erow == nullptr || (reinterpret_cast<uintptr_t>(erow) & 15) == 0);
      const bool trans = p.d_trans && p.splits == 1;
      mbar_wait(smem_u32(&acc_full[buf]), (it >> 1) & 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (2 * G3_BN);
#pragma unroll 1
      for (int chunk = 0; chunk < G3_BN / 32; ++chunk) {
        const int64_t nb = wk.n0 + chunk * 32;
        if (nb >= p.N) break;  // warp-uniform
        uint32_t r[32], rc[32];
        DS_LD32(taddr + chunk * 32, r);
        DS_LD32(taddr + G3_BN + chunk * 32, rc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (chunk == G3_BN / 32 - 1 || nb + 32 >= p.N) {
          // last TMEM read of this work item: hand the accumulator pair back before the global stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        }
        if (p.tma_epi) {
          // results go through the chunk ring: this thread's row of the [128 x 32] chunk (128-byte rows, 16-byte units
          // XOR-swizzled by row % 8 exactly as TMA laid E down) is combined in place, then each warp TMA-stores its own
          // 32-row box -- full lines, clipped at the edges of D by the tensor map
          mbar_wait(smem_u32(&e_full[eslot]), (gch / p.epi_slots) & 1);
          const int rl = q * 32 + lane;
          unsigned char* rowp = epi + eslot * G3_EPI_BYTES + rl * 128;
          const bool use_e = erow != nullptr;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4* cell = reinterpret_cast<float4*>(rowp + ((j ^ (rl & 7)) << 4));
            float4 v = make_float4((__uint_as_float(r[4 * j]) + __uint_as_float(rc[4 * j])) * ra,
                                   (__uint_as_float(r[4 * j + 1]) + __uint_as_float(rc[4 * j + 1])) * ra,
                                   (__uint_as_float(r[4 * j + 2]) + __uint_as_float(rc[4 * j + 2])) * ra,
                                   (__uint_as_float(r[4 * j + 3]) + __uint_as_float(rc[4 * j + 3])) * ra);
            if (use_e) {
              const float4 e = *cell;
              v.x = fmaf(rb, e.x, v.x); v.y = fmaf(rb, e.y, v.y); v.z = fmaf(rb, e.z, v.z); v.w = fmaf(rb, e.w, v.w);
            }
            *cell = v;
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (wk.m0 + q * 32 < p.M) {
              const int zb = (p.splits > 1) ? (int)(wk.batch * p.splits + wk.split) : (int)wk.batch;
              asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                           ::"l"(&tmD), "r"(smem_u32(epi + eslot * G3_EPI_BYTES + q * 4096)), "r"((int)nb),
                             "r"((int)(wk.m0 + q * 32)), "r"(zb)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            mbar_arrive(smem_u32(&e_empty[eslot]));
          }
          ++gch;
          if (++eslot == p.epi_slots) eslot = 0;
          continue;
        }
        if (!row_ok) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(rc[i]));
        if (trans) {
          // element (m, n) -> D[n * ldd + m]: for every n the 32 lanes (consecutive m) write one 128-byte line
          float* Db = p.D + wk.batch * p.d_bs + m;
          const float* Eb = p.E ? p.E + wk.batch * p.e_bs + m : nullptr;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int64_t n = nb + i;
            if (n < p.N) {
              const float fa = p.row_alpha ? __ldg(p.row_alpha + wk.batch * p.ra_bs + n) : p.alpha;
              float v = __uint_as_float(r[i]) * fa;
              if (Eb) {
                const float fb = p.row_beta ? __ldg(p.row_beta + wk.batch * p.rb_bs + n) : 1.f;
                v = fmaf(fb, Eb[n * p.lde], v);
              }
              Db[n * p.ldd] = v;
            }
          }
        } else if (vec_ok && nb + 32 <= p.N) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 v = make_float4(__uint_as_float(r[i]) * ra, __uint_as_float(r[i + 1]) * ra,
                                   __uint_as_float(r[i + 2]) * ra, __uint_as_float(r[i + 3]) * ra);
            if (erow) {
              const float4 e = *reinterpret_cast<const float4*>(erow + nb + i);
              v.x = fmaf(rb, e.x, v.x); v.y = fmaf(rb, e.y, v.y); v.z = fmaf(rb, e.z, v.z); v.w = fmaf(rb, e.w, v.w);
            }
            *reinterpret_cast<float4*>(dst + nb + i) = v;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (nb + i < p.N) {
              float v = __uint_as_float(r[i]) * ra;
              if (erow) v = fmaf(rb, erow[nb + i], v);
              dst[nb + i] = v;
            }
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== converters: lo = x - trunc_tf32(x), both operand tiles =====================
    const int ct = threadIdx.x - (G3_THREADS - G3_CONV_THREADS);
    constexpr int NV = 2 * G3_TILE_BYTES / 16 / G3_CONV_THREADS;  // 8 x 16 bytes per thread per k block
    int s = 0;
    uint32_t ph = 0;
    G3Work wk;
    for (int64_t wi = 0; g3_decode(p, wi, wk); ++wi) {
      const int nkb = wk.nkb;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        const uint4* src = reinterpret_cast<const uint4*>(smem + s * G3_STAGE_BYTES) + ct;
        uint4* dstv = reinterpret_cast<uint4*>(smem + s * G3_STAGE_BYTES + 2 * G3_TILE_BYTES) + ct;
        uint4 v[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = src[i * G3_CONV_THREADS];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          uint4 o;
          o.x = __float_as_uint(__uint_as_float(v[i].x) - __uint_as_float(v[i].x & 0xFFFFE000u));
          o.y = __float_as_uint(__uint_as_float(v[i].y) - __uint_as_float(v[i].y & 0xFFFFE000u));
          o.z = __float_as_uint(__uint_as_float(v[i].z) - __uint_as_float(v[i].z & 0xFFFFE000u));
          o.w = __float_as_uint(__uint_as_float(v[i].w) - __uint_as_float(v[i].w & 0xFFFFE000u));
          dstv[i * G3_CONV_THREADS] = o;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&lo_full[s]));
        if (++s == G3_STAGES) { s = 0; ph ^= 1; }
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

// out[b][m][n] = ra * sum_s partial[b][s][m][n]  (+ rb * E), splits added in a fixed order in double
template <typename TO>
__global__ void __launch_bounds__(256)
k_g3_reduce(G3Params p, TO* __restrict__ out) {
  const int64_t total = p.M * p.N;
  const int64_t gidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= total * p.nbatch) return;
  const int64_t batch = gidx / total, idx = gidx - batch * total;
  const int64_t m = idx / p.N, n = idx - m * p.N;
  const float* src = p.partial + batch * p.splits * total + idx;
  double s = 0.0;
  for (int i = 0; i < p.splits; ++i) s += (double)src[(int64_t)i * total];
  const int64_t f = p.d_trans ? n : m;  // index of the row factors
  const int64_t o_r = p.d_trans ? n : m, o_c = p.d_trans ? m : n;
  double ra = p.row_alpha ? (double)p.row_alpha[batch * p.ra_bs + f] : (double)p.alpha;
  double v = s * ra;
  if (p.E) v += (p.row_beta ? (double)p.row_beta[batch * p.rb_bs + f] : 1.0) * (double)p.E[batch * p.e_bs + o_r * p.lde + o_c];
  out[batch * p.d_bs + o_r * p.ldd + o_c] = (TO)v;
}

typedef CUresult (*PFN_encodeTiled_g3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled_g3 g3_encode_fn() {
  static PFN_encodeTiled_g3 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled_g3)ptr;
  }
  return fn;
}

// operand (rows x K) with `rows` = M or N.  K-major: element (r, k) at ptr[r * ld + k]; MN-major: ptr[k * ld + r].
static bool g3_make_map(CUtensorMap* tm, const float* ptr, int mn, int64_t rows, int64_t K, int64_t ld, int64_t bs,
                        int64_t nbatch) {
  PFN_encodeTiled_g3 enc = g3_encode_fn();
  if (!enc) return false;
  if ((ld % 4) != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
  if (nbatch > 1 && (bs % 4) != 0) return false;
  if (rows >= (1LL << 31) || K >= (1LL << 31) || nbatch >= (1LL << 31)) return false;
  const int64_t inner = mn ? rows : K, outer = mn ? K : rows;
  const int64_t bstride = (nbatch > 1) ? bs : outer * ld;
  if ((uint64_t)ld * 4 >= (1ULL << 40) || (uint64_t)bstride * 4 >= (1ULL << 40)) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)nbatch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)bstride * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)(mn ? G3_BK : 128), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// row-major (rows x cols) matrices with a batch dimension, boxes of box_rows x 32 columns, SWIZZLE_128B: the epilogue's
// chunk ring (E loads: 128-row boxes, D stores: one 32-row box per epilogue warp)
static bool g3_make_rm_map(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int64_t bs,
                           int64_t nbatch, int box_rows) {
  PFN_encodeTiled_g3 enc = g3_encode_fn();
  if (!enc || !ptr) return false;
  if ((ld % 4) != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
  if (nbatch > 1 && (bs % 4) != 0) return false;
  if (rows >= (1LL << 31) || cols >= (1LL << 31) || nbatch >= (1LL << 31)) return false;
  const int64_t bstride = (nbatch > 1) ? bs : rows * ld;
  if ((uint64_t)ld * 4 >= (1ULL << 40) || (uint64_t)bstride * 4 >= (1ULL << 40)) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nbatch};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * 4, (cuuint64_t)bstride * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Split-K policy.  Two reasons to split: (1) parallelism when there are fewer tiles than SMs; (2) accuracy -- the
// tensor core adds into its fp32 accumulator with truncation, an error that grows linearly with the number of k steps
// accumulated in TMEM (measured at K = 10^7: 6e-4 relative with 37 splits, 2e-5 with 1024), so one split never
// accumulates more than G3_MAX_K_PER_SPLIT contraction indices; the partial sums are added in double.
constexpr int64_t G3_MAX_K_PER_SPLIT = 4096;
constexpr int64_t G3_MAX_PARTIAL_BYTES = 4LL << 30;

static int g3_pick_splits(int64_t batch, int64_t M, int64_t N, int64_t K, int requested) {
  const int64_t tiles = batch * cdiv(M, G3_BM) * cdiv(N, G3_BN);
  const int64_t max_splits = std::max<int64_t>(1, cdiv(K, 8 * G3_BK));  // at least 8 k blocks per split
  int64_t s = requested;
  if (s <= 0) {
    s = 1;
    if (tiles < kNumSMs) s = cdiv(2 * kNumSMs, tiles);
    s = std::max<int64_t>(s, cdiv(K, G3_MAX_K_PER_SPLIT));
    const int64_t cap = std::max<int64_t>(1, G3_MAX_PARTIAL_BYTES / (batch * M * N * (int64_t)sizeof(float)));
    if (s > cap) s = cap;
  }
  if (s > max_splits) s = max_splits;
  if (s > 65535) s = 65535;
  if (s < 1) s = 1;
  // every split must own at least one k block
  int64_t kps = cdiv(cdiv(K, s), G3_BK) * G3_BK;
  s = cdiv(K, kps);
  return (int)s;
}

}  // namespace lob

using namespace lob;

extern "C" int32_t lob_gemm3x_splits(int64_t batch, int64_t M, int64_t N, int64_t K, int32_t requested) {
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return 1;
  return g3_pick_splits(batch, M, N, K, requested);
}

extern "C" size_t lob_gemm3x_workspace_bytes(int64_t batch, int64_t M, int64_t N, int64_t K, int32_t requested_splits) {
  if (batch <= 0 || M <= 0 || N <= 0 || K <= 0) return 0;
  const int s = g3_pick_splits(batch, M, N, K, requested_splits);
  return s > 1 ? (size_t)batch * s * M * N * sizeof(float) : 0;
}

extern "C" int lob_gemm3x(int64_t batch, int64_t M, int64_t N, int64_t K, const void* A, int32_t a_mn, int64_t lda,
                          int64_t a_batch_stride, int64_t a_batch_div, const void* B, int32_t b_mn, int64_t ldb,
                          int64_t b_batch_stride, int64_t b_batch_div, void* D, int32_t d_dtype, int64_t ldd,
                          int64_t d_batch_stride, double alpha, const void* row_alpha, int64_t row_alpha_batch_stride,
                          const void* E, int64_t lde, int64_t e_batch_stride, const void* row_beta,
                          int64_t row_beta_batch_stride, int32_t d_trans, int32_t requested_splits, void* ws,
                          size_t ws_bytes, void* stream) {
  LOB_REQUIRE(batch > 0 && M > 0 && N > 0 && K > 0, "lob_gemm3x: sizes must be positive");
  LOB_REQUIRE(A && B && D, "lob_gemm3x: NULL pointer");
  LOB_REQUIRE(a_batch_div >= 1 && b_batch_div >= 1, "lob_gemm3x: batch divisors must be >= 1");
  LOB_REQUIRE(d_dtype == LOB_F32 || d_dtype == LOB_F64, "lob_gemm3x: output dtype must be LOB_F32 or LOB_F64");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = g3_pick_splits(batch, M, N, K, requested_splits);
  if (d_dtype == LOB_F64 && splits == 1) return fail(LOB_ERR_UNSUPPORTED, "lob_gemm3x: fp64 output needs the split-K path");
  const int64_t tiles_m = cdiv(M, G3_BM), tiles_n = cdiv(N, G3_BN);
  if (tiles_m >= (1LL << 31) || tiles_n >= (1LL << 31)) return fail(LOB_ERR_UNSUPPORTED, "lob_gemm3x: too many tiles");
  CUtensorMap tmA, tmB;
  if (!g3_make_map(&tmA, (const float*)A, a_mn, M, K, lda, a_batch_stride, cdiv(batch, a_batch_div)) ||
      !g3_make_map(&tmB, (const float*)B, b_mn, N, K, ldb, b_batch_stride, cdiv(batch, b_batch_div)))
    return fail(LOB_ERR_UNSUPPORTED, "lob_gemm3x: operand layout not TMA-addressable (16-byte alignment of base / strides)");
  G3Params p;
  p.D = (float*)D;
  p.ldd = ldd;
  p.d_bs = d_batch_stride;
  p.E = (const float*)E;
  p.lde = lde;
  p.e_bs = e_batch_stride;
  p.row_alpha = (const float*)row_alpha;
  p.ra_bs = row_alpha_batch_stride;
  p.row_beta = (const float*)row_beta;
  p.rb_bs = row_beta_batch_stride;
  p.alpha = (float)alpha;
  p.partial = nullptr;
  p.M = M;
  p.N = N;
  p.K = K;
  p.tiles_m = (int)tiles_m;
  p.tiles_n = (int)tiles_n;
  p.nbatch = batch;
  p.total = batch * splits * tiles_m * tiles_n;
  p.jobs = batch * splits * tiles_n;
  p.sched_flat = (tiles_m > 8 || p.jobs < kNumSMs) ? 1 : 0;
  p.splits = splits;
  p.kps = cdiv(cdiv(K, splits), G3_BK) * G3_BK;
  p.a_div = a_batch_div;
  p.b_div = b_batch_div;
  p.a_mn = a_mn ? 1 : 0;
  p.b_mn = b_mn ? 1 : 0;
  p.d_trans = d_trans ? 1 : 0;
  if (splits > 1) {
    const size_t need = (size_t)batch * splits * M * N * sizeof(float);
    LOB_REQUIRE(ws && ws_bytes >= need, "lob_gemm3x: workspace too small for the split-K partial sums");
    p.partial = (float*)ws;
  }
  // epilogue through shared memory + TMA when the output (or the partial-sum buffer) and E are TMA-addressable
  CUtensorMap tmD, tmE;
  memset(&tmD, 0, sizeof(tmD));
  memset(&tmE, 0, sizeof(tmE));
  p.tma_epi = 0;
  if (!p.d_trans) {
    bool ok;
    if (splits > 1)
      ok = g3_make_rm_map(&tmD, p.partial, M, N, N, M * N, batch * splits, 32);
    else
      ok = (d_dtype == LOB_F32) && g3_make_rm_map(&tmD, p.D, M, N, ldd, d_batch_stride, batch, 32);
    if (ok && splits == 1 && p.E) ok = g3_make_rm_map(&tmE, p.E, M, N, lde, e_batch_stride, batch, 128);
    p.tma_epi = ok ? 1 : 0;
  }
  // short contractions are epilogue-bound: trade one operand stage for four more chunks in flight
  const int64_t max_kb = cdiv(std::min<int64_t>(K, p.kps), G3_BK);
  if (p.tma_epi && max_kb <= 16) {
    p.stages = 2;
    p.epi_slots = 6;
  } else {
    p.stages = 3;
    p.epi_slots = 2;
  }
  LOB_CUDA(cudaFuncSetAttribute(k_gemm3x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G3_SMEM));
  const unsigned grid = (unsigned)std::min<int64_t>(p.sched_flat ? p.total : p.jobs, kNumSMs);
  k_gemm3x<<<grid, G3_THREADS, G3_SMEM, st>>>(tmA, tmB, tmD, tmE, p);
  LOB_TRY(check_launch("k_gemm3x"));
  if (splits > 1) {
    const unsigned rgrid = (unsigned)cdiv(batch * M * N, 256);
    if (d_dtype == LOB_F64)
      k_g3_reduce<double><<<rgrid, 256, 0, st>>>(p, (double*)D);
    else
      k_g3_reduce<float><<<rgrid, 256, 0, st>>>(p, (float*)D);
    LOB_TRY(check_launch("k_g3_reduce"));
  }
  return LOB_OK;
}

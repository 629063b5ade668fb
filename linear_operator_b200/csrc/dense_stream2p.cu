// dense_stream2p.cu -- CTA-pair (tcgen05 cta_group::2) version of the streaming kernel in dense_stream2.cu for the fp32
// dense operator matmul  Y = alpha (A X) + d (.) E  with the fused <E, Y> partials.
// Reference arithmetic: operators/dense_linear_operator.py:60-64, operators/added_diag_linear_operator.py:72-76,135-140,
// utils/linear_cg.py:250-251.
//
// dense_stream2.cu is bound by the shared-memory data pipe (ncu: 89.6 % busy): per 32-column k block a CTA moves 32 KB
// of operator three times (TMA write, tensor-core read of A_hi, converter read for A_lo) plus 42 KB for the X operand
// (10 KB of TMA writes, 32 KB of tensor-core reads: the [X_hi ; X_lo] tile is re-read by every one of the 16 MMAs).
// Here the two CTAs of a cluster (the two SMs of a TPC) each own 256 operator rows of a 512-row pair tile and the MMAs
// are M = 256 pair instructions issued by the leader CTA: every CTA supplies its own 128 rows of A (shared memory for
// A_hi, its own TMEM operand slot for A_lo) and only HALF of the B operand, so the X traffic through each SM's shared
// memory halves (6 KB written, 18 KB read per k block; 120 KB instead of 138 KB per 32 KB of operator).
//
// Column layout.  With h = the half width (C / 2 rounded up to 8; 24 at C = 33) the pre-split operand has 4 h rows
//   [ X_hi[0:h] ; X_lo[h:2h] | X_hi[h:2h] ; X_lo[0:h] ]        (CTA 0 loads the first 2 h rows, CTA 1 the last 2 h)
// so that ONE descriptor serves both instructions of a k step:
//   hi MMA  (A_hi from shared memory, N = 4 h): D[:, 0:h]   = A_hi X_hi[0:h]     D[:, h:2h]  = A_hi X_lo[h:2h]
//                                               D[:, 2h:3h] = A_hi X_hi[h:2h]    D[:, 3h:4h] = A_hi X_lo[0:h]
//   lo MMA  (A_lo from TMEM, N = 2 h: the first h rows of either CTA's half):
//                                               D[:, 0:h]  += A_lo X_hi[0:h]     D[:, h:2h] += A_lo X_hi[h:2h]
//   y[:, c] = D[:, c] + D[:, 3h + c]  (c < h),   y[:, c] = D[:, c] + D[:, h + c]  (c >= h)        (3xTF32)
// Column counts just above a multiple of 16 (C = 2 g + e, g % 8 == 0, 1 <= e <= 4: the solver's 32 probes + 1 right-hand
// side) take a tighter layout: 2 g + 8 rows per CTA instead of 2 h = 2 (g + 8),
//   [ X_hi[0:g] ; extras (8 rows: x_hi, x_lo of column 2g + j at rows 2j, 2j + 1, zeros) ; X_lo[g:2g] |
//     X_hi[g:2g] ; zeros (8) ; X_lo[0:g] ]
// hi MMA N = 4 g + 16 (80 instead of 96 at C = 33), lo MMA N = 2 g + 16 over the first g + 8 rows of either half: it
// adds A_lo x_hi (and a harmless A_lo x_lo) to the extras' own columns and A_lo 0 to the columns behind CTA 1's hi rows.
//   y[:, c] = D[c] + D[3g+16+c] (c < g),  D[c+8] + D[c+8+g] (g <= c < 2g),  D[g+2j] + D[g+2j+1] (c = 2g + j)
// TMEM per CTA (512 columns): 4 accumulators x N_hi columns (2 M tiles x 2 buffers) + A_lo operand slots of 2 BK columns.
// Synchronisation: TMA -> converters and MMA commit -> TMA / converters / epilogue are CTA-local barriers (the commits
// are multicast to both CTAs); converters -> MMA and epilogue -> MMA arrive on the LEADER's barriers from both CTAs.
// Warp roles (512 threads per CTA): 0 TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM allocator, 4-7 epilogue,
// 8-15 converters.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05_util.cuh"

namespace lob {

constexpr int P2_ROWS = 256;  // operator rows per CTA and tile (its halves of two M = 256 pair MMAs)
constexpr int P2_THREADS = 512;
constexpr int P2_MAX_ST = 12;
constexpr int P2_RED_DOUBLES = 2 * 2 * 4 * 48;  // [buffer][tile][quarter][column]

struct P2Params {
  float* Y;
  const float* E;
  const float* alpha;
  int64_t alpha_bs;
  const float* dg;
  int64_t d_bs, d_st;
  double* dots;
  int64_t M, K, C;
  int n_parts;
  int G;         // half width g of the regular columns (multiple of 8): columns [0, 2 g)
  int MX;        // 0, or 8: extras block (columns [2 g, C), at most 4) in the middle of CTA 0's tile
  int R;         // rows of one CTA's X tile: 2 g + MX
  int xbytes;    // bytes of one CTA's X tile (R x BK x 4, rounded up to 1 KB)
  int SA;        // ring stages
  int MTP;       // pair tiles (512 rows) per batch element
  int64_t nptiles;
  int a_shared;
  uint32_t idesc_hi, idesc_lo;
  int acc_stride;  // TMEM columns per accumulator (4 h)
  int slot_base;   // first A_lo operand slot column
  int nslot;       // A_lo operand slots (each 2 * BK columns)
  int acc_bufs;
  int seg;         // k blocks per accumulator segment (see "segments" in the kernel); >= number of k blocks: one segment
  int ysum_ld;     // leading dimension (floats) of the running-sum tile in shared memory, 0 when there is one segment
  int dbg;         // harness experiments (LOB_DIAG builds): 1 skip lo MMA, 2 skip conversion, 4 skip all MMAs
};

template <int BK>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_dense_stream2p(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX, P2Params p) {
  using namespace ds;
  constexpr int A_STAGE = P2_ROWS * BK * 4;
  constexpr int ROW_BYTES = BK * 4;
  constexpr int NU = BK / 4;  // 16-byte units per operator row
  const int NSLOT = p.nslot;

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = A_STAGE + p.xbytes;
  unsigned char* sRing = smem;
  // Segments.  The tensor core truncates when it adds into its fp32 accumulator: a bias that grows with the number of
  // k steps (1.5e-8 per contraction index on all-positive data).  Long contractions are therefore cut into segments of
  // p.seg k blocks that alternate between the two TMEM accumulator buffers; the epilogue warps drain a finished segment
  // into a running sum in shared memory (round-to-nearest adds) while the MMAs of the next one run, and only the last
  // segment goes through the output epilogue.  The bias no longer depends on K.
  float* ysum = reinterpret_cast<float*>(smem + p.SA * stage_bytes);  // [P2_ROWS][ysum_ld]
  double* dred = reinterpret_cast<double*>(smem + p.SA * stage_bytes + (size_t)P2_ROWS * p.ysum_ld * 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dred + P2_RED_DOUBLES);
  uint64_t* full = bars;                   // [MAX_ST] TMA -> converters                      (local)
  uint64_t* empty = full + P2_MAX_ST;      // [MAX_ST] MMA commit (multicast) -> TMA           (local)
  uint64_t* lo_full = empty + P2_MAX_ST;   // [NSLOT]  converters of both CTAs -> MMA          (leader's copy is used)
  uint64_t* lo_empty = lo_full + 8;        // [NSLOT]  MMA commit (multicast) -> converters    (local)
  uint64_t* acc_full = lo_empty + 8;       // [2]      MMA commit (multicast) -> epilogue      (local)
  uint64_t* acc_empty = acc_full + 2;      // [2]      epilogue of both CTAs -> MMA            (leader's copy is used)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int64_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb = (int)((p.K + BK - 1) / BK);
  const int nseg = (nkb + p.seg - 1) / p.seg;
  const int G = p.G, MX = p.MX;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int i = 0; i < p.SA; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
    }
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(smem_u32(&lo_full[i]), 16);
      mbar_init(smem_u32(&lo_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_holder))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer (each CTA loads its 256 operator rows and its half of the X operand) ==========
    if (elect_one()) {
      uint64_t pol_stream, pol_keep;
      asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
      int s = 0;
      uint32_t ph = 0;
      const uint32_t xtx = (uint32_t)(p.R * BK * 4);
      for (int64_t pt = pair; pt < p.nptiles; pt += npairs) {
        const int b = (int)(pt / p.MTP);
        const int m0 = ((int)(pt - (int64_t)b * p.MTP) * 2 + (int)rank) * P2_ROWS;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t dst = smem_u32(sRing + s * stage_bytes);
          const uint32_t bar = smem_u32(&full[s]);
          mbar_arrive_expect_tx(bar, (uint32_t)A_STAGE + xtx);
          tma_load_3d(dst + A_STAGE, &tmX, bar, kb * BK, (int)rank * p.R, b, pol_keep);
          tma_load_3d(dst, &tmA, bar, kb * BK, m0, p.a_shared ? 0 : b, pol_stream);
          if (++s == p.SA) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA drives both SMs =====================
    if (rank == 0) {
      int s = 0, sl = 0;
      uint32_t phl = 0, it = 0;
      for (int64_t pt = pair; pt < p.nptiles; pt += npairs) {
        for (int kb0 = 0; kb0 < nkb; kb0 += p.seg, ++it) {  // one accumulator buffer per segment
          const int kb1 = min(nkb, kb0 + p.seg);
          const uint32_t buf = (p.acc_bufs == 2) ? (it & 1) : 0u;
          const uint32_t accph = (p.acc_bufs == 2) ? ((it >> 1) & 1) : (it & 1);
          mbar_wait(smem_u32(&acc_empty[buf]), accph ^ 1);
          const uint32_t d0 = tmem_base + buf * 2 * p.acc_stride;
          for (int kb = kb0; kb < kb1; ++kb) {
            // the converters of a CTA pass its full[s] before they arrive here: lo_full also says "both tiles landed"
            mbar_wait(smem_u32(&lo_full[sl]), phl);
            __syncwarp();
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_addr = smem_u32(sRing + s * stage_bytes);
              const uint64_t xdesc = make_kmajor_desc<BK>(a_addr + A_STAGE);
              const uint32_t lo_slot = tmem_base + p.slot_base + sl * (2 * BK);
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const uint64_t adesc = make_kmajor_desc<BK>(a_addr + t * (128 * ROW_BYTES));
                const uint32_t d_addr = d0 + t * p.acc_stride;
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                  const uint64_t adv = (uint64_t)((k * 32) >> 4);  // 8 tf32 = 32 bytes inside the swizzle span
                  if (!(p.dbg & 4))
                    umma_tf32_ss_pair(d_addr, adesc + adv, xdesc + adv, p.idesc_hi, ((kb - kb0) | k) ? 1u : 0u);
                  if (!(p.dbg & 5)) umma_tf32_ts_pair(d_addr, lo_slot + t * BK + k * 8, xdesc + adv, p.idesc_lo, 1u);
                }
              }
              umma_commit_pair(smem_u32(&empty[s]));
              umma_commit_pair(smem_u32(&lo_empty[sl]));
              if (kb == kb1 - 1) umma_commit_pair(smem_u32(&acc_full[buf]));
            }
            __syncwarp();
            if (++s == p.SA) s = 0;
            if (++sl == NSLOT) { sl = 0; phl ^= 1; }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: thread = operator row of this CTA's 256 =====================
    const int q = warp & 3;
    const int C = (int)p.C;
    const bool need_e = (p.dg != nullptr) || (p.dots != nullptr);
    const int et = threadIdx.x - 128;  // 0..127 inside the epilogue group
    const uint32_t acc_empty_leader = map_to_cta(smem_u32(&acc_empty[0]), 0);
    uint32_t it = 0, tile_no = 0;
    for (int64_t pt = pair; pt < p.nptiles; pt += npairs, ++tile_no) {
      const int64_t b = pt / p.MTP;
      const int mt = (int)(pt - b * p.MTP) * 2 + (int)rank;
      const int64_t m0 = (int64_t)mt * P2_ROWS;
      const float alpha_b = p.alpha ? p.alpha[b * p.alpha_bs] : 1.0f;
      const float* Eb = p.E + b * p.M * C;
      float* Yb = p.Y + b * p.M * C;
      double* red = dred + (tile_no & 1) * (2 * 4 * 48);
#pragma unroll 1
      for (int seg = 0; seg < nseg; ++seg, ++it) {
        const bool first = seg == 0, last = seg == nseg - 1;
        const uint32_t buf = (p.acc_bufs == 2) ? (it & 1) : 0u;
        const uint32_t accph = (p.acc_bufs == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(smem_u32(&acc_full[buf]), accph);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const int64_t row = m0 + t * 128 + q * 32 + lane;
          const bool rok = row < p.M;
          const float dv = (p.dg && rok && last) ? __ldg(p.dg + b * p.d_bs + row * p.d_st) : 0.f;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + t) * p.acc_stride;
          float* ys = ysum + (t * 128 + q * 32 + lane) * p.ysum_ld;  // this row's running sum (odd pitch: no conflicts)
#pragma unroll 1
          for (int c0 = 0; c0 < 2 * G + 2 * MX; c0 += 16) {
            float e[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              e[i] = (need_e && last && rok && c0 + i < C) ? __ldg(Eb + row * C + c0 + i) : 0.f;
            uint32_t hi[16], lo[16];
            if (c0 < 2 * G) {
              // two 8-column groups (G % 8 == 0: a group never straddles the halves); partner columns per the layout
              const int ca = c0, cb = c0 + 8;
              DS_LD8(taddr + (ca < G ? ca : ca + MX), hi);
              DS_LD8(taddr + (ca < G ? 3 * G + 2 * MX + ca : ca + MX + G), lo);
              DS_LD8(taddr + (cb < G ? cb : cb + MX), (hi + 8));
              DS_LD8(taddr + (cb < G ? 3 * G + 2 * MX + cb : cb + MX + G), (lo + 8));
            } else {
              // the extras: column 2 g + j sits at D[g + 2 j] (A_hi x_hi + A_lo x_hi) and D[g + 2 j + 1] (A_hi x_lo)
              uint32_t v[8];
              DS_LD8(taddr + G, v);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int i = 0; i < 16; ++i) hi[i] = lo[i] = 0u;
#pragma unroll
              for (int i = 0; i < 4; ++i) hi[i] = v[2 * i], lo[i] = v[2 * i + 1];
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              acc[i] = __uint_as_float(hi[i]) + __uint_as_float(lo[i]);
              if (!first && c0 + i < C) acc[i] += ys[c0 + i];
            }
            if (!last) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < C) ys[c0 + i] = acc[i];
              continue;
            }
            double pd[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float y = fmaf(dv, e[i], acc[i] * alpha_b);
              const bool ok = rok && c0 + i < C;
              if (ok) Yb[row * C + c0 + i] = y;
              pd[i] = ok ? (double)e[i] * (double)y : 0.0;
            }
            if (p.dots) {
              // column sums over the 32 rows of this warp as a transpose-reduce (fixed tree: deterministic); see
              // dense_stream2.cu
#pragma unroll
              for (int hh = 8, o = 16; hh >= 1; hh >>= 1, o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int j = 0; j < hh; ++j) {
                  const double send = up ? pd[j] : pd[j + hh];
                  const double keep = up ? pd[j + hh] : pd[j];
                  pd[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
              }
              pd[0] += __shfl_xor_sync(0xffffffffu, pd[0], 1);
              const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
              if ((lane & 1) == 0) red[(t * 4 + q) * 48 + c0 + col] = pd[0];
            }
          }
        }
        // accumulators drained: hand the TMEM buffer of this CTA back to the leader's MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader + buf * 8);
      }
      if (p.dots) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int idx = et; idx < 2 * C; idx += 128) {
          const int t = idx / C, c = idx - t * C;
          const int64_t pi = (int64_t)mt * 2 + t;
          if (pi < p.n_parts) {
            const double* r4 = red + t * 4 * 48 + c;
            p.dots[(b * p.n_parts + pi) * C + c] = (r4[0] + r4[48]) + (r4[96] + r4[144]);
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ===================== converters: A_lo rows -> this CTA's TMEM operand slot =====================
    const int q = warp & 3;
    const int t = (warp - 8) >> 2;
    const int row = t * 128 + q * 32 + lane;
    const uint32_t swz = (BK == 32) ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + p.slot_base + t * BK;
    const uint32_t lo_full_leader = map_to_cta(smem_u32(&lo_full[0]), 0);
    int s = 0, sl = 0;
    uint32_t ph = 0, phl = 0;
    for (int64_t pt = pair; pt < p.nptiles; pt += npairs) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        const unsigned char* arow = sRing + s * stage_bytes + row * ROW_BYTES;
        uint32_t lo[BK];
        if (!(p.dbg & 2)) {
#pragma unroll
          for (int u = 0; u < NU; ++u) {
            const uint4 v = *reinterpret_cast<const uint4*>(arow + ((u ^ swz) << 4));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float l = __uint_as_float(w[j]) - __uint_as_float(w[j] & 0xFFFFE000u);
              lo[u * 4 + j] = (__float_as_uint(l) + 0x1000u) & 0xFFFFE000u;  // tf32, round to nearest
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < BK; ++j) lo[j] = 0u;
        }
        mbar_wait(smem_u32(&lo_empty[sl]), phl ^ 1);
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = lane_addr + sl * (2 * BK);
        DS_ST16(taddr, lo);
        if constexpr (BK == 32) DS_ST16(taddr + 16, (lo + 16));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lo_full_leader + sl * 8);
        if (++s == p.SA) { s = 0; ph ^= 1; }
        if (++sl == NSLOT) { sl = 0; phl ^= 1; }
      }
    }
  }

  // ---- teardown: neither CTA may exit (or free its TMEM) while the pair's MMAs / remote arrivals are in flight ----
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Xs (B, 4 h, Kp), K-major, tf32-rounded:  rows [0,h) x_hi[:, 0:h] | [h,2h) x_lo[:, h:2h] | [2h,3h) x_hi[:, h:2h] |
// [3h,4h) x_lo[:, 0:h];  k >= K and c >= C are zeros.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SPLITP_KT = 64;

__global__ void __launch_bounds__(256)
k_split_x2p(const float* __restrict__ X, float* __restrict__ Xs, int64_t K, int64_t Kp, int C, int G, int MX) {
  extern __shared__ __align__(16) float xsp[];  // [SPLITP_KT][C | 1]
  const int ldx = C | 1;
  const int64_t b = blockIdx.y;
  const int64_t k0 = (int64_t)blockIdx.x * SPLITP_KT;
  const int kvalid = (int)max((int64_t)0, min((int64_t)SPLITP_KT, K - k0));
  const float* src = X + (b * K + k0) * C;
  // the 64 x C block is one contiguous run: 16-byte loads when the pitch needs no padding (odd C) and the run is
  // 16-byte addressable, else element-wise with (row, column) advanced without divisions
  if (ldx == C && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int nv = (kvalid * C) >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(xsp);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = s4[i];
    for (int i = (nv << 2) + threadIdx.x; i < kvalid * C; i += blockDim.x) xsp[i] = src[i];
  } else {
    const int dq = blockDim.x / C, dr = blockDim.x - dq * C;
    int r = threadIdx.x / C, c = threadIdx.x - r * C;
    for (int e = threadIdx.x; e < kvalid * C; e += blockDim.x) {
      xsp[r * ldx + c] = src[e];
      r += dq;
      c += dr;
      if (c >= C) { c -= C; ++r; }
    }
  }
  __syncthreads();
  // each thread writes one 16-byte word (4 consecutive k of one operand row): 16 lanes cover the 64 k of a row
  const int q4 = threadIdx.x & 15;
  const int kw = (int)min((int64_t)SPLITP_KT, Kp - k0);
  const int RT = 2 * G + MX, R = 2 * RT;  // rows of one CTA's tile, of both
  for (int r = threadIdx.x >> 4; r < R; r += blockDim.x >> 4) {
    const int cta = r >= RT, j = r - cta * RT;
    int part, c;  // 0: hi rows, 1: lo rows; c >= C: zero row
    if (j < G) {
      part = 0, c = cta * G + j;
    } else if (j < G + MX) {
      const int jj = j - G;
      part = jj & 1, c = cta ? C : 2 * G + (jj >> 1);  // extras live in CTA 0's tile only
    } else {
      part = 1, c = (1 - cta) * G + (j - G - MX);
    }
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = 4 * q4 + i;
        if (kk < kvalid) {
          const float x = xsp[kk * ldx + c];
          const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
          o[i] = part == 0 ? hi : __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
        }
      }
    }
    if (4 * q4 < kw) *reinterpret_cast<float4*>(Xs + (b * R + r) * Kp + k0 + 4 * q4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled_p2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled_p2 p2_encode_fn() {
  static PFN_encodeTiled_p2 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled_p2)ptr;
  }
  return fn;
}

struct P2Config {
  int acc_bufs;  // 0/2: double-buffered accumulators; 1: single (more A_lo operand slots in TMEM)
  int bk;        // 16 (SWIZZLE_64B) or 32 (SWIZZLE_128B)
  int sa;        // ring depth (0 = deepest that fits)
  int grid;      // CTAs (even; 0 = one per SM)
  int dbg;
};

static P2Config p2_default_config() {
  static P2Config cfg = [] {
    P2Config c{0, 32, 0, 0, 0};
#ifdef LOB_DIAG  // tuning / bottleneck-experiment knobs exist only in the harness build
    if (const char* e = getenv("LOB_P2_ACC_BUFS")) c.acc_bufs = atoi(e);
    if (const char* e = getenv("LOB_P2_DBG")) c.dbg = atoi(e);
    if (const char* e = getenv("LOB_P2_BK")) c.bk = atoi(e);
    if (const char* e = getenv("LOB_P2_SA")) c.sa = atoi(e);
    if (const char* e = getenv("LOB_P2_GRID")) c.grid = atoi(e);
#endif
    return c;
  }();
  return cfg;
}

constexpr size_t P2_SMEM_MAX = 232448;
constexpr size_t P2_SMEM_FIXED = 1024 /*alignment*/ + P2_RED_DOUBLES * 8 + 512 /*barriers*/;

// column layout: regular half width g and extras block (0 or 8 rows); rows of one CTA's operand tile = 2 g + mx
struct P2Layout {
  int g, mx;
  int rows() const { return 2 * g + mx; }
};
static P2Layout p2_layout(int64_t C) {
  const int gl = (int)((C - 1) / 16 * 8);  // largest g (multiple of 8) with 2 g < C
  if (gl >= 8 && C - 2 * gl <= 4) return {gl, 8};
  return {(int)(((C + 1) / 2 + 7) / 8 * 8), 0};
}

size_t dense_stream2p_workspace_bytes(int64_t B, int64_t K, int64_t C) {
  if (B <= 0 || K <= 0 || C <= 0 || C > 48) return 0;
  const int64_t Kp = (K + 3) / 4 * 4;
  return (size_t)B * 2 * p2_layout(C).rows() * Kp * sizeof(float);
}

// returns LOB_ERR_UNSUPPORTED when the shape does not qualify (caller falls back to dense_stream2.cu)
int dense_matmul_stream2p_f32_cfg(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                                  const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                                  const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                                  cudaStream_t st, P2Config cfg) {
  if (C > 48 || C < 1) return LOB_ERR_UNSUPPORTED;
  if ((lda % 4) != 0 || (a_bs % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0) return LOB_ERR_UNSUPPORTED;
  if (M >= (1LL << 31) - 512 || K >= (1LL << 31) || B >= (1LL << 31)) return LOB_ERR_UNSUPPORTED;
  if (!ws || (reinterpret_cast<uintptr_t>(ws) & 15) != 0 || ws_bytes < dense_stream2p_workspace_bytes(B, K, C))
    return LOB_ERR_UNSUPPORTED;
  if ((d || dots) && !E && M != K) return LOB_ERR_UNSUPPORTED;
  PFN_encodeTiled_p2 enc = p2_encode_fn();
  if (!enc) return LOB_ERR_UNSUPPORTED;
  const int BK = (cfg.bk == 16) ? 16 : 32;
  const P2Layout lay = p2_layout(C);
  const int R = lay.rows();
  const int64_t Kp = (K + 3) / 4 * 4;
  const bool shared = (a_bs == 0);
  const CUtensorMapSwizzle swz = (BK == 32) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  CUtensorMap tmA, tmX;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)(shared ? 1 : B)};
    cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)(shared ? (cuuint64_t)M * lda : a_bs) * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, P2_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)(2 * R), (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)(2 * R) * Kp * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)R, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ws, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  }

  const int a_stage = P2_ROWS * BK * 4;
  const int xbytes = (int)align_up((size_t)R * BK * 4, 1024);
  const int stage = a_stage + xbytes;
  // accumulator segments of 64 k blocks (2048 contraction indices); one segment needs no running-sum tile
  const int nkb_host = (int)cdiv(K, BK);
  const int seg = (cfg.dbg & 4096) ? (1 << 30) : 64;
  const int ysum_ld = nkb_host > seg ? ((int)C | 1) : 0;
  const size_t ysum_bytes = align_up((size_t)P2_ROWS * ysum_ld * 4, 16);
  int sa = cfg.sa > 0 ? cfg.sa : (int)((P2_SMEM_MAX - P2_SMEM_FIXED - ysum_bytes) / stage);
  if (sa > P2_MAX_ST) sa = P2_MAX_ST;
  if (sa < 2) return LOB_ERR_UNSUPPORTED;
  const size_t smem = P2_SMEM_FIXED + (size_t)sa * stage + ysum_bytes;
  if (smem > P2_SMEM_MAX) return LOB_ERR_UNSUPPORTED;

  P2Params p;
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
  p.G = lay.g;
  p.MX = lay.mx;
  p.R = R;
  p.xbytes = xbytes;
  p.SA = sa;
  p.MTP = (int)cdiv(M, 2 * P2_ROWS);
  p.nptiles = B * p.MTP;
  p.a_shared = shared ? 1 : 0;
  p.idesc_hi = ds::make_idesc_tf32(256, 2 * R);                  // either CTA supplies its R rows
  p.idesc_lo = ds::make_idesc_tf32(256, 2 * (lay.g + lay.mx));   // ... resp. its first g + mx rows
  p.dbg = cfg.dbg;
  p.seg = seg;
  p.ysum_ld = ysum_ld;
  p.acc_bufs = (cfg.acc_bufs == 1) ? 1 : 2;
  p.acc_stride = 2 * R;
  p.slot_base = 2 * p.acc_bufs * p.acc_stride;
  p.nslot = (512 - p.slot_base) / (2 * BK);
  if (p.nslot > 8) p.nslot = 8;
  if (p.nslot < 2) return LOB_ERR_UNSUPPORTED;

  {
    dim3 grid((unsigned)cdiv(Kp, SPLITP_KT), (unsigned)B);
    const size_t sm = (size_t)SPLITP_KT * ((int)C | 1) * sizeof(float);
    k_split_x2p<<<grid, 256, sm, st>>>(X, (float*)ws, K, Kp, (int)C, lay.g, lay.mx);
    LOB_TRY(check_launch("k_split_x2p"));
  }

  int64_t npairs = std::min<int64_t>(p.nptiles, (cfg.grid > 0 ? cfg.grid : kNumSMs) / 2);
  if (npairs < 1) npairs = 1;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(2 * npairs));
  lc.blockDim = dim3(P2_THREADS);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  if (BK == 32) {
    auto kern = k_dense_stream2p<32>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P2_SMEM_MAX));
    LOB_CUDA(cudaLaunchKernelEx(&lc, kern, tmA, tmX, p));
  } else {
    auto kern = k_dense_stream2p<16>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P2_SMEM_MAX));
    LOB_CUDA(cudaLaunchKernelEx(&lc, kern, tmA, tmX, p));
  }
  return check_launch("k_dense_stream2p");
}

int dense_matmul_stream2p_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                              const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                              const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                              cudaStream_t st) {
  return dense_matmul_stream2p_f32_cfg(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, ws,
                                       ws_bytes, st, p2_default_config());
}

}  // namespace lob

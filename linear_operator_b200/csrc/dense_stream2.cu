// dense_stream2.cu -- second-generation streaming kernel for the fp32 dense operator matmul  Y = alpha (A X) + d (.) E.
// Reference arithmetic: operators/dense_linear_operator.py:60-64, operators/added_diag_linear_operator.py:72-76,135-140,
// utils/linear_cg.py:250-251 (fused <p, Ap>).
//
// Same persistent skeleton as dense_stream.cu (one CTA per SM, static tile schedule, TMA ring, elect.sync'd UMMA issue,
// double-buffered TMEM accumulators drained by dedicated epilogue warps), but "A-stationary": ncu showed the first
// kernel saturating the shared-memory data pipe (tensor-core operand reads 36 % + converter LDS/STS 39 % + TMA writes
// 16 % at 0.63 of the HBM roofline), so this one moves less through shared memory:
//   * the operator tile (256 rows x BK columns, as TMA lands it) is the UMMA *A* operand of two M = 128 MMAs per k step,
//     read raw from shared memory (the tensor core drops the 13 low mantissa bits: A_hi for free);
//   * A_lo = A - tf32(A) never goes back to shared memory: each converter thread owns one operator row, reads its BK
//     values (swizzle-aware, conflict-free LDS.128), and stores the correction straight into a TMEM operand slot
//     (tcgen05.st); a second, narrower MMA per k step takes its A operand from TMEM;
//   * the right-hand side is the UMMA *B* operand, N = 2 CP rows [X_hi ; X_lo] (CP = C rounded up to 8), K-major,
//     pre-split into a workspace by k_split_x2 and TMA-loaded beside the A tile.  Only 2 CP x BK x 4 bytes of it are
//     read per MMA (vs. 8 KB of A_lo + 4 KB of X in the first kernel), and the tensor time per k step drops from
//     2 x 128 to 2 x (CP/2.67 + CP/5.3) cycles.
//   D[:, c] = A_hi X_hi + A_lo X_hi,  D[:, CP + c] = A_hi X_lo:  y = D[:, c] + D[:, CP + c]   (3xTF32, fp32 accumulate)
// TMEM (512 columns): 4 accumulators x 2 CP columns (2 M tiles x 2 buffers: the epilogue of tile i overlaps the main
// loop of tile i + 1) + the remaining columns as A_lo operand slots of 2 BK columns (3 slots at C = 33, BK = 32).
// Warp roles (512 threads): 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 2-3 in-kernel X producers (only when no
// workspace was supplied), 4-7 epilogue (thread = operator row: tcgen05.ld, hi + lo, alpha, + d (.) E, fp64 <E, Y>
// partials by a transpose-reduce, direct stores), 8-15 converters.
// Measured at BASELINE config 2 (B = 1024, N = 5000, C = 33): 19.9 - 21.7 ms per launch inside the power-capped solve
// (0.74 - 0.81 of the measured HBM peak), DRAM traffic 1.025 x algorithmic, shared-memory data pipe 89.6 % busy (ncu,
// profiles/r1_dense_stream2_ncu.md).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05_util.cuh"

namespace lob {

constexpr int D2_ROWS = 256;  // operator rows per tile (two UMMA M = 128 tiles)
constexpr int D2_THREADS = 512;
constexpr int D2_ACC_COLS = 96;
constexpr int D2_SLOT_BASE = 4 * D2_ACC_COLS;  // 384
constexpr int D2_MAX_ST = 12;
constexpr int D2_RED_DOUBLES = 2 * 2 * 4 * 48;  // [buffer][tile][quarter][column]

struct D2Params {
  float* Y;
  const float* X;      // (B, K, C) right-hand side (read by the in-kernel X producers when xmode == 1)
  const float* E;
  const float* alpha;
  int64_t alpha_bs;
  const float* dg;
  int64_t d_bs, d_st;
  double* dots;
  int64_t M, K, C;
  int n_parts;
  int CP;        // C rounded up to 8: X_hi rows [0, CP), X_lo rows [CP, 2 CP)
  int xbytes;    // bytes of one X tile (2 CP x BK x 4, rounded up to 1 KB)
  int SA;        // ring stages
  int MT;
  int64_t ntiles;
  int a_shared;
  uint32_t idesc_hi, idesc_lo;
  int acc_stride;  // TMEM columns per accumulator
  int slot_base;   // first A_lo operand slot column
  int nslot;       // A_lo operand slots (each 2 * BK columns)
  int acc_bufs;    // 2: accumulators double-buffered (epilogue overlaps the next tile); 1: single, more operand slots
  int xmode;     // 0: X operand tiles TMA-loaded from the pre-split workspace; 1: produced in the kernel (warps 2-3)
  int dbg;       // harness experiments: 1 skip lo MMA, 2 skip conversion, 4 skip all MMAs
  int estage;    // bytes of one staged E / Y tile (256 rows x C floats) or 0: see "staged epilogue" below
};

template <int BK>
__global__ void __launch_bounds__(D2_THREADS, 1)
k_dense_stream2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX, D2Params p) {
  using namespace ds;
  constexpr int A_STAGE = D2_ROWS * BK * 4;
  const int NSLOT = p.nslot;  // A_lo operand slots in TMEM (each: 2 M tiles x BK columns)
  constexpr int ROW_BYTES = BK * 4;
  constexpr int NU = BK / 4;  // 16-byte units per operator row

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = A_STAGE + p.xbytes;
  unsigned char* sRing = smem;
  // staged epilogue (short contractions, where the epilogue IS the kernel): the E rows of a tile are one contiguous
  // run of global memory, so the TMA warp bulk-copies them into shared memory a tile ahead, the epilogue threads
  // overwrite them in place with Y and one thread bulk-stores the tile -- no dependent, 132-byte-strided global loads
  // and stores in the per-row epilogue threads.  Two buffers.
  float* sE = reinterpret_cast<float*>(smem + p.SA * stage_bytes);
  double* dred = reinterpret_cast<double*>(smem + p.SA * stage_bytes + 2 * p.estage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dred + D2_RED_DOUBLES);
  uint64_t* full = bars;                    // [MAX_ST] TMA -> converters, MMA
  uint64_t* empty = full + D2_MAX_ST;       // [MAX_ST] MMA (commit) -> TMA
  uint64_t* x_full = empty + D2_MAX_ST;     // [MAX_ST] X producers -> MMA (xmode 1)
  uint64_t* lo_full = x_full + D2_MAX_ST;   // [NSLOT]  converters -> MMA
  uint64_t* lo_empty = lo_full + 8;         // [NSLOT]  MMA (commit) -> converters
  uint64_t* acc_full = lo_empty + 8;        // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint64_t* e_full = acc_empty + 2;         // [2] bulk copy of the E tile -> epilogue
  uint64_t* e_empty = e_full + 2;           // [2] bulk store of the Y tile has read the buffer -> TMA warp
  uint64_t* y_ready = e_empty + 2;          // [2] epilogue warps -> store warp
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(y_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (int)((p.K + BK - 1) / BK);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    for (int i = 0; i < p.SA; ++i) {
      mbar_init(smem_u32(&full[i]), 1);
      mbar_init(smem_u32(&empty[i]), 1);
      mbar_init(smem_u32(&x_full[i]), 2);
    }
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(smem_u32(&lo_full[i]), 8);
      mbar_init(smem_u32(&lo_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&acc_full[i]), 1);
      mbar_init(smem_u32(&acc_empty[i]), 4);
      mbar_init(smem_u32(&e_full[i]), 1);
      mbar_init(smem_u32(&e_empty[i]), 1);
      mbar_init(smem_u32(&y_ready[i]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_holder))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.xmode == 1) {
    // X tiles are written by the producer warps; their padding rows (c >= C) stay zero for the whole kernel
    for (int s = 0; s < p.SA; ++s) {
      uint4* xt = reinterpret_cast<uint4*>(smem + s * (A_STAGE + p.xbytes) + A_STAGE);
      for (int i = threadIdx.x; i < p.xbytes / 16; i += D2_THREADS) xt[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async_smem();
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
      // measured (ncu dram__bytes_read): evict_first on the operator stream makes L2 drop the promoted 256-byte lines
      // before the next k-block of the same rows needs their other half: +8.7 % DRAM reads.  evict_normal avoids it.
      if (!(p.dbg & 512)) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
      if (p.dbg & 256) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_stream));
      int s = 0;
      uint32_t ph = 0, it = 0;
      const uint32_t xtx = (uint32_t)(2 * p.CP * BK * 4);
      for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const int b = (int)(tile / p.MT);
        const int m0 = (int)(tile - (int64_t)b * p.MT) * D2_ROWS;
        if (p.estage) {
          const uint32_t eb = it & 1;
          mbar_wait(smem_u32(&e_empty[eb]), ((it >> 1) & 1) ^ 1);
          const uint32_t ebytes = (uint32_t)(min((int64_t)D2_ROWS, p.M - m0) * p.C * 4);
          const uint32_t bar = smem_u32(&e_full[eb]);
          mbar_arrive_expect_tx(bar, ebytes);
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(smem_u32(reinterpret_cast<unsigned char*>(sE) + eb * p.estage)),
              "l"(p.E + ((int64_t)b * p.M + m0) * p.C), "r"(ebytes), "r"(bar)
              : "memory");
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(&empty[s]), ph ^ 1);
          const uint32_t dst = smem_u32(sRing + s * stage_bytes);
          const uint32_t bar = smem_u32(&full[s]);
          if (p.xmode == 1) {
            mbar_arrive_expect_tx(bar, (uint32_t)A_STAGE);
          } else {
            mbar_arrive_expect_tx(bar, (uint32_t)A_STAGE + xtx);
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
      const uint32_t buf = (p.acc_bufs == 2) ? (it & 1) : 0u;
      const uint32_t accph = (p.acc_bufs == 2) ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(smem_u32(&acc_empty[buf]), accph ^ 1);
      const uint32_t d0 = tmem_base + buf * 2 * p.acc_stride;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(smem_u32(&full[s]), ph);
        if (p.xmode == 1) mbar_wait(smem_u32(&x_full[s]), ph);
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
              if (!(p.dbg & 4)) umma_tf32_ss(d_addr, adesc + adv, xdesc + adv, p.idesc_hi, (kb | k) ? 1u : 0u);
              if (!(p.dbg & 5)) umma_tf32_ts(d_addr, lo_slot + t * BK + k * 8, xdesc + adv, p.idesc_lo, 1u);
            }
          }
          umma_commit(smem_u32(&empty[s]));
          umma_commit(smem_u32(&lo_empty[sl]));
          if (kb == nkb - 1) umma_commit(smem_u32(&acc_full[buf]));
        }
        __syncwarp();
        if (++s == p.SA) { s = 0; ph ^= 1; }
        if (++sl == NSLOT) { sl = 0; phl ^= 1; }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== X producers (xmode 1): [X_hi ; X_lo] tile, K-major, hardware swizzle =====================
    // Item e = (k-quad j, column c), e = j * C + c: consecutive threads read consecutive addresses of the BK x C block of
    // X (one contiguous run of BK * C floats, L2 resident: 20 row tiles share it) and write one 16-byte unit of row c
    // (X_hi) and of row CP + c (X_lo).  Values are prefetched one k-block ahead in registers.
    if (p.estage && warp == 3) {
      // ---- store warp (staged epilogue): Y tile, shared -> global as one bulk copy ----
      if (elect_one()) {
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
          const uint32_t eb = it & 1;
          const int64_t b = tile / p.MT;
          const int64_t m0 = (tile - b * p.MT) * D2_ROWS;
          const uint32_t ybytes = (uint32_t)(min((int64_t)D2_ROWS, p.M - m0) * p.C * 4);
          mbar_wait(smem_u32(&y_ready[eb]), (it >> 1) & 1);
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(p.Y + (b * p.M + m0) * p.C),
                       "r"(smem_u32(reinterpret_cast<unsigned char*>(sE) + eb * p.estage)), "r"(ybytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(smem_u32(&e_empty[eb]));
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
    } else if (p.xmode == 1) {
      constexpr int NQ = BK / 4;
      constexpr int NI = (48 * NQ + 63) / 64;
      const int tid2 = threadIdx.x - 64;
      const int C = (int)p.C;
      const int nitems = C * NQ;
      uint32_t off[NI];
      int jq[NI], cc[NI];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int e = tid2 + 64 * i;
        const int j = e / C, c = e - j * C;
        jq[i] = (e < nitems) ? j : -1;
        cc[i] = c;
        const uint32_t sw = (BK == 32) ? (uint32_t)(c & 7) : (uint32_t)((c >> 1) & 3);
        off[i] = (uint32_t)c * ROW_BYTES + ((((uint32_t)j) ^ sw) << 4);  // CP % 8 == 0: rows c and CP + c swizzle alike
      }
      const int64_t my_tiles = (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
      const int64_t total = my_tiles * nkb;
      float xa[NI][4], xb[NI][4];
      auto load_x = [&](float (&reg)[NI][4], int64_t lin) {
        if (lin >= total) return;
        const int64_t ti = lin / nkb;
        const int kb = (int)(lin - ti * nkb);
        const int64_t tile = blockIdx.x + ti * gridDim.x;
        const int64_t b = tile / p.MT;
        const float* Xb = p.X + b * p.K * C;
        const int64_t k0 = (int64_t)kb * BK;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            const int64_t k = k0 + 4 * jq[i] + qd;
            reg[i][qd] = (jq[i] >= 0 && k < p.K) ? __ldg(Xb + k * C + cc[i]) : 0.f;
          }
        }
      };
      int s = 0;
      uint32_t ph = 0;
      auto store_x = [&](const float (&reg)[NI][4]) {
        mbar_wait(smem_u32(&empty[s]), ph ^ 1);
        unsigned char* xt = sRing + s * stage_bytes + A_STAGE;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          if (jq[i] >= 0) {
            uint4 h, l;
            uint32_t* hp = reinterpret_cast<uint32_t*>(&h);
            uint32_t* lp = reinterpret_cast<uint32_t*>(&l);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const uint32_t hb = (__float_as_uint(reg[i][qd]) + 0x1000u) & 0xFFFFE000u;
              hp[qd] = hb;
              lp[qd] = (__float_as_uint(reg[i][qd] - __uint_as_float(hb)) + 0x1000u) & 0xFFFFE000u;
            }
            *reinterpret_cast<uint4*>(xt + off[i]) = h;
            *reinterpret_cast<uint4*>(xt + off[i] + p.CP * ROW_BYTES) = l;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&x_full[s]));
        if (++s == p.SA) { s = 0; ph ^= 1; }
      };
      load_x(xa, 0);
      for (int64_t lin = 0; lin < total; lin += 2) {
        load_x(xb, lin + 1);
        store_x(xa);
        if (lin + 1 < total) {
          load_x(xa, lin + 2);
          store_x(xb);
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== epilogue: thread = operator row =====================
    const int q = warp & 3;
    const int C = (int)p.C, CP = p.CP;
    const bool need_e = (p.dg != nullptr) || (p.dots != nullptr);
    const int et = threadIdx.x - 128;  // 0..127 inside the epilogue group
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint32_t buf = (p.acc_bufs == 2) ? (it & 1) : 0u;
      const uint32_t accph = (p.acc_bufs == 2) ? ((it >> 1) & 1) : (it & 1);
      const int64_t b = tile / p.MT;
      const int mt = (int)(tile - b * p.MT);
      const int64_t m0 = (int64_t)mt * D2_ROWS;
      const float alpha_b = p.alpha ? p.alpha[b * p.alpha_bs] : 1.0f;
      const float* Eb = p.E + b * p.M * C;
      float* Yb = p.Y + b * p.M * C;
      double* red = dred + (it & 1) * (2 * 4 * 48);
      const bool staged = p.estage != 0;
      float* sEt = sE + (it & 1) * (p.estage >> 2);  // staged: E rows of this tile, overwritten in place with Y
      if (staged) mbar_wait(smem_u32(&e_full[it & 1]), (it >> 1) & 1);
      mbar_wait(smem_u32(&acc_full[buf]), accph);
      __syncwarp();
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int64_t row = m0 + t * 128 + q * 32 + lane;
        const bool rok = row < p.M;
        const float dv = (p.dg && rok) ? __ldg(p.dg + b * p.d_bs + row * p.d_st) : 0.f;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * 2 + t) * p.acc_stride;
        float* srow = sEt + (t * 128 + q * 32 + lane) * C;  // odd C: consecutive rows fall into consecutive banks
#pragma unroll 1
        for (int c0 = 0; c0 < CP; c0 += 16) {
          float e[16];
          if (staged) {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = (rok && c0 + i < C) ? srow[c0 + i] : 0.f;
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = (need_e && rok && c0 + i < C) ? __ldg(Eb + row * C + c0 + i) : 0.f;
          }
          uint32_t hi[16], lo[16];
          DS_LD16(taddr + c0, hi);
          DS_LD16(taddr + CP + c0, lo);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          double pd[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float y = fmaf(dv, e[i], (__uint_as_float(hi[i]) + __uint_as_float(lo[i])) * alpha_b);
            const bool ok = rok && c0 + i < C;
            if (ok) {
              if (staged) srow[c0 + i] = y;
              else Yb[row * C + c0 + i] = y;
            }
            pd[i] = ok ? (double)e[i] * (double)y : 0.0;
          }
          if (p.dots) {
            // Column sums of e * y over the 32 rows of this warp as a transpose-reduce: at each step a lane keeps half of
            // its columns, ships the other half to its partner and adds what it receives -- 16 fp64 adds and 16 64-bit
            // shuffles per chunk instead of 80 and 80 for sixteen independent butterflies (the fp64 pipe is what bounds
            // the epilogue, ncu).  The summation tree is fixed: deterministic.
#pragma unroll
            for (int h = 8, o = 16; h >= 1; h >>= 1, o >>= 1) {
              const bool up = (lane & o) != 0;
#pragma unroll
              for (int j = 0; j < h; ++j) {
                const double send = up ? pd[j] : pd[j + h];
                const double keep = up ? pd[j + h] : pd[j];
                pd[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
              }
            }
            pd[0] += __shfl_xor_sync(0xffffffffu, pd[0], 1);
            // lane l now holds column 8 b4 + 4 b3 + 2 b2 + b1 of the chunk (b_k = bit k of l); lanes l and l^1 agree
            const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            if ((lane & 1) == 0) red[(t * 4 + q) * 48 + c0 + col] = pd[0];
          }
        }
      }
      // accumulators drained: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      if (staged) fence_async_smem();  // the Y tile in shared memory is read by the bulk store (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&acc_empty[buf]));
        if (staged) mbar_arrive(smem_u32(&y_ready[it & 1]));
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
    // ===================== converters: A_lo rows -> TMEM operand slot =====================
    const int q = warp & 3;
    const int t = (warp - 8) >> 2;
    const int row = t * 128 + q * 32 + lane;
    const uint32_t swz = (BK == 32) ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + p.slot_base + t * BK;
    int s = 0, sl = 0;
    uint32_t ph = 0, phl = 0;
    for (int64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
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
        if (lane == 0) mbar_arrive(smem_u32(&lo_full[sl]));
        if (++s == p.SA) { s = 0; ph ^= 1; }
        if (++sl == NSLOT) { sl = 0; phl ^= 1; }
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
// Xs (B, 2 CP, Kp): rows [0, CP) = tf32(x[:, c]), rows [CP, 2 CP) = tf32(x - hi); k >= K and c >= C are zeros.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SPLIT2_KT = 64;

__global__ void __launch_bounds__(256)
k_split_x2(const float* __restrict__ X, float* __restrict__ Xs, int64_t K, int64_t Kp, int C, int CP) {
  extern __shared__ float xs2[];  // [SPLIT2_KT][C | 1]
  const int ldx = C | 1;
  const int64_t b = blockIdx.y;
  const int64_t k0 = (int64_t)blockIdx.x * SPLIT2_KT;
  const int kvalid = (int)max((int64_t)0, min((int64_t)SPLIT2_KT, K - k0));
  const float* src = X + (b * K + k0) * C;
  for (int e = threadIdx.x; e < kvalid * C; e += blockDim.x) xs2[(e / C) * ldx + (e % C)] = src[e];
  __syncthreads();
  // each thread writes one 16-byte word (4 consecutive k of one operand row): 16 lanes cover the 64 k of a row with
  // one contiguous 256-byte run (Kp % 4 == 0 and k0 % 64 == 0 keep the stores aligned)
  const int q4 = threadIdx.x & 15;
  const int kw = (int)min((int64_t)SPLIT2_KT, Kp - k0);
  const int R = 2 * CP;
  for (int r = threadIdx.x >> 4; r < R; r += blockDim.x >> 4) {
    const int part = r >= CP;
    const int c = r - part * CP;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = 4 * q4 + i;
        if (kk < kvalid) {
          const float x = xs2[kk * ldx + c];
          const float hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
          o[i] = part == 0 ? hi : __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
        }
      }
    }
    if (4 * q4 < kw) *reinterpret_cast<float4*>(Xs + (b * R + r) * Kp + k0 + 4 * q4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled_d2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled_d2 d2_encode_fn() {
  static PFN_encodeTiled_d2 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled_d2)ptr;
  }
  return fn;
}

struct D2Config {
  int acc_bufs;  // 0/2: double-buffered accumulators; 1: single (more A_lo operand slots in TMEM)
  int bk;    // 16 (SWIZZLE_64B) or 32 (SWIZZLE_128B)
  int sa;    // ring depth (0 = deepest that fits)
  int grid;  // 0 = one CTA per SM
  int dbg;
  int xmode; // 0: pre-split workspace + TMA (default); 1: X operand produced in the kernel by warps 2-3: needs no
             // workspace but is ~1.8x slower (two warps cannot hide the L2 latency of the strided X reads) -- used when
             // the caller passes no workspace
};

static D2Config d2_default_config() {
  static D2Config cfg = [] {
    D2Config c{0, 32, 0, 0, 0, 0};
#ifdef LOB_DIAG  // tuning / bottleneck-experiment knobs exist only in the harness build
    if (const char* e = getenv("LOB_D2_ACC_BUFS")) c.acc_bufs = atoi(e);
    if (const char* e = getenv("LOB_D2_XMODE")) c.xmode = atoi(e);
    if (const char* e = getenv("LOB_D2_DBG")) c.dbg = atoi(e);
    if (const char* e = getenv("LOB_D2_BK")) c.bk = atoi(e);
    if (const char* e = getenv("LOB_D2_SA")) c.sa = atoi(e);
    if (const char* e = getenv("LOB_D2_GRID")) c.grid = atoi(e);
#endif
    return c;
  }();
  return cfg;
}

constexpr size_t D2_SMEM_MAX = 232448;
constexpr size_t D2_SMEM_FIXED = 1024 /*alignment*/ + D2_RED_DOUBLES * 8 + 512 /*barriers*/;
constexpr int64_t D2_ESTAGE_MAX_K = 256;  // staged epilogue: contractions this short are epilogue-bound

size_t dense_stream2_workspace_bytes(int64_t B, int64_t K, int64_t C) {
  if (B <= 0 || K <= 0 || C <= 0 || C > 48) return 0;
  const int64_t CP = (C + 7) / 8 * 8;
  const int64_t Kp = (K + 3) / 4 * 4;
  return (size_t)B * 2 * CP * Kp * sizeof(float);
}

// returns LOB_ERR_UNSUPPORTED when the shape does not qualify (caller falls back to another CUDA kernel)
int dense_matmul_stream2_f32_cfg(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                                 const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                                 const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                                 cudaStream_t st, D2Config cfg) {
  if (C > 48 || C < 1) return LOB_ERR_UNSUPPORTED;
  if ((lda % 4) != 0 || (a_bs % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0) return LOB_ERR_UNSUPPORTED;
  if (M >= (1LL << 31) || K >= (1LL << 31) || B >= (1LL << 31)) return LOB_ERR_UNSUPPORTED;
  const bool have_ws = ws && (reinterpret_cast<uintptr_t>(ws) & 15) == 0 &&
                       ws_bytes >= dense_stream2_workspace_bytes(B, K, C);
  const int xmode = (cfg.xmode || !have_ws) ? 1 : 0;
  if ((d || dots) && !E && M != K) return LOB_ERR_UNSUPPORTED;
  PFN_encodeTiled_d2 enc = d2_encode_fn();
  if (!enc) return LOB_ERR_UNSUPPORTED;
  const int BK = (cfg.bk == 32) ? 32 : 16;
  const int CP = (int)((C + 7) / 8 * 8);
  const int64_t Kp = (K + 3) / 4 * 4;
  const bool shared = (a_bs == 0);
  const CUtensorMapSwizzle swz = (BK == 32) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;

  CUtensorMap tmA, tmX;
  {
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)(shared ? 1 : B)};
    cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)(shared ? (cuuint64_t)M * lda : a_bs) * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, D2_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     (cfg.dbg & 32) ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                    : ((cfg.dbg & 64) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_256B),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  }
  if (xmode == 1) {
    tmX = tmA;  // unused by the kernel
  } else {
    cuuint64_t gdim[3] = {(cuuint64_t)Kp, (cuuint64_t)(2 * CP), (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)Kp * 4, (cuuint64_t)(2 * CP) * Kp * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(2 * CP), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ws, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  }

  if (xmode == 0 && !(cfg.dbg & 128)) {
    dim3 grid((unsigned)cdiv(Kp, SPLIT2_KT), (unsigned)B);
    const size_t sm = (size_t)SPLIT2_KT * ((int)C | 1) * sizeof(float);
    k_split_x2<<<grid, 256, sm, st>>>(X, (float*)ws, K, Kp, (int)C, CP);
    LOB_TRY(check_launch("k_split_x2"));
  }

  const int a_stage = D2_ROWS * BK * 4;
  const int xbytes = (int)align_up((size_t)2 * CP * BK * 4, 1024);
  const int stage = a_stage + xbytes;
  // staged epilogue: E given, rows of 132-byte-odd pitch (bank-conflict-free per-row access), tiles 16-byte addressable
  int estage = 0;
  if (xmode == 0 && K <= D2_ESTAGE_MAX_K && E && (d || dots) && (C & 1) && ((M * C) % 4) == 0 &&
      (reinterpret_cast<uintptr_t>(E) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0 && !(cfg.dbg & 8))
    estage = (int)align_up((size_t)D2_ROWS * C * 4, 16);
  int sa = cfg.sa > 0 ? cfg.sa : (int)((D2_SMEM_MAX - D2_SMEM_FIXED - 2 * estage) / stage);
  if (sa > D2_MAX_ST) sa = D2_MAX_ST;
  if (sa < 2) return LOB_ERR_UNSUPPORTED;
  const size_t smem = D2_SMEM_FIXED + (size_t)sa * stage + 2 * (size_t)estage;
  if (smem > D2_SMEM_MAX) return LOB_ERR_UNSUPPORTED;

  D2Params p;
  p.estage = estage;
  p.Y = Y;
  p.X = X;
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
  p.CP = CP;
  p.xbytes = xbytes;
  p.SA = sa;
  p.MT = (int)cdiv(M, D2_ROWS);
  p.ntiles = B * p.MT;
  p.a_shared = shared ? 1 : 0;
  p.idesc_hi = ds::make_idesc_tf32(128, 2 * CP);
  // UMMA M = 128 needs N % 16 == 0: when CP is an odd multiple of 8 the narrow MMA also covers the first 8 X_lo rows,
  // which only adds the (otherwise dropped, O(2^-22)) a_lo * x_lo terms of those columns
  p.idesc_lo = ds::make_idesc_tf32(128, (CP + 15) / 16 * 16);
  p.dbg = cfg.dbg;
  p.xmode = xmode;
  p.acc_bufs = (cfg.acc_bufs == 1) ? 1 : 2;
  p.acc_stride = 2 * CP;
  p.slot_base = 2 * p.acc_bufs * p.acc_stride;
  p.nslot = (512 - p.slot_base) / (2 * BK);
  if (p.nslot > 8) p.nslot = 8;
  const int64_t grid = std::min<int64_t>(p.ntiles, cfg.grid > 0 ? cfg.grid : kNumSMs);
  if (BK == 32) {
    auto kern = k_dense_stream2<32>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D2_SMEM_MAX));
    kern<<<(unsigned)grid, D2_THREADS, smem, st>>>(tmA, tmX, p);
  } else {
    auto kern = k_dense_stream2<16>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D2_SMEM_MAX));
    kern<<<(unsigned)grid, D2_THREADS, smem, st>>>(tmA, tmX, p);
  }
  return check_launch("k_dense_stream2");
}

int dense_matmul_stream2_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                             const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs,
                             const float* d, int64_t d_bs, int64_t d_st, double* dots, void* ws, size_t ws_bytes,
                             cudaStream_t st) {
  return dense_matmul_stream2_f32_cfg(B, M, K, C, A, lda, a_bs, X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, ws,
                                      ws_bytes, st, d2_default_config());
}

}  // namespace lob

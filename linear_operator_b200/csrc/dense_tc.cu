// dense_tc.cu -- fp32 dense operator matmul  Y = A X (+ d (.) X)  on the 5th-generation tensor cores (tcgen05),
// fp32-accurate through an error-compensated 3xTF32 split, with the <X, Y> column partials of CG fused in the epilogue.
// Reference arithmetic: operators/dense_linear_operator.py:60-64 (torch.matmul -> cuBLAS SGEMM),
// operators/added_diag_linear_operator.py:72-76 (addcmul), utils/linear_cg.py:250-251 (sum(p * Ap)).
//
// Why this shape: A is (B, N, N) fp32 (102 GB at BASELINE config 2) and X is skinny (N x 33), so the kernel is bound by
// streaming A from HBM once; arithmetic intensity 16.5 flop/byte needs ~110 TFLOP/s of fp32-accurate math to keep up
// with 6.5 TB/s, which CUDA cores cannot deliver (the SIMT kernel in matmul_simt.cu reaches ~20 % of the roofline).
//
// One CTA computes 256 rows of Y for one batch element (two M=128 UMMA tiles sharing one X operand), K-loop over
// 32-column blocks of A.  Warp roles (12 warps):
//   warp 0      : TMA producer -- one 256 x 32 fp32 box of A per k-block (32 KB, 128B-swizzled) into a 5-stage ring
//   warp 1      : MMA issuer   -- one elected lane issues tcgen05.mma.kind::tf32 (A from TMEM, B from smem)
//   warps 2-3   : X producers  -- load the 32 x C block of X, split it into tf32 hi / lo, write it transposed
//                                 (K-major, 128B swizzle) into a 3-stage ring as the stacked operand [X_hi ; X_lo]
//   warps 4-11  : converters   -- read their row of the A tile from smem, split into hi = tf32(a), lo = a - hi, and
//                                 tcgen05.st both halves into a double-buffered TMEM operand slot; afterwards the
//                                 same warps run the epilogue (tcgen05.ld, + d x, <x,y> partials, coalesced store)
// Per k-block and M tile:  D[:, 0:2Cp] += A_hi [X_hi ; X_lo]^T   (one N = 2*Cp MMA per 8-wide k step)
//                          D[:, 0:Cp]  += A_lo  X_hi^T           (one N = Cp MMA)
// and y = D[:, c] + D[:, Cp + c]: the three products of the 3xTF32 scheme with fp32 accumulation in TMEM.
// The dropped a_lo * x_lo term is O(2^-22) relative.
#include <cuda.h>

#include "common.cuh"

namespace lob {

constexpr int TC_ROWS = 256;       // rows of Y per CTA (two UMMA M=128 tiles)
constexpr int TC_BK = 32;          // k-block: 32 fp32 = one 128-byte swizzle span
constexpr int TC_NSA = 5;          // A stages (32 KB each)
constexpr int TC_NSB = 3;          // X-operand stages
constexpr int TC_A_STAGE = TC_ROWS * TC_BK * 4;  // 32768
constexpr int TC_THREADS = 384;
constexpr long long TC_SPIN_CYCLES = 4000000000LL;  // ~2 s: a protocol bug traps instead of hanging

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > TC_SPIN_CYCLES) __trap();
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// tf32 with round-to-nearest (ties away from zero) as two integer ops: add half an ulp of the 10-bit mantissa to the
// sign-magnitude bit pattern, clear the 13 low bits (same result as cvt.rna.tf32.f32, which issues at conversion rate).
// hi = rna(a), lo = rna(a - hi): an unbiased split; a - hi is exact in fp32.
__device__ __forceinline__ uint32_t tf32_rna(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address
  desc |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  desc |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset: 8 rows * 128 B
  desc |= (uint64_t)1 << 46;                      // descriptor version (Blackwell)
  desc |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return desc;
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4)                      // D format: f32
         | (2u << 7)                    // A format: tf32
         | (2u << 10)                   // B format: tf32
         | ((uint32_t)(N >> 3) << 17)   // N / 8
         | ((uint32_t)(M >> 4) << 24);  // M / 16
}

#define TC_ST32(taddr, r)                                                                                            \
  asm volatile(                                                                                                      \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "  \
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                   \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),          \
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),      \
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),    \
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                         \
      : "memory")

#define TC_LD16(taddr, r)                                                                                             \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, " \
      "[%16];"                                                                                                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),   \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                      \
      : "r"(taddr)                                                                                                    \
      : "memory")

struct TcParams {
  const float* X;
  float* Y;
  const float* E;      // (B, M, C) matrix multiplied by the diagonal term and dotted with Y (NULL: X, needs M == K)
  const float* alpha;  // per-batch scale of the product (NULL: 1)
  int64_t alpha_bs;
  const float* dg;
  int64_t d_bs, d_st;
  double* dots;
  int64_t M, K, C;
  int n_parts;
};

// CP = padded column count of X (16, 32 or 48); the accumulator of one M tile is 2*CP TMEM columns.
template <int CP>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_dense_tc(const __grid_constant__ CUtensorMap tmA, TcParams p) {
  constexpr int B_STAGE = 2 * CP * 128;  // stacked [X_hi ; X_lo], CP rows each, 128 B per row
  constexpr int D_COLS = 2 * CP;
  constexpr int D_STRIDE = 128;          // TMEM column offset between the two accumulators
  constexpr int SLOT0 = 256;             // first A-operand slot
  constexpr int SLOT_COLS = 128;         // tile0 hi | tile0 lo | tile1 hi | tile1 lo, 32 columns each
  constexpr uint32_t IDESC_WIDE = make_idesc_tf32(128, 2 * CP);
  constexpr uint32_t IDESC_NARROW = make_idesc_tf32(128, CP);

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // pad to a 1024-byte boundary with pointer arithmetic on the __shared__ array (keeps LDS/STS code generation)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                  // TC_NSA * 32 KB
  unsigned char* sB = smem + TC_NSA * TC_A_STAGE;            // TC_NSB * B_STAGE
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + TC_NSB * B_STAGE);
  uint64_t* full_a = bars;                   // [NSA]  TMA -> converters
  uint64_t* empty_a = full_a + TC_NSA;       // [NSA]  converters -> TMA
  uint64_t* full_b = empty_a + TC_NSA;       // [NSB]  X producers -> MMA
  uint64_t* empty_b = full_b + TC_NSB;       // [NSB]  MMA -> X producers
  uint64_t* tm_full = empty_b + TC_NSB;      // [2]    converters -> MMA
  uint64_t* tm_empty = tm_full + 2;          // [2]    MMA -> converters
  uint64_t* acc_full = tm_empty + 2;         // [1]    MMA -> epilogue
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = blockIdx.y;
  const int64_t m0 = (int64_t)blockIdx.x * TC_ROWS;
  const int nkb = (int)((p.K + TC_BK - 1) / TC_BK);

  // ---- one-time setup ----
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    for (int i = 0; i < TC_NSA; ++i) {
      mbar_init(smem_u32(&full_a[i]), 1);
      mbar_init(smem_u32(&empty_a[i]), 8);
    }
    for (int i = 0; i < TC_NSB; ++i) {
      mbar_init(smem_u32(&full_b[i]), 2);
      mbar_init(smem_u32(&empty_b[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tm_full[i]), 8);
      mbar_init(smem_u32(&tm_empty[i]), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_holder))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the X-operand ring once: padding rows (c >= C) are never written again
  for (int i = threadIdx.x; i < TC_NSB * B_STAGE / 16; i += TC_THREADS)
    reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TC_NSA;
        const uint32_t ph = (kb / TC_NSA) & 1;
        mbar_wait(smem_u32(&empty_a[s]), ph ^ 1);
        mbar_arrive_expect_tx(smem_u32(&full_a[s]), TC_A_STAGE);
        tma_load_3d(smem_u32(sA + s * TC_A_STAGE), &tmA, smem_u32(&full_a[s]), kb * TC_BK, (int)m0, (int)b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Unrolled by 6 = lcm(2 TMEM slots, 3 X stages): every TMEM address and smem descriptor below is then a
    // loop-invariant base plus a compile-time offset, which keeps the per-MMA issue sequence short.
    static_assert(TC_NSB == 3, "the MMA issue loop is unrolled for 3 X stages");
    const uint64_t bdesc0 = make_kmajor_sw128_desc(smem_u32(sB));
    for (int kb0 = 0; kb0 < nkb; kb0 += 6) {
#pragma unroll
      for (int u = 0; u < 6; ++u) {
        const int kb = kb0 + u;
        if (kb >= nkb) break;
        constexpr int dummy = 0;
        (void)dummy;
        const int t = u & 1;
        const int sb = u % 3;
        const uint32_t pht = (kb >> 1) & 1;
        const uint32_t phb = (kb / TC_NSB) & 1;
        mbar_wait(smem_u32(&tm_full[t]), pht);
        mbar_wait(smem_u32(&full_b[sb]), phb);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t bdesc = bdesc0 + (uint64_t)((sb * B_STAGE) >> 4);
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const uint32_t d_addr = tmem_base + tile * D_STRIDE;
            const uint32_t a_hi = tmem_base + SLOT0 + t * SLOT_COLS + tile * 64;
            const uint32_t a_lo = a_hi + 32;
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint64_t bd = bdesc + (uint64_t)((k * 32) >> 4);  // 8 tf32 = 32 bytes inside the swizzle span
              umma_tf32_ts(d_addr, a_hi + k * 8, bd, IDESC_WIDE, (kb | k) ? 1u : 0u);
              umma_tf32_ts(d_addr, a_lo + k * 8, bd, IDESC_NARROW, 1u);
            }
          }
          umma_commit(smem_u32(&tm_empty[t]));
          umma_commit(smem_u32(&empty_b[sb]));
          if (kb == nkb - 1) umma_commit(smem_u32(acc_full));
        }
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    // ===================== X producers: split + transpose into [X_hi ; X_lo] =====================
    // The 32 x C block of X is one contiguous run of 32*C floats.  Thread i of the 64 owns elements i, i+64, ...;
    // their (k, c) coordinates -- hence the swizzled destination offsets -- do not depend on the k-block, so they are
    // computed once.  Loads of block kb+1 are issued before block kb is written (register prefetch) so the global /
    // L2 latency hides behind one k-block of work.
    constexpr int NE = (TC_BK * CP + 63) / 64;
    const int tid2 = threadIdx.x - 64;  // 0..63
    const float* Xb = p.X + b * p.K * p.C;
    const int C = (int)p.C;
    const int total = TC_BK * C;
    uint32_t off[NE];
    int kk[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
      const int e = tid2 + i * 64;
      const int k = e / C, c = e - k * C;
      kk[i] = (e < total) ? k : TC_BK;  // TC_BK = never valid
      off[i] = (uint32_t)c * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)c & 7u)) << 4) + (((uint32_t)k & 3u) << 2);
    }
    float r0[NE], r1[NE];
    auto load_block = [&](float (&reg)[NE], int kb) {
      if (kb < nkb) {
        const int64_t k1 = (int64_t)kb * TC_BK;
        const int kvalid = (int)min((int64_t)TC_BK, p.K - k1);
        const float* src = Xb + k1 * C + tid2;
#pragma unroll
        for (int i = 0; i < NE; ++i) reg[i] = (kk[i] < kvalid) ? __ldg(src + i * 64) : 0.f;
      }
    };
    auto store_block = [&](const float (&reg)[NE], int kb) {
      const int sb = kb % TC_NSB;
      const uint32_t phb = (kb / TC_NSB) & 1;
      mbar_wait(smem_u32(&empty_b[sb]), phb ^ 1);
      unsigned char* dst = sB + sb * B_STAGE;
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        if (kk[i] < TC_BK) {
          const uint32_t hi_bits = tf32_rna(reg[i]);
          const float lo = reg[i] - __uint_as_float(hi_bits);
          *reinterpret_cast<uint32_t*>(dst + off[i]) = hi_bits;
          // rows CP.. hold X_lo; CP is a multiple of 8 so the swizzle phase (row & 7) is unchanged
          *reinterpret_cast<uint32_t*>(dst + off[i] + CP * 128) = tf32_rna(lo);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&full_b[sb]));
    };
    load_block(r0, 0);
    load_block(r1, 1);
    for (int kb = 0; kb < nkb; kb += 2) {
      store_block(r0, kb);
      load_block(r0, kb + 2);
      if (kb + 1 < nkb) {
        store_block(r1, kb + 1);
        load_block(r1, kb + 3);
      }
    }
  } else {
    // ===================== converters (then epilogue) =====================
    const int cw = warp - 4;            // 0..7
    const int tile = cw >> 2;           // M tile of this warp
    const int quarter = warp & 3;       // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;  // row inside the M tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_NSA;
      const uint32_t ph = (kb / TC_NSA) & 1;
      const int t = kb & 1;
      const uint32_t pht = (kb >> 1) & 1;
      mbar_wait(smem_u32(&full_a[s]), ph);
      const unsigned char* arow = sA + s * TC_A_STAGE + tile * (128 * 128) + row * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 v = *reinterpret_cast<const uint4*>(arow + ((j ^ (row & 7)) << 4));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t h = tf32_rna(__uint_as_float(w[q]));
          hi[j * 4 + q] = h;
          lo[j * 4 + q] = tf32_rna(__uint_as_float(w[q]) - __uint_as_float(h));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty_a[s]));  // smem stage consumed (values are in registers)
      mbar_wait(smem_u32(&tm_empty[t]), pht ^ 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + SLOT0 + t * SLOT_COLS + tile * 64;
      TC_ST32(taddr, hi);
      TC_ST32(taddr + 32, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tm_full[t]));
    }

    // ---------------- epilogue ----------------
    mbar_wait(smem_u32(acc_full), 0);
    tc_fence_after();
    const int C = (int)p.C;
    // staging (aliases the A ring, which is fully consumed): per tile xs[128][C], ys[128][C]
    float* xs = reinterpret_cast<float*>(sA) + tile * (2 * 128 * 49);
    float* ys = xs + 128 * 49;
    const int64_t tile_m0 = m0 + tile * 128;
    const int rows_valid = (int)max((int64_t)0, min((int64_t)128, p.M - tile_m0));
    const int ct = threadIdx.x - 128 - tile * 128;  // 0..127 within the tile's 4 warps
    const float* Xb = p.X + b * p.K * p.C;
    float* Yb = p.Y + b * p.M * p.C;
    const bool need_x = (p.dg != nullptr) || (p.dots != nullptr);
    if (need_x) {
      const float* Eb = p.E ? p.E + b * p.M * p.C : Xb;
      for (int e = ct; e < rows_valid * C; e += 128) xs[(e / C) * 49 + (e % C)] = Eb[tile_m0 * C + e];
    }
    const float alpha_b = p.alpha ? p.alpha[b * p.alpha_bs] : 1.0f;
    // named barrier per tile (ids 1, 2), 128 threads
    asm volatile("bar.sync %0, 128;" ::"r"(1 + tile) : "memory");
    float dv = 0.f;
    if (p.dg && row < rows_valid) dv = p.dg[b * p.d_bs + (tile_m0 + row) * p.d_st];
    const uint32_t d_addr = tmem_base + lane_base + tile * D_STRIDE;
#pragma unroll
    for (int c0 = 0; c0 < CP; c0 += 16) {
      uint32_t a[16], bb[16];
      TC_LD16(d_addr + c0, a);
      TC_LD16(d_addr + CP + c0, bb);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = c0 + i;
        if (c < C && row < rows_valid) {
          float y = (__uint_as_float(a[i]) + __uint_as_float(bb[i])) * alpha_b;
          if (p.dg) y += dv * xs[row * 49 + c];
          ys[row * 49 + c] = y;
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + tile) : "memory");
    for (int e = ct; e < rows_valid * C; e += 128) Yb[tile_m0 * C + e] = ys[(e / C) * 49 + (e % C)];
    if (p.dots) {
      // <x, y> per column over the rows of this tile: 3 row groups x C columns, then a fixed-order combine
      double* redt = reinterpret_cast<double*>(reinterpret_cast<float*>(sA) + 2 * (2 * 128 * 49)) + tile * (3 * 48);
      for (int idx = ct; idx < 3 * C; idx += 128) {
        const int grp = idx / C, c = idx - grp * C;
        const int r0 = grp * 43, r1 = min(r0 + 43, rows_valid);
        double s = 0.0;
        for (int r = r0; r < r1; ++r) s += (double)xs[r * 49 + c] * (double)ys[r * 49 + c];
        redt[grp * 48 + c] = s;
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + tile) : "memory");
      if (ct < C) {
        const int64_t part = (int64_t)blockIdx.x * 2 + tile;
        if (part < p.n_parts) p.dots[(b * p.n_parts + part) * p.C + ct] = redt[ct] + redt[48 + ct] + redt[96 + ct];
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)ptr;
  }
  return fn;
}

template <int CP>
static int launch_tc(const CUtensorMap& tm, const TcParams& p, int64_t B, cudaStream_t st) {
  constexpr int B_STAGE = 2 * CP * 128;
  const size_t smem = 1024 + (size_t)TC_NSA * TC_A_STAGE + (size_t)TC_NSB * B_STAGE + 256;
  auto kern = k_dense_tc<CP>;
  static bool attr_set = false;
  if (!attr_set) {
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((unsigned)cdiv(p.M, TC_ROWS), (unsigned)B);
  kern<<<grid, TC_THREADS, smem, st>>>(tm, p);
  return check_launch("k_dense_tc");
}

// returns LOB_ERR_UNSUPPORTED when the shape does not qualify (caller falls back to the CUDA-core kernel)
int dense_matmul_tc_f32(int64_t B, int64_t M, int64_t K, int64_t C, const float* A, int64_t lda, int64_t a_bs,
                        const float* X, float* Y, const float* E, const float* alpha, int64_t alpha_bs, const float* d,
                        int64_t d_bs, int64_t d_st, double* dots, cudaStream_t st) {
  if (C > 48 || M < 32 || K < 32) return LOB_ERR_UNSUPPORTED;
  if ((lda % 4) != 0 || (a_bs % 4) != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0) return LOB_ERR_UNSUPPORTED;
  if (M >= (1LL << 31) || K >= (1LL << 31)) return LOB_ERR_UNSUPPORTED;
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return LOB_ERR_UNSUPPORTED;
  CUtensorMap tm;
  const bool shared = (a_bs == 0);
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)M, (cuuint64_t)(shared ? 1 : B)};
  cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)(shared ? (cuuint64_t)M * lda : a_bs) * 4};
  cuuint32_t box[3] = {TC_BK, TC_ROWS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(A), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return LOB_ERR_UNSUPPORTED;
  if (shared) return LOB_ERR_UNSUPPORTED;  // broadcast operator: batch coordinate would have to be pinned to 0
  TcParams p{X, Y, E, alpha, alpha_bs, d, d_bs, d_st, dots, M, K, C, (int)cdiv(M, 128)};
  if (C <= 16) return launch_tc<16>(tm, p, B, st);
  if (C <= 32) return launch_tc<32>(tm, p, B, st);
  return launch_tc<48>(tm, p, B, st);
}

}  // namespace lob

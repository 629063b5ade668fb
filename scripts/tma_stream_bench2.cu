// Micro-benchmark 2: what limits a persistent 148-CTA TMA stream of a (B, N, N) fp32 operator when the ring is
// embedded in a warp-specialised kernel?  Sweeps: box width (16 / 32 floats), ring depth, number of hand-off hops
// between "landed" and "slot free" (1 = consumer releases directly; 2, 3 = relayed through further warps, as the
// converter -> MMA -> commit chain of k_dense_stream does), extra warps spinning on mbarriers, and a second, L2-resident
// operand stream (the X operand tile).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench2 tma_stream_bench2.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void k_fill_rand(float* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = i * 0x9E3779B97F4A7C15ULL; x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
    p[i] = (float)(unsigned)(x >> 40) * (1.0f / 8388608.0f) - 1.0f;
  }
}

struct P {
  int cols, rows, nstage, hops, spin_warps, xrows, nkb, tiles_per_batch, ntiles;
};

// warp 0: producer; warps 1..hops: relay chain (warp h waits on bar[h-1][s], arrives on bar[h][s]; the last one's
// barrier is the "empty" barrier); further warps spin on a barrier that completes only at the end.
__global__ void __launch_bounds__(512, 1) k_stream2(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmx, P p, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int a_bytes = p.rows * p.cols * 4;
  const int x_bytes = (p.xrows * p.cols * 4 + 1023) / 1024 * 1024;
  const int stage_bytes = a_bytes + x_bytes;
  uint64_t* bars = (uint64_t*)(smem + p.nstage * stage_bytes);  // [hops + 1][nstage], bars[0] = full, bars[hops] = empty
  uint64_t* never = bars + 4 * 16;
  if (threadIdx.x == 0) {
    for (int h = 0; h <= p.hops; ++h)
      for (int i = 0; i < p.nstage; ++i) mbar_init(smem_u32(&bars[h * 16 + i]), 1);
    mbar_init(smem_u32(never), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / p.tiles_per_batch, m0 = (tile % p.tiles_per_batch) * p.rows;
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(smem_u32(&bars[p.hops * 16 + s]), ph ^ 1);
          const uint32_t bar = smem_u32(&bars[s]);
          mbar_expect(bar, a_bytes + (p.xrows ? p.xrows * p.cols * 4 : 0));
          if (p.xrows) tma3(smem_u32(smem + s * stage_bytes + a_bytes), &tmx, bar, kb * p.cols, 0, b);
          tma3(smem_u32(smem + s * stage_bytes), &tm, bar, kb * p.cols, m0, b);
          if (++s == p.nstage) { s = 0; ph ^= 1; }
        }
      }
      mbar_arrive(smem_u32(never));
    }
  } else if (warp <= p.hops) {
    float acc = 0.f;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(smem_u32(&bars[(warp - 1) * 16 + s]), ph);
        if (warp == 1) acc += ((float*)(smem + s * stage_bytes))[lane];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[warp * 16 + s]));
        if (++s == p.nstage) { s = 0; ph ^= 1; }
      }
    }
    if (acc == 123.456f) sink[0] = acc;
  } else if (warp <= p.hops + p.spin_warps) {
    mbar_wait(smem_u32(never), 0);
  }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int64_t B = 256, N = 5000;
  float *A, *X, *sink;
  cudaMalloc(&A, B * N * N * 4);
  cudaMalloc(&X, B * 128 * N * 4);
  cudaMalloc(&sink, 4);
  if (argc > 1) { k_fill_rand<<<1184, 256>>>(A, (size_t)B * N * N); printf("operator filled with random data\n"); }
  else cudaMemset(A, 0x3c, B * N * N * 4);
  cudaMemset(X, 0x3c, B * 128 * N * 4);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  cudaFuncSetAttribute(k_stream2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  struct Cfg { int cols, nstage, hops, spin, xrows; } cfgs[] = {
      {32, 5, 1, 0, 0},  {32, 4, 1, 0, 0},  {32, 3, 1, 0, 0},  {16, 8, 1, 0, 0},
      {32, 5, 1, 0, 80}, {32, 4, 1, 0, 80}, {32, 5, 3, 12, 80}};
  for (auto c : cfgs) {
    CUtensorMap tm, tmx;
    const int rows = 256;
    CUtensorMapSwizzle sw = c.cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
      cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
      cuuint64_t gstr[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
      cuuint32_t box[3] = {(cuuint32_t)c.cols, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    }
    {
      const int xr = c.xrows ? c.xrows : 8;
      cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)xr, (cuuint64_t)B};
      cuuint64_t gstr[2] = {(cuuint64_t)N * 4, (cuuint64_t)xr * N * 4};
      cuuint32_t box[3] = {(cuuint32_t)c.cols, (cuuint32_t)xr, 1}, es[3] = {1, 1, 1};
      CUresult r = enc(&tmx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, X, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode x failed %d\n", (int)r); continue; }
    }
    P p;
    p.cols = c.cols; p.rows = rows; p.nstage = c.nstage; p.hops = c.hops; p.spin_warps = c.spin; p.xrows = c.xrows;
    p.nkb = (int)((N + c.cols - 1) / c.cols);
    p.tiles_per_batch = (int)((N + rows - 1) / rows);
    p.ntiles = p.tiles_per_batch * (int)B;
    const int x_bytes = (c.xrows * c.cols * 4 + 1023) / 1024 * 1024;
    size_t smem = 1024 + (size_t)c.nstage * (rows * c.cols * 4 + x_bytes) + 1024;
    if (smem > 227 * 1024) { printf("cols=%d stages=%d xrows=%d: smem %zu too large\n", c.cols, c.nstage, c.xrows, smem); continue; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_stream2<<<148, 512, smem>>>(tm, tmx, p, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    double gb = (double)B * N * N * 4 / 1e9;
    printf("cols=%2d stages=%2d (%3d KB ring) hops=%d spin_warps=%2d xrows=%2d : %.3f ms  %.0f GB/s of A (%s)\n", c.cols, c.nstage, (int)(c.nstage * (rows * c.cols * 4 + x_bytes) / 1024), c.hops, c.spin, c.xrows, best, gb / (best * 1e-3), cudaGetErrorString(err));
  }
  return 0;
}

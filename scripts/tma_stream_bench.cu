// Micro-benchmark: how fast can 148 CTAs stream a (B, N, N) fp32 operator through TMA boxes of ROWS x 32 floats
// (128-byte row segments, 128B swizzle), as a function of rows per box, ring depth and how many adjacent k-blocks are
// issued back to back?  Used to separate the DRAM access-pattern ceiling from the matmul pipeline.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_bench tma_stream_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0; long long t0 = clock64();
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// persistent: CTA c handles tiles c, c+grid, ...; tile = (batch, row block); streams all k-blocks of the tile
__global__ void __launch_bounds__(64, 1) k_stream(const __grid_constant__ CUtensorMap tm, int rows, int nstage, int group, int nkb, int tiles_per_batch, int ntiles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = rows * 128;
  uint64_t* full = (uint64_t*)(smem + nstage * stage_bytes);
  uint64_t* empty = full + nstage;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long it = 0;
  if (warp == 0) {
    if (lane == 0) {
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / tiles_per_batch, m0 = (tile % tiles_per_batch) * rows;
        for (int kb = 0; kb < nkb; kb += group) {
          for (int j = 0; j < group && kb + j < nkb; ++j) { long long i2 = it + j; mbar_wait(smem_u32(&empty[i2 % nstage]), ((i2 / nstage) & 1) ^ 1); }
          for (int j = 0; j < group && kb + j < nkb; ++j) {
            long long i2 = it + j; int s = i2 % nstage;
            mbar_expect(smem_u32(&full[s]), stage_bytes);
            tma3(smem_u32(smem + s * stage_bytes), &tm, smem_u32(&full[s]), (kb + j) * 32, m0, b);
          }
          it += (kb + group <= nkb) ? group : (nkb - kb);
        }
      }
    }
  } else {
    float acc = 0.f;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        int s = it % nstage;
        mbar_wait(smem_u32(&full[s]), (it / nstage) & 1);
        acc += ((float*)(smem + s * stage_bytes))[lane];
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
      }
    }
    if (acc == 123.456f) sink[0] = acc;
  }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int64_t B = 256, N = 5000;
  float* A; float* sink;
  cudaMalloc(&A, B * N * N * 4); cudaMalloc(&sink, 4);
  cudaMemset(A, 0, B * N * N * 4);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  PFN enc = (PFN)fp;
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024);
  int promos[2] = {CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B};
  struct Cfg { int rows, nstage, group, promo, contig; } cfgs[] = {
      {256, 5, 1, 1, 0}, {256, 6, 2, 1, 0}, {256, 6, 3, 1, 0}, {256, 6, 2, 0, 0}, {256, 5, 1, 0, 0},
      {128, 12, 4, 1, 0}, {128, 12, 2, 1, 0}, {128, 12, 1, 1, 0}, {128, 12, 6, 1, 0}, {64, 24, 8, 1, 0},
      {256, 5, 1, 1, 1}, {256, 6, 2, 1, 1}};
  for (auto c : cfgs) {
    CUtensorMap tm;
    int64_t lda = c.contig ? 32 : N;            // contig: pretend rows are packed 128 B apart (pure streaming pattern)
    int64_t rows_total = c.contig ? N * (N / 32) : N;
    cuuint64_t gdim[3] = {(cuuint64_t)(c.contig ? 32 : N), (cuuint64_t)rows_total, (cuuint64_t)B};
    cuuint64_t gstr[2] = {(cuuint64_t)lda * 4, (cuuint64_t)N * N * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)c.rows, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, A, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, (CUtensorMapL2promotion)promos[c.promo], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    int nkb, tiles_per_batch;
    if (c.contig) { nkb = 1; tiles_per_batch = (int)(rows_total / c.rows); }
    else { nkb = (int)((N + 31) / 32); tiles_per_batch = (int)((N + c.rows - 1) / c.rows); }
    int ntiles = tiles_per_batch * (int)B;
    size_t smem = 1024 + (size_t)c.nstage * c.rows * 128 + 16 * c.nstage + 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (c.contig) k_stream<<<148, 64, smem>>>(tm, c.rows, c.nstage, c.group, 1, tiles_per_batch, ntiles, sink);
      else k_stream<<<148, 64, smem>>>(tm, c.rows, c.nstage, c.group, nkb, tiles_per_batch, ntiles, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    double gb = (double)B * N * N * 4 / 1e9;
    printf("rows=%3d stages=%2d group=%d promo=%s contig=%d : %.3f ms  %.0f GB/s  (%s)\n", c.rows, c.nstage, c.group, c.promo ? "256B" : "128B", c.contig, best, gb / (best * 1e-3), cudaGetErrorString(err));
  }
  return 0;
}

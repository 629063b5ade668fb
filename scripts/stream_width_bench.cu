// Does the width of the per-thread global load matter for a 4-read / 2-write streaming update (the shape of k_step_xr:
// r -= a*q, x += a*p)?  Same bytes, same thread count; 4-byte loads with 4 independent rows per trip (the shipped
// kernel's shape) against 16-byte loads.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/stream_width_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

__global__ void k_w4(const float* __restrict__ q, const float* __restrict__ p, float* __restrict__ r, float* __restrict__ x,
                     size_t n, float a) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    float q0 = q[i], q1 = q[i + stride], q2 = q[i + 2 * stride], q3 = q[i + 3 * stride];
    float r0 = r[i], r1 = r[i + stride], r2 = r[i + 2 * stride], r3 = r[i + 3 * stride];
    float p0 = p[i], p1 = p[i + stride], p2 = p[i + 2 * stride], p3 = p[i + 3 * stride];
    float x0 = x[i], x1 = x[i + stride], x2 = x[i + 2 * stride], x3 = x[i + 3 * stride];
    r[i] = r0 - a * q0; r[i + stride] = r1 - a * q1; r[i + 2 * stride] = r2 - a * q2; r[i + 3 * stride] = r3 - a * q3;
    x[i] = x0 + a * p0; x[i + stride] = x1 + a * p1; x[i + 2 * stride] = x2 + a * p2; x[i + 3 * stride] = x3 + a * p3;
  }
  for (; i < n; i += stride) { r[i] -= a * q[i]; x[i] += a * p[i]; }
}
template <int U>
__global__ void k_w16(const float4* __restrict__ q, const float4* __restrict__ p, float4* __restrict__ r,
                      float4* __restrict__ x, size_t n4, float a) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 qq[U], rr[U], pp[U], xx[U];
#pragma unroll
    for (int k = 0; k < U; ++k) { qq[k] = q[i + k * stride]; rr[k] = r[i + k * stride]; pp[k] = p[i + k * stride]; xx[k] = x[i + k * stride]; }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      rr[k].x -= a * qq[k].x; rr[k].y -= a * qq[k].y; rr[k].z -= a * qq[k].z; rr[k].w -= a * qq[k].w;
      xx[k].x += a * pp[k].x; xx[k].y += a * pp[k].y; xx[k].z += a * pp[k].z; xx[k].w += a * pp[k].w;
      r[i + k * stride] = rr[k]; x[i + k * stride] = xx[k];
    }
  }
  for (; i < n4; i += stride) {
    float4 qq = q[i], rr = r[i], pp = p[i], xx = x[i];
    rr.x -= a * qq.x; rr.y -= a * qq.y; rr.z -= a * qq.z; rr.w -= a * qq.w;
    xx.x += a * pp.x; xx.y += a * pp.y; xx.z += a * pp.z; xx.w += a * pp.w;
    r[i] = rr; x[i] = xx;
  }
}
int main() {
  const size_t n = (size_t)8 * 1000000 * 33;  // config 3's vector
  float *q, *p, *r, *x;
  CK(cudaMalloc(&q, n * 4)); CK(cudaMalloc(&p, n * 4)); CK(cudaMalloc(&r, n * 4)); CK(cudaMalloc(&x, n * 4));
  CK(cudaMemset(q, 0, n * 4)); CK(cudaMemset(p, 0, n * 4)); CK(cudaMemset(r, 0, n * 4)); CK(cudaMemset(x, 0, n * 4));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double bytes = 6.0 * n * 4;
  for (int grid : {148 * 4, 148 * 8, 148 * 16}) {
    for (int variant = 0; variant < 4; ++variant) {
      float best = 1e30f;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        if (variant == 0) k_w4<<<grid, 256>>>(q, p, r, x, n, 0.5f);
        if (variant == 1) k_w16<1><<<grid, 256>>>((const float4*)q, (const float4*)p, (float4*)r, (float4*)x, n / 4, 0.5f);
        if (variant == 2) k_w16<2><<<grid, 256>>>((const float4*)q, (const float4*)p, (float4*)r, (float4*)x, n / 4, 0.5f);
        if (variant == 3) k_w16<4><<<grid, 256>>>((const float4*)q, (const float4*)p, (float4*)r, (float4*)x, n / 4, 0.5f);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
      }
      const char* names[] = {"4-byte loads x4 rows", "16-byte loads x1", "16-byte loads x2", "16-byte loads x4"};
      printf("grid %5d  %-22s %.3f ms  %.0f GB/s\n", grid, names[variant], best, bytes / best * 1e-6);
    }
  }
  return 0;
}

// backward.cu -- kernels of the backward pass of the Krylov path.
//
//   k_bilinear_dense  : DenseLinearOperator._bilinear_derivative (dense_linear_operator.py:69-71)
//                       G[b,i,j] (+)= sum_c w[b,c] * L[b,i,c] * R[b,j,c]        -- the (B, N, N) operator gradient of
//                       InvQuadLogdet / InvQuad / Solve backward (functions/_inv_quad_logdet.py:193-206): a rank-C
//                       outer-product update that WRITES B*N*N elements, i.e. the forward matmul's roofline on the store
//                       side.  The per-column weights w fold the `mul(coef).mul(norms).mul(grad_output)` passes of
//                       :181-183,:200 into the tile load.
//   k_bilinear_diag   : DiagLinearOperator._bilinear_derivative (diag_linear_operator.py:37-45)
//                       g[b,n] = sum_c w[b,c] * L[b,n,c] * R[b,n,c]
//   k_tri_inverse     : inverse of a (batch of) lower-triangular k x k blocks, the `solve_triangular` adjoint inside
//                       PivotedCholesky.backward (functions/_pivoted_cholesky.py:128-137)
//   k_cross_spectrum  : sum over columns of 2 Re(conj(Fu) Fv): the spectrum of sym_toeplitz_derivative_quadratic_form
//                       (utils/toeplitz.py:164-204), which the reference evaluates with two Toeplitz products per column
#include "common.cuh"

namespace lob {

// CTA tile (16*TM) x (16*TN), 256 threads, thread tile TM x TN; operands staged c-major in shared memory, CK columns
// per stage (CK >= C for the path's 33-column blocks: one load phase, one barrier pair per tile).
template <typename T, int TM, int TN, int CK>
__global__ void __launch_bounds__(256)
k_bilinear_dense(int64_t N, int64_t M, int C, const T* __restrict__ Lf, const T* __restrict__ Rt,
                 const T* __restrict__ w, T* __restrict__ G, int accumulate) {
  constexpr int BM = 16 * TM, BN = 16 * TN;
  constexpr int LDL = BM + 4, LDR = BN + 4;
  extern __shared__ __align__(16) unsigned char bl_smem[];
  T* Ls = reinterpret_cast<T*>(bl_smem);  // [CK][LDL]
  T* Rs = Ls + CK * LDL;                  // [CK][LDR]
  const int64_t b = blockIdx.z;
  const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const T* Lb = Lf + (b * N) * C;
  const T* Rb = Rt + (b * M) * C;
  const T* wb = w ? w + b * C : nullptr;
  T acc[TM][TN];
#pragma unroll
  for (int a = 0; a < TM; ++a)
#pragma unroll
    for (int q = 0; q < TN; ++q) acc[a][q] = (T)0;

  for (int c0 = 0; c0 < C; c0 += CK) {
    const int cw = min(CK, C - c0);
    __syncthreads();
    // rows i0..i0+BM of L are BM runs of cw contiguous elements (one contiguous chunk when cw == C): read them in
    // element order, scatter c-major.  (i, c) come from the element index by a multiply-shift (exact for e < 2^20 / cw),
    // so the loads of the unrolled iterations are independent and stay in flight together.
    {
      const uint32_t inv = (1u << 20) / (uint32_t)cw + 1u;
#pragma unroll 4
      for (int e = tid; e < BM * cw; e += 256) {
        const int i = (int)(((uint32_t)e * inv) >> 20), c = e - i * cw;
        T v = (T)0;
        if (i0 + i < N) {
          v = Lb[(i0 + i) * C + c0 + c];
          if (wb) v *= wb[c0 + c];
        }
        Ls[c * LDL + i] = v;
      }
#pragma unroll 4
      for (int e = tid; e < BN * cw; e += 256) {
        const int i = (int)(((uint32_t)e * inv) >> 20), c = e - i * cw;
        Rs[c * LDR + i] = (j0 + i < M) ? Rb[(j0 + i) * C + c0 + c] : (T)0;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < cw; ++c) {
      T a[TM], q[TN];
#pragma unroll
      for (int u = 0; u < TM; ++u) a[u] = Ls[c * LDL + ty * TM + u];
#pragma unroll
      for (int u = 0; u < TN; ++u) q[u] = Rs[c * LDR + tx * TN + u];
#pragma unroll
      for (int u = 0; u < TM; ++u)
#pragma unroll
        for (int v = 0; v < TN; ++v) acc[u][v] = fma(a[u], q[v], acc[u][v]);
    }
  }
  T* Gb = G + b * N * M;
  const int64_t jb = j0 + tx * TN;
  // 16-byte stores when the row segment is aligned and complete (fp32: 2 x float4 per row, fp64: 2 x double2)
  constexpr int VEC = 16 / sizeof(T);
  const bool vec = (jb + TN <= M) && ((M % VEC) == 0) && ((reinterpret_cast<uintptr_t>(Gb) & 15) == 0) && (TN % VEC) == 0;
#pragma unroll
  for (int u = 0; u < TM; ++u) {
    const int64_t i = i0 + ty * TM + u;
    if (i >= N) continue;
    T* row = Gb + i * M + jb;
    if (vec) {
#pragma unroll
      for (int v = 0; v < TN; v += VEC) {
        if constexpr (sizeof(T) == 4) {
          float4 o = make_float4(acc[u][v], acc[u][v + 1], acc[u][v + 2], acc[u][v + 3]);
          if (accumulate) {
            const float4 g = *reinterpret_cast<const float4*>(row + v);
            o.x += g.x; o.y += g.y; o.z += g.z; o.w += g.w;
          }
          *reinterpret_cast<float4*>(row + v) = o;
        } else {
          double2 o = make_double2(acc[u][v], acc[u][v + 1]);
          if (accumulate) {
            const double2 g = *reinterpret_cast<const double2*>(row + v);
            o.x += g.x; o.y += g.y;
          }
          *reinterpret_cast<double2*>(row + v) = o;
        }
      }
    } else {
#pragma unroll
      for (int v = 0; v < TN; ++v) {
        if (jb + v < M) {
          T r = acc[u][v];
          if (accumulate) r += row[v];
          row[v] = r;
        }
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_bilinear_diag(int64_t BN, int64_t N, int C, const T* __restrict__ Lf, const T* __restrict__ Rt,
                const T* __restrict__ w, T* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= BN) return;
  const int64_t b = r / N;
  const T* l = Lf + r * C;
  const T* q = Rt + r * C;
  double s = 0.0;
  for (int c = 0; c < C; ++c) {
    const double t = (double)l[c] * (double)q[c];
    s += w ? t * (double)w[b * C + c] : t;
  }
  out[r] = (T)s;
}

// one CTA per batch element; thread j solves C x = e_j by forward substitution (double accumulation)
template <typename T>
__global__ void __launch_bounds__(256)
k_tri_inverse(int k, const T* __restrict__ Cm, int64_t ldc, int64_t c_bs, T* __restrict__ out) {
  extern __shared__ unsigned char smem_raw[];
  T* Cs = reinterpret_cast<T*>(smem_raw);  // k x k
  const int64_t b = blockIdx.x;
  const T* Cb = Cm + b * c_bs;
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) {
    const int i = e / k, j = e - i * k;
    Cs[e] = Cb[i * ldc + j];
  }
  __syncthreads();
  T* Ob = out + (int64_t)b * k * k;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    for (int i = 0; i < j; ++i) Ob[i * k + j] = (T)0;
    for (int i = j; i < k; ++i) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int t = j; t < i; ++t) s -= (double)Cs[i * k + t] * (double)Ob[t * k + j];
      Ob[i * k + j] = (T)(s / (double)Cs[i * k + i]);
    }
  }
}

// S[b,h] = sum_c w[b,c] * 2 Re(conj(Fu[b,c,h]) Fv[b,c,h])   (real, stored as complex with zero imaginary part)
template <typename T>
__global__ void __launch_bounds__(256)
k_cross_spectrum(int64_t B, int C, int64_t H, const T* __restrict__ fu, const T* __restrict__ fv,
                 const T* __restrict__ w, T* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int64_t b = idx / H, h = idx - b * H;
  double s = 0.0;
  for (int c = 0; c < C; ++c) {
    const int64_t o = ((b * C + c) * H + h) * 2;
    const double t = (double)fu[o] * (double)fv[o] + (double)fu[o + 1] * (double)fv[o + 1];
    s += w ? t * (double)w[b * C + c] : t;
  }
  out[idx * 2] = (T)(2.0 * s);
  out[idx * 2 + 1] = (T)0;
}

// res[b,i] = scale * y[b,i] for i < N (y has row length L), res[b,0] additionally halved (the i = 0 term of the
// derivative counts the main diagonal once, utils/toeplitz.py:201)
template <typename T>
__global__ void k_toeplitz_deriv_finish(int64_t B, int64_t N, int64_t L, const T* __restrict__ y, double scale,
                                        T* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  const int64_t b = idx / N, i = idx - b * N;
  const double v = (double)y[b * L + i] * scale;
  out[idx] = (T)(i == 0 ? 0.5 * v : v);
}

}  // namespace lob

using namespace lob;

extern "C" int lob_bilinear_dense(int32_t dtype, int64_t B, int64_t N, int64_t M, int64_t C, const void* left,
                                  const void* right, const void* w, void* out, int32_t accumulate, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && M > 0 && C > 0, "lob_bilinear_dense: sizes must be positive");
  LOB_REQUIRE(left && right && out, "lob_bilinear_dense: NULL pointer");
  LOB_REQUIRE(B <= 65535 && C < (1 << 30), "lob_bilinear_dense: batch > 65535 not supported");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LOB_F32) {
    dim3 grid((unsigned)cdiv(M, 128), (unsigned)cdiv(N, 128), (unsigned)B);
    constexpr int CK = 48;
    const size_t sm = (size_t)CK * (132 + 132) * sizeof(float);
    auto kern = k_bilinear_dense<float, 8, 8, CK>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<grid, 256, sm, st>>>(N, M, (int)C, (const float*)left, (const float*)right, (const float*)w, (float*)out,
                                accumulate);
  } else if (dtype == LOB_F64) {
    dim3 grid((unsigned)cdiv(M, 64), (unsigned)cdiv(N, 64), (unsigned)B);
    constexpr int CK = 48;
    const size_t sm = (size_t)CK * (68 + 68) * sizeof(double);
    auto kern = k_bilinear_dense<double, 4, 4, CK>;
    LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<grid, 256, sm, st>>>(N, M, (int)C, (const double*)left, (const double*)right, (const double*)w,
                                (double*)out, accumulate);
  } else {
    return fail(LOB_ERR_ARG, "dtype must be LOB_F32 or LOB_F64");
  }
  return check_launch("k_bilinear_dense");
}

extern "C" int lob_bilinear_diag(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* left, const void* right,
                                 const void* w, void* out, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0, "lob_bilinear_diag: sizes must be positive");
  LOB_REQUIRE(left && right && out, "lob_bilinear_diag: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  LOB_DISPATCH_DTYPE(dtype, {
    k_bilinear_diag<scalar_t><<<(unsigned)cdiv(B * N, 256), 256, 0, st>>>(
        B * N, N, (int)C, (const scalar_t*)left, (const scalar_t*)right, (const scalar_t*)w, (scalar_t*)out);
  });
  return check_launch("k_bilinear_diag");
}

extern "C" int lob_tri_inverse(int32_t dtype, int64_t B, int32_t k, const void* Cm, int64_t ldc,
                               int64_t c_batch_stride, void* out, void* stream) {
  LOB_REQUIRE(B > 0 && k > 0, "lob_tri_inverse: sizes must be positive");
  LOB_REQUIRE(Cm && out, "lob_tri_inverse: NULL pointer");
  const size_t smem = (size_t)k * k * dsize(dtype);
  LOB_REQUIRE(smem <= 200 * 1024, "lob_tri_inverse: block too large for shared memory (k <= 226 fp32, 160 fp64)");
  cudaStream_t st = (cudaStream_t)stream;
  LOB_DISPATCH_DTYPE(dtype, {
    LOB_CUDA(cudaFuncSetAttribute(k_tri_inverse<scalar_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tri_inverse<scalar_t><<<(unsigned)B, 256, smem, st>>>(k, (const scalar_t*)Cm, ldc, c_batch_stride, (scalar_t*)out);
  });
  return check_launch("k_tri_inverse");
}

extern "C" int lob_toeplitz_cross_spectrum(int32_t dtype, int64_t B, int64_t C, int64_t H, const void* fu,
                                           const void* fv, const void* w, void* out, void* stream) {
  LOB_REQUIRE(B > 0 && C > 0 && H > 0, "lob_toeplitz_cross_spectrum: sizes must be positive");
  LOB_REQUIRE(fu && fv && out, "lob_toeplitz_cross_spectrum: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  LOB_DISPATCH_DTYPE(dtype, {
    k_cross_spectrum<scalar_t><<<(unsigned)cdiv(B * H, 256), 256, 0, st>>>(
        B, (int)C, H, (const scalar_t*)fu, (const scalar_t*)fv, (const scalar_t*)w, (scalar_t*)out);
  });
  return check_launch("k_cross_spectrum");
}

extern "C" int lob_toeplitz_deriv_finish(int32_t dtype, int64_t B, int64_t N, int64_t L, const void* y, double scale,
                                         void* out, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && L >= N, "lob_toeplitz_deriv_finish: bad sizes");
  LOB_REQUIRE(y && out, "lob_toeplitz_deriv_finish: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  LOB_DISPATCH_DTYPE(dtype, {
    k_toeplitz_deriv_finish<scalar_t><<<(unsigned)cdiv(B * N, 256), 256, 0, st>>>(B, N, L, (const scalar_t*)y, scale,
                                                                                 (scalar_t*)out);
  });
  return check_launch("k_toeplitz_deriv_finish");
}

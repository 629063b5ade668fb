// structured.cu -- structured-operator matmul pieces and the low-rank Woodbury capacitance solve.
//   Kronecker : kronecker_product_linear_operator.py:34-45   (one fused mode product per factor)
//   Toeplitz  : utils/toeplitz.py:131-149                    (pad / embed / pointwise-multiply / unpad around cuFFT)
//   Low rank  : low_rank_root_added_diag_linear_operator.py:36-47,62-101
#include <algorithm>

#include "common.cuh"
#include "simt_tile.cuh"

namespace lob {

// Y[b, q, i, c] = sum_j K[b, i, j] X[b, j, q, c].   X is (B, n, Q*C) row-major, so this is the GEMM
// (n x n)(n x QC) with the output tile scattered as (Q, n, C): the reference's bmm + transposing copy in one pass.
template <typename T, int RN>
__global__ void __launch_bounds__(256)
k_kron_mode(int64_t n, int64_t Q, int64_t C, const T* __restrict__ K, int64_t k_bs, const T* __restrict__ X,
            T* __restrict__ Y) {
  constexpr int TK = TileK<T>::value;
  constexpr int CP = 8 * RN;
  __shared__ __align__(16) T As[TK * LDA_S];
  __shared__ __align__(16) T Bs[TK * CP];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int64_t b = blockIdx.y;
  const int64_t QC = Q * C;
  const int64_t f0 = (int64_t)blockIdx.x * CP;
  const int64_t m0 = (int64_t)blockIdx.z * TM;
  const T* Kb = K + b * k_bs;
  const T* Xb = X + b * n * QC;
  T* Yb = Y + b * n * QC;
  const int ncol = (int)min((int64_t)CP, QC - f0);
  T acc[4][RN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = (T)0;
  for (int64_t k0 = 0; k0 < n; k0 += TK) {
    for (int e = tid; e < TM * TK; e += 256) {
      const int row = e / TK, kk = e % TK;
      T v = (T)0;
      if (m0 + row < n && k0 + kk < n) v = __ldg(Kb + (m0 + row) * n + k0 + kk);
      As[kk * LDA_S + row] = v;
    }
    for (int e = tid; e < TK * CP; e += 256) {
      const int kk = e / CP, cc = e % CP;
      T v = (T)0;
      if (k0 + kk < n && cc < ncol) v = __ldg(Xb + (k0 + kk) * QC + f0 + cc);
      Bs[e] = v;
    }
    __syncthreads();
    tile_fma<T, T, RN, TK>(As, Bs, CP, ty, tx, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = m0 + ty * 4 + i;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int cc = tx * RN + j;
      if (cc >= ncol) continue;
      const int64_t f = f0 + cc, q = f / C, c = f % C;
      Yb[(q * n + row) * C + c] = acc[i][j];
    }
  }
}

// xt (B, C, L) = [X^T, 0]
template <typename T>
__global__ void __launch_bounds__(256)
k_toeplitz_pad(int64_t N, int64_t C, int64_t L, const T* __restrict__ X, T* __restrict__ xt) {
  __shared__ T tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t l0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int nl = ty; nl < 32; nl += 8) {
    const int64_t n = l0 + nl, c = c0 + tx;
    tile[nl][tx] = (n < N && c < C) ? X[(b * N + n) * C + c] : (T)0;
  }
  __syncthreads();
  for (int cl = ty; cl < 32; cl += 8) {
    const int64_t c = c0 + cl, l = l0 + tx;
    if (c < C && l < L) xt[(b * C + c) * L + l] = tile[tx][cl];
  }
}

// c (B, L) = [col_0..col_{N-1}, 0, ..., 0, col_{N-1}..col_1]   (utils/toeplitz.py:131-139 with padding in the middle)
template <typename T>
__global__ void k_toeplitz_embed(int64_t N, int64_t L, const T* __restrict__ col, int64_t col_bs, T* __restrict__ c,
                                 int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / L, l = idx % L;
  T v = (T)0;
  if (l < N) v = col[b * col_bs + l];
  else if (L - l < N) v = col[b * col_bs + (L - l)];
  c[idx] = v;
}

// fx[b,c,h] *= fc[b,h]  (complex)
template <typename T2>
__global__ void k_toeplitz_mul(int64_t C, int64_t H, const T2* __restrict__ fc, int64_t fc_bs, T2* __restrict__ fx,
                               int64_t total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int64_t b = idx / (C * H), h = idx % H;
  const T2 a = fx[idx], w = fc[b * fc_bs + h];
  T2 r;
  r.x = a.x * w.x - a.y * w.y;
  r.y = a.x * w.y + a.y * w.x;
  fx[idx] = r;
}

// Y (B,N,C) = scale * yt[b,c,n] (+ d (.) X)
template <typename T>
__global__ void __launch_bounds__(256)
k_toeplitz_unpad(int64_t N, int64_t C, int64_t L, const T* __restrict__ yt, T scale, const T* __restrict__ X,
                 const T* __restrict__ d, int64_t d_bs, int64_t d_st, T* __restrict__ Y) {
  __shared__ T tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t n0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int cl = ty; cl < 32; cl += 8) {
    const int64_t c = c0 + cl, n = n0 + tx;
    tile[cl][tx] = (c < C && n < N) ? yt[(b * C + c) * L + n] : (T)0;
  }
  __syncthreads();
  for (int nl = ty; nl < 32; nl += 8) {
    const int64_t n = n0 + nl, c = c0 + tx;
    if (n < N && c < C) {
      const int64_t idx = (b * N + n) * C + c;
      T v = scale * tile[tx][nl];
      if (d) v += d[b * d_bs + n * d_st] * X[idx];
      Y[idx] = v;
    }
  }
}


// Copies `count` consecutive elements global -> shared with 16-byte loads when both sides allow it (every thread of the
// block takes part; no barrier inside).  Wide loads are what puts enough bytes in flight: with 4-byte loads the
// write-back passes below ran at 3 TB/s, stalled on the long scoreboard with DRAM a third busy (ncu).
template <typename T>
__device__ __forceinline__ void stage_flat(T* __restrict__ dst, const T* __restrict__ src, int count, int tid, int nthr) {
  constexpr int V = 16 / sizeof(T);
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int nv = count / V;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = tid; i < nv; i += nthr) d4[i] = s4[i];
    for (int i = nv * V + tid; i < count; i += nthr) dst[i] = src[i];
  } else {
    for (int i = tid; i < count; i += nthr) dst[i] = src[i];
  }
}

// Flat-coalesced versions of the two transposing passes (used when a row block of all C columns fits in shared memory):
// a block of `rows` vector rows is ONE contiguous run of rows * C elements in (B, N, C), read / written with
// consecutive threads on consecutive addresses; the plane side moves `rows` consecutive samples per column.  The
// 32 x 32-tile kernels above re-read every 132-byte row twice at C = 33 (ncu: 2 x over-fetch, 2.1 TB/s).
template <typename T>
__global__ void __launch_bounds__(256)
k_planes_in(int64_t N, int64_t C, int64_t L, int rows, const T* __restrict__ X, T* __restrict__ xt) {
  extern __shared__ __align__(16) unsigned char pl_smem[];
  T* tile = reinterpret_cast<T*>(pl_smem);  // [rows][C | 1]
  const int ld = (int)C | 1;
  const int64_t b = blockIdx.y, n0 = (int64_t)blockIdx.x * rows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  T* xb = xt + b * C * L;
  const int span = (int)min((int64_t)rows, L - n0);  // samples of this block inside the padded length
  if (n0 >= N) {
    for (int c = warp; c < C; c += 8)
      for (int l = lane; l < span; l += 32) xb[(int64_t)c * L + n0 + l] = (T)0;
    return;
  }
  const int valid = (int)min((int64_t)rows, N - n0);
  const T* src = X + (b * N + n0) * C;
  {
    const int dq = 256 / (int)C, dr = 256 - dq * (int)C;
    int r = tid / (int)C, c = tid - r * (int)C;
    for (int e = tid; e < valid * (int)C; e += 256) {
      tile[r * ld + c] = src[e];
      r += dq;
      c += dr;
      if (c >= (int)C) { c -= (int)C; ++r; }
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += 8)
    for (int l = lane; l < span; l += 32) xb[(int64_t)c * L + n0 + l] = l < valid ? tile[l * ld + c] : (T)0;
}

// blockDim.x = C * ry threads (ry = 256 / C): in the write phase thread (tx = column, ty) owns the elements
// (row ty + k ry, column tx) -- consecutive threads still touch consecutive addresses (a row block is one contiguous
// run), and every thread stays in ONE column, so the optional <X, Y> partial sums are a register accumulation plus one
// reduction over ty at the end.
template <typename T>
__global__ void __launch_bounds__(256)
k_planes_out(int64_t N, int64_t C, int64_t L, int rows, const T* __restrict__ yt, T scale, const T* __restrict__ X,
             const T* __restrict__ d, int64_t d_bs, int64_t d_st, T* __restrict__ Y, double* __restrict__ dots) {
  extern __shared__ __align__(16) unsigned char pl_smem[];
  T* tile = reinterpret_cast<T*>(pl_smem);             // [rows][C | 1]  gathered planes
  const int ld = (int)C | 1;
  T* xs = tile + (rows * ld + 3) / 4 * 4;              // [rows][C]      X rows of this block (same pitch as global)
  T* ds = xs + (rows * (int)C + 3) / 4 * 4;            // [rows]         diagonal
  const int64_t b = blockIdx.y, n0 = (int64_t)blockIdx.x * rows;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int valid = (int)min((int64_t)rows, N - n0);
  const T* yb = yt + b * C * L;
  const int64_t base = (b * N + n0) * C;
  const bool need_x = d || dots;
  // every global read of the block is requested before the one barrier
  const int shift = 31 - __clz(rows);  // rows is a power of two (planes_rows)
  if (sizeof(T) == 4 && (L & 3) == 0 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(yt) & 15) == 0) {
    // four consecutive samples of a plane per load (n0 and L are multiples of 4: aligned)
    for (int idx = tid; idx < (int)C * (rows >> 2); idx += nthr) {
      const int c = idx >> (shift - 2), l = (idx & ((rows >> 2) - 1)) << 2;
      if (l < valid) {
        const float4 q = *reinterpret_cast<const float4*>(yb + (int64_t)c * L + n0 + l);
        const T v[4] = {(T)q.x, (T)q.y, (T)q.z, (T)q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (l + i < valid) tile[(l + i) * ld + c] = v[i];
      }
    }
  } else {
    for (int idx = tid; idx < (int)C * rows; idx += nthr) {
      const int c = idx >> shift, l = idx & (rows - 1);
      if (l < valid) tile[l * ld + c] = yb[(int64_t)c * L + n0 + l];
    }
  }
  if (need_x) stage_flat(xs, X + base, valid * (int)C, tid, nthr);
  if (d)
    for (int l = tid; l < valid; l += nthr) ds[l] = d[b * d_bs + (n0 + l) * d_st];
  __syncthreads();
  const int ry = nthr / (int)C;
  const int ty = tid / (int)C, tx = tid - ty * (int)C;
  double acc = 0.0;
  for (int r = ty; r < valid; r += ry) {
    const T x = need_x ? xs[r * (int)C + tx] : (T)0;
    T v = scale * tile[r * ld + tx];
    if (d) v += ds[r] * x;
    Y[base + (int64_t)r * C + tx] = v;
    acc += (double)x * (double)v;
  }
  if (dots) {
    // partial <X, Y> of this row block per column (linear_cg.py:250-251 fused into the last pass of the product)
    __syncthreads();
    double* red = reinterpret_cast<double*>(pl_smem);  // [ry][C] <= 2 KB (the launcher sizes the buffer for both uses)
    red[ty * C + tx] = acc;
    __syncthreads();
    if (ty == 0) {
      double a = 0.0;
      for (int i = 0; i < ry; ++i) a += red[i * C + tx];
      dots[(b * gridDim.x + blockIdx.x) * C + tx] = a;
    }
  }
}

// shared memory of k_planes_out / k_toeplitz_unpack for `rows` rows: plane tile + X rows + diagonal (+ column scales)
static size_t out_tile_bytes(int rows, int64_t C, size_t elem) {
  return ((size_t)(rows * (C | 1) + 3) / 4 * 4 + (size_t)(rows * C + 3) / 4 * 4 + rows + C) * elem;
}

// rows per CTA of pack / unpack (the tile also holds the C column scales); 0: too many columns
static int pair_rows(int64_t C, size_t elem) {
  if (C > 256) return 0;
  for (int r = 128; r >= 32; r >>= 1)
    if (out_tile_bytes(r, C, elem) <= 48 * 1024) return r;
  return 0;
}

// rows per CTA of the flat kernels: 128, 64 or 32 (a power of two), tile within 48 KB; 0: use the 32 x 32-tile kernels
static int planes_rows(int64_t C, size_t elem) {
  if (C > 256) return 0;
  for (int r = 128; r >= 32; r >>= 1)
    if (out_tile_bytes(r, C, elem) <= 48 * 1024) return r;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Symmetric Toeplitz product through COMPLEX FFTs of column pairs (utils/toeplitz.py:131-149).  The circulant embedding
// of a symmetric Toeplitz matrix has a REAL spectrum, so  ifft(F . fft(x_a + i x_b)) = y_a + i y_b : two real columns
// ride one complex transform, with none of the real-to-complex pre/post-processing passes and no separate transposes:
//   colmax (per-column max |x|) -> pack (transpose + pair + zero-pad, columns scaled by an exact power of two to a
//   common magnitude so the rounding of one column cannot leak into a much smaller partner) -> cuFFT C2C (in place) ->
//   mulr (real spectrum) -> cuFFT C2C inverse (in place) -> unpack (first N samples, 1/L, un-scale, + d (.) X).
// ---------------------------------------------------------------------------------------------------------------
template <typename T> struct PairOf;
template <> struct PairOf<float> { using type = float2; using bits = unsigned int; };
template <> struct PairOf<double> { using type = double2; using bits = unsigned long long; };
__device__ __forceinline__ unsigned int abs_bits(float v) { return __float_as_uint(fabsf(v)); }
__device__ __forceinline__ unsigned long long abs_bits(double v) { return (unsigned long long)__double_as_longlong(fabs(v)); }
// exact power-of-two scale that brings a column with max |x| = m to [1, 2); 1 for an all-zero (or non-finite) column
__device__ __forceinline__ float col_scale(unsigned int mb) {
  const int e = (int)(mb >> 23);
  if (e == 0 || e >= 254) return 1.f;
  return __uint_as_float((unsigned int)(254 - e) << 23);
}
__device__ __forceinline__ double col_scale(unsigned long long mb) {
  const int e = (int)(mb >> 52);
  if (e == 0 || e >= 2046) return 1.0;
  return __longlong_as_double((long long)(2046 - e) << 52);
}


// maxbits[b, c] = bit pattern of max_n |X[b, n, c]|  (non-negative IEEE values order like unsigned integers; the
// result of an atomic max does not depend on the order of arrival).  block (cx = min(C, 128), ry), grid (chunks, B)
template <typename T>
__global__ void k_toeplitz_colmax(int64_t N, int64_t C, int64_t rows_per_cta, const T* __restrict__ X,
                                  typename PairOf<T>::bits* __restrict__ maxbits) {
  using bits_t = typename PairOf<T>::bits;
  extern __shared__ unsigned long long cm_red[];  // [ry][cx]
  const int64_t b = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(N, r0 + rows_per_cta);
  const int tx = threadIdx.x, ty = threadIdx.y, CX = blockDim.x, RY = blockDim.y;
  for (int64_t c0 = 0; c0 < C; c0 += CX) {
    const int64_t c = c0 + tx;
    bits_t m = 0;
    if (c < C) {
      const T* col = X + b * N * C + c;
      int64_t r = r0 + ty;
      for (; r + 3 * RY < r1; r += 4 * RY) {
        const T v0 = col[r * C], v1 = col[(r + RY) * C], v2 = col[(r + 2 * RY) * C], v3 = col[(r + 3 * RY) * C];
        m = max(max(m, abs_bits(v0)), max(abs_bits(v1), max(abs_bits(v2), abs_bits(v3))));
      }
      for (; r < r1; r += RY) m = max(m, abs_bits(col[r * C]));
    }
    cm_red[ty * CX + tx] = m;
    __syncthreads();
    if (ty == 0 && c < C) {
      unsigned long long mm = 0;
      for (int i = 0; i < RY; ++i) mm = max(mm, cm_red[i * CX + tx]);
      if (mm) atomicMax(maxbits + b * C + c, (bits_t)mm);
    }
    __syncthreads();
  }
}

// zt (B, P, L) complex, P = ceil(C / 2):  zt[b, p, l] = (s_2p X[b, l, 2p], s_2p+1 X[b, l, 2p+1]) for l < N, 0 above
template <typename T>
__global__ void __launch_bounds__(256)
k_toeplitz_pack(int64_t N, int64_t C, int64_t L, int rows_cta, const T* __restrict__ X,
                const typename PairOf<T>::bits* __restrict__ maxbits, typename PairOf<T>::type* __restrict__ zt) {
  using T2 = typename PairOf<T>::type;
  extern __shared__ __align__(16) unsigned char tp_smem[];
  T* tile = reinterpret_cast<T*>(tp_smem);  // [rows_cta][C | 1], rows_cta a power of two (planes_rows)
  const int ld = (int)C | 1;
  T* sc = tile + rows_cta * ld;             // [C]
  const int shift = 31 - __clz(rows_cta);
  const int64_t b = blockIdx.y, n0 = (int64_t)blockIdx.x * rows_cta;
  const int P = (int)((C + 1) / 2);
  const int tid = threadIdx.x;
  T2* zb = zt + b * P * L;
  if (n0 >= N) {  // zero padding
    for (int idx = tid; idx < P * rows_cta; idx += 256) {
      const int pp = idx >> shift, l = idx & (rows_cta - 1);
      if (n0 + l < L) zb[(int64_t)pp * L + n0 + l] = T2{(T)0, (T)0};
    }
    return;
  }
  for (int c = tid; c < C; c += 256) sc[c] = col_scale(maxbits[b * C + c]);
  const int rows = (int)min((int64_t)rows_cta, N - n0);
  const T* src = X + (b * N + n0) * C;
  {  // the rows of this block are one contiguous run: flat coalesced loads, (row, column) advanced without divisions
    const int dq = 256 / (int)C, dr = 256 - dq * (int)C;
    int r = tid / (int)C, c = tid - r * (int)C;
    for (int e = tid; e < rows * (int)C; e += 256) {
      tile[r * ld + c] = src[e];
      r += dq;
      c += dr;
      if (c >= (int)C) { c -= (int)C; ++r; }
    }
  }
  __syncthreads();
  for (int idx = tid; idx < P * rows_cta; idx += 256) {
    const int pp = idx >> shift, l = idx & (rows_cta - 1);
    if (n0 + l >= L) continue;
    T2 z{(T)0, (T)0};
    if (l < rows) {
      z.x = tile[l * ld + 2 * pp] * sc[2 * pp];
      if (2 * pp + 1 < C) z.y = tile[l * ld + 2 * pp + 1] * sc[2 * pp + 1];
    }
    zb[(int64_t)pp * L + n0 + l] = z;
  }
}

// zt[b, p, k] *= fr[b, min(k, L - k)]   (fr: real spectrum of the embedding, L / 2 + 1 entries; two samples per thread)
template <typename T>
__global__ void k_toeplitz_mulr(int64_t P, int64_t L, const T* __restrict__ fr, int64_t fr_bs,
                                typename PairOf<T>::type* __restrict__ zt, int64_t total_pairs) {
  using T2 = typename PairOf<T>::type;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_pairs) return;
  const int64_t e = 2 * idx;  // L is even: a pair of samples never straddles two sequences
  const int64_t seq = e / L, k = e - seq * L;
  const int64_t b = seq / P;
  const T* f = fr + b * fr_bs;
  const T f0 = f[k <= L - k ? k : L - k];
  const T f1 = f[k + 1 <= L - k - 1 ? k + 1 : L - k - 1];
  T2 a = zt[e], c = zt[e + 1];
  a.x *= f0, a.y *= f0, c.x *= f1, c.y *= f1;
  zt[e] = a;
  zt[e + 1] = c;
}

// Y[b, n, c] = (scale / s_c) * part_c(zt[b, c / 2, n]) (+ d (.) X); structure of k_planes_out
template <typename T>
__global__ void __launch_bounds__(256)
k_toeplitz_unpack(int64_t N, int64_t C, int64_t L, int rows_cta, const typename PairOf<T>::type* __restrict__ zt,
                  T scale, const typename PairOf<T>::bits* __restrict__ maxbits, const T* __restrict__ X,
                  const T* __restrict__ d, int64_t d_bs, int64_t d_st, T* __restrict__ Y, double* __restrict__ dots) {
  using T2 = typename PairOf<T>::type;
  extern __shared__ __align__(16) unsigned char tp_smem[];
  T* tile = reinterpret_cast<T*>(tp_smem);                 // [rows][C | 1]
  const int ld = (int)C | 1;
  T* xs = tile + (rows_cta * ld + 3) / 4 * 4;              // [rows][C]
  T* ds = xs + (rows_cta * (int)C + 3) / 4 * 4;            // [rows]
  T* sc = ds + rows_cta;                                   // [C]
  const int shift = 31 - __clz(rows_cta);
  const int64_t b = blockIdx.y, n0 = (int64_t)blockIdx.x * rows_cta;
  const int P = (int)((C + 1) / 2);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int rows = (int)min((int64_t)rows_cta, N - n0);
  const int64_t base = (b * N + n0) * C;
  const bool need_x = d || dots;
  // an all-zero input column gives exactly zero (not its partner's rounding noise)
  for (int c = tid; c < C; c += nthr) sc[c] = maxbits[b * C + c] ? scale / col_scale(maxbits[b * C + c]) : (T)0;
  const T2* zb = zt + b * P * L;
  // two samples (32 bytes of fp64, 16 of fp32) per load: L and n0 are even, so the pair is aligned
  for (int idx = tid; idx < P * (rows_cta >> 1); idx += nthr) {
    const int pp = idx >> (shift - 1), l = (idx & ((rows_cta >> 1) - 1)) << 1;
    if (l < rows) {
      const T2* zp = zb + (int64_t)pp * L + n0 + l;
      T2 z0, z1;
      if constexpr (sizeof(T) == 4) {
        const float4 q = *reinterpret_cast<const float4*>(zp);
        z0 = T2{q.x, q.y}, z1 = T2{q.z, q.w};
      } else {
        z0 = zp[0], z1 = zp[1];
      }
      tile[l * ld + 2 * pp] = z0.x;
      if (2 * pp + 1 < C) tile[l * ld + 2 * pp + 1] = z0.y;
      if (l + 1 < rows) {
        tile[(l + 1) * ld + 2 * pp] = z1.x;
        if (2 * pp + 1 < C) tile[(l + 1) * ld + 2 * pp + 1] = z1.y;
      }
    }
  }
  if (need_x) stage_flat(xs, X + base, rows * (int)C, tid, nthr);
  if (d)
    for (int l = tid; l < rows; l += nthr) ds[l] = d[b * d_bs + (n0 + l) * d_st];
  __syncthreads();
  const int ry = nthr / (int)C;
  const int ty = tid / (int)C, tx = tid - ty * (int)C;
  const T scl = sc[tx];
  double acc = 0.0;
  for (int r = ty; r < rows; r += ry) {
    const T x = need_x ? xs[r * (int)C + tx] : (T)0;
    T v = tile[r * ld + tx] * scl;
    if (d) v += ds[r] * x;
    Y[base + (int64_t)r * C + tx] = v;
    acc += (double)x * (double)v;
  }
  if (dots) {
    __syncthreads();
    double* red = reinterpret_cast<double*>(tp_smem);  // [ry][C] <= 2 KB
    red[ty * C + tx] = acc;
    __syncthreads();
    if (ty == 0) {
      double a = 0.0;
      for (int i = 0; i < ry; ++i) a += red[i * C + tx];
      dots[(b * gridDim.x + blockIdx.x) * C + tx] = a;
    }
  }
}

// cap = I + G; W <- cap^-1 W via Cholesky in a double workspace held in global memory (L2 resident), one CTA per
// batch element.  logdet_cap = 2 sum log diag chol(cap).
template <typename T>
__global__ void __launch_bounds__(256)
k_cap_solve(int k, int64_t C, const double* __restrict__ G, int64_t g_bs, T* __restrict__ W, T* __restrict__ logdet,
            int32_t* __restrict__ info, double* __restrict__ ws) {
  __shared__ double scratch[32];
  __shared__ int bad;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  double* a = ws + b * ((int64_t)k * k + (int64_t)k * C);
  double* w = a + (int64_t)k * k;
  const double* g = G + b * g_bs;
  if (tid == 0) bad = 0;
  for (int e = tid; e < k * k; e += blockDim.x) a[e] = g[e] + ((e / k) == (e % k) ? 1.0 : 0.0);
  for (int64_t e = tid; e < (int64_t)k * C; e += blockDim.x) w[e] = (double)W[b * k * C + e];
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    if (tid == 0) {
      const double p = a[j * k + j];
      if (!(p > 0.0)) bad = 1;
      a[j * k + j] = sqrt(p);
    }
    __syncthreads();
    const double djj = a[j * k + j];
    for (int i = j + 1 + tid; i < k; i += blockDim.x) a[i * k + j] /= djj;
    __syncthreads();
    const int rem = k - j - 1;
    for (int e = tid; e < rem * rem; e += blockDim.x) {
      const int i = j + 1 + e / rem, m = j + 1 + e % rem;
      if (m <= i) a[i * k + m] -= a[i * k + j] * a[m * k + j];
    }
    __syncthreads();
  }
  double acc = 0.0;
  for (int i = tid; i < k; i += blockDim.x) acc += log(a[i * k + i]);
  const double ls = block_sum(acc, scratch);
  if (tid == 0) {
    logdet[b] = (T)(2.0 * ls);
    info[b] = bad;
  }
  // forward then backward substitution, one thread per right-hand-side column
  for (int64_t c = tid; c < C; c += blockDim.x) {
    for (int i = 0; i < k; ++i) {
      double s = w[i * C + c];
      for (int m = 0; m < i; ++m) s -= a[i * k + m] * w[m * C + c];
      w[i * C + c] = s / a[i * k + i];
    }
    for (int i = k - 1; i >= 0; --i) {
      double s = w[i * C + c];
      for (int m = i + 1; m < k; ++m) s -= a[m * k + i] * w[m * C + c];
      w[i * C + c] = s / a[i * k + i];
    }
    for (int i = 0; i < k; ++i) W[b * k * C + i * C + c] = (T)w[i * C + c];
  }
}

}  // namespace lob

namespace lob {
template <typename T>
static int launch_kron(int64_t B, int64_t n, int64_t Q, int64_t C, const T* K, int64_t k_bs, const T* X, T* Y,
                       cudaStream_t st) {
  const int64_t QC = Q * C;
  const int rn = pick_rn(QC);
  dim3 grid((unsigned)cdiv(QC, 8 * rn), (unsigned)B, (unsigned)cdiv(n, TM));
#define LOB_KR_CASE(R)                                                    \
  case R:                                                                 \
    k_kron_mode<T, R><<<grid, 256, 0, st>>>(n, Q, C, K, k_bs, X, Y);      \
    break;
  switch (rn) {
    LOB_KR_CASE(1) LOB_KR_CASE(2) LOB_KR_CASE(3) LOB_KR_CASE(4) LOB_KR_CASE(5) LOB_KR_CASE(6) LOB_KR_CASE(7)
    LOB_KR_CASE(8)
  }
#undef LOB_KR_CASE
  return check_launch("k_kron_mode");
}
}  // namespace lob

using namespace lob;

extern "C" int lob_kron_mode_matmul(int32_t dtype, int64_t B, int64_t n, int64_t Q, int64_t C, const void* K,
                                    int64_t k_batch_stride, const void* X, void* Y, void* stream) {
  LOB_REQUIRE(B > 0 && n > 0 && Q > 0 && C > 0, "lob_kron_mode_matmul: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_kron_mode_matmul: flattened batch > 65535 not supported");
  LOB_REQUIRE(K && X && Y && X != Y, "lob_kron_mode_matmul: NULL or aliased pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    return launch_kron<scalar_t>(B, n, Q, C, (const scalar_t*)K, k_batch_stride, (const scalar_t*)X, (scalar_t*)Y,
                                 (cudaStream_t)stream);
  });
}

extern "C" int lob_toeplitz_pad(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* X, void* xt,
                                void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0 && L >= N, "lob_toeplitz_pad: bad sizes");
  LOB_REQUIRE(B <= 65535, "lob_toeplitz_pad: flattened batch > 65535 not supported");
  LOB_REQUIRE(X && xt, "lob_toeplitz_pad: NULL pointer");
  dim3 grid((unsigned)cdiv(L, 32), (unsigned)cdiv(C, 32), (unsigned)B);
  LOB_DISPATCH_DTYPE(dtype, {
    const int rows = planes_rows(C, sizeof(scalar_t));
    if (rows) {
      const size_t smem = sizeof(scalar_t) * (size_t)rows * ((int)C | 1);
      k_planes_in<scalar_t><<<dim3((unsigned)cdiv(L, rows), (unsigned)B), 256, smem, (cudaStream_t)stream>>>(
          N, C, L, rows, (const scalar_t*)X, (scalar_t*)xt);
      return check_launch("k_planes_in");
    }
    k_toeplitz_pad<scalar_t><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(N, C, L, (const scalar_t*)X,
                                                                            (scalar_t*)xt);
    return check_launch("k_toeplitz_pad");
  });
}

extern "C" int lob_toeplitz_embed(int32_t dtype, int64_t B, int64_t N, int64_t L, const void* col,
                                  int64_t col_batch_stride, void* c, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && L >= 2 * N - 1, "lob_toeplitz_embed: need L >= 2N-1");
  LOB_REQUIRE(col && c, "lob_toeplitz_embed: NULL pointer");
  const int64_t total = B * L;
  LOB_DISPATCH_DTYPE(dtype, {
    k_toeplitz_embed<scalar_t><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        N, L, (const scalar_t*)col, col_batch_stride, (scalar_t*)c, total);
    return check_launch("k_toeplitz_embed");
  });
}

extern "C" int lob_toeplitz_mul(int32_t dtype, int64_t B, int64_t C, int64_t H, const void* fc, int64_t fc_batch_stride,
                                void* fx, void* stream) {
  LOB_REQUIRE(B > 0 && C > 0 && H > 0, "lob_toeplitz_mul: sizes must be positive");
  LOB_REQUIRE(fc && fx, "lob_toeplitz_mul: NULL pointer");
  const int64_t total = B * C * H;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == LOB_F32) {
    k_toeplitz_mul<float2><<<(unsigned)cdiv(total, 256), 256, 0, st>>>(C, H, (const float2*)fc, fc_batch_stride,
                                                                      (float2*)fx, total);
  } else if (dtype == LOB_F64) {
    k_toeplitz_mul<double2><<<(unsigned)cdiv(total, 256), 256, 0, st>>>(C, H, (const double2*)fc, fc_batch_stride,
                                                                       (double2*)fx, total);
  } else {
    return fail(LOB_ERR_ARG, "dtype must be LOB_F32 or LOB_F64");
  }
  return check_launch("k_toeplitz_mul");
}

extern "C" int lob_toeplitz_colmax(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* X, void* maxbits,
                                   void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0, "lob_toeplitz_colmax: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_toeplitz_colmax: flattened batch > 65535 not supported");
  LOB_REQUIRE(X && maxbits, "lob_toeplitz_colmax: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  LOB_CUDA(cudaMemsetAsync(maxbits, 0, (size_t)B * C * dsize(dtype), st));
  const int cx = (int)(C < 128 ? C : 128), ry = 256 / cx > 0 ? 256 / cx : 1;
  int64_t chunks = cdiv((int64_t)kNumSMs * 8, B);
  const int64_t maxch = cdiv(N, (int64_t)ry * 4);
  if (chunks > maxch) chunks = maxch;
  const int64_t rows = cdiv(N, chunks);
  dim3 grid((unsigned)cdiv(N, rows), (unsigned)B);
  const size_t smem = sizeof(unsigned long long) * cx * ry;
  LOB_DISPATCH_DTYPE(dtype, {
    k_toeplitz_colmax<scalar_t><<<grid, dim3(cx, ry), smem, st>>>(N, C, rows, (const scalar_t*)X,
                                                                   (typename PairOf<scalar_t>::bits*)maxbits);
    return check_launch("k_toeplitz_colmax");
  });
}

extern "C" int lob_toeplitz_pack(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* X,
                                 const void* maxbits, void* zt, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0 && L >= N, "lob_toeplitz_pack: bad sizes");
  LOB_REQUIRE(B <= 65535 && C <= 4096, "lob_toeplitz_pack: flattened batch > 65535 or more than 4096 columns not supported");
  LOB_REQUIRE(X && maxbits && zt, "lob_toeplitz_pack: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    const int rows = pair_rows(C, sizeof(scalar_t));
    LOB_REQUIRE(rows > 0, "lob_toeplitz_pack: too many columns for one shared-memory tile (split the column block)");
    dim3 grid((unsigned)cdiv(L, rows), (unsigned)B);
    const size_t smem = sizeof(scalar_t) * ((size_t)rows * ((int)C | 1) + C);
    k_toeplitz_pack<scalar_t><<<grid, 256, smem, (cudaStream_t)stream>>>(
        N, C, L, rows, (const scalar_t*)X, (const typename PairOf<scalar_t>::bits*)maxbits,
        (typename PairOf<scalar_t>::type*)zt);
    return check_launch("k_toeplitz_pack");
  });
}

extern "C" int lob_toeplitz_mulr(int32_t dtype, int64_t B, int64_t P, int64_t L, const void* fr, int64_t fr_batch_stride,
                                 void* zt, void* stream) {
  LOB_REQUIRE(B > 0 && P > 0 && L > 0 && (L % 2) == 0, "lob_toeplitz_mulr: sizes must be positive, L even");
  LOB_REQUIRE(fr && zt, "lob_toeplitz_mulr: NULL pointer");
  const int64_t total = B * P * (L / 2);
  LOB_DISPATCH_DTYPE(dtype, {
    k_toeplitz_mulr<scalar_t><<<(unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
        P, L, (const scalar_t*)fr, fr_batch_stride, (typename PairOf<scalar_t>::type*)zt, total);
    return check_launch("k_toeplitz_mulr");
  });
}

extern "C" int32_t lob_toeplitz_unpack_parts(int32_t dtype, int64_t N, int64_t C) {
  const int rows = pair_rows(C, dsize(dtype));
  return rows ? (int32_t)cdiv(N, rows) : 0;
}

extern "C" int lob_toeplitz_unpack(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* zt,
                                   double scale, const void* maxbits, const void* X, const void* d,
                                   int64_t d_batch_stride, int64_t d_stride, void* Y, double* dots, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0 && L >= N, "lob_toeplitz_unpack: bad sizes");
  LOB_REQUIRE(B <= 65535 && C <= 4096, "lob_toeplitz_unpack: flattened batch > 65535 or more than 4096 columns not supported");
  LOB_REQUIRE(zt && maxbits && Y && ((!d && !dots) || X), "lob_toeplitz_unpack: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    const int rows = pair_rows(C, sizeof(scalar_t));
    LOB_REQUIRE(rows > 0, "lob_toeplitz_unpack: too many columns for one shared-memory tile (split the column block)");
    dim3 grid((unsigned)cdiv(N, rows), (unsigned)B);
    const size_t smem = out_tile_bytes(rows, C, sizeof(scalar_t));
    const int ry = std::max<int>(1, 256 / (int)C);
    k_toeplitz_unpack<scalar_t><<<grid, (unsigned)(ry * C), std::max(smem, sizeof(double) * ry * (size_t)C),
                                  (cudaStream_t)stream>>>(
        N, C, L, rows, (const typename PairOf<scalar_t>::type*)zt, (scalar_t)scale,
        (const typename PairOf<scalar_t>::bits*)maxbits, (const scalar_t*)X, (const scalar_t*)d, d_batch_stride,
        d_stride, (scalar_t*)Y, dots);
    return check_launch("k_toeplitz_unpack");
  });
}

// row blocks (= partial <X, Y> sums per column) of lob_toeplitz_unpad; 0: this shape takes the tile kernel, no dots
extern "C" int32_t lob_toeplitz_unpad_parts(int32_t dtype, int64_t N, int64_t C) {
  const int rows = planes_rows(C, dsize(dtype));
  return rows ? (int32_t)cdiv(N, rows) : 0;
}

extern "C" int lob_toeplitz_unpad(int32_t dtype, int64_t B, int64_t N, int64_t C, int64_t L, const void* yt,
                                  double scale, const void* X, const void* d, int64_t d_batch_stride, int64_t d_stride,
                                  void* Y, double* dots, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0 && L >= N, "lob_toeplitz_unpad: bad sizes");
  LOB_REQUIRE(B <= 65535, "lob_toeplitz_unpad: flattened batch > 65535 not supported");
  LOB_REQUIRE(yt && Y && ((!d && !dots) || X), "lob_toeplitz_unpad: NULL pointer");
  LOB_REQUIRE(!dots || planes_rows(C, dsize(dtype)) > 0, "lob_toeplitz_unpad: no fused dots for this column count");
  dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(C, 32), (unsigned)B);
  LOB_DISPATCH_DTYPE(dtype, {
    const int rows = planes_rows(C, sizeof(scalar_t));
    if (rows) {
      const int ry = std::max<int>(1, 256 / (int)C);
      const size_t smem = std::max(out_tile_bytes(rows, C, sizeof(scalar_t)), sizeof(double) * ry * (size_t)C);
      k_planes_out<scalar_t><<<dim3((unsigned)cdiv(N, rows), (unsigned)B), (unsigned)(ry * C), smem,
                               (cudaStream_t)stream>>>(
          N, C, L, rows, (const scalar_t*)yt, (scalar_t)scale, (const scalar_t*)X, (const scalar_t*)d, d_batch_stride,
          d_stride, (scalar_t*)Y, dots);
      return check_launch("k_planes_out");
    }
    k_toeplitz_unpad<scalar_t><<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        N, C, L, (const scalar_t*)yt, (scalar_t)scale, (const scalar_t*)X, (const scalar_t*)d, d_batch_stride,
        d_stride, (scalar_t*)Y);
    return check_launch("k_toeplitz_unpad");
  });
}

extern "C" size_t lob_cap_solve_workspace_bytes(int64_t B, int32_t k, int64_t C) {
  if (B <= 0 || k <= 0 || C <= 0) return 0;
  return (size_t)B * ((size_t)k * k + (size_t)k * C) * sizeof(double);
}

extern "C" int lob_cap_solve(int32_t dtype, int64_t B, int32_t k, int64_t C, const double* G, int64_t g_batch_stride,
                             void* W, void* logdet_cap, int32_t* info, void* ws, void* stream) {
  LOB_REQUIRE(B > 0 && k > 0 && C > 0, "lob_cap_solve: sizes must be positive");
  LOB_REQUIRE(G && W && logdet_cap && info && ws, "lob_cap_solve: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    k_cap_solve<scalar_t><<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(k, C, G, g_batch_stride, (scalar_t*)W,
                                                                        (scalar_t*)logdet_cap, info, (double*)ws);
    return check_launch("k_cap_solve");
  });
}

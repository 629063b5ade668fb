// pivchol.cu -- batched greedy pivoted (partial) Cholesky, reference: functions/_pivoted_cholesky.py:13-105.
//
// The reference runs ~30 gather/scatter launches and one host synchronisation (`torch.max(errors) > tol`) per
// pivot step.  Here a step is two launches and no host round trip:
//   k_pc_pivot  (one CTA)          : finishes the arg-max of the residual diagonal per batch element from the
//                                    per-chunk candidates, evaluates the batch-global stop rule, swaps the permutation
//   k_pc_update (nchunks x B CTAs) : fetches the pivot row from the operator (dense / Kronecker / Toeplitz functor,
//                                    the `_get_indices` of the reference), applies the rank-m correction
//                                    row - sum_t L[t,pi] L[t,:], rescales, updates the residual diagonal and emits
//                                    the arg-max / error candidates of the next step.
// Index semantics reproduced exactly: candidates are compared in *permuted order* (first maximal position wins, like
// torch.max on the gathered diagonal, :61-63), the swap is the reference's (:67-70), perm is int64.
#include <math.h>

#include "common.cuh"

namespace lob {

struct PcControl {
  int32_t active;  // loop still running
  int32_t m;       // steps taken
};

struct PcLayout {
  int nchunks;
  int64_t cols_per_chunk;
  size_t off_ctrl, off_diag, off_pos, off_cand_val, off_cand_pos, off_cand_idx, off_errsum, off_pi, off_piv, off_orig,
      total;
};

static PcLayout pc_layout(int64_t B, int64_t N, int rank, size_t es) {
  PcLayout L;
  // Column chunks per batch element.  Small batches: enough CTAs for ~4 per SM (under one wave whatever the row
  // functor's register count).  Large batches, where one chunk per element already overfills the GPU (B = 1024,
  // N = 5000 ran 1.15 waves of 6 CTAs per SM -- the second 15 % full): the count that maximises (fill of the last
  // wave) x (threads of a 256-thread CTA that own a 16-byte column group).
  int64_t maxch = cdiv(N, 256);
  if (maxch > 256) maxch = 256;
  if (maxch < 1) maxch = 1;
  int64_t nch = cdiv((int64_t)kNumSMs * 4, B);
  if (nch > maxch) nch = maxch;
  if (nch < 1) nch = 1;
  if (B >= (int64_t)kNumSMs * 4) {
    const double slots = (double)kNumSMs * 6;
    double best = -1.0;
    for (int64_t c = 1; c <= maxch && c <= 16; ++c) {
      const int64_t cols = cdiv(cdiv(N, c), 4) * 4, groups = cols / 4;
      const double waves = (double)(B * cdiv(N, cols)) / slots;
      const double eff = (waves / ceil(waves)) * ((double)groups / (double)(cdiv(groups, 256) * 256));
      if (eff > best + 1e-9) {
        best = eff;
        nch = c;
      }
    }
  }
  L.cols_per_chunk = cdiv(cdiv(N, nch), 4) * 4;  // multiple of 4: lets the fp32 update use 16-byte words
  L.nchunks = (int)cdiv(N, L.cols_per_chunk);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  L.off_ctrl = take(sizeof(PcControl));
  L.off_diag = take((size_t)B * N * es);
  L.off_pos = take((size_t)B * N * 4);
  L.off_cand_val = take((size_t)B * L.nchunks * es);
  L.off_cand_pos = take((size_t)B * L.nchunks * 4);
  L.off_cand_idx = take((size_t)B * L.nchunks * 4);
  L.off_errsum = take((size_t)B * L.nchunks * 8);
  L.off_pi = take((size_t)B * 4);
  L.off_piv = take((size_t)B * es);
  L.off_orig = take((size_t)B * es);
  L.total = o;
  (void)rank;
  return L;
}

// ---- row sources ("_get_indices" of each operator class) ---------------------------------------------------
template <typename T>
struct DenseSrc {  // dense_linear_operator.py:47-50, :37-40
  const T* A;
  int64_t lda, bs;
  __device__ __forceinline__ T diag(int64_t b, int64_t i) const { return A[b * bs + i * lda + i]; }
  __device__ __forceinline__ T entry(int64_t b, int64_t r, int64_t i) const { return A[b * bs + r * lda + i]; }
};

template <typename T>
struct KronSrc {  // kronecker_product_linear_operator.py:198-216, :20-27 (first factor slowest)
  const T* f[4];
  int64_t n[4], bs[4], stride[4];
  int nf;
  __device__ __forceinline__ T entry(int64_t b, int64_t r, int64_t i) const {
    T v = (T)1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (q < nf) {
        const int64_t rq = (r / stride[q]) % n[q], iq = (i / stride[q]) % n[q];
        const T e = f[q][b * bs[q] + rq * n[q] + iq];
        v = (q == 0) ? e : v * e;
      }
    }
    return v;
  }
  __device__ __forceinline__ T diag(int64_t b, int64_t i) const { return entry(b, i, i); }
};

template <typename T>
struct ToeplitzSrc {  // toeplitz_linear_operator.py:38-40, :25-31
  const T* col;
  int64_t bs;
  __device__ __forceinline__ T diag(int64_t b, int64_t) const { return col[b * bs]; }
  __device__ __forceinline__ T entry(int64_t b, int64_t r, int64_t i) const {
    const int64_t d = r > i ? r - i : i - r;
    return col[b * bs + d];
  }
};

template <typename T>
struct RowBufSrc {  // generic operators: the host fetched the pivot row (or the diagonal) through `_get_indices`
  const T* buf;     // (B, n): entry(b, r, i) = buf[b, i] whatever r is -- the caller guarantees buf holds row r = pi[b]
  int64_t n;
  __device__ __forceinline__ T diag(int64_t b, int64_t i) const { return buf[b * n + i]; }
  __device__ __forceinline__ T entry(int64_t b, int64_t, int64_t i) const { return buf[b * n + i]; }
};

// candidate ordering: NaN beats everything (torch.max propagates NaN), then larger value, then smaller position
template <typename T>
__device__ __forceinline__ bool cand_better(T v1, int p1, T v2, int p2) {
  const bool n1 = v1 != v1, n2 = v2 != v2;
  if (p2 < 0) return p1 >= 0;
  if (p1 < 0) return false;
  if (n1 || n2) {
    if (n1 && n2) return p1 < p2;
    return n1;
  }
  if (v1 > v2) return true;
  if (v1 < v2) return false;
  return p1 < p2;
}

template <typename T>
__device__ __forceinline__ void block_argmax(T& v, int& p, int& idx, T* sv, int* sp, int* si) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int op = __shfl_xor_sync(0xffffffffu, p, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (cand_better<T>(ov, op, v, p)) {
      v = ov; p = op; idx = oi;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) {
    sv[warp] = v; sp[warp] = p; si[warp] = idx;
  }
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? sv[lane] : (T)0;
    p = lane < nw ? sp[lane] : -1;
    idx = lane < nw ? si[lane] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int op = __shfl_xor_sync(0xffffffffu, p, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (cand_better<T>(ov, op, v, p)) {
        v = ov; p = op; idx = oi;
      }
    }
  }
}

template <typename T, typename Src>
__global__ void __launch_bounds__(256)
k_pc_init(Src src, int64_t N, int nchunks, int64_t cpc, T* __restrict__ diag, int* __restrict__ pos,
          int64_t* __restrict__ perm, T* __restrict__ cval, int* __restrict__ cpos, int* __restrict__ cidx,
          double* __restrict__ errsum, PcControl* ctrl) {
  __shared__ T sv[32];
  __shared__ int sp[32], si[32];
  __shared__ double scratch[32];
  const int64_t b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t i0 = (int64_t)chunk * cpc, i1 = min(i0 + cpc, N);
  T bv = (T)0;
  int bp = -1, bi = -1;
  double es = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const T dv = src.diag(b, i);
    diag[b * N + i] = dv;
    pos[b * N + i] = (int)i;
    perm[b * N + i] = i;
    es += fabs((double)dv);
    if (cand_better<T>(dv, (int)i, bv, bp)) {
      bv = dv; bp = (int)i; bi = (int)i;
    }
  }
  block_argmax<T>(bv, bp, bi, sv, sp, si);
  const double tot = block_sum(es, scratch);
  if (threadIdx.x == 0) {
    cval[b * nchunks + chunk] = bv;
    cpos[b * nchunks + chunk] = bp;
    cidx[b * nchunks + chunk] = bi;
    errsum[b * nchunks + chunk] = tot;
    if (b == 0 && chunk == 0) {
      ctrl->active = 1;
      ctrl->m = 0;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(1024)
k_pc_pivot(int64_t B, int64_t N, int rankmax, int m, int nchunks, double tol, const T* __restrict__ cval,
           const int* __restrict__ cpos, const int* __restrict__ cidx, const double* __restrict__ errsum,
           T* __restrict__ orig, int* __restrict__ pos, int64_t* __restrict__ perm, int* __restrict__ pi,
           T* __restrict__ piv, T* __restrict__ Lt, PcControl* ctrl, int64_t* __restrict__ pi64 = nullptr) {
  __shared__ int any_gt, any_nan;
  if (!ctrl->active) return;
  if (threadIdx.x == 0) {
    any_gt = 0;
    any_nan = 0;
  }
  __syncthreads();
  if (m > 0) {
    // errors = || diag[perm[m:]] ||_1 / orig_error per batch element; loop while max over the batch > tol (:57)
    int gt = 0, nn = 0;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
      double s = 0.0;
      for (int c = 0; c < nchunks; ++c) s += errsum[b * nchunks + c];
      const T err = (T)s / orig[b];
      if (err != err) nn = 1;
      if (err > (T)tol) gt = 1;
    }
    if (gt) atomicOr(&any_gt, 1);
    if (nn) atomicOr(&any_nan, 1);
    __syncthreads();
    if (!(m < rankmax && any_gt && !any_nan)) {
      __syncthreads();
      if (threadIdx.x == 0) ctrl->active = 0;
      return;
    }
  }
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    T bv = (T)0;
    int bp = -1, bi = -1;
    for (int c = 0; c < nchunks; ++c) {
      const T v = cval[b * nchunks + c];
      const int p = cpos[b * nchunks + c], i = cidx[b * nchunks + c];
      if (cand_better<T>(v, p, bv, bp)) {
        bv = v; bp = p; bi = i;
      }
    }
    if (m == 0) orig[b] = bv;
    // swap perm[m] <-> perm[bp]  (:67-70)
    const int64_t old = perm[b * N + m];
    perm[b * N + m] = bi;
    perm[b * N + bp] = old;
    pos[b * N + bi] = m;
    if (bp != m) pos[b * N + old] = bp;
    const T root = (T)sqrt((double)bv);
    pi[b] = bi;
    if (pi64) pi64[b] = bi;
    piv[b] = root;
    Lt[(b * rankmax + m) * N + bi] = root;  // :73-74
  }
  if (threadIdx.x == 0) ctrl->m = m + 1;
}

template <typename T, typename Src>
__global__ void __launch_bounds__(256)
k_pc_update(Src src, int64_t N, int rankmax, int m, int nchunks, int64_t cpc, T* __restrict__ diag,
            const int* __restrict__ pos, const int* __restrict__ pi, const T* __restrict__ piv, T* __restrict__ Lt,
            T* __restrict__ cval, int* __restrict__ cpos, int* __restrict__ cidx, double* __restrict__ errsum,
            const PcControl* ctrl) {
  extern __shared__ unsigned char smem_raw[];
  T* u = reinterpret_cast<T*>(smem_raw);  // u[t] = L[t, pi], t < m
  __shared__ T sv[32];
  __shared__ int sp[32], si[32];
  __shared__ double scratch[32];
  if (!ctrl->active || ctrl->m != m + 1) return;
  const int64_t b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t i0 = (int64_t)chunk * cpc, i1 = min(i0 + cpc, N);
  const int64_t r = pi[b];
  const T pv = piv[b];
  T* Lb = Lt + b * rankmax * N;
  for (int t = threadIdx.x; t < m; t += blockDim.x) u[t] = Lb[(int64_t)t * N + r];
  __syncthreads();
  T bv = (T)0;
  int bp = -1, bi = -1;
  double es = 0.0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int p = pos[b * N + i];
    if (p <= m) continue;  // already pivoted (includes the pivot itself)
    T s = (T)0;
    int t = 0;
    for (; t + 3 < m; t += 4) {
      const T l0 = Lb[(int64_t)t * N + i], l1 = Lb[(int64_t)(t + 1) * N + i], l2 = Lb[(int64_t)(t + 2) * N + i],
              l3 = Lb[(int64_t)(t + 3) * N + i];
      if constexpr (sizeof(T) == 4) {
        s = __fadd_rn(s, __fmul_rn(u[t], l0));
        s = __fadd_rn(s, __fmul_rn(u[t + 1], l1));
        s = __fadd_rn(s, __fmul_rn(u[t + 2], l2));
        s = __fadd_rn(s, __fmul_rn(u[t + 3], l3));
      } else {
        s = __dadd_rn(s, __dmul_rn(u[t], l0));
        s = __dadd_rn(s, __dmul_rn(u[t + 1], l1));
        s = __dadd_rn(s, __dmul_rn(u[t + 2], l2));
        s = __dadd_rn(s, __dmul_rn(u[t + 3], l3));
      }
    }
    for (; t < m; ++t) {
      const T l0 = Lb[(int64_t)t * N + i];
      if constexpr (sizeof(T) == 4) s = __fadd_rn(s, __fmul_rn(u[t], l0));
      else s = __dadd_rn(s, __dmul_rn(u[t], l0));
    }
    T v = src.entry(b, r, i);
    if (m > 0) v = v - s;
    v = v / pv;
    Lb[(int64_t)m * N + i] = v;
    T sq;
    if constexpr (sizeof(T) == 4) sq = __fmul_rn(v, v);
    else sq = __dmul_rn(v, v);
    const T dn = diag[b * N + i] - sq;
    diag[b * N + i] = dn;
    es += fabs((double)dn);
    if (cand_better<T>(dn, p, bv, bp)) {
      bv = dn; bp = p; bi = (int)i;
    }
  }
  block_argmax<T>(bv, bp, bi, sv, sp, si);
  const double tot = block_sum(es, scratch);
  if (threadIdx.x == 0) {
    cval[b * nchunks + chunk] = bv;
    cpos[b * nchunks + chunk] = bp;
    cidx[b * nchunks + chunk] = bi;
    errsum[b * nchunks + chunk] = tot;
  }
}

// fp32 variant with four consecutive columns per thread (N % 4 == 0): every load of a previous row of L is a 16-byte
// word and a CTA touches 4 KB of it at a time instead of 1 KB -- the update reads m rows of N floats per step, each
// 4 N bytes apart, so longer contiguous runs are what the DRAM pages want.  Arithmetic and its order per column are
// those of k_pc_update (one rounding per multiply and per add), candidates are merged in the same order.
template <typename Src>
__global__ void __launch_bounds__(256)
k_pc_update_v4(Src src, int64_t N, int rankmax, int m, int nchunks, int64_t cpc, float* __restrict__ diag,
               const int* __restrict__ pos, const int* __restrict__ pi, const float* __restrict__ piv,
               float* __restrict__ Lt, float* __restrict__ cval, int* __restrict__ cpos, int* __restrict__ cidx,
               double* __restrict__ errsum, const PcControl* ctrl) {
  extern __shared__ unsigned char smem_raw[];
  float* u = reinterpret_cast<float*>(smem_raw);
  __shared__ float sv[32];
  __shared__ int sp[32], si[32];
  __shared__ double scratch[32];
  if (!ctrl->active || ctrl->m != m + 1) return;
  const int64_t b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int64_t i0 = (int64_t)chunk * cpc, i1 = min(i0 + cpc, N);  // cpc % 4 == 0
  const int64_t r = pi[b];
  const float pv = piv[b];
  float* Lb = Lt + b * rankmax * N;
  for (int t = threadIdx.x; t < m; t += blockDim.x) u[t] = Lb[(int64_t)t * N + r];
  __syncthreads();
  float bv = 0.f;
  int bp = -1, bi = -1;
  double es = 0.0;
  for (int64_t i = i0 + 4 * (int64_t)threadIdx.x; i < i1; i += 4 * (int64_t)blockDim.x) {
    const int4 p4 = *reinterpret_cast<const int4*>(pos + b * N + i);
    const int pp[4] = {p4.x, p4.y, p4.z, p4.w};
    if (pp[0] <= m && pp[1] <= m && pp[2] <= m && pp[3] <= m) continue;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    int t = 0;
    for (; t + 7 < m; t += 8) {  // eight rows of L requested before the first is used (same order of the additions)
      float4 l[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) l[q] = *reinterpret_cast<const float4*>(Lb + (int64_t)(t + q) * N + i);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float uq = u[t + q];
        s[0] = __fadd_rn(s[0], __fmul_rn(uq, l[q].x));
        s[1] = __fadd_rn(s[1], __fmul_rn(uq, l[q].y));
        s[2] = __fadd_rn(s[2], __fmul_rn(uq, l[q].z));
        s[3] = __fadd_rn(s[3], __fmul_rn(uq, l[q].w));
      }
    }
    for (; t + 3 < m; t += 4) {
      float4 l[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) l[q] = *reinterpret_cast<const float4*>(Lb + (int64_t)(t + q) * N + i);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float uq = u[t + q];
        s[0] = __fadd_rn(s[0], __fmul_rn(uq, l[q].x));
        s[1] = __fadd_rn(s[1], __fmul_rn(uq, l[q].y));
        s[2] = __fadd_rn(s[2], __fmul_rn(uq, l[q].z));
        s[3] = __fadd_rn(s[3], __fmul_rn(uq, l[q].w));
      }
    }
    for (; t < m; ++t) {
      const float4 l0 = *reinterpret_cast<const float4*>(Lb + (int64_t)t * N + i);
      const float uq = u[t];
      s[0] = __fadd_rn(s[0], __fmul_rn(uq, l0.x));
      s[1] = __fadd_rn(s[1], __fmul_rn(uq, l0.y));
      s[2] = __fadd_rn(s[2], __fmul_rn(uq, l0.z));
      s[3] = __fadd_rn(s[3], __fmul_rn(uq, l0.w));
    }
    float4 d4 = *reinterpret_cast<const float4*>(diag + b * N + i);
    float dd[4] = {d4.x, d4.y, d4.z, d4.w};
    float out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      out[q] = 0.f;
      if (pp[q] <= m) {  // already pivoted: leave L and diag as they are
        out[q] = Lb[(int64_t)m * N + i + q];
        continue;
      }
      float v = src.entry(b, r, i + q);
      if (m > 0) v = v - s[q];
      v = v / pv;
      out[q] = v;
      const float dn = dd[q] - __fmul_rn(v, v);
      dd[q] = dn;
      es += fabs((double)dn);
      if (cand_better<float>(dn, pp[q], bv, bp)) {
        bv = dn; bp = pp[q]; bi = (int)(i + q);
      }
    }
    *reinterpret_cast<float4*>(Lb + (int64_t)m * N + i) = make_float4(out[0], out[1], out[2], out[3]);
    *reinterpret_cast<float4*>(diag + b * N + i) = make_float4(dd[0], dd[1], dd[2], dd[3]);
  }
  block_argmax<float>(bv, bp, bi, sv, sp, si);
  const double tot = block_sum(es, scratch);
  if (threadIdx.x == 0) {
    cval[b * nchunks + chunk] = bv;
    cpos[b * nchunks + chunk] = bp;
    cidx[b * nchunks + chunk] = bi;
    errsum[b * nchunks + chunk] = tot;
  }
}

__global__ void k_pc_finish(const PcControl* ctrl, int32_t* m_out) { *m_out = ctrl->m; }
__global__ void k_pc_status(const PcControl* ctrl, int32_t* m_out, int32_t* active_out) {
  *m_out = ctrl->m;
  if (active_out) *active_out = ctrl->active;
}

template <typename T, typename Src>
static int run_pivchol(Src src, int64_t B, int64_t N, int rank, double tol, T* Lt, int64_t* perm, int32_t* m_out,
                       void* ws, cudaStream_t st) {
  LOB_REQUIRE(B > 0 && N > 0 && rank > 0, "lob_pivchol: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_pivchol: flattened batch > 65535 not supported");
  LOB_REQUIRE(N < (1LL << 31), "lob_pivchol: N must fit in int32");
  LOB_REQUIRE(Lt && perm && m_out && ws, "lob_pivchol: NULL pointer");
  const int rankmax = (int)(rank < N ? rank : N);
  PcLayout L = pc_layout(B, N, rankmax, sizeof(T));
  char* w = (char*)ws;
  PcControl* ctrl = (PcControl*)(w + L.off_ctrl);
  T* diag = (T*)(w + L.off_diag);
  int* pos = (int*)(w + L.off_pos);
  T* cval = (T*)(w + L.off_cand_val);
  int* cpos = (int*)(w + L.off_cand_pos);
  int* cidx = (int*)(w + L.off_cand_idx);
  double* errsum = (double*)(w + L.off_errsum);
  int* pi = (int*)(w + L.off_pi);
  T* piv = (T*)(w + L.off_piv);
  T* orig = (T*)(w + L.off_orig);
  dim3 grid((unsigned)L.nchunks, (unsigned)B);
  k_pc_init<T, Src><<<grid, 256, 0, st>>>(src, N, L.nchunks, L.cols_per_chunk, diag, pos, perm, cval, cpos, cidx,
                                          errsum, ctrl);
  LOB_TRY(check_launch("k_pc_init"));
  for (int m = 0; m < rankmax; ++m) {
    k_pc_pivot<T><<<1, 1024, 0, st>>>(B, N, rankmax, m, L.nchunks, tol, cval, cpos, cidx, errsum, orig, pos, perm, pi,
                                      piv, Lt, ctrl);
    LOB_TRY(check_launch("k_pc_pivot"));
    if (m + 1 < N) {  // :77
      bool done = false;
      if constexpr (sizeof(T) == 4) {
        if ((N % 4) == 0 && (L.cols_per_chunk % 4) == 0 && (reinterpret_cast<uintptr_t>(Lt) & 15) == 0 &&
            (reinterpret_cast<uintptr_t>(diag) & 15) == 0 && (reinterpret_cast<uintptr_t>(pos) & 15) == 0) {
          k_pc_update_v4<Src><<<grid, 256, sizeof(float) * (size_t)rankmax, st>>>(
              src, N, rankmax, m, L.nchunks, L.cols_per_chunk, (float*)diag, pos, pi, (const float*)piv, (float*)Lt,
              (float*)cval, cpos, cidx, errsum, ctrl);
          LOB_TRY(check_launch("k_pc_update_v4"));
          done = true;
        }
      }
      if (!done) {
        k_pc_update<T, Src><<<grid, 256, sizeof(T) * (size_t)(rankmax > 0 ? rankmax : 1), st>>>(
            src, N, rankmax, m, L.nchunks, L.cols_per_chunk, diag, pos, pi, piv, Lt, cval, cpos, cidx, errsum, ctrl);
        LOB_TRY(check_launch("k_pc_update"));
      }
    }
  }
  k_pc_finish<<<1, 1, 0, st>>>(ctrl, m_out);
  return check_launch("k_pc_finish");
}

}  // namespace lob

using namespace lob;

extern "C" size_t lob_pivchol_workspace_bytes(int64_t B, int64_t N, int32_t rank) {
  if (B <= 0 || N <= 0 || rank <= 0) return 0;
  return pc_layout(B, N, rank, 8).total;
}

extern "C" int lob_pivchol_dense(int32_t dtype, int64_t B, int64_t N, int32_t rank, double error_tol, const void* A,
                                 int64_t lda, int64_t a_batch_stride, void* Lt, int64_t* perm, int32_t* m_out,
                                 void* ws, void* stream) {
  LOB_REQUIRE(A != nullptr, "lob_pivchol_dense: A is NULL");
  LOB_DISPATCH_DTYPE(dtype, {
    DenseSrc<scalar_t> src{(const scalar_t*)A, lda, a_batch_stride};
    return run_pivchol<scalar_t>(src, B, N, rank, error_tol, (scalar_t*)Lt, perm, m_out, ws, (cudaStream_t)stream);
  });
}

extern "C" int lob_pivchol_kron(int32_t dtype, int64_t B, int32_t n_factors, const int64_t* host_sizes,
                                const void* const* host_factors, const int64_t* host_batch_strides, int32_t rank,
                                double error_tol, void* Lt, int64_t* perm, int32_t* m_out, void* ws, void* stream) {
  LOB_REQUIRE(n_factors >= 1 && n_factors <= 4, "lob_pivchol_kron: 1..4 factors supported");
  LOB_REQUIRE(host_sizes && host_factors && host_batch_strides, "lob_pivchol_kron: NULL pointer");
  int64_t N = 1;
  for (int q = 0; q < n_factors; ++q) N *= host_sizes[q];
  LOB_DISPATCH_DTYPE(dtype, {
    KronSrc<scalar_t> src;
    src.nf = n_factors;
    int64_t stride = N;
    for (int q = 0; q < 4; ++q) {
      if (q < n_factors) {
        stride /= host_sizes[q];
        src.f[q] = (const scalar_t*)host_factors[q];
        src.n[q] = host_sizes[q];
        src.bs[q] = host_batch_strides[q];
        src.stride[q] = stride;
      } else {
        src.f[q] = nullptr;
        src.n[q] = 1;
        src.bs[q] = 0;
        src.stride[q] = 1;
      }
    }
    return run_pivchol<scalar_t>(src, B, N, rank, error_tol, (scalar_t*)Lt, perm, m_out, ws, (cudaStream_t)stream);
  });
}

extern "C" int lob_pivchol_toeplitz(int32_t dtype, int64_t B, int64_t N, const void* col, int64_t col_batch_stride,
                                    int32_t rank, double error_tol, void* Lt, int64_t* perm, int32_t* m_out, void* ws,
                                    void* stream) {
  LOB_REQUIRE(col != nullptr, "lob_pivchol_toeplitz: col is NULL");
  LOB_DISPATCH_DTYPE(dtype, {
    ToeplitzSrc<scalar_t> src{(const scalar_t*)col, col_batch_stride};
    return run_pivchol<scalar_t>(src, B, N, rank, error_tol, (scalar_t*)Lt, perm, m_out, ws, (cudaStream_t)stream);
  });
}

// ---- generic operators: the pivot row of every step comes from the operator's own `_get_indices`, called by the host
// with the device-resident pivot indices (functions/_pivoted_cholesky.py:57-98 + utils/permutation.py:9-88).  Same
// kernels, same index semantics; no host read per step (the indices never leave the device).
namespace lob {
template <typename T>
static int pc_rows_step(int phase, int64_t B, int64_t N, int rank, int m, double tol, const T* buf, T* Lt,
                        int64_t* perm, int64_t* pi64, int32_t* m_out, int32_t* active_out, void* ws, cudaStream_t st) {
  LOB_REQUIRE(B > 0 && N > 0 && rank > 0, "lob_pivchol_rows: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_pivchol_rows: flattened batch > 65535 not supported");
  LOB_REQUIRE(N < (1LL << 31), "lob_pivchol_rows: N must fit in int32");
  LOB_REQUIRE(ws != nullptr, "lob_pivchol_rows: NULL workspace");
  const int rankmax = (int)(rank < N ? rank : N);
  PcLayout L = pc_layout(B, N, rankmax, sizeof(T));
  char* w = (char*)ws;
  PcControl* ctrl = (PcControl*)(w + L.off_ctrl);
  T* diag = (T*)(w + L.off_diag);
  int* pos = (int*)(w + L.off_pos);
  T* cval = (T*)(w + L.off_cand_val);
  int* cpos = (int*)(w + L.off_cand_pos);
  int* cidx = (int*)(w + L.off_cand_idx);
  double* errsum = (double*)(w + L.off_errsum);
  int* pi = (int*)(w + L.off_pi);
  T* piv = (T*)(w + L.off_piv);
  T* orig = (T*)(w + L.off_orig);
  dim3 grid((unsigned)L.nchunks, (unsigned)B);
  RowBufSrc<T> src{buf, N};
  switch (phase) {
    case 0:
      LOB_REQUIRE(buf && perm, "lob_pivchol_rows_begin: NULL pointer");
      k_pc_init<T, RowBufSrc<T>><<<grid, 256, 0, st>>>(src, N, L.nchunks, L.cols_per_chunk, diag, pos, perm, cval,
                                                       cpos, cidx, errsum, ctrl);
      return check_launch("k_pc_init");
    case 1:
      LOB_REQUIRE(Lt && perm && pi64, "lob_pivchol_rows_pivot: NULL pointer");
      LOB_REQUIRE(m >= 0 && m < rankmax, "lob_pivchol_rows_pivot: step out of range");
      k_pc_pivot<T><<<1, 1024, 0, st>>>(B, N, rankmax, m, L.nchunks, tol, cval, cpos, cidx, errsum, orig, pos, perm,
                                        pi, piv, Lt, ctrl, pi64);
      return check_launch("k_pc_pivot");
    case 2:
      LOB_REQUIRE(buf && Lt, "lob_pivchol_rows_update: NULL pointer");
      LOB_REQUIRE(m >= 0 && m < rankmax, "lob_pivchol_rows_update: step out of range");
      if (m + 1 < N) {
        k_pc_update<T, RowBufSrc<T>><<<grid, 256, sizeof(T) * (size_t)rankmax, st>>>(
            src, N, rankmax, m, L.nchunks, L.cols_per_chunk, diag, pos, pi, piv, Lt, cval, cpos, cidx, errsum, ctrl);
        return check_launch("k_pc_update");
      }
      return LOB_OK;
    default:
      LOB_REQUIRE(m_out, "lob_pivchol_rows_status: NULL pointer");
      k_pc_status<<<1, 1, 0, st>>>(ctrl, m_out, active_out);
      return check_launch("k_pc_status");
  }
}
}  // namespace lob

extern "C" int lob_pivchol_rows_begin(int32_t dtype, int64_t B, int64_t N, int32_t rank, const void* diag,
                                      int64_t* perm, void* ws, void* stream) {
  LOB_DISPATCH_DTYPE(dtype, {
    return pc_rows_step<scalar_t>(0, B, N, rank, 0, 0.0, (const scalar_t*)diag, nullptr, perm, nullptr, nullptr,
                                  nullptr, ws, (cudaStream_t)stream);
  });
}
extern "C" int lob_pivchol_rows_pivot(int32_t dtype, int64_t B, int64_t N, int32_t rank, int32_t m, double error_tol,
                                      void* Lt, int64_t* perm, int64_t* pivot_rows, void* ws, void* stream) {
  LOB_DISPATCH_DTYPE(dtype, {
    return pc_rows_step<scalar_t>(1, B, N, rank, m, error_tol, nullptr, (scalar_t*)Lt, perm, pivot_rows, nullptr,
                                  nullptr, ws, (cudaStream_t)stream);
  });
}
extern "C" int lob_pivchol_rows_update(int32_t dtype, int64_t B, int64_t N, int32_t rank, int32_t m, const void* rows,
                                       void* Lt, void* ws, void* stream) {
  LOB_DISPATCH_DTYPE(dtype, {
    return pc_rows_step<scalar_t>(2, B, N, rank, m, 0.0, (const scalar_t*)rows, (scalar_t*)Lt, nullptr, nullptr,
                                  nullptr, nullptr, ws, (cudaStream_t)stream);
  });
}
extern "C" int lob_pivchol_rows_status(int32_t dtype, int64_t B, int64_t N, int32_t rank, int32_t* m_out,
                                       int32_t* active_out, void* ws, void* stream) {
  LOB_DISPATCH_DTYPE(dtype, {
    return pc_rows_step<scalar_t>(3, B, N, rank, 0, 0.0, nullptr, nullptr, nullptr, nullptr, m_out, active_out, ws,
                                  (cudaStream_t)stream);
  });
}

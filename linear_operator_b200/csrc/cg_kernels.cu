// cg_kernels.cu -- fused per-iteration updates of modified batched CG (reference: utils/linear_cg.py:98-359).
//
// Layout: every vector is (B, N, C) row-major, C fastest.  A column reduction is therefore a strided reduction
// over N.  One CTA owns a contiguous slab of rows of one batch element ("chunk"); thread (tx, ty) owns column
// c0+tx and rows ty, ty+RY, ... so a warp always touches consecutive addresses (fully coalesced 4/8-byte
// accesses).  Cross-chunk reductions are two-stage and deterministic: each CTA writes its per-column partial sum
// (double) to (B, nchunks, C); every consumer CTA re-adds the nchunks partials in a fixed order.  No atomics on the
// data path, no host synchronisation: the stop test / tridiagonal bookkeeping of the reference runs in a one-CTA
// control kernel that sets device-side control words, and every kernel early-exits once `stop` is set.
#include <math.h>

#include <algorithm>
#include <initializer_list>

#include "common.cuh"

namespace lob {

constexpr int kFoldThreshold = 64;  // caller-supplied partial lists longer than this are folded first
constexpr int kFoldParts = 16;      // ... to this many partials per column (one CTA each)

struct CgLayout {
  int64_t B, N, C;
  int nchunks;
  int64_t rows_per_chunk;
  int cx, ry;
  // byte offsets into the workspace
  size_t off_status, off_rhs_norm, off_rhs_zero, off_conv, off_alpha, off_beta, off_rz, off_resid, off_prev_ar,
      off_prev_beta, off_parts_a, off_parts_rr, off_parts_rz, off_fold, total;
};

static CgLayout cg_layout(const lob_cg_params* p) {
  CgLayout L;
  L.B = p->B;
  L.N = p->N;
  L.C = p->C;
  L.cx = (int)(p->C < 128 ? p->C : 128);
  L.ry = 256 / L.cx;
  if (L.ry < 1) L.ry = 1;
  // Row chunks per batch element: the vector kernels keep 5 CTAs per SM resident (registers), so the grid B * nch is
  // chosen to fill whole waves of 5 * 148 CTAs (a 1.7-wave grid ran its second wave 70 % full; small batches with the
  // former cap of 64 chunks left a third of the slots empty).  Consumers re-add the nch partials per column in a fixed
  // order, so nch stays <= 256.
  const int64_t slots = (int64_t)kNumSMs * 5;
  int64_t maxch = cdiv(p->N, (int64_t)L.ry * 4);
  if (maxch > 256) maxch = 256;
  int64_t nch = 1;
  double best_eff = -1.0;
  for (int64_t c = std::max<int64_t>(1, cdiv(2 * slots, p->B)); c <= maxch; ++c) {
    const double waves = (double)(p->B * c) / (double)slots;
    const double eff = waves / ceil(waves);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      nch = c;
    }
    if (eff >= 0.97 || c >= 8 * cdiv(2 * slots, p->B)) break;
  }
  if (best_eff < 0) nch = maxch;  // fewer rows than two waves' worth: as many chunks as the rows allow
  L.rows_per_chunk = (cdiv(p->N, nch) + 3) / 4 * 4;  // whole groups of four rows: the float4 kernels' unit
  L.nchunks = (int)cdiv(p->N, L.rows_per_chunk);
  const size_t bc = (size_t)p->B * p->C;
  const size_t bs = (size_t)p->B * (p->n_tridiag > 0 ? p->n_tridiag : 1);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  L.off_status = take(sizeof(lob_cg_status));
  L.off_rhs_norm = take(bc * 8);
  L.off_rhs_zero = take(bc);
  L.off_conv = take(bc);
  L.off_alpha = take(bc * 8);
  L.off_beta = take(bc * 8);
  L.off_rz = take(2 * bc * 8);
  L.off_resid = take(bc * 8);
  L.off_prev_ar = take(bs * 8);
  L.off_prev_beta = take(bs * 8);
  L.off_parts_a = take(bc * L.nchunks * 8);
  L.off_parts_rr = take(bc * L.nchunks * 8);
  L.off_parts_rz = take(bc * L.nchunks * 8);
  L.off_fold = take(2 * bc * kFoldParts * 8);
  L.total = o;
  return L;
}

struct CgPtrs {
  lob_cg_status* status;
  double* rhs_norm;
  uint8_t* rhs_zero;
  uint8_t* conv;
  double* alpha;
  double* beta;
  double* rz;  // [2][B*C]
  double* resid;
  double* prev_ar;
  double* prev_beta;
  double* parts_a;
  double* parts_rr;
  double* parts_rz;
  double* fold;  // [2][B*C]: caller-supplied partial sums folded to one value per column (k_fold_parts)
};

static CgPtrs cg_ptrs(const CgLayout& L, void* ws) {
  char* b = (char*)ws;
  CgPtrs P;
  P.status = (lob_cg_status*)(b + L.off_status);
  P.rhs_norm = (double*)(b + L.off_rhs_norm);
  P.rhs_zero = (uint8_t*)(b + L.off_rhs_zero);
  P.conv = (uint8_t*)(b + L.off_conv);
  P.alpha = (double*)(b + L.off_alpha);
  P.beta = (double*)(b + L.off_beta);
  P.rz = (double*)(b + L.off_rz);
  P.resid = (double*)(b + L.off_resid);
  P.prev_ar = (double*)(b + L.off_prev_ar);
  P.prev_beta = (double*)(b + L.off_prev_beta);
  P.parts_a = (double*)(b + L.off_parts_a);
  P.parts_rr = (double*)(b + L.off_parts_rr);
  P.parts_rz = (double*)(b + L.off_parts_rz);
  P.fold = (double*)(b + L.off_fold);
  return P;
}

struct Dims {
  int64_t N, C;
  int nchunks;
  int64_t rows_per_chunk;
};

// per-column reduction over ty; result valid for ty == 0
__device__ __forceinline__ double reduce_over_ty(double v, double* red) {
  const int tx = threadIdx.x, ty = threadIdx.y, cx = blockDim.x, ry = blockDim.y;
  __syncthreads();
  red[ty * cx + tx] = v;
  __syncthreads();
  double s = 0.0;
  if (ty == 0)
    for (int i = 0; i < ry; ++i) s += red[i * cx + tx];
  return s;
}

__device__ __forceinline__ double sum_parts(const double* parts, int64_t b, int nparts, int64_t C, int64_t c) {
  const double* q = parts + (b * nparts) * C + c;
  double s = 0.0;
  for (int i = 0; i < nparts; ++i) s += q[(int64_t)i * C];
  return s;
}

#define CG_PROLOGUE                                                              \
  extern __shared__ double red[];                                                \
  const int64_t b = blockIdx.y;                                                  \
  const int chunk = blockIdx.x;                                                  \
  const int64_t row0 = (int64_t)chunk * d.rows_per_chunk;                        \
  const int64_t row1 = min(row0 + d.rows_per_chunk, d.N);                        \
  const int tx = threadIdx.x, ty = threadIdx.y, RY = blockDim.y, CX = blockDim.x; \
  const int64_t base = b * d.N * d.C;                                            \
  (void)red; (void)row0; (void)row1; (void)tx; (void)ty; (void)RY; (void)CX; (void)base;

// ---- out[b,c] = sum_i parts[b,i,c]: folds long lists of caller-supplied partials (the fused epilogues of the matmul
// kernels emit one per 128 operator rows: 7813 at N = 10^6) once, instead of in the prologue of every consumer CTA ----
__global__ void k_fold_parts(const double* __restrict__ parts, int n_parts, int C, double* __restrict__ out,
                             const lob_cg_status* status) {
  if (status && status->stop) return;
  extern __shared__ double red[];
  const int64_t b = blockIdx.y;
  const int slice = blockIdx.x;  // of kFoldParts
  const int per = (n_parts + kFoldParts - 1) / kFoldParts;
  const int i0 = slice * per, i1 = min(n_parts, i0 + per);
  const int tx = threadIdx.x, ty = threadIdx.y, RY = blockDim.y, CX = blockDim.x;
  for (int c0 = 0; c0 < C; c0 += CX) {
    const int c = c0 + tx;
    double acc = 0.0;
    if (c < C)
      for (int i = i0 + ty; i < i1; i += RY) acc += parts[(b * n_parts + i) * C + c];
    const double s = reduce_over_ty(acc, red);
    if (ty == 0 && c < C) out[(b * kFoldParts + slice) * C + c] = s;
  }
}

// ---- generic: parts[b,chunk,c] = sum_rows u*v -------------------------------------------------------------
template <typename T>
__global__ void k_dots_partials(Dims d, const T* __restrict__ u, const T* __restrict__ v, double* __restrict__ parts,
                                const lob_cg_status* status) {
  if (status && status->stop) return;
  CG_PROLOGUE
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    double acc = 0.0;
    if (c < d.C) {
      int64_t row = row0 + ty;
      for (; row + 3 * RY < row1; row += 4 * RY) {
        const int64_t i0 = base + row * d.C + c, st = (int64_t)RY * d.C;
        T u0 = u[i0], u1 = u[i0 + st], u2 = u[i0 + 2 * st], u3 = u[i0 + 3 * st];
        T v0 = v[i0], v1 = v[i0 + st], v2 = v[i0 + 2 * st], v3 = v[i0 + 3 * st];
        acc += (double)u0 * (double)v0;
        acc += (double)u1 * (double)v1;
        acc += (double)u2 * (double)v2;
        acc += (double)u3 * (double)v3;
      }
      for (; row < row1; row += RY) {
        const int64_t i0 = base + row * d.C + c;
        acc += (double)u[i0] * (double)v[i0];
      }
    }
    double s = reduce_over_ty(acc, red);
    if (ty == 0 && c < d.C) parts[(b * d.nchunks + chunk) * d.C + c] = s;
  }
}

// ---- setup: normalise rhs / x0 (linear_cg.py:177-183) -----------------------------------------------------
template <typename T>
__global__ void k_setup_normalize(Dims d, const T* __restrict__ rhs, const T* __restrict__ x0, T* __restrict__ rhs_n,
                                  T* __restrict__ x, const double* __restrict__ parts, double* __restrict__ rhs_norm,
                                  uint8_t* __restrict__ rhs_zero, double eps) {
  CG_PROLOGUE
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    if (c >= d.C) continue;
    T nrm = (T)sqrt(sum_parts(parts, b, d.nchunks, d.C, c));
    const bool zero = nrm < (T)eps;
    if (zero) nrm = (T)1;
    if (chunk == 0 && ty == 0) {
      rhs_norm[b * d.C + c] = (double)nrm;
      rhs_zero[b * d.C + c] = zero ? 1 : 0;
    }
    for (int64_t row = row0 + ty; row < row1; row += RY) {
      const int64_t i0 = base + row * d.C + c;
      rhs_n[i0] = rhs[i0] / nrm;
      x[i0] = x0 ? (T)(x0[i0] / nrm) : (T)0;
    }
  }
}

// ---- r = rhs_n - A x0, partial <r,r>, NaN check (linear_cg.py:186,199,204) --------------------------------
template <typename T>
__global__ void k_residual_init(Dims d, const T* __restrict__ rhs_n, const T* __restrict__ ax0, T* __restrict__ r,
                                double* __restrict__ parts_rr, lob_cg_status* status) {
  CG_PROLOGUE
  bool saw_nan = false;
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    double acc = 0.0;
    if (c < d.C) {
      for (int64_t row = row0 + ty; row < row1; row += RY) {
        const int64_t i0 = base + row * d.C + c;
        T v = rhs_n[i0];
        if (ax0) v = v - ax0[i0];
        r[i0] = v;
        saw_nan |= (v != v);
        acc += (double)v * (double)v;
      }
    }
    double s = reduce_over_ty(acc, red);
    if (ty == 0 && c < d.C) parts_rr[(b * d.nchunks + chunk) * d.C + c] = s;
  }
  if (saw_nan) atomicOr(&status->nan_detected, 1);
}

// one CTA: residual norms, has_converged, "skip the iteration" test (linear_cg.py:204-208)
__global__ void k_residual_flags(int64_t BC, int64_t C, int nchunks, const double* __restrict__ parts_rr,
                                 double* __restrict__ resid, uint8_t* __restrict__ conv, lob_cg_status* status,
                                 double sua, int n_tridiag, int is_f32) {
  __shared__ int not_conv;
  if (threadIdx.x == 0) not_conv = 0;
  __syncthreads();
  int local = 0;
  for (int64_t i = threadIdx.x; i < BC; i += blockDim.x) {
    const int64_t b = i / C, c = i % C;
    double nrm = sqrt(sum_parts(parts_rr, b, nchunks, C, c));
    if (is_f32) nrm = (double)(float)nrm;
    resid[i] = nrm;
    const bool cv = is_f32 ? ((float)nrm < (float)sua) : (nrm < sua);
    conv[i] = cv ? 1 : 0;
    if (!cv) local = 1;
  }
  if (local) atomicOr(&not_conv, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int all_conv = not_conv ? 0 : 1;
    status->all_converged_at_start = all_conv;
    if (status->nan_detected) status->stop = 1;
    if (all_conv && !n_tridiag) {
      status->stop = 1;
      status->iterations = 0;
    }
  }
}

// ---- rz0 = sum z*r (from partials), p = z (linear_cg.py:213-215) ------------------------------------------
template <typename T>
__global__ void k_direction_init(Dims d, const T* __restrict__ z, T* __restrict__ pvec,
                                 const double* __restrict__ parts_rz, int n_rz_parts, double* __restrict__ rz0,
                                 int is_f32) {
  CG_PROLOGUE
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    if (c >= d.C) continue;
    if (chunk == 0 && ty == 0) {
      double s = sum_parts(parts_rz, b, n_rz_parts, d.C, c);
      rz0[b * d.C + c] = is_f32 ? (double)(float)s : s;
    }
    for (int64_t row = row0 + ty; row < row1; row += RY) {
      const int64_t i0 = base + row * d.C + c;
      pvec[i0] = z[i0];
    }
  }
}

// ---- alpha, r -= alpha Ap, x += alpha p, partial <r,r> (linear_cg.py:250-264 + :31) -----------------------
template <typename T>
__global__ void k_step_xr(Dims d, const T* __restrict__ ap, const T* __restrict__ pvec, T* __restrict__ x,
                          T* __restrict__ r, const double* __restrict__ pap_parts, int n_pap_parts,
                          const double* __restrict__ rz, const uint8_t* __restrict__ conv, double* __restrict__ alpha_out,
                          double* __restrict__ parts_rr, lob_cg_status* status, double eps, int check_nan) {
  if (status->stop) return;
  CG_PROLOGUE
  bool saw_nan = false;
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    double acc = 0.0;
    if (c < d.C) {
      // alpha = rz / <p,Ap>; denominator < eps -> 0; converged columns -> 0
      T den = (T)sum_parts(pap_parts, b, n_pap_parts, d.C, c);
      const bool is_zero = den < (T)eps;
      T a = is_zero ? (T)0 : (T)((T)rz[b * d.C + c] / den);
      if (conv[b * d.C + c]) a = (T)0;
      if (chunk == 0 && ty == 0) alpha_out[b * d.C + c] = (double)a;
      int64_t row = row0 + ty;
      const int64_t st = (int64_t)RY * d.C;
      for (; row + 3 * RY < row1; row += 4 * RY) {
        const int64_t i0 = base + row * d.C + c;
        T q0 = ap[i0], q1 = ap[i0 + st], q2 = ap[i0 + 2 * st], q3 = ap[i0 + 3 * st];
        T r0 = r[i0], r1 = r[i0 + st], r2 = r[i0 + 2 * st], r3 = r[i0 + 3 * st];
        T p0 = pvec[i0], p1 = pvec[i0 + st], p2 = pvec[i0 + 2 * st], p3 = pvec[i0 + 3 * st];
        T x0 = x[i0], x1 = x[i0 + st], x2 = x[i0 + 2 * st], x3 = x[i0 + 3 * st];
        if (check_nan) saw_nan |= (q0 != q0) | (q1 != q1) | (q2 != q2) | (q3 != q3);
        r0 -= a * q0; r1 -= a * q1; r2 -= a * q2; r3 -= a * q3;
        x0 += a * p0; x1 += a * p1; x2 += a * p2; x3 += a * p3;
        r[i0] = r0; r[i0 + st] = r1; r[i0 + 2 * st] = r2; r[i0 + 3 * st] = r3;
        x[i0] = x0; x[i0 + st] = x1; x[i0 + 2 * st] = x2; x[i0 + 3 * st] = x3;
        acc += (double)r0 * (double)r0;
        acc += (double)r1 * (double)r1;
        acc += (double)r2 * (double)r2;
        acc += (double)r3 * (double)r3;
      }
      for (; row < row1; row += RY) {
        const int64_t i0 = base + row * d.C + c;
        T q0 = ap[i0];
        if (check_nan) saw_nan |= (q0 != q0);
        T r0 = r[i0] - a * q0;
        r[i0] = r0;
        x[i0] = x[i0] + a * pvec[i0];
        acc += (double)r0 * (double)r0;
      }
    }
    double s = reduce_over_ty(acc, red);
    if (ty == 0 && c < d.C) parts_rr[(b * d.nchunks + chunk) * d.C + c] = s;
  }
  if (saw_nan) atomicOr(&status->nan_detected, 1);
}

// ---- beta, p = z + beta p, residual norm, converged flags (linear_cg.py:31-46, :298-300) -------------------
template <typename T>
__global__ void k_step_p(Dims d, const T* __restrict__ z, T* __restrict__ pvec, const double* __restrict__ parts_rz,
                         int n_rz_parts, const double* __restrict__ parts_rr, const double* __restrict__ rz_old,
                         double* __restrict__ rz_new, double* __restrict__ beta_out, double* __restrict__ resid,
                         uint8_t* __restrict__ conv, const uint8_t* __restrict__ rhs_zero,
                         const lob_cg_status* status, double eps, double sua) {
  if (status->stop) return;
  CG_PROLOGUE
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    if (c >= d.C) continue;
    const T rzn = (T)sum_parts(parts_rz, b, n_rz_parts, d.C, c);
    const T rzo = (T)rz_old[b * d.C + c];
    const bool is_zero = rzo < (T)eps;
    const T be = is_zero ? (T)0 : (T)(rzn / rzo);
    if (chunk == 0 && ty == 0) {
      const int64_t i = b * d.C + c;
      rz_new[i] = (double)rzn;
      beta_out[i] = (double)be;
      T nrm = (T)sqrt(sum_parts(parts_rr, b, d.nchunks, d.C, c));
      if (rhs_zero[i]) nrm = (T)0;
      resid[i] = (double)nrm;
      conv[i] = (nrm < (T)sua) ? 1 : 0;
    }
    int64_t row = row0 + ty;
    const int64_t st = (int64_t)RY * d.C;
    for (; row + 3 * RY < row1; row += 4 * RY) {
      const int64_t i0 = base + row * d.C + c;
      T z0 = z[i0], z1 = z[i0 + st], z2 = z[i0 + 2 * st], z3 = z[i0 + 3 * st];
      T p0 = pvec[i0], p1 = pvec[i0 + st], p2 = pvec[i0 + 2 * st], p3 = pvec[i0 + 3 * st];
      pvec[i0] = p0 * be + z0;
      pvec[i0 + st] = p1 * be + z1;
      pvec[i0 + 2 * st] = p2 * be + z2;
      pvec[i0 + 3 * st] = p3 * be + z3;
    }
    for (; row < row1; row += RY) {
      const int64_t i0 = base + row * d.C + c;
      pvec[i0] = pvec[i0] * be + z[i0];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// float4 forms of the two per-iteration vector kernels (fp32, C <= 128, N * C % 4 == 0, 16-byte aligned vectors).
// A group of four rows is C consecutive float4s; thread (tx, ty) always takes float4 tx of the groups ty, ty + ry, ...:
// its four lanes are the FIXED columns (4 tx + i) mod C, so the per-column scalars sit in four registers and the
// per-column partial sums in four accumulators.  Four times the bytes in flight per load instruction of the scalar
// kernels, which ran at 4.4 - 4.8 TB/s where a plain float4 stream of the same shape reaches 6.5 (scripts/
// stream_width_bench.cu).  Same arithmetic per element as k_step_xr / k_step_p; the per-column sums are added in a
// different (still fixed) order.
// ---------------------------------------------------------------------------------------------------------------
struct V4Map {
  int c[4];   // column of lane i
  int dr[4];  // row of lane i inside the group of four
};
__device__ __forceinline__ V4Map v4_map(int tx, int C) {
  V4Map m;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = 4 * tx + i;
    m.dr[i] = f / C;
    m.c[i] = f - m.dr[i] * C;
  }
  return m;
}
// column sums of per-lane accumulators: red[(ty * C + tx) * 4 + i]; column c collects the lanes at flat positions
// c, c + C, c + 2 C, c + 3 C of the group, over all ty -- fixed order
__device__ __forceinline__ double v4_column_sum(const double* red, int c, int C, int ry) {
  double s = 0.0;
  for (int m = 0; m < 4; ++m) {
    const int f = c + m * C;
    for (int i = 0; i < ry; ++i) s += red[(i * C + (f >> 2)) * 4 + (f & 3)];
  }
  return s;
}

__global__ void k_step_xr_v4(Dims d, const float* __restrict__ ap, const float* __restrict__ pvec, float* __restrict__ x,
                             float* __restrict__ r, const double* __restrict__ pap_parts, int n_pap_parts,
                             const double* __restrict__ rz, const uint8_t* __restrict__ conv,
                             double* __restrict__ alpha_out, double* __restrict__ parts_rr, lob_cg_status* status,
                             double eps, int check_nan) {
  if (status->stop) return;
  extern __shared__ double red[];  // [ry][C][4] doubles, then C floats (alpha)
  const int64_t b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int C = (int)d.C, tx = threadIdx.x, ty = threadIdx.y, ry = blockDim.y;
  float* alpha_s = reinterpret_cast<float*>(red + (size_t)ry * C * 4);
  const int64_t row0 = (int64_t)chunk * d.rows_per_chunk;
  const int64_t nrows = min(row0 + d.rows_per_chunk, d.N) - row0;
  if (ty == 0) {
    // alpha = rz / <p,Ap>; denominator < eps -> 0; converged columns -> 0  (as k_step_xr)
    const float den = (float)sum_parts(pap_parts, b, n_pap_parts, C, tx);
    float a = (den < (float)eps) ? 0.f : (float)((float)rz[b * C + tx] / den);
    if (conv[b * C + tx]) a = 0.f;
    if (chunk == 0) alpha_out[b * C + tx] = (double)a;
    alpha_s[tx] = a;
  }
  __syncthreads();
  const V4Map m = v4_map(tx, C);
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = alpha_s[m.c[i]];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  bool saw_nan = false;
  const int64_t base = (b * d.N + row0) * C;  // multiple of 4 elements (launcher)
  const int64_t ngroups = (nrows + 3) / 4;
  const float4* ap4 = reinterpret_cast<const float4*>(ap + base);
  const float4* p4 = reinterpret_cast<const float4*>(pvec + base);
  float4* r4 = reinterpret_cast<float4*>(r + base);
  float4* x4 = reinterpret_cast<float4*>(x + base);
  const int64_t full = nrows / 4;  // groups whose four rows all exist
  int64_t g = ty;
  for (; g + ry < full; g += 2 * ry) {
    const int64_t i0 = g * C + tx, i1 = (g + ry) * C + tx;
    const float4 q0 = ap4[i0], q1 = ap4[i1], rr0 = r4[i0], rr1 = r4[i1], pp0 = p4[i0], pp1 = p4[i1], xx0 = x4[i0],
                 xx1 = x4[i1];
    float qv[2][4] = {{q0.x, q0.y, q0.z, q0.w}, {q1.x, q1.y, q1.z, q1.w}};
    float rv[2][4] = {{rr0.x, rr0.y, rr0.z, rr0.w}, {rr1.x, rr1.y, rr1.z, rr1.w}};
    float pv[2][4] = {{pp0.x, pp0.y, pp0.z, pp0.w}, {pp1.x, pp1.y, pp1.z, pp1.w}};
    float xv[2][4] = {{xx0.x, xx0.y, xx0.z, xx0.w}, {xx1.x, xx1.y, xx1.z, xx1.w}};
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (check_nan) saw_nan |= (qv[u][i] != qv[u][i]);
        rv[u][i] -= a[i] * qv[u][i];
        xv[u][i] += a[i] * pv[u][i];
        acc[i] += (double)rv[u][i] * (double)rv[u][i];
      }
    r4[i0] = make_float4(rv[0][0], rv[0][1], rv[0][2], rv[0][3]);
    r4[i1] = make_float4(rv[1][0], rv[1][1], rv[1][2], rv[1][3]);
    x4[i0] = make_float4(xv[0][0], xv[0][1], xv[0][2], xv[0][3]);
    x4[i1] = make_float4(xv[1][0], xv[1][1], xv[1][2], xv[1][3]);
  }
  for (; g < ngroups; g += ry) {
    if (g < full) {
      const int64_t i0 = g * C + tx;
      const float4 q0 = ap4[i0], rr0 = r4[i0], pp0 = p4[i0], xx0 = x4[i0];
      float qv[4] = {q0.x, q0.y, q0.z, q0.w}, rv[4] = {rr0.x, rr0.y, rr0.z, rr0.w};
      float pv[4] = {pp0.x, pp0.y, pp0.z, pp0.w}, xv[4] = {xx0.x, xx0.y, xx0.z, xx0.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (check_nan) saw_nan |= (qv[i] != qv[i]);
        rv[i] -= a[i] * qv[i];
        xv[i] += a[i] * pv[i];
        acc[i] += (double)rv[i] * (double)rv[i];
      }
      r4[i0] = make_float4(rv[0], rv[1], rv[2], rv[3]);
      x4[i0] = make_float4(xv[0], xv[1], xv[2], xv[3]);
    } else {  // the ragged last group (N % 4 rows): element by element
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (g * 4 + m.dr[i] < nrows) {
          const int64_t e = base + (g * 4 + m.dr[i]) * C + m.c[i];
          const float q = ap[e];
          if (check_nan) saw_nan |= (q != q);
          const float rn = r[e] - a[i] * q;
          r[e] = rn;
          x[e] = x[e] + a[i] * pvec[e];
          acc[i] += (double)rn * (double)rn;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) red[(ty * C + tx) * 4 + i] = acc[i];
  __syncthreads();
  if (ty == 0) parts_rr[(b * d.nchunks + chunk) * C + tx] = v4_column_sum(red, tx, C, ry);
  if (saw_nan) atomicOr(&status->nan_detected, 1);
}

__global__ void k_step_p_v4(Dims d, const float* __restrict__ z, float* __restrict__ pvec,
                            const double* __restrict__ parts_rz, int n_rz_parts, const double* __restrict__ parts_rr,
                            const double* __restrict__ rz_old, double* __restrict__ rz_new, double* __restrict__ beta_out,
                            double* __restrict__ resid, uint8_t* __restrict__ conv, const uint8_t* __restrict__ rhs_zero,
                            const lob_cg_status* status, double eps, double sua) {
  if (status->stop) return;
  extern __shared__ double red[];
  const int64_t b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int C = (int)d.C, tx = threadIdx.x, ty = threadIdx.y, ry = blockDim.y;
  float* beta_s = reinterpret_cast<float*>(red);
  const int64_t row0 = (int64_t)chunk * d.rows_per_chunk;
  const int64_t nrows = min(row0 + d.rows_per_chunk, d.N) - row0;
  if (ty == 0) {  // beta, residual norm, converged flags (as k_step_p)
    const float rzn = (float)sum_parts(parts_rz, b, n_rz_parts, C, tx);
    const float rzo = (float)rz_old[b * C + tx];
    const float be = (rzo < (float)eps) ? 0.f : (float)(rzn / rzo);
    beta_s[tx] = be;
    if (chunk == 0) {
      const int64_t i = b * C + tx;
      rz_new[i] = (double)rzn;
      beta_out[i] = (double)be;
      float nrm = (float)sqrt(sum_parts(parts_rr, b, d.nchunks, C, tx));
      if (rhs_zero[i]) nrm = 0.f;
      resid[i] = (double)nrm;
      conv[i] = (nrm < (float)sua) ? 1 : 0;
    }
  }
  __syncthreads();
  const V4Map m = v4_map(tx, C);
  float be[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) be[i] = beta_s[m.c[i]];
  const int64_t base = (b * d.N + row0) * C;
  const int64_t ngroups = (nrows + 3) / 4, full = nrows / 4;
  const float4* z4 = reinterpret_cast<const float4*>(z + base);
  float4* p4 = reinterpret_cast<float4*>(pvec + base);
  int64_t g = ty;
  for (; g + 3 * ry < full; g += 4 * ry) {
    float4 zz[4], pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      zz[u] = z4[(g + u * ry) * C + tx];
      pp[u] = p4[(g + u * ry) * C + tx];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      p4[(g + u * ry) * C + tx] = make_float4(pp[u].x * be[0] + zz[u].x, pp[u].y * be[1] + zz[u].y,
                                              pp[u].z * be[2] + zz[u].z, pp[u].w * be[3] + zz[u].w);
  }
  for (; g < ngroups; g += ry) {
    if (g < full) {
      const int64_t i0 = g * C + tx;
      const float4 zz = z4[i0], pp = p4[i0];
      p4[i0] = make_float4(pp.x * be[0] + zz.x, pp.y * be[1] + zz.y, pp.z * be[2] + zz.z, pp.w * be[3] + zz.w);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (g * 4 + m.dr[i] < nrows) {
          const int64_t e = base + (g * 4 + m.dr[i]) * C + m.c[i];
          pvec[e] = pvec[e] * be[i] + z[e];
        }
      }
    }
  }
}

// ---- one CTA: stop test + tridiagonal update (linear_cg.py:302-332) ---------------------------------------
template <typename T>
__global__ void k_control(lob_cg_params p, int k, const double* __restrict__ alpha, const double* __restrict__ beta,
                          const double* __restrict__ resid, double* __restrict__ prev_ar,
                          double* __restrict__ prev_beta, T* __restrict__ t_mat, lob_cg_status* status) {
  __shared__ double scratch[32];
  __shared__ int s_flag;
  if (status->stop) return;
  const int64_t BC = p.B * p.C;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < BC; i += blockDim.x) acc += resid[i];
  const double mean = block_sum(acc, scratch) / (double)BC;
  const int kmin = min(10, p.max_iter - 1);
  const int ktri = min(p.n_tridiag_iter, p.max_iter - 1);
  const bool tol_ok = (sizeof(T) == 4) ? ((float)mean < (float)p.tolerance) : (mean < p.tolerance);
  const bool stop_now = (k >= kmin) && tol_ok && !(p.n_tridiag > 0 && k < ktri);
  if (stop_now || k + 1 >= p.n_iter) {
    __syncthreads();
    if (threadIdx.x == 0) {
      status->iterations = k + 1;
      status->residual_norm_mean = mean;
      if (stop_now) status->tolerance_reached = 1;
      status->stop = 1;
    }
    if (stop_now) return;
  } else if (threadIdx.x == 0) {
    status->iterations = k + 1;
    status->residual_norm_mean = mean;
  }
  if (!(p.n_tridiag > 0 && k < p.n_tridiag_iter && status->update_tridiag)) return;
  // tridiagonal entries for the first S columns of every batch element
  const int S = p.n_tridiag, Tn = p.n_tridiag_iter;
  const int64_t BS = p.B * S;
  bool keep = false;  // any off-diagonal entry that is NOT < 1e-6 (NaN counts: torch.max propagates NaN)
  for (int64_t i = threadIdx.x; i < BS; i += blockDim.x) {
    const int64_t b = i / S, j = i % S;
    const double a = alpha[b * p.C + j];
    const double be = beta[b * p.C + j];
    const T ar = (T)1 / ((a == 0.0) ? (T)1 : (T)a);
    T* t = t_mat + ((j * p.B + b) * Tn) * Tn;
    if (k == 0) {
      t[0] = ar;
    } else {
      const T pb = (T)prev_beta[i], par = (T)prev_ar[i];
      t[(int64_t)k * Tn + k] = ar + pb * par;
      const T off = (T)sqrt((double)pb) * par;
      t[(int64_t)k * Tn + k - 1] = off;
      t[(int64_t)(k - 1) * Tn + k] = off;
      if (!(off < (T)1e-6)) keep = true;
    }
    prev_ar[i] = (double)ar;
    prev_beta[i] = (double)(T)be;
  }
  if (k > 0) {
    if (threadIdx.x == 0) s_flag = 0;
    __syncthreads();
    if (keep) atomicOr(&s_flag, 1);
    __syncthreads();
    if (threadIdx.x == 0 && !s_flag) status->update_tridiag = 0;
  }
  if (threadIdx.x == 0) status->last_tridiag_iter = k;
}

__global__ void k_status_reset(lob_cg_status* s) {
  s->stop = 0;
  s->tolerance_reached = 0;
  s->iterations = 0;
  s->update_tridiag = 1;
  s->last_tridiag_iter = 0;
  s->nan_detected = 0;
  s->all_converged_at_start = 0;
  s->reserved = 0;
  s->residual_norm_mean = 0.0;
}

// x *= rhs_norm (linear_cg.py:335)
template <typename T>
__global__ void k_finish(Dims d, T* __restrict__ x, const double* __restrict__ rhs_norm) {
  CG_PROLOGUE
  for (int64_t c0 = 0; c0 < d.C; c0 += CX) {
    const int64_t c = c0 + tx;
    if (c >= d.C) continue;
    const T nrm = (T)rhs_norm[b * d.C + c];
    for (int64_t row = row0 + ty; row < row1; row += RY) {
      const int64_t i0 = base + row * d.C + c;
      x[i0] = x[i0] * nrm;
    }
  }
}

static int check_params(const lob_cg_params* p) {
  LOB_REQUIRE(p != nullptr, "lob_cg: params is NULL");
  LOB_REQUIRE(p->B > 0 && p->N > 0 && p->C > 0, "lob_cg: B, N, C must be positive");
  LOB_REQUIRE(p->B <= 65535 * 1LL * 65535, "lob_cg: batch too large");
  LOB_REQUIRE(p->dtype == LOB_F32 || p->dtype == LOB_F64, "lob_cg: bad dtype");
  LOB_REQUIRE(p->n_tridiag >= 0 && p->n_tridiag <= p->C, "lob_cg: n_tridiag must be in [0, C]");
  LOB_REQUIRE(p->n_tridiag == 0 || p->n_tridiag_iter > 0, "lob_cg: n_tridiag_iter must be positive");
  return LOB_OK;
}

struct Launch {
  dim3 grid, block;
  size_t smem;
  Dims d;
};
static Launch make_launch(const CgLayout& L) {
  Launch l;
  l.grid = dim3((unsigned)L.nchunks, (unsigned)L.B, 1);
  l.block = dim3((unsigned)L.cx, (unsigned)L.ry, 1);
  l.smem = sizeof(double) * L.cx * L.ry;
  l.d = Dims{L.N, L.C, L.nchunks, L.rows_per_chunk};
  return l;
}

// the float4 kernels need fp32, one thread per column, and 16-byte addressable batch elements / vectors
static bool v4_ok(const lob_cg_params* p, const CgLayout& L, std::initializer_list<const void*> ptrs) {
  if (p->dtype != LOB_F32 || p->C > 128 || L.cx != p->C || ((p->N * p->C) % 4) != 0) return false;
  for (const void* q : ptrs)
    if (q && (reinterpret_cast<uintptr_t>(q) & 15) != 0) return false;
  return true;
}
static size_t v4_smem(const CgLayout& L) { return sizeof(double) * L.cx * L.ry * 4 + sizeof(float) * L.cx; }

// long caller-supplied partial lists are folded once into ws (slot 0: <p,Ap>, slot 1: <r,z>)
static int fold_parts(const CgLayout& L, const CgPtrs& P, const Launch& l, const double*& parts, int& n_parts, int slot,
                      const lob_cg_status* status, cudaStream_t st) {
  if (!parts || n_parts <= kFoldThreshold) return LOB_OK;
  double* out = P.fold + (size_t)slot * L.B * L.C * kFoldParts;
  k_fold_parts<<<dim3(kFoldParts, (unsigned)L.B), l.block, l.smem, st>>>(parts, n_parts, (int)L.C, out, status);
  LOB_TRY(check_launch("k_fold_parts"));
  parts = out;
  n_parts = kFoldParts;
  return LOB_OK;
}

}  // namespace lob

using namespace lob;

extern "C" size_t lob_cg_workspace_bytes(const lob_cg_params* p) {
  if (!p || p->B <= 0 || p->N <= 0 || p->C <= 0) return 0;
  return cg_layout(p).total;
}

extern "C" int lob_cg_setup(const lob_cg_params* p, void* ws, const void* rhs, const void* x0, void* rhs_n, void* x,
                            void* t_mat, void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(p->B <= 65535, "lob_cg: flattened batch > 65535 not supported yet");
  LOB_REQUIRE(ws && rhs && rhs_n && x, "lob_cg_setup: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  Launch l = make_launch(L);
  k_status_reset<<<1, 1, 0, st>>>(P.status);
  LOB_TRY(check_launch("k_status_reset"));
  if (p->n_tridiag > 0 && t_mat) {
    size_t bytes = (size_t)p->n_tridiag * p->B * p->n_tridiag_iter * p->n_tridiag_iter * dsize(p->dtype);
    LOB_CUDA(cudaMemsetAsync(t_mat, 0, bytes, st));
  }
  LOB_DISPATCH_DTYPE(p->dtype, {
    k_dots_partials<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)rhs, (const scalar_t*)rhs,
                                                               P.parts_a, nullptr);
    LOB_TRY(check_launch("k_dots_partials"));
    k_setup_normalize<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)rhs, (const scalar_t*)x0,
                                                                 (scalar_t*)rhs_n, (scalar_t*)x, P.parts_a, P.rhs_norm,
                                                                 P.rhs_zero, p->eps);
    LOB_TRY(check_launch("k_setup_normalize"));
  });
  return LOB_OK;
}

extern "C" int lob_cg_residual_init(const lob_cg_params* p, void* ws, const void* rhs_n, const void* ax0, void* r,
                                    void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(ws && rhs_n && r, "lob_cg_residual_init: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  Launch l = make_launch(L);
  LOB_DISPATCH_DTYPE(p->dtype, {
    k_residual_init<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)rhs_n, (const scalar_t*)ax0,
                                                               (scalar_t*)r, P.parts_rr, P.status);
    LOB_TRY(check_launch("k_residual_init"));
  });
  k_residual_flags<<<1, 1024, 0, st>>>(p->B * p->C, p->C, L.nchunks, P.parts_rr, P.resid, P.conv, P.status,
                                       p->stop_updating_after, p->n_tridiag, p->dtype == LOB_F32);
  LOB_TRY(check_launch("k_residual_flags"));
  return LOB_OK;
}

extern "C" int lob_cg_direction_init(const lob_cg_params* p, void* ws, const void* r, const void* z, void* pvec,
                                     const double* rz_partials, int32_t n_rz_parts, void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(ws && r && z && pvec, "lob_cg_direction_init: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  Launch l = make_launch(L);
  LOB_DISPATCH_DTYPE(p->dtype, {
    const double* parts = P.parts_rr;  // z aliases r: <r,z> = <r,r> already accumulated by residual_init
    int n_rz = L.nchunks;
    if (z != r) {
      if (rz_partials) {
        parts = rz_partials;
        n_rz = n_rz_parts;
        LOB_TRY(fold_parts(L, P, l, parts, n_rz, 1, nullptr, st));
      } else {
        k_dots_partials<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)z, (const scalar_t*)r,
                                                                   P.parts_rz, nullptr);
        LOB_TRY(check_launch("k_dots_partials"));
        parts = P.parts_rz;
      }
    }
    k_direction_init<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)z, (scalar_t*)pvec, parts, n_rz,
                                                                P.rz, p->dtype == LOB_F32);
    LOB_TRY(check_launch("k_direction_init"));
  });
  return LOB_OK;
}

extern "C" int lob_cg_step_xr(const lob_cg_params* p, void* ws, int32_t k, const void* ap, const void* pvec, void* x,
                              void* r, const double* pap_partials, int32_t n_parts, void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(ws && ap && pvec && x && r, "lob_cg_step_xr: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  Launch l = make_launch(L);
  const size_t bc = (size_t)p->B * p->C;
  LOB_DISPATCH_DTYPE(p->dtype, {
    const double* parts = pap_partials;
    int nparts = n_parts;
    LOB_TRY(fold_parts(L, P, l, parts, nparts, 0, P.status, st));
    if (!parts) {
      k_dots_partials<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)pvec, (const scalar_t*)ap,
                                                                 P.parts_a, P.status);
      LOB_TRY(check_launch("k_dots_partials"));
      parts = P.parts_a;
      nparts = L.nchunks;
    }
    if (v4_ok(p, L, {ap, pvec, x, r})) {
      k_step_xr_v4<<<l.grid, l.block, v4_smem(L), st>>>(l.d, (const float*)ap, (const float*)pvec, (float*)x, (float*)r,
                                                        parts, nparts, P.rz + (size_t)(k & 1) * bc, P.conv, P.alpha,
                                                        P.parts_rr, P.status, p->eps, k == 0);
      LOB_TRY(check_launch("k_step_xr_v4"));
      return LOB_OK;
    }
    k_step_xr<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)ap, (const scalar_t*)pvec,
                                                         (scalar_t*)x, (scalar_t*)r, parts, nparts,
                                                         P.rz + (size_t)(k & 1) * bc, P.conv, P.alpha, P.parts_rr,
                                                         P.status, p->eps, k == 0);
    LOB_TRY(check_launch("k_step_xr"));
  });
  return LOB_OK;
}

extern "C" int lob_cg_step_p(const lob_cg_params* p, void* ws, int32_t k, const void* z, const void* r, void* pvec,
                             void* t_mat, const double* rz_partials, int32_t n_rz_parts, void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(ws && r && pvec, "lob_cg_step_p: NULL pointer");
  LOB_REQUIRE(p->n_tridiag == 0 || t_mat, "lob_cg_step_p: t_mat is NULL but n_tridiag > 0");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  Launch l = make_launch(L);
  const size_t bc = (size_t)p->B * p->C;
  LOB_DISPATCH_DTYPE(p->dtype, {
    const double* parts_rz = P.parts_rr;
    int n_rz = L.nchunks;
    const scalar_t* zz = (const scalar_t*)r;
    if (z && z != r) {
      zz = (const scalar_t*)z;
      if (rz_partials) {  // <r,z> came out of the preconditioner's fused epilogue
        parts_rz = rz_partials;
        n_rz = n_rz_parts;
        LOB_TRY(fold_parts(L, P, l, parts_rz, n_rz, 1, P.status, st));
      } else {
        k_dots_partials<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (const scalar_t*)z, (const scalar_t*)r,
                                                                   P.parts_rz, P.status);
        LOB_TRY(check_launch("k_dots_partials"));
        parts_rz = P.parts_rz;
      }
    }
    if (v4_ok(p, L, {zz, pvec})) {
      k_step_p_v4<<<l.grid, l.block, v4_smem(L), st>>>(l.d, (const float*)zz, (float*)pvec, parts_rz, n_rz, P.parts_rr,
                                                       P.rz + (size_t)(k & 1) * bc, P.rz + (size_t)((k + 1) & 1) * bc,
                                                       P.beta, P.resid, P.conv, P.rhs_zero, P.status, p->eps,
                                                       p->stop_updating_after);
      LOB_TRY(check_launch("k_step_p_v4"));
    } else {
      k_step_p<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, zz, (scalar_t*)pvec, parts_rz, n_rz, P.parts_rr,
                                                          P.rz + (size_t)(k & 1) * bc,
                                                          P.rz + (size_t)((k + 1) & 1) * bc, P.beta, P.resid, P.conv,
                                                          P.rhs_zero, P.status, p->eps, p->stop_updating_after);
      LOB_TRY(check_launch("k_step_p"));
    }
    k_control<scalar_t><<<1, 1024, 0, st>>>(*p, k, P.alpha, P.beta, P.resid, P.prev_ar, P.prev_beta,
                                            (scalar_t*)t_mat, P.status);
    LOB_TRY(check_launch("k_control"));
  });
  return LOB_OK;
}

extern "C" int lob_cg_poll_sync(const lob_cg_params* p, void* ws, lob_cg_status* host_status, void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(ws && host_status, "lob_cg_poll_sync: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  LOB_CUDA(cudaMemcpyAsync(host_status, P.status, sizeof(lob_cg_status), cudaMemcpyDeviceToHost, st));
  LOB_CUDA(cudaStreamSynchronize(st));
  return LOB_OK;
}

extern "C" int lob_cg_finish(const lob_cg_params* p, void* ws, void* x, void* stream) {
  LOB_TRY(check_params(p));
  LOB_REQUIRE(ws && x, "lob_cg_finish: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CgLayout L = cg_layout(p);
  CgPtrs P = cg_ptrs(L, ws);
  Launch l = make_launch(L);
  LOB_DISPATCH_DTYPE(p->dtype, {
    k_finish<scalar_t><<<l.grid, l.block, l.smem, st>>>(l.d, (scalar_t*)x, P.rhs_norm);
    LOB_TRY(check_launch("k_finish"));
  });
  return LOB_OK;
}

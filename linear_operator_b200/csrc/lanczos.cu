// lanczos.cu -- one fused kernel per Lanczos iteration with full re-orthogonalisation (utils/lanczos.py:9-164).
//
// The reference spends ~15 ATen launches and >= 2 host synchronisations per iteration and reads the whole Q panel three
// times through broadcasting temporaries of size (k+1, *batch, N, C).  Every (batch element, column) pair is an
// independent recurrence, so here one CTA owns a group of up to 8 columns of one batch element and runs the whole
// iteration for them -- alpha = <q_k, w - beta q_{k-1}>, r = w - beta q_{k-1} - alpha q_k, the Gram-Schmidt pass
// r -= Q (Q^T r), the norm / beta, the renormalisation and the <q_j, r> > tol check -- with block-level reductions only:
//   * r lives in shared memory for the whole iteration when N x 8 columns fit (else in its q_mat slot, L2-resident);
//   * the k+1 dots of Q^T r are split over the 8 warps (warp w owns panel vectors j = w, w+8, ...), each a
//     shuffle-reduced strided dot with double accumulation;
//   * the correction r -= sum_j corr_j q_j and the norm are one row-partitioned pass;
//   * all reductions have a fixed order: bit-reproducible.
// The panel is read three times per iteration (dots, correction, check), as in the reference, but in ONE launch and with
// no temporaries.  The only host traffic is the two decision words of the reference's own control flow
// ("some <q_j, r> > tol -> re-orthogonalise again", "all |beta| <= 1e-6 -> stop", lanczos.py:136-150).
#include "common.cuh"

namespace lob {

constexpr int LZ_CW = 8;        // columns per CTA
constexpr int LZ_THREADS = 256;
constexpr int LZ_ROWL = LZ_THREADS / LZ_CW;  // 32 row lanes in the row-partitioned passes

template <typename T>
struct LzParams {
  const T* w;    // (B, N, C)  A q_k
  T* q_mat;      // (Tcap, B, N, C)
  T* t_mat;      // (Tcap, Tcap, B, C)
  int32_t* flags;  // [0] some <q_j, r> > tol   [1] some |beta| > 1e-6
  int64_t B, N, C;
  int Tcap, k, mode;  // mode 0: first iteration, 1: iteration k >= 1, 2: one more re-orthogonalisation of q_{k+1}
  int r_in_smem;
  double tol;
};

// per-column block sum of one double per thread (thread = (row lane, column)); result broadcast through smem
__device__ __forceinline__ void lz_block_colsum(double v, double* scratch /*[ROWL][CW]*/, double* out /*[CW]*/) {
  const int col = threadIdx.x % LZ_CW, rl = threadIdx.x / LZ_CW;
  __syncthreads();
  scratch[rl * LZ_CW + col] = v;
  __syncthreads();
  if (threadIdx.x < LZ_CW) {
    double s = 0.0;
    for (int i = 0; i < LZ_ROWL; ++i) s += scratch[i * LZ_CW + threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(LZ_THREADS)
k_lanczos_step(LzParams<T> p) {
  extern __shared__ __align__(16) unsigned char lz_smem[];
  double* scratch = reinterpret_cast<double*>(lz_smem);                 // [ROWL][CW]
  double* colv = scratch + LZ_ROWL * LZ_CW;                             // [4][CW]: alpha, norm, spare
  double* corr = colv + 4 * LZ_CW;                                      // [Tcap][CW]
  T* rsm = reinterpret_cast<T*>(corr + (size_t)p.Tcap * LZ_CW);         // [N][CW] when r_in_smem

  const int64_t b = blockIdx.y;
  const int c0 = blockIdx.x * LZ_CW;
  const int cw = (int)min((int64_t)LZ_CW, p.C - c0);
  const int tid = threadIdx.x, col = tid % LZ_CW, rl = tid / LZ_CW;
  const int warp = tid >> 5, lane = tid & 31;
  const bool cok = col < cw;
  const int64_t N = p.N, C = p.C;
  const int k = p.k;
  const int64_t vec = p.B * N * C;                   // elements of one q vector slot
  const int64_t base = (b * N) * C + c0;             // offset of (b, row 0, c0) inside a slot
  const T* wv = p.w + base;
  T* qn = p.q_mat + (int64_t)(k + 1) * vec + base;   // q_{k+1}
  const T* qk = p.q_mat + (int64_t)k * vec + base;
  const T* qp = p.q_mat + (int64_t)(k > 0 ? k - 1 : 0) * vec + base;
  // r: shared memory [n][CW] or its final place in q_mat
  T* rb = p.r_in_smem ? rsm : qn;
  const int64_t rstride = p.r_in_smem ? LZ_CW : C;
  auto tm = [&](int i, int j) -> T* { return p.t_mat + (((int64_t)i * p.Tcap + j) * p.B + b) * C + c0; };

  if (p.mode != 2) {
    // ---- alpha = <q_k, w - beta_prev q_{k-1}> ----
    const T bprev = (p.mode == 1 && cok) ? *(tm(k, k - 1) + col) : (T)0;
    double s = 0.0;
    if (cok)
      for (int64_t n = rl; n < N; n += LZ_ROWL) {
        const T r0 = wv[n * C + col] - bprev * qp[n * C + col];
        s += (double)qk[n * C + col] * (double)r0;
      }
    lz_block_colsum(s, scratch, colv);
    if (tid < cw) *(tm(k, k) + tid) = (T)colv[tid];
    if (k + 1 >= p.Tcap) return;  // last iteration: only alpha (lanczos.py:112-116)
    // ---- r = w - beta_prev q_{k-1} - alpha q_k ----
    const T alpha = cok ? (T)colv[col] : (T)0;
    if (cok)
      for (int64_t n = rl; n < N; n += LZ_ROWL) {
        T r0 = wv[n * C + col] - bprev * qp[n * C + col];
        r0 -= alpha * qk[n * C + col];
        rb[n * rstride + col] = r0;
      }
    __syncthreads();
  } else if (p.r_in_smem) {
    if (cok)
      for (int64_t n = rl; n < N; n += LZ_ROWL) rb[n * rstride + col] = qn[n * C + col];
    __syncthreads();
  }

  const int nq = k + 1;  // panel vectors q_0 .. q_k
  const int l_col = lane % LZ_CW, l_row = lane / LZ_CW;  // warp-per-vector passes: 4 row lanes x 8 columns
  const bool lok = l_col < cw;

  if (p.mode != 0) {
    // ---- corr_j = <q_j, r>   (lanczos.py:120 / :142) ----
    for (int j = warp; j < nq; j += LZ_THREADS / 32) {
      const T* qj = p.q_mat + (int64_t)j * vec + base;
      double s = 0.0;
      if (lok)
        for (int64_t n = l_row; n < N; n += 4) s += (double)qj[n * C + l_col] * (double)rb[n * rstride + l_col];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (lane < LZ_CW) corr[j * LZ_CW + lane] = s;
    }
    __syncthreads();
    // ---- r -= sum_j corr_j q_j   (:121-122 / :143-144) ----
    if (cok)
      for (int64_t n = rl; n < N; n += LZ_ROWL) {
        T acc = (T)0;
        for (int j = 0; j < nq; ++j)
          acc += p.q_mat[(int64_t)j * vec + base + n * C + col] * (T)corr[j * LZ_CW + col];
        rb[n * rstride + col] -= acc;
      }
    __syncthreads();
  }

  // ---- norm, beta, normalise   (:123-129 / :145-146; first iteration :90-97) ----
  {
    double s = 0.0;
    if (cok)
      for (int64_t n = rl; n < N; n += LZ_ROWL) {
        const double v = (double)rb[n * rstride + col];
        s += v * v;
      }
    lz_block_colsum(s, scratch, colv + LZ_CW);
    const T nrm = cok ? (T)sqrt(colv[LZ_CW + col]) : (T)1;
    if (p.mode != 2 && tid < cw) {
      const T bt = (T)sqrt(colv[LZ_CW + tid]);
      *(tm(k, k + 1) + tid) = bt;
      *(tm(k + 1, k) + tid) = bt;
      if (fabs((double)bt) > 1e-6) atomicOr(p.flags + 1, 1);
    }
    if (cok)
      for (int64_t n = rl; n < N; n += LZ_ROWL) {
        const T v = rb[n * rstride + col] / nrm;
        rb[n * rstride + col] = v;
        if (p.r_in_smem) qn[n * C + col] = v;
      }
    __syncthreads();
  }

  if (p.mode != 0) {
    // ---- does any <q_j, r> exceed tol?  (signed comparison, as the reference: :133 / :147) ----
    int bad = 0;
    for (int j = warp; j < nq; j += LZ_THREADS / 32) {
      const T* qj = p.q_mat + (int64_t)j * vec + base;
      double s = 0.0;
      if (lok)
        for (int64_t n = l_row; n < N; n += 4) s += (double)qj[n * C + l_col] * (double)rb[n * rstride + l_col];
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      if (lane < cw && (double)(T)s > p.tol) bad = 1;
    }
    if (bad) atomicOr(p.flags, 1);
  }
}

// q0 = init / ||init||_2 per column (lanczos.py:83-84)
template <typename T>
__global__ void __launch_bounds__(LZ_THREADS)
k_lanczos_init(const T* __restrict__ init, T* __restrict__ q0, int64_t N, int64_t C) {
  __shared__ double scratch[LZ_ROWL * LZ_CW];
  __shared__ double colv[LZ_CW];
  const int64_t b = blockIdx.y;
  const int c0 = blockIdx.x * LZ_CW;
  const int cw = (int)min((int64_t)LZ_CW, C - c0);
  const int col = threadIdx.x % LZ_CW, rl = threadIdx.x / LZ_CW;
  const bool cok = col < cw;
  const int64_t base = b * N * C + c0;
  double s = 0.0;
  if (cok)
    for (int64_t n = rl; n < N; n += LZ_ROWL) {
      const double v = (double)init[base + n * C + col];
      s += v * v;
    }
  lz_block_colsum(s, scratch, colv);
  if (cok) {
    const T nrm = (T)sqrt(colv[col]);
    for (int64_t n = rl; n < N; n += LZ_ROWL) q0[base + n * C + col] = init[base + n * C + col] / nrm;
  }
}

template <typename T>
static int lanczos_step_t(int mode, int64_t B, int64_t N, int64_t C, int Tcap, int k, const void* w, void* q_mat,
                          void* t_mat, int32_t* flags, double tol, cudaStream_t st) {
  LzParams<T> p;
  p.w = (const T*)w;
  p.q_mat = (T*)q_mat;
  p.t_mat = (T*)t_mat;
  p.flags = flags;
  p.B = B;
  p.N = N;
  p.C = C;
  p.Tcap = Tcap;
  p.k = k;
  p.mode = mode;
  p.tol = tol;
  const size_t fixed = (size_t)(LZ_ROWL * LZ_CW + 4 * LZ_CW + (size_t)Tcap * LZ_CW) * sizeof(double);
  const size_t rbytes = (size_t)N * LZ_CW * sizeof(T);
  p.r_in_smem = (fixed + rbytes <= 200 * 1024) ? 1 : 0;
  const size_t smem = fixed + (p.r_in_smem ? rbytes : 0);
  if (fixed > 200 * 1024) return fail(LOB_ERR_UNSUPPORTED, "lob_lanczos_step: too many iterations for the scratch");
  auto kern = k_lanczos_step<T>;
  LOB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // decision words of this launch
  if (mode != 2) LOB_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), st));
  else LOB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
  dim3 grid((unsigned)cdiv(C, LZ_CW), (unsigned)B);
  kern<<<grid, LZ_THREADS, smem, st>>>(p);
  return check_launch("k_lanczos_step");
}

}  // namespace lob

using namespace lob;

extern "C" int lob_lanczos_init(int32_t dtype, int64_t B, int64_t N, int64_t C, const void* init, void* q0,
                                void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0, "lob_lanczos_init: sizes must be positive");
  LOB_REQUIRE(B <= 65535, "lob_lanczos_init: flattened batch > 65535 not supported");
  LOB_REQUIRE(init && q0, "lob_lanczos_init: NULL pointer");
  dim3 grid((unsigned)cdiv(C, LZ_CW), (unsigned)B);
  LOB_DISPATCH_DTYPE(dtype, {
    k_lanczos_init<scalar_t><<<grid, LZ_THREADS, 0, (cudaStream_t)stream>>>((const scalar_t*)init, (scalar_t*)q0, N, C);
    return check_launch("k_lanczos_init");
  });
}

extern "C" int lob_lanczos_step(int32_t dtype, int32_t mode, int64_t B, int64_t N, int64_t C, int32_t t_cap, int32_t k,
                                const void* w, void* q_mat, void* t_mat, int32_t* flags, double tol, void* stream) {
  LOB_REQUIRE(B > 0 && N > 0 && C > 0 && t_cap >= 2, "lob_lanczos_step: sizes must be positive, t_cap >= 2");
  LOB_REQUIRE(B <= 65535, "lob_lanczos_step: flattened batch > 65535 not supported");
  LOB_REQUIRE(mode >= 0 && mode <= 2, "lob_lanczos_step: mode must be 0, 1 or 2");
  LOB_REQUIRE(k >= 0 && k < t_cap && (mode != 0 || k == 0) && (mode == 0 || k >= 1),
              "lob_lanczos_step: iteration index out of range for this mode");
  LOB_REQUIRE(mode != 2 || k + 1 < t_cap, "lob_lanczos_step: nothing to re-orthogonalise after the last iteration");
  LOB_REQUIRE((mode == 2 || w) && q_mat && t_mat && flags, "lob_lanczos_step: NULL pointer");
  LOB_DISPATCH_DTYPE(dtype, {
    return lanczos_step_t<scalar_t>(mode, B, N, C, t_cap, k, w, q_mat, t_mat, flags, tol, (cudaStream_t)stream);
  });
}

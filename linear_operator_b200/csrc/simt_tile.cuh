// simt_tile.cuh -- shared register-tile inner product of the CUDA-core matmul kernels.
#pragma once
#include "common.cuh"

namespace lob {

constexpr int TM = 128;      // output rows per CTA
constexpr int LDA_S = TM + 4;  // smem leading dimension of the k-major A tile (keeps 16-byte alignment)

template <typename T>
struct TileK {
  static constexpr int value = sizeof(T) == 4 ? 32 : 16;
};

// acc[i][j] += As[k][ty*4+i] * Bs[k][tx*RN+j]
template <typename T, typename ACC, int RN, int TK>
__device__ __forceinline__ void tile_fma(const T* __restrict__ As, const T* __restrict__ Bs, int cp, int ty, int tx,
                                         ACC (&acc)[4][RN]) {
#pragma unroll 4
  for (int k = 0; k < TK; ++k) {
    T a[4];
    if constexpr (sizeof(T) == 4) {
      const float4 v = *reinterpret_cast<const float4*>(As + k * LDA_S + ty * 4);
      a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    } else {
      const double2 v0 = *reinterpret_cast<const double2*>(As + k * LDA_S + ty * 4);
      const double2 v1 = *reinterpret_cast<const double2*>(As + k * LDA_S + ty * 4 + 2);
      a[0] = v0.x; a[1] = v0.y; a[2] = v1.x; a[3] = v1.y;
    }
    T bb[RN];
#pragma unroll
    for (int j = 0; j < RN; ++j) bb[j] = Bs[k * cp + tx * RN + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < RN; ++j) acc[i][j] += (ACC)a[i] * (ACC)bb[j];
  }
}


inline int pick_rn(int64_t C) {
  int rn = (int)cdiv(C, 8);
  if (rn > 8) rn = 8;
  if (rn < 1) rn = 1;
  return rn;
}

}  // namespace lob

// Register-tiled fp32 GEMM building block shared by engine A (ecnf_solve_impl.cuh) and engine B (ecnf_train.cu).
#pragma once
#include "ecnf_common.cuh"

namespace ecnf_tile {

constexpr int NTHREADS = 256;
constexpr int WCHUNK = 4096;  // floats per weight chunk (16 KB), double buffered

// ------------------------------------------------------------------------------------------------
// [TR x K] (smem) x [K x N] (global, streamed) -> register tile
// thread (ty, tx): rows 2*RT*warp + 2*rr + (ty&1), cols (cc>>2)*64 + tx*4 + (cc&3)
// ------------------------------------------------------------------------------------------------
template <int N>
struct ColT {
  static constexpr int CT = (N >= 64) ? N / 16 : 4;
};

template <int K, int N, int TR, int LD, bool ZERO>
__device__ __forceinline__ void tile_gemm(const float* Xs, const float* __restrict__ Wg, float* Wb,
                                          float (&acc)[TR / 16][ColT<N>::CT], int nrows) {
  constexpr int RT = TR / 16, CT = ColT<N>::CT, NSEG = CT / 4;
  constexpr int KC = (WCHUNK / N < K) ? WCHUNK / N : K;
  constexpr int NCH = K / KC;
  static_assert(K % KC == 0 && KC % 4 == 0, "chunking");
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, warp = tid >> 5;
  const bool active = (2 * RT * warp < nrows) && (tx * 4 < N);
  const int row0 = 2 * RT * warp + (ty & 1);
  if (ZERO) {
#pragma unroll
    for (int rr = 0; rr < RT; ++rr)
#pragma unroll
      for (int cc = 0; cc < CT; ++cc) acc[rr][cc] = 0.f;
  }
  auto load_chunk = [&](int c, int buf) {
    const float4* src = reinterpret_cast<const float4*>(Wg + (size_t)c * KC * N);
    float4* dst = reinterpret_cast<float4*>(Wb + buf * WCHUNK);
    for (int i = tid; i < KC * N / 4; i += NTHREADS) cp_async16(dst + i, src + i);
    cp_async_commit();
  };
  load_chunk(0, 0);
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (c + 1 < NCH) {
      load_chunk(c + 1, (c + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (active) {
      const float* Wc = Wb + (c & 1) * WCHUNK + tx * 4;
      const float* Xc = Xs + row0 * LD + c * KC;
#pragma unroll 2
      for (int k4 = 0; k4 < KC / 4; ++k4) {
        float4 a[RT];
#pragma unroll
        for (int rr = 0; rr < RT; ++rr) a[rr] = *reinterpret_cast<const float4*>(Xc + 2 * rr * LD + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float w[CT];
#pragma unroll
          for (int sg = 0; sg < NSEG; ++sg) {
            float4 wv = *reinterpret_cast<const float4*>(Wc + (4 * k4 + kk) * N + sg * 64);
            w[4 * sg + 0] = wv.x; w[4 * sg + 1] = wv.y; w[4 * sg + 2] = wv.z; w[4 * sg + 3] = wv.w;
          }
#pragma unroll
          for (int rr = 0; rr < RT; ++rr) {
            const float av = kk == 0 ? a[rr].x : kk == 1 ? a[rr].y : kk == 2 ? a[rr].z : a[rr].w;
#pragma unroll
            for (int cc = 0; cc < CT; ++cc) acc[rr][cc] = fmaf(av, w[cc], acc[rr][cc]);
          }
        }
      }
    }
    __syncthreads();
  }
}


}  // namespace ecnf_tile

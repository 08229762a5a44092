// Engine B on the tensor cores: the big row-major GEMMs of the flow-matching step
//     C[M, N] = epilogue( op(A)[M, K] x W[K, N] ),   M = B*n*(n-1) edge rows (175 k for QM9 at batch 512), K = N = mlp_units
// (forward Dense layers and, with the pre-transposed weights, the backward-data GEMMs dA = dZ W^T) as tcgen05.mma with
// the same 3-pass bf16 split as engine A (hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM, ~4e-6 relative error).
//
//   * one persistent CTA per SM; a 128-row tile = the 128 TMEM lanes, thread (row, kh) owns row `row`;
//   * A: coalesced loads (64 columns at a time) -> op (identity / SiLU) -> shared-memory stage -> the row's owner splits
//     it to bf16 (hi, lo) and writes it into TENSOR MEMORY as the A operand (lane = row, one column = two consecutive k);
//   * W: one 128-column half at a time, converted once per CTA into shared memory in the no-swizzle K-major canonical
//     layout (the B operand); the row tiles then stream past it.  The CTAs of the two halves walk the tiles together, so
//     A comes from HBM once (second read from L2);
//   * epilogue: accumulator rows -> shared-memory stage -> coalesced pass (one warp = 512 contiguous bytes of a row):
//     + bias, + per-graph row vector, + residual, + add, x silu'(z), fp32 store.
#pragma once
#include "ecnf_tc.cuh"

namespace ecnf_train_tc {

using namespace ecnf_tc;

struct Args {
  const float* A;
  const float* W;
  const float* bias;
  const float* rowvec;
  int rows_per_vec;
  const float* resid;
  const float* add;
  const float* mulz;
  float* C;
  int M;
  int a_op;
};

__device__ __forceinline__ float silu(float z) { return z * (1.0f / (1.0f + expf(-z))); }
__device__ __forceinline__ float dsilu(float z) {
  const float s = 1.0f / (1.0f + expf(-z));
  return s * (1.f + z * (1.f - s));
}

constexpr int STAGE_LD = 132;   // floats per staged row: 16-byte aligned, rows 8 apart share banks => float4 accesses by
                                // 32 rows take the minimal 4 wavefronts

template <int K, int N>
__global__ void __launch_bounds__(256, 1) gemm_rows_tc_kernel(const __grid_constant__ Args g) {
  static_assert((K == 128 || K == 256) && (N == 128 || N == 256), "shapes");
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* Whi = reinterpret_cast<__nv_bfloat16*>(smem);            // [128 n][K] canonical K-major
  __nv_bfloat16* Wlo = Whi + 128 * K;
  float* stage = reinterpret_cast<float*>(smem + (size_t)4 * 128 * K);    // [128 rows][STAGE_LD]: A chunks in, C tile out
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, kh = tid >> 7;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) mbar_init(&mbar, 1);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
  const uint32_t a_hi = tmem, a_lo = tmem + K / 2, acc = tmem + K;       // A: K columns, accumulator: 128 columns
  uint32_t phase = 0;
  const int ntiles = (g.M + 127) / 128;

  // CTA c works on column half (c % NH) of the row tiles c / NH, c / NH + grid / NH, ...: the CTAs of a group read the same
  // A tiles at about the same time, so A comes from HBM once and from L2 for the other half
  constexpr int NH = N / 128;
  const int tile_step = gridDim.x / NH;
  {
    const int nh = blockIdx.x % NH;
    // ---- W[:, 128 nh .. +128) -> bf16 hi / lo, canonical K-major with rows = n
    for (int idx = tid; idx < K * 128; idx += 256) {
      const int k = idx >> 7, n = idx & 127;
      const float w = g.W[(size_t)k * N + 128 * nh + n];
      const __nv_bfloat16 h = __float2bfloat16(w);
      const int d = canon_index(n, k, 128);
      Whi[d] = h;
      Wlo[d] = __float2bfloat16(w - __bfloat162float(h));
    }
    fence_proxy_async();
    __syncthreads();
    // A chunks (128 rows x 64 columns) are prefetched into registers one step ahead: the loads of chunk c+1 (or of the
    // next tile's first chunk) are in flight while chunk c is converted / multiplied / written out
    float4 nx[8];
    auto fetch = [&](int tile_, int cb_) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = tid + 256 * i, r = idx >> 4, c4 = idx & 15;
        nx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile_ < ntiles && tile_ * 128 + r < g.M)
          nx[i] = *reinterpret_cast<const float4*>(g.A + (size_t)(tile_ * 128 + r) * K + cb_ * 64 + 4 * c4);
      }
    };
    fetch(blockIdx.x / NH, 0);
    for (int tile = blockIdx.x / NH; tile < ntiles; tile += tile_step) {
      const int r0 = tile * 128;
      // ---- A operand, 64 columns at a time: coalesced global loads (16 lanes = 256 B of one row) -> op -> stage;
      //      then thread (row, kh) takes its row's 32 values, splits them and writes them into tensor memory
#pragma unroll 1
      for (int cb = 0; cb < K / 64; ++cb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = tid + 256 * i, r = idx >> 4, c4 = idx & 15;
          float4 x = nx[i];
          if (g.a_op) { x.x = silu(x.x); x.y = silu(x.y); x.z = silu(x.z); x.w = silu(x.w); }
          *reinterpret_cast<float4*>(stage + r * STAGE_LD + 4 * c4) = x;
        }
        if (cb + 1 < K / 64) fetch(tile, cb + 1);
        else fetch(tile + tile_step, 0);
        __syncthreads();
        uint32_t h[16], l[16];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 x = *reinterpret_cast<const float4*>(stage + row * STAGE_LD + 32 * kh + 4 * c4);
          split_pack(x.x, x.y, h[2 * c4], l[2 * c4]);
          split_pack(x.z, x.w, h[2 * c4 + 1], l[2 * c4 + 1]);
        }
        const uint32_t col = (uint32_t)(cb * 32 + kh * 16);
        tmem_st16(a_hi + lane_addr + col, h);
        tmem_st16(a_lo + lane_addr + col, l);
        __syncthreads();     // stage is rewritten by the next chunk
      }
      tmem_wait_st();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 128);
        const uint32_t lbo = (128 / 8) * 128;
        for (int ks = 0; ks < K / 16; ++ks) {
          const uint64_t bh = make_sdesc(smem_u32(Whi) + ks * 2 * lbo, lbo, 128);
          const uint64_t bl = make_sdesc(smem_u32(Wlo) + ks * 2 * lbo, lbo, 128);
          mma_ts(acc, a_hi + ks * 8, bh, idesc, ks > 0 ? 1u : 0u);
          mma_ts(acc, a_lo + ks * 8, bh, idesc, 1u);
          mma_ts(acc, a_hi + ks * 8, bl, idesc, 1u);
        }
        mma_commit(&mbar);
      }
      mbar_wait(&mbar, phase & 1u);
      ++phase;
      tc_fence_after();
      // ---- epilogue: accumulator row segments -> stage, then coalesced (one warp = one 512-byte row segment)
      {
        uint32_t x0[32], x1[32];
        tmem_ld32(acc + lane_addr + 64u * kh, x0);
        tmem_ld32(acc + lane_addr + 64u * kh + 32u, x1);
        tmem_wait_ld();
        float* dst = stage + row * STAGE_LD + 64 * kh;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          *reinterpret_cast<uint4*>(dst + 4 * c4) = make_uint4(x0[4 * c4], x0[4 * c4 + 1], x0[4 * c4 + 2], x0[4 * c4 + 3]);
          *reinterpret_cast<uint4*>(dst + 32 + 4 * c4) = make_uint4(x1[4 * c4], x1[4 * c4 + 1], x1[4 * c4 + 2], x1[4 * c4 + 3]);
        }
      }
      tc_fence_before();
      __syncthreads();     // accumulator and A operand are free for the next tile; the C tile is staged
      tc_fence_after();
#pragma unroll 8
      for (int i = 0; i < 16; ++i) {
        const int idx = tid + 256 * i, r = idx >> 5, c4 = idx & 31;
        if (r0 + r >= g.M) continue;
        const size_t grow = (size_t)(r0 + r);
        const int col = 128 * nh + 4 * c4;
        float4 v = *reinterpret_cast<const float4*>(stage + r * STAGE_LD + 4 * c4);
        if (g.bias) {
          const float4 b = *reinterpret_cast<const float4*>(g.bias + col);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (g.rowvec) {
          const float4 b = *reinterpret_cast<const float4*>(g.rowvec + (grow / g.rows_per_vec) * N + col);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (g.resid) {
          const float4 b = *reinterpret_cast<const float4*>(g.resid + grow * N + col);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (g.add) {
          const float4 b = *reinterpret_cast<const float4*>(g.add + grow * N + col);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (g.mulz) {
          const float4 z = *reinterpret_cast<const float4*>(g.mulz + grow * N + col);
          v.x *= dsilu(z.x); v.y *= dsilu(z.y); v.z *= dsilu(z.z); v.w *= dsilu(z.w);
        }
        *reinterpret_cast<float4*>(g.C + grow * N + col) = v;
      }
      __syncthreads();     // stage is rewritten by the next tile's A chunks
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int K, int N>
cudaError_t launch(const Args& g, int num_sms, cudaStream_t st) {
  const size_t smem = (size_t)2 * 128 * K * sizeof(__nv_bfloat16) + (size_t)128 * STAGE_LD * sizeof(float);
  auto kern = gemm_rows_tc_kernel<K, N>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int ntiles = (g.M + 127) / 128;
  constexpr int NH = N / 128;
  int groups = num_sms / NH;
  if (groups > ntiles) groups = ntiles;
  kern<<<groups * NH, 256, smem, st>>>(g);
  return cudaGetLastError();
}

}  // namespace ecnf_train_tc

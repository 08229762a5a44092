// Hardware probe for the "feature-on-lane" formulation used by the tensor-core engine (ecnf_solve_tc.cuh):
//     D^T[out feature m (TMEM lane), row n (TMEM column)] = W^T[m][k] (A operand) x Act^T[k][n] (B operand)
//   * A (weights, bf16 hi/lo) lives in TENSOR MEMORY (lane = m, one 32-bit column = two consecutive k), written with
//     tcgen05.st by the thread that owns the lane;
//   * B (activations, bf16 hi/lo) lives in shared memory in the no-swizzle MN-MAJOR canonical layout:
//         byte offset(n, k) = (n/8) * SBO + (k/8) * LBO + (k%8) * 16 + (n%8) * 2,   LBO = 128, SBO = 16 * 128
//     so that the thread owning feature k writes 8 consecutive rows with one 16-byte store and a warp writes 512
//     contiguous bytes.
// Checks the product against fp64 (3-pass split), for N = 128 and N = 112 and K = 128 / 64, then times the MMA rate
// for TS (A in TMEM) and SS (A in shared memory, K-major) issue, alone and with shared-memory store traffic from the
// other warps, and the TMEM load / store cost of an epilogue.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tc2 tools/probe_tc2.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ecnf_b200/csrc/ecnf_tc.cuh"

using namespace ecnf_tc;

constexpr int M = 128, KMAX = 128, NMAX = 128;
constexpr uint32_t B_LBO = 128, B_SBO = 2048;

__device__ __forceinline__ uint32_t idesc_mn(int m, int n) { return make_idesc_bf16(m, n) | (1u << 16); }

// mode 0: correctness (TS, 3-pass).  mode 1: time TS.  mode 2: time SS.  mode 3/4: as 1/2 with STS traffic from warps 1..7.
// mode 5: time an epilogue-like TMEM ld (128 x 128 fp32) + st (2 x 128 x 64 packed) by all 8 warps, no MMA.
__global__ void __launch_bounds__(256, 1) probe2(const float* __restrict__ Wt, const float* __restrict__ Act,
                                                 float* __restrict__ D, int N, int K, int mode, int reps,
                                                 long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sBhi = smem;                 // 32 KB
  unsigned char* sBlo = smem + 32768;         // 32 KB
  __nv_bfloat16* sAhi = reinterpret_cast<__nv_bfloat16*>(smem + 65536);   // 32 KB (SS modes)
  __nv_bfloat16* sAlo = sAhi + M * KMAX;                                   // 32 KB
  unsigned char* junk = smem + 131072;        // 32 KB of store target for the traffic modes
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, f = tid & 127, hh = tid >> 7;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { mbar_init(&mbar, 1); stop_flag = 0; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t acc = tmem, a_hi = tmem + 256, a_lo = tmem + 256 + 64;
  const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
  // ---- A: thread (m = f, hh) packs k in [64 hh, +64) -> 32 words
  {
    uint32_t vh[32], vl[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int k = 64 * hh + 2 * c;
      const float x0 = k < K ? Wt[f * KMAX + k] : 0.f, x1 = k + 1 < K ? Wt[f * KMAX + k + 1] : 0.f;
      split_pack(x0, x1, vh[c], vl[c]);
    }
    tmem_st32(a_hi + lane_addr + 32 * hh, vh);
    tmem_st32(a_lo + lane_addr + 32 * hh, vl);
    tmem_wait_st();
  }
  for (int idx = tid; idx < M * KMAX; idx += blockDim.x) {
    const int r = idx / KMAX, k = idx % KMAX;
    const float a = Wt[idx];
    const __nv_bfloat16 h = __float2bfloat16(a);
    sAhi[canon_index(r, k, M)] = h;
    sAlo[canon_index(r, k, M)] = __float2bfloat16(a - __bfloat162float(h));
  }
  // ---- B: thread (k = f, hh) writes row groups [8 hh, +8), 8 rows per 16-byte store
  for (int g8 = 8 * hh; g8 < 8 * hh + 8; ++g8) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int n0 = 8 * g8 + 2 * p;
      const float x0 = (n0 < N && f < K) ? Act[n0 * KMAX + f] : 0.f, x1 = (n0 + 1 < N && f < K) ? Act[(n0 + 1) * KMAX + f] : 0.f;
      split_pack(x0, x1, h[p], l[p]);
    }
    *reinterpret_cast<uint4*>(sBhi + g8 * B_SBO + f * 16) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(sBlo + g8 * B_SBO + f * 16) = make_uint4(l[0], l[1], l[2], l[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  if (mode == 5) {
    __syncthreads();
    t0 = clock64();
    float sink = 0.f;
    for (int r = 0; r < reps; ++r) {
      uint32_t va[32], vb[32];
      tmem_ld32(acc + lane_addr + 64 * hh, va);
      tmem_ld32(acc + lane_addr + 64 * hh + 32, vb);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; ++c) { va[c] += vb[c]; }
      tmem_st32(tmem + 128 + lane_addr + 32 * hh, va);
      tmem_st32(tmem + 192 + lane_addr + 32 * hh, vb);
      tmem_wait_st();
      sink += __uint_as_float(va[0]);
    }
    __syncthreads();
    t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    if (sink == 123.456f) D[0] = sink;
  } else if (tid == 0) {
    const uint32_t idesc = idesc_mn(M, N);
    const uint32_t idesc_ss = idesc;
    const uint32_t lboA = (M / 8) * 128;
    const int nk = K / 16;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int ks = 0; ks < nk; ++ks) {
        const uint64_t bh = make_sdesc(smem_u32(sBhi) + ks * 2 * B_LBO, B_LBO, B_SBO);
        const uint64_t bl = make_sdesc(smem_u32(sBlo) + ks * 2 * B_LBO, B_LBO, B_SBO);
        const uint32_t first = (r == 0 && ks == 0) ? 0u : 1u;
        if (mode == 2 || mode == 4) {
          const uint64_t ah = make_sdesc(smem_u32(sAhi) + ks * 2 * lboA, lboA, 128);
          const uint64_t al = make_sdesc(smem_u32(sAlo) + ks * 2 * lboA, lboA, 128);
          mma_ss(acc, ah, bh, idesc_ss, first);
          mma_ss(acc, al, bh, idesc_ss, 1);
          mma_ss(acc, ah, bl, idesc_ss, 1);
        } else {
          mma_ts(acc, a_hi + ks * 8, bh, idesc, first);
          mma_ts(acc, a_lo + ks * 8, bh, idesc, 1);
          mma_ts(acc, a_hi + ks * 8, bl, idesc, 1);
        }
      }
    }
    mma_commit(&mbar);
    mbar_wait(&mbar, 0);
    t1 = clock64();
    cycles[0] = t1 - t0;
    stop_flag = 1;
  } else if ((mode == 3 || mode == 4) && warp > 0) {
    // shared-memory store traffic while the MMAs run: every thread streams 16-byte stores (conflict-free)
    uint4 val = make_uint4(tid, tid, tid, tid);
    int it = 0;
    while (!stop_flag) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(junk + ((it + u) & 7) * 4096 + (tid & 255) * 16)),
                     "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w)
                     : "memory");
      ++it;
    }
  }
  __syncthreads();
  tc_fence_after();
  if (mode == 0) {
    for (int q = 0; q < 2; ++q) {
      uint32_t v[32];
      tmem_ld32(acc + lane_addr + 64 * hh + 32 * q, v);
      tmem_wait_ld();
      for (int c = 0; c < 32; ++c) D[f * NMAX + 64 * hh + 32 * q + c] = __uint_as_float(v[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<float> hW(M * KMAX), hA(NMAX * KMAX), hD(M * NMAX);
  srand(7);
  for (auto& v : hW) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
  for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dW, *dA, *dD;
  long long* dC;
  cudaMalloc(&dW, hW.size() * 4); cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dD, hD.size() * 4); cudaMalloc(&dC, 64);
  cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 5 * 32768;
  cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int bad = 0;
  const int cases[4][2] = {{128, 128}, {112, 128}, {128, 64}, {80, 64}};
  for (auto& cs : cases) {
    const int N = cs[0], K = cs[1];
    cudaMemset(dD, 0, hD.size() * 4);
    probe2<<<1, 256, smem>>>(dW, dA, dD, N, K, 0, 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d: CUDA error %s\n", N, K, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)hW[m * KMAX + k] * (double)hA[n * KMAX + k];
        max_err = fmax(max_err, fabs(ref - hD[m * NMAX + n]));
        max_ref = fmax(max_ref, fabs(ref));
      }
    const double rel = max_err / max_ref;
    printf("TS (A=W^T in TMEM, B MN-major smem) 3-pass N=%d K=%d: max abs err %.3e, max |ref| %.3e, rel %.3e  %s\n", N, K,
           max_err, max_ref, rel, rel < 3e-5 ? "OK" : "MISMATCH");
    bad += rel >= 3e-5;
  }
  const char* names[6] = {"", "TS", "SS", "TS + STS traffic", "SS + STS traffic", "epilogue ld+st"};
  for (int mode = 1; mode <= 5; ++mode) {
    for (int N : {128, 64}) {
      if (mode == 5 && N != 128) continue;
      const int reps = mode == 5 ? 64 : 32;
      probe2<<<1, 256, smem>>>(dW, dA, dD, N, 128, mode, reps, dC);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long cyc;
      cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
      if (mode == 5) printf("%-18s: %lld cycles for %d x (ld 128x128 fp32 + st 128x128 b32): %.0f cyc each\n", names[mode], cyc, reps, (double)cyc / reps);
      else printf("%-18s N=%3d: %lld cycles for %d x 24 MMAs (K=128, 3-pass): %.1f cyc / MMA, %.0f cyc / layer\n", names[mode], N, cyc, reps,
                  (double)cyc / (reps * 24), (double)cyc / reps);
    }
  }
  return bad;
}

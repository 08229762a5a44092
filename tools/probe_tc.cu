// Hardware probe for the tcgen05 building blocks used by ecnf_solve_tc: validates, against a CPU product,
//   (1) SS mode: A and B from shared memory in the no-swizzle K-major canonical layout (8x16B core matrices),
//   (2) TS mode: A from tensor memory (lane = row, 2 bf16 per 32-bit column), written with tcgen05.st.32x32b,
//   (3) the 3-pass bf16 split (hi*hi + lo*hi + hi*lo) against an fp64 product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_tc tools/probe_tc.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ecnf_b200/csrc/ecnf_tc.cuh"

using namespace ecnf_tc;

constexpr int M = 128, N = 128, K = 128;

// mode 0: SS single pass; 1: TS single pass; 2: TS 3-pass split
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                       float* __restrict__ D, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __nv_bfloat16* sAhi = reinterpret_cast<__nv_bfloat16*>(smem);              // 32 KB
  __nv_bfloat16* sBhi = sAhi + M * K;                                         // 32 KB
  __nv_bfloat16* sBlo = sBhi + N * K;                                         // 32 KB
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // canonical images: elem (row, k) at ((k/8) * (ROWS/8) + row/8) * 64 + (row%8) * 8 + (k%8)
  for (int idx = tid; idx < M * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    const float a = A[idx];
    sAhi[canon_index(r, k, M)] = __float2bfloat16(a);
  }
  for (int idx = tid; idx < N * K; idx += blockDim.x) {
    const int n = idx / K, k = idx % K;   // B given as [N][K] (K contiguous)
    const float b = B[idx];
    const __nv_bfloat16 hi = __float2bfloat16(b);
    sBhi[canon_index(n, k, N)] = hi;
    sBlo[canon_index(n, k, N)] = __float2bfloat16(b - __bfloat162float(hi));
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) mbar_init(&mbar, 1);
  fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t acc = tmem;            // columns [0,128)
  const uint32_t a_hi = tmem + 128;     // columns [128,192): A hi, 2 bf16 per column
  const uint32_t a_lo = tmem + 192;     // columns [192,256)
  if (mode >= 1) {
    // each thread owns row `tid`: pack (k, k+1) into one 32-bit column and store 32 columns at a time
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int half = 0; half < 2; ++half) {
      uint32_t vh[32], vl[32];
      for (int c = 0; c < 32; ++c) {
        const int k = 2 * (half * 32 + c);
        const float x0 = A[tid * K + k], x1 = A[tid * K + k + 1];
        split_pack(x0, x1, vh[c], vl[c]);
      }
      tmem_st32(a_hi + lane_addr + half * 32, vh);
      tmem_st32(a_lo + lane_addr + half * 32, vl);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint32_t sbo = 128, lboA = (M / 8) * 128, lboB = (N / 8) * 128;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bd_hi = make_sdesc(smem_u32(sBhi) + ks * 2 * lboB, lboB, sbo);
      const uint64_t bd_lo = make_sdesc(smem_u32(sBlo) + ks * 2 * lboB, lboB, sbo);
      if (mode == 0) {
        const uint64_t ad = make_sdesc(smem_u32(sAhi) + ks * 2 * lboA, lboA, sbo);
        mma_ss(acc, ad, bd_hi, idesc, ks > 0);
      } else if (mode == 1) {
        mma_ts(acc, a_hi + ks * 8, bd_hi, idesc, ks > 0);
      } else {
        mma_ts(acc, a_hi + ks * 8, bd_hi, idesc, ks > 0);
        mma_ts(acc, a_lo + ks * 8, bd_hi, idesc, 1);
        mma_ts(acc, a_hi + ks * 8, bd_lo, idesc, 1);
      }
    }
    mma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after();
  {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int q = 0; q < 4; ++q) {
      uint32_t v[32];
      tmem_ld32(acc + lane_addr + q * 32, v);
      tmem_wait_ld();
      for (int c = 0; c < 32; ++c) D[tid * N + q * 32 + c] = __uint_as_float(v[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  std::vector<float> hA(M * K), hB(N * K), hD(M * N);
  srand(1);
  for (auto& v : hA) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : hB) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 3 * 32 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  auto bf = [](float x) { return __bfloat162float(__float2bfloat16(x)); };
  int bad = 0;
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(dD, 0, hD.size() * 4);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) {
          const double a = mode == 2 ? (double)hA[m * K + k] : (double)bf(hA[m * K + k]);
          const double b = mode == 2 ? (double)hB[n * K + k] : (double)bf(hB[n * K + k]);
          ref += a * b;
        }
        max_err = fmax(max_err, fabs(ref - hD[m * N + n]));
        max_ref = fmax(max_ref, fabs(ref));
      }
    const double rel = max_err / max_ref;
    const double tol = mode == 2 ? 3e-5 : 2e-6;
    printf("mode %d (%s): max abs err %.3e, max |ref| %.3e, rel %.3e  %s\n", mode,
           mode == 0 ? "SS bf16" : mode == 1 ? "TS bf16" : "TS 3-pass split vs fp64", max_err, max_ref, rel,
           rel < tol ? "OK" : "MISMATCH");
    bad += rel >= tol;
  }
  return bad;
}

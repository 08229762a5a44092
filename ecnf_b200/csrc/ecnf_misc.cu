// Small bandwidth-bound kernels around the hot path: base distribution, noise, Adam, ESS statistics, targets.
#include <cmath>

#include "ecnf_common.cuh"

namespace {

// ---------------- Philox4x32-10 (counter based; stream = (seed, substream), counter = (row, chunk)) --------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t row, uint32_t chunk, uint32_t sub, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)row, (uint32_t)(row >> 32), chunk, sub};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}
__device__ __forceinline__ float u01_open(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)
__device__ __forceinline__ float u01_half(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }          // [0,1)

// standard normal #i of row `row` (Box-Muller on a Philox block; 4 normals per block)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t row, int i, uint32_t sub) {
  uint32_t r[4];
  philox4x32(seed, row, (uint32_t)(i >> 2), sub, r);
  const int p = (i >> 1) & 1;
  const float u1 = u01_open(r[2 * p]), u2 = u01_open(r[2 * p + 1]);
  const float rad = sqrtf(-2.f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  return (i & 1) ? rad * s : rad * c;
}

// x0 = s * (eps - mean_nodes eps)      zero_com_base.py:88-93 + ScalarAffine (build_cnf.py:46-48)
__global__ void base_sample_kernel(uint64_t seed, int64_t goff, int64_t B, int n, int dim, float scale,
                                   const float* __restrict__ eps_in, float* __restrict__ out, uint32_t sub) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (b >= B) return;
  const int D = n * dim, lane = threadIdx.x;
  float v[3] = {0.f, 0.f, 0.f};  // up to 96 values per row, 32 lanes
  for (int q = 0; q < 3; ++q) {
    const int i = lane + 32 * q;
    if (i < D) v[q] = eps_in ? eps_in[b * D + i] : philox_normal(seed, (uint64_t)(goff + b), i, sub);
  }
  // per-coordinate mean over nodes: element i belongs to coordinate i % dim
  for (int c = 0; c < dim; ++c) {
    float s = 0.f;
    for (int q = 0; q < 3; ++q) {
      const int i = lane + 32 * q;
      if (i < D && i % dim == c) s += v[q];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)n;
    for (int q = 0; q < 3; ++q) {
      const int i = lane + 32 * q;
      if (i < D && i % dim == c) v[q] -= mean;
    }
  }
  for (int q = 0; q < 3; ++q) {
    const int i = lane + 32 * q;
    if (i < D) out[b * D + i] = scale * v[q];
  }
}

// plain N(0, 1) draws keyed by the global sample index (Hutchinson probes)
__global__ void normal_kernel(uint64_t seed, int64_t goff, int64_t B, int D, float* __restrict__ out, uint32_t sub) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int64_t b = idx / D;
  out[idx] = philox_normal(seed, (uint64_t)(goff + b), (int)(idx - b * D), sub);
}

__global__ void uniform_kernel(uint64_t seed, int64_t goff, int64_t B, float* __restrict__ out, uint32_t sub) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint32_t r[4];
  philox4x32(seed, (uint64_t)(goff + b), 0u, sub, r);
  out[b] = u01_half(r[0]);
}

// log p0(x) = -1/2 |rm(x/s)|^2 - 1/2 (n-1) dim log 2pi - (n-1) dim log s
__global__ void base_log_prob_kernel(const float* __restrict__ x, int64_t B, int n, int dim, float scale,
                                     float* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (b >= B) return;
  const int D = n * dim, lane = threadIdx.x;
  float v[3] = {0.f, 0.f, 0.f};
  for (int q = 0; q < 3; ++q) {
    const int i = lane + 32 * q;
    if (i < D) v[q] = x[b * D + i] / scale;
  }
  float r2 = 0.f;
  for (int c = 0; c < dim; ++c) {
    float s = 0.f;
    for (int q = 0; q < 3; ++q) {
      const int i = lane + 32 * q;
      if (i < D && i % dim == c) s += v[q];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)n;
    for (int q = 0; q < 3; ++q) {
      const int i = lane + 32 * q;
      if (i < D && i % dim == c) { const float z = v[q] - mean; r2 = fmaf(z, z, r2); }
    }
  }
  for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
  if (lane == 0) {
    const float dof = (float)((n - 1) * dim);
    out[b] = -0.5f * r2 - 0.5f * dof * 1.8378770664093453f - dof * logf(scale);
  }
}

// ---------------- Adam + norms + EMA (gradient_step.py:39-50; optax.adam) ---------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mu,
                            float* __restrict__ nu, float* __restrict__ ema, int64_t count, float lr, float b1,
                            float b2, float eps, float bc1, float bc2, float ema_beta, float* __restrict__ sums) {
  float gs = 0.f, us = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float m = b1 * mu[i] + (1.f - b1) * gi;
    const float v = b2 * nu[i] + (1.f - b2) * gi * gi;
    mu[i] = m;
    nu[i] = v;
    const float upd = -lr * (m / bc1) / (sqrtf(v / bc2) + eps);
    const float np = p[i] + upd;
    p[i] = np;
    if (ema) ema[i] = ema[i] * ema_beta + (1.f - ema_beta) * np;
    gs = fmaf(gi, gi, gs);
    us = fmaf(upd, upd, us);
  }
  for (int o = 16; o > 0; o >>= 1) {
    gs += __shfl_xor_sync(0xffffffffu, gs, o);
    us += __shfl_xor_sync(0xffffffffu, us, o);
  }
  __shared__ float sg[32], su[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { sg[w] = gs; su[w] = us; }
  __syncthreads();
  if (w == 0) {
    gs = l < (blockDim.x >> 5) ? sg[l] : 0.f;
    us = l < (blockDim.x >> 5) ? su[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) {
      gs += __shfl_xor_sync(0xffffffffu, gs, o);
      us += __shfl_xor_sync(0xffffffffu, us, o);
    }
    if (l == 0) { atomicAdd(&sums[0], gs); atomicAdd(&sums[1], us); }
  }
}
__global__ void zero2_kernel(float* s) { s[0] = 0.f; s[1] = 0.f; }
__global__ void sqrt2_kernel(float* s) { s[0] = sqrtf(s[0]); s[1] = sqrtf(s[1]); }

// ---------------- ESS sufficient statistics ---------------------------------------------------------------
// single CTA, two passes (N is at most a few 1e5..1e6 floats); out5 = {max w, S1, S2, max(-w), S1'}
__global__ void ess_stats_kernel(const float* __restrict__ w, int64_t N, float* __restrict__ out5) {
  __shared__ float sa[32], sb[32], sc[32];
  __shared__ float s_max, s_nmax;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  float mx = -INFINITY, nmx = -INFINITY;
  for (int64_t i = tid; i < N; i += blockDim.x) { mx = fmaxf(mx, w[i]); nmx = fmaxf(nmx, -w[i]); }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nmx = fmaxf(nmx, __shfl_xor_sync(0xffffffffu, nmx, o));
  }
  if (lane == 0) { sa[warp] = mx; sb[warp] = nmx; }
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < nw; ++i) { mx = fmaxf(mx, sa[i]); nmx = fmaxf(nmx, sb[i]); }
    s_max = mx; s_nmax = nmx;
  }
  __syncthreads();
  mx = s_max; nmx = s_nmax;
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int64_t i = tid; i < N; i += blockDim.x) {
    const float e = expf(w[i] - mx);
    s1 += e;
    s2 = fmaf(e, e, s2);
    s3 += expf(-w[i] - nmx);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  __syncthreads();
  if (lane == 0) { sa[warp] = s1; sb[warp] = s2; sc[warp] = s3; }
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < nw; ++i) { s1 += sa[i]; s2 += sb[i]; s3 += sc[i]; }
    out5[0] = mx; out5[1] = s1; out5[2] = s2; out5[3] = nmx; out5[4] = s3;
  }
}

// ---------------- target energies: one warp per sample ----------------------------------------------------
__global__ void target_log_prob_kernel(int kind, const float* __restrict__ x, int64_t B, int n, int dim,
                                       float* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
  if (b >= B) return;
  const int lane = threadIdx.x;
  const float* xb = x + b * n * dim;
  const int E = n * (n - 1);
  float acc = 0.f;
  for (int e = lane; e < E; e += 32) {
    const int i = e / (n - 1);
    int j = i + 1 + (e - i * (n - 1));
    if (j >= n) j -= n;
    float s = 0.f;
    for (int c = 0; c < dim; ++c) { const float v = xb[j * dim + c] - xb[i * dim + c]; s = fmaf(v, v, s); }
    const float d = sqrtf(s == 0.f ? 1.f : s);
    if (kind == ECNF_TARGET_LJ) {
      const float ir = 1.f / d, ir2 = ir * ir, ir6 = ir2 * ir2 * ir2;
      acc += ir6 * ir6 - 2.f * ir6;                    // leonard_jones.py:19 (r = 1)
    } else {
      const float dd = d - 4.f, d2 = dd * dd;
      acc += -4.f * d2 + 0.9f * d2 * d2;               // double_well.py:16 (a=0, b=-4, c=0.9)
    }
  }
  float harm = 0.f;
  if (kind == ECNF_TARGET_LJ) {
    for (int c = 0; c < dim; ++c) {
      float mean = 0.f;
      for (int i = 0; i < n; ++i) mean += xb[i * dim + c];
      mean /= (float)n;
      for (int i = lane; i < n; i += 32) { const float z = xb[i * dim + c] - mean; harm = fmaf(z, z, harm); }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    harm += __shfl_xor_sync(0xffffffffu, harm, o);
  }
  if (lane == 0) out[b] = (kind == ECNF_TARGET_LJ) ? -(0.5f * acc + 0.5f * harm) : -(0.5f * acc);
}

}  // namespace

extern "C" {

int ecnf_base_sample(const ecnf_model* m, uint64_t seed, int64_t goff, int64_t B, float* out, void* stream) {
  if (!m || !out || B < 0) { ecnf_set_error("ecnf_base_sample: bad argument"); return ECNF_ERR_INVALID; }
  if (B == 0) return ECNF_OK;
  dim3 blk(32, 8);
  base_sample_kernel<<<(unsigned)((B + 7) / 8), blk, 0, (cudaStream_t)stream>>>(seed, goff, B, m->cfg.n_frames, m->cfg.dim,
                                                                              m->cfg.base_scale, nullptr, out, 0u);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_normal_noise(const ecnf_model* m, uint64_t seed, int64_t goff, int64_t B, uint32_t substream, float* out, void* stream) {
  if (!m || !out || B < 0) { ecnf_set_error("ecnf_normal_noise: bad argument"); return ECNF_ERR_INVALID; }
  if (B == 0) return ECNF_OK;
  const int D = m->cfg.n_frames * m->cfg.dim;
  const int64_t tot = B * D;
  normal_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, goff, B, D, out, substream);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_base_sample_from_noise(const ecnf_model* m, const float* eps, int64_t B, float* out, void* stream) {
  if (!m || !out || !eps || B < 0) { ecnf_set_error("ecnf_base_sample_from_noise: bad argument"); return ECNF_ERR_INVALID; }
  if (B == 0) return ECNF_OK;
  dim3 blk(32, 8);
  base_sample_kernel<<<(unsigned)((B + 7) / 8), blk, 0, (cudaStream_t)stream>>>(0, 0, B, m->cfg.n_frames, m->cfg.dim,
                                                                              m->cfg.base_scale, eps, out, 0u);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_base_log_prob(const ecnf_model* m, const float* x, int64_t B, float* out, void* stream) {
  if (!m || !out || !x || B < 0) { ecnf_set_error("ecnf_base_log_prob: bad argument"); return ECNF_ERR_INVALID; }
  if (B == 0) return ECNF_OK;
  dim3 blk(32, 8);
  base_log_prob_kernel<<<(unsigned)((B + 7) / 8), blk, 0, (cudaStream_t)stream>>>(x, B, m->cfg.n_frames, m->cfg.dim,
                                                                                m->cfg.base_scale, out);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_fm_draw_noise(const ecnf_model* m, uint64_t seed, int64_t goff, int64_t B, float* out_x0, float* out_t,
                       void* stream) {
  if (!m || !out_x0 || !out_t || B < 0) { ecnf_set_error("ecnf_fm_draw_noise: bad argument"); return ECNF_ERR_INVALID; }
  if (B == 0) return ECNF_OK;
  dim3 blk(32, 8);
  base_sample_kernel<<<(unsigned)((B + 7) / 8), blk, 0, (cudaStream_t)stream>>>(seed, goff, B, m->cfg.n_frames, m->cfg.dim,
                                                                              m->cfg.base_scale, nullptr, out_x0, 0u);
  uniform_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, goff, B, out_t, 1u);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_adam_step(float* params, const float* grad, float* mu, float* nu, float* ema, int64_t count, int64_t step,
                   float lr, float b1, float b2, float eps, float ema_beta, float* out_norms, void* stream) {
  if (!params || !grad || !mu || !nu || !out_norms || count <= 0) { ecnf_set_error("ecnf_adam_step: bad argument"); return ECNF_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const float bc1 = 1.f - (float)pow((double)b1, (double)(step + 1));
  const float bc2 = 1.f - (float)pow((double)b2, (double)(step + 1));
  zero2_kernel<<<1, 1, 0, st>>>(out_norms);
  int grid = (int)((count + 256 * 4 - 1) / (256 * 4));
  if (grid > 148 * 8) grid = 148 * 8;
  adam_kernel<<<grid, 256, 0, st>>>(params, grad, mu, nu, ema, count, lr, b1, b2, eps, bc1, bc2, ema_beta, out_norms);
  sqrt2_kernel<<<1, 1, 0, st>>>(out_norms);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_ess_stats(const float* log_w, int64_t N, float* out5, void* stream) {
  if (!log_w || !out5 || N <= 0) { ecnf_set_error("ecnf_ess_stats: bad argument"); return ECNF_ERR_INVALID; }
  ess_stats_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(log_w, N, out5);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

int ecnf_target_log_prob(int kind, const float* x, int64_t B, int n_frames, int dim, float* out, void* stream) {
  if (!x || !out || B < 0 || (kind != ECNF_TARGET_LJ && kind != ECNF_TARGET_DW)) { ecnf_set_error("ecnf_target_log_prob: bad argument"); return ECNF_ERR_INVALID; }
  if (B == 0) return ECNF_OK;
  dim3 blk(32, 8);
  target_log_prob_kernel<<<(unsigned)((B + 7) / 8), blk, 0, (cudaStream_t)stream>>>(kind, x, B, n_frames, dim, out);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

}  // extern "C"

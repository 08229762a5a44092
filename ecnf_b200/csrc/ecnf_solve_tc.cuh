// Engine A on the 5th-generation tensor cores (tcgen05 + TMEM), for mlp_units = 128, n_invariant_feat_hidden = 64
// (DW4 / LJ13 shapes) with the exact divergence.
//
// Same algorithm, tile structure and tangent-row scheme as the fp32 SIMT engine (ecnf_solve_impl.cuh); what changes
// is where the rows live and who multiplies:
//   * a row tile is 128 rows = the 128 lanes of tensor memory; thread t owns row (t & 127) and the 64 columns
//     [64*(t>>7), +64) of it -- the natural tcgen05.ld/st 32x32b ownership, so no shuffles anywhere;
//   * activations never touch shared memory between layers: accumulator (TMEM, fp32) -> registers -> bias/SiLU or the
//     tangent rule -> split into bf16 (hi, lo) -> A operand written back to TMEM with tcgen05.st (2 bf16 per column);
//   * every Dense layer is 3 x (K/16) tcgen05.mma (A from TMEM, B from shared memory): hi*hi + lo*hi + hi*lo with
//     fp32 accumulation, i.e. ~2^-16 relative error per product (measured 4e-6 by tools/probe_tc.cu) -- single-pass
//     bf16/tf32 misses the 1e-4 tolerance on log q;
//   * weights are pre-split into bf16 (hi, lo) images in the MMA's canonical no-swizzle K-major layout by a prep kernel
//     and streamed from L2 with cp.async.bulk into two 64 KB buffers, one layer ahead;
//   * two row tiles are in flight (TMEM holds 2 x (128 accumulator + 128 operand columns)): while the 256 threads run
//     the epilogue of one tile, the tensor core multiplies the other.
#pragma once
#include "ecnf_solve_impl.cuh"
#include "ecnf_tc.cuh"

namespace ecnf_solve_detail {

using namespace ecnf_tc;

constexpr int TCU = 128, TCH = 64;
constexpr int TC_GMAX = 32;        // groups (edges / nodes) per tile
constexpr int TC_WBYTES = 65536;   // one weight buffer: hi + lo image of a 128 x 128 layer
constexpr int TC_SLD = 65;          // staging row stride (floats): odd, so lanes = rows is conflict-free

__host__ __device__ inline TcSmemLayout make_tc_layout(int n, int dim, int n_layers) {
  const int D = n * dim, S = D + 1;
  TcSmemLayout L;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) & ~127; return r; };
  L.Wb0 = take(TC_WBYTES);
  L.Wb1 = take(TC_WBYTES);
  L.stage = take(128 * TC_SLD * 4);        // staging [128 rows][65]; also the primal-activation hand-over buffer
  L.G = take(TC_GMAX * TCU * 4);
  L.macc = take((1 + D) * TCU * 4);       // aggregated messages of the receiver being processed
  L.vecs = take((2 * n_layers - 1 + 3) * TCU * 4);       // per-block vectors: layer biases, w_d, attention / head weights
  L.xt = take(D * D * 4);
  L.xtacc = take(D * D * 4);
  L.dacc = take(D * 4);
  L.xs = take(D * 4);
  L.xs0 = take(D * 4);
  L.xacc = take(D * 4);
  L.mu = take(16);
  L.tau = take(ECNF_MAX_T * 4);
  L.cvec = take(64 * 4);
  L.ode = take(11 * S * 4);
  L.red = take((S + 16) * 4);
  L.rowsd = take(2 * 128 * 4);
  L.rowslot = take(2 * 128 * 4);
  L.rowgrp = take(2 * 128 * 4);
  L.rdot = take(2 * 2 * 128 * 4);
  L.gi = take(2 * TC_GMAX * 4);
  L.gj = take(2 * TC_GMAX * 4);
  L.giz = take(2 * TC_GMAX * 4);
  L.gv = take(2 * TC_GMAX * 3 * 4);
  L.gs1 = take(2 * TC_GMAX * 4);
  L.glen = take(2 * TC_GMAX * 4);
  L.ginv = take(2 * TC_GMAX * 4);
  L.ge = take(2 * TC_GMAX * 4);
  L.bars = take(64);
  L.prof = take(32 * 8);
  L.total_bytes = o;
  return L;
}

extern __shared__ __align__(128) unsigned char smem_tc[];

// Every buffer is addressed as (shared-memory base + offset from the kernel parameters): the offsets are constant-bank
// loads, so no pointer has to be kept alive in (or spilled from) registers across the very large inlined body.
#define TCF(field) (reinterpret_cast<float*>(smem_tc + a.lay.field))
#define TCI(field) (reinterpret_cast<int*>(smem_tc + a.lay.field))
#define stage TCF(stage)
#define G TCF(G)
#define macc TCF(macc)
#define vecs TCF(vecs)
#define xt TCF(xt)
#define xtacc TCF(xtacc)
#define dacc TCF(dacc)
#define xs TCF(xs)
#define xs0 TCF(xs0)
#define xacc TCF(xacc)
#define mu TCF(mu)
#define tau TCF(tau)
#define cvec TCF(cvec)
#define rowsd TCF(rowsd)
#define rdot TCF(rdot)
#define gv TCF(gv)
#define gs1 TCF(gs1)
#define glen TCF(glen)
#define ginv TCF(ginv)
#define ge TCF(ge)
#define rowslot TCI(rowslot)
#define rowgrp TCI(rowgrp)
#define gi TCI(gi)
#define gj TCI(gj)
#define giz TCI(giz)
#define mbar_mma (reinterpret_cast<uint64_t*>(smem_tc + a.lay.bars))
#define mbar_w (reinterpret_cast<uint64_t*>(smem_tc + a.lay.bars) + 2)
#define tmem_slot (reinterpret_cast<uint32_t*>(smem_tc + a.lay.bars + 32))
#define prof_s (reinterpret_cast<long long*>(smem_tc + a.lay.prof))
#define hA (a.scratch + (size_t)blockIdx.x * a.scratch_stride)
#define hB (hA + (size_t)n * ND * H)
#define Ps (hB + (size_t)n * ND * H)
#define Pr (Ps + (size_t)n * ND * U)
#define Mg (Pr + (size_t)n * ND * U)

struct EngineTC {
  static constexpr int U = TCU, H = TCH;
  static constexpr int NT = NTHREADS;
  const KernelArgs& a;     // the __grid_constant__ kernel parameter
  const EcnfModelDev& m;
  const TcImages& img;
  const int n, dim, D, ND, E;
  const int tid, row, ch, warp;
  uint32_t tmem;         // TMEM base address
  uint32_t lane_addr;    // (32 * (warp & 3)) << 16
  uint32_t ph_mma0, ph_mma1;   // completed-phase counters of the two MMA barriers
  uint32_t wq_head;      // weight loads issued so far (monotonic; buffer = seq & 1, parity = (seq >> 1) & 1)
  // coarse cycle counters (compile with -DECNF_TC_PROFILE; CTA 0 dumps them into the workspace header, tools/tc_profile.py)
  enum { P_NODE_PRE, P_NODE_POST, P_META, P_BUILD, P_WAIT_MMA, P_EPI, P_MSG, P_SYNC_ISSUE, P_COORDS, P_EDGE_INIT, P_EVAL_MISC,
         P_EPI_LD, P_EPI_ACT, P_EPI_ST, P_MSG_DOT, P_MSG_STAGE, P_MSG_LOOP, P_MSG_SEG, P_MSG_FLUSH, P_BUILD_GATHER, P_BUILD_ACT, P_NCOUNT };
#ifdef ECNF_TC_PROFILE
  long long prof_t0, prof_t1;
  __device__ __forceinline__ void pbeg() { prof_t0 = clock64(); }
  __device__ __forceinline__ void pend(int k) { const long long t1 = clock64(); if (tid == 0) prof_s[k] += t1 - prof_t0; prof_t0 = t1; }
  __device__ __forceinline__ void qbeg() { prof_t1 = clock64(); }
  __device__ __forceinline__ void qend(int k) { const long long t1 = clock64(); if (tid == 0) prof_s[k] += t1 - prof_t1; prof_t1 = t1; }
#else
  __device__ __forceinline__ void pbeg() {}
  __device__ __forceinline__ void pend(int) {}
  __device__ __forceinline__ void qbeg() {}
  __device__ __forceinline__ void qend(int) {}
#endif

  __device__ __forceinline__ float* ode_ptr() const { return TCF(ode); }
  __device__ __forceinline__ float* red_ptr() const { return TCF(red); }

  __device__ __forceinline__ EngineTC(const KernelArgs& a_)
      : a(a_), m(a_.m), img(a_.img), n(a_.m.n), dim(a_.m.dim), D(a_.m.n * a_.m.dim), ND(1 + a_.m.n * a_.m.dim),
        E(a_.m.n * (a_.m.n - 1)), tid(threadIdx.x), row(threadIdx.x & 127), ch(threadIdx.x >> 7), warp(threadIdx.x >> 5) {
    if (tid == 0) {
      mbar_init(&mbar_mma[0], 1); mbar_init(&mbar_mma[1], 1);
      mbar_init(&mbar_w[0], 1); mbar_init(&mbar_w[1], 1);
      for (int k = 0; k < P_NCOUNT; ++k) prof_s[k] = 0;
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem = *tmem_slot;
    lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    ph_mma0 = ph_mma1 = 0;
    wq_head = 0;
  }
  __device__ __forceinline__ void finish(long long* out) {
    if (out && tid == 0 && blockIdx.x == 0)
      for (int k = 0; k < P_NCOUNT; ++k) out[k] = prof_s[k];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
  }

  // ---- TMEM map: slot s -> accumulator [256 s, +128), A hi [256 s + 128, +64), A lo [256 s + 192, +64)
  __device__ __forceinline__ uint32_t acc_of(int s) const { return tmem + 256u * s; }
  __device__ __forceinline__ uint32_t ahi_of(int s) const { return tmem + 256u * s + 128u; }
  __device__ __forceinline__ uint32_t alo_of(int s) const { return tmem + 256u * s + 192u; }

  // ---- weight queue ------------------------------------------------------------------------------------------
  // Issue the load of one weight image (hi + lo) into buffer (seq & 1).  The caller guarantees that every MMA that
  // read the previous content of that buffer has completed.
  __device__ __forceinline__ void wq_load(int img_off, int bytes) {
    const uint32_t seq = wq_head++;
    if (tid == 0) {
      uint64_t* bar = &mbar_w[seq & 1];
      mbar_expect_tx(bar, (uint32_t)bytes);
      unsigned char* dst = smem_tc + a.lay.Wb0 + (seq & 1) * TC_WBYTES;
      const unsigned char* src = img.base + img_off;
      for (int o = 0; o < bytes; o += 16384) bulk_g2s(dst + o, src + o, (uint32_t)min(16384, bytes - o), bar);
    }
    __syncwarp();
  }
  // thread 0 only: block until the image of sequence number `seq` has landed; returns its shared-memory address
  __device__ __forceinline__ uint32_t wq_wait(uint32_t seq) {
    mbar_wait(&mbar_w[seq & 1], (seq >> 1) & 1);
    return smem_u32(smem_tc + a.lay.Wb0) + (seq & 1) * TC_WBYTES;
  }

  // ---- MMA issue (thread 0): acc (+)= A[128 x K] (TMEM hi/lo) x W[K x N] (3-pass split) -------------------------
  // Called by ALL lanes of warp 0 (warp-uniform operands stay in uniform registers); one elected lane issues.
  __device__ __forceinline__ void issue_mma(uint32_t acc, uint32_t a_hi, uint32_t a_lo, uint32_t wsm, int K, int N,
                                            bool accumulate, uint64_t* commit_bar) {
    __syncwarp();
    if (!elect_one()) return;
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint32_t lbo = (uint32_t)(N / 8) * 128u;
    uint64_t bh = make_sdesc(wsm, lbo, 128);
    uint64_t bl = make_sdesc(wsm + (uint32_t)(K * N * 2), lbo, 128);
    const uint64_t step = (uint64_t)((2u * lbo) >> 4);   // the start-address field advances by two K chunks per MMA
    const int nk = K / 16;
    mma_ts(acc, a_hi, bh, idesc, accumulate ? 1u : 0u);
    mma_ts(acc, a_lo, bh, idesc, 1u);
    mma_ts(acc, a_hi, bl, idesc, 1u);
    for (int ks = 1; ks < nk; ++ks) {
      bh += step; bl += step; a_hi += 8; a_lo += 8;
      mma_ts(acc, a_hi, bh, idesc, 1u);
      mma_ts(acc, a_lo, bh, idesc, 1u);
      mma_ts(acc, a_hi, bl, idesc, 1u);
    }
    if (commit_bar) mma_commit(commit_bar);   // same thread as the MMAs: tracks their completion
  }
  // all threads: make this thread's TMEM stores/loads visible, CTA barrier, then the issuer may proceed
  __device__ __forceinline__ void tc_sync() {
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __device__ __forceinline__ void wait_mma(int s) {
    if (s == 0) { mbar_wait(&mbar_mma[0], ph_mma0 & 1); ph_mma0++; }
    else { mbar_wait(&mbar_mma[1], ph_mma1 & 1); ph_mma1++; }
    tc_fence_after();
  }

  // ---- per-thread row helpers ----------------------------------------------------------------------------------
  __device__ __forceinline__ void ld_acc64(uint32_t acc, float (&v)[64]) {
    uint32_t a[32], b[32];
    tmem_ld32(acc + lane_addr + 64u * ch, a);
    tmem_ld32(acc + lane_addr + 64u * ch + 32u, b);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 32; ++c) { v[c] = __uint_as_float(a[c]); v[32 + c] = __uint_as_float(b[c]); }
  }
  // write this thread's 64 K-values [64 ch, +64) of its row as bf16 (hi, lo) into the A operand of a slot
  __device__ __forceinline__ void st_a64(uint32_t a_hi, uint32_t a_lo, const float (&v)[64]) {
    uint32_t h[32], l[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) split_pack(v[2 * c], v[2 * c + 1], h[c], l[c]);
    tmem_st32(a_hi + lane_addr + 32u * ch, h);
    tmem_st32(a_lo + lane_addr + 32u * ch, l);
  }
  // bias + SiLU on primal rows / silu'(z_primal) * z on tangent rows.  The few primal rows of a tile hand their
  // pre-activations over through shared memory (PA = the staging buffer) so that ALL threads share the transcendental
  // work; G returns silu'.  Must be called by all 256 threads (CTA barriers inside).  valid = row < nrows.
  __device__ __forceinline__ void act_rule(float (&v)[64], const float* bias, bool valid, bool primal, int grp,
                                           int ngroups) {
    float* PA = stage;
    if (valid && primal) {
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b = *reinterpret_cast<const float4*>(bias + 64 * ch + 4 * c4);
        *reinterpret_cast<float4*>(PA + grp * U + 64 * ch + 4 * c4) =
            make_float4(v[4 * c4] + b.x, v[4 * c4 + 1] + b.y, v[4 * c4 + 2] + b.z, v[4 * c4 + 3] + b.w);
      }
    }
    __syncthreads();
    for (int idx = tid; idx < ngroups * U; idx += NTHREADS) {
      const float z = PA[idx];
      const float s = __fdividef(1.f, 1.f + __expf(-z));
      PA[idx] = z * s;
      G[idx] = s * (1.f + z * (1.f - s));
    }
    __syncthreads();
    if (valid) {
      const float* src = (primal ? PA : G) + grp * U + 64 * ch;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 g = *reinterpret_cast<const float4*>(src + 4 * c4);
        if (primal) {
          v[4 * c4] = g.x; v[4 * c4 + 1] = g.y; v[4 * c4 + 2] = g.z; v[4 * c4 + 3] = g.w;
        } else {
          v[4 * c4] *= g.x; v[4 * c4 + 1] *= g.y; v[4 * c4 + 2] *= g.z; v[4 * c4 + 3] *= g.w;
        }
      }
    }
  }
  __device__ __forceinline__ float dot64(const float (&v)[64], const float* w) const {
    float s = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 ww = *reinterpret_cast<const float4*>(w + 64 * ch + 4 * c4);
      s = fmaf(v[4 * c4], ww.x, s); s = fmaf(v[4 * c4 + 1], ww.y, s);
      s = fmaf(v[4 * c4 + 2], ww.z, s); s = fmaf(v[4 * c4 + 3], ww.w, s);
    }
    return s;
  }
  __device__ __forceinline__ void load_row64(const float* src, float (&v)[64]) const {
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 x = *reinterpret_cast<const float4*>(src + 4 * c4);
      v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
    }
  }
  __device__ __forceinline__ void store_row64(float* dst, const float (&v)[64]) const {
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4)
      *reinterpret_cast<float4*>(dst + 4 * c4) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
  }

  // =============================================================================================================
  // node phase 1: h_in = [h | tau] Wd + bd ; P_s = h_in We0[0:H] ; P_r = h_in We0[H:2H] + be0
  // =============================================================================================================
  __device__ __forceinline__ void node_pre(int b, bool htan) {
    const EcnfBlockParams& bp = m.blk[b];
    const TcImgBlock& ib = img.blk[b];
    if (tid < H) {
      float cv = bp.bd[tid];
      for (int k = 0; k < m.T; ++k) cv = fmaf(tau[k], bp.Wd[(H + k) * H + tid], cv);
      cvec[tid] = cv;
    }
    const int r = htan ? ND : 1;
    const int gpt = min(128 / r, TC_GMAX);
    for (int node0 = 0; node0 < n; node0 += gpt) {
      const int nn = min(gpt, n - node0), nrows = nn * r;
      const bool valid = row < nrows;
      const int g = valid ? row / r : 0, q = row - g * r;
      const bool primal = (q == 0);
      const uint32_t s0 = wq_head;
      wq_load(ib.Wd, 2 * H * H * 2);
      wq_load(ib.We0s, 2 * H * U * 2);
      float v[64];
      // A <- h rows (K = 64: only the ch == 0 half of the threads carries data)
      if (ch == 0) {
        if (valid) load_row64(hA + ((size_t)(node0 + g) * ND + q) * H, v);
        else {
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] = 0.f;
        }
        st_a64(ahi_of(0), alo_of(0), v);
      }
      tc_sync();
      if (warp == 0) { const uint32_t w = wq_wait(s0); issue_mma(acc_of(0), ahi_of(0), alo_of(0), w, H, H, false, &mbar_mma[0]); }
      wait_mma(0);
      wq_load(ib.We0r, 2 * H * U * 2);   // buffer of Wd is free again
      if (ch == 0) {
        ld_acc64(acc_of(0), v);
        if (valid && primal) {
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] += cvec[c];
        }
        if (valid) store_row64(hB + ((size_t)(node0 + g) * ND + q) * H, v);
        st_a64(ahi_of(0), alo_of(0), v);
      }
      tc_sync();
      if (warp == 0) {
        const uint32_t w1 = wq_wait(s0 + 1);
        issue_mma(acc_of(0), ahi_of(0), alo_of(0), w1, H, U, false, nullptr);
        const uint32_t w2 = wq_wait(s0 + 2);
        issue_mma(acc_of(1), ahi_of(0), alo_of(0), w2, H, U, false, &mbar_mma[0]);
      }
      wait_mma(0);
      ld_acc64(acc_of(0), v);
      if (valid) store_row64(Ps + ((size_t)(node0 + g) * ND + q) * U + 64 * ch, v);
      ld_acc64(acc_of(1), v);
      if (valid) {
        if (primal) {
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] += bp.be[0][64 * ch + c];
        }
        store_row64(Pr + ((size_t)(node0 + g) * ND + q) * U + 64 * ch, v);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }

  // =============================================================================================================
  // node phase 2: h <- phi_h([M | h_in]) + h_in
  // =============================================================================================================
  __device__ __forceinline__ void node_post(int b, bool htan) {
    const EcnfBlockParams& bp = m.blk[b];
    const TcImgBlock& ib = img.blk[b];
    const int L = m.L;
    const int r = ND;
    const int gpt = min(128 / r, TC_GMAX);
    for (int node0 = 0; node0 < n; node0 += gpt) {
      const int nn = min(gpt, n - node0), nrows = nn * r;
      const bool valid = row < nrows;
      const int g = valid ? row / r : 0, q = row - g * r;
      const bool primal = (q == 0);
      uint32_t seq = wq_head;
      wq_load(ib.Wh0m, 2 * U * U * 2);
      wq_load(ib.Wh0h, 2 * H * U * 2);
      float v[64];
      if (valid) load_row64(Mg + ((size_t)(node0 + g) * ND + q) * U + 64 * ch, v);
      else {
#pragma unroll
        for (int c = 0; c < 64; ++c) v[c] = 0.f;
      }
      st_a64(ahi_of(0), alo_of(0), v);
      tc_sync();
      if (warp == 0) { const uint32_t w = wq_wait(seq); issue_mma(acc_of(0), ahi_of(0), alo_of(0), w, U, U, false, &mbar_mma[0]); }
      wait_mma(0);
      wq_load(L > 1 ? ib.Wh[1] : ib.WhL, L > 1 ? 2 * U * U * 2 : 2 * U * H * 2);
      if (ch == 0) {
        if (valid && (primal || htan)) load_row64(hB + ((size_t)(node0 + g) * ND + q) * H, v);
        else {
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] = 0.f;
        }
        st_a64(ahi_of(0), alo_of(0), v);
      }
      tc_sync();
      if (warp == 0) { const uint32_t w = wq_wait(seq + 1); issue_mma(acc_of(0), ahi_of(0), alo_of(0), w, H, U, true, &mbar_mma[0]); }
      wait_mma(0);
      seq += 2;
      // hidden layers: epilogue of layer l-1, then MMA with Wh[l] (l = 1..L-1), finally WhL
      for (int l = 1; l <= L; ++l) {
        // the buffer of the weights consumed two steps ago is free: prefetch the image after the next one
        if (l + 1 <= L) wq_load(l + 1 < L ? ib.Wh[l + 1] : ib.WhL, l + 1 < L ? 2 * U * U * 2 : 2 * U * H * 2);
        ld_acc64(acc_of(0), v);
        act_rule(v, bp.bh[l - 1], valid, primal, g, nn);
        st_a64(ahi_of(0), alo_of(0), v);
        tc_sync();
        if (warp == 0) {
          const uint32_t w = wq_wait(seq);
          issue_mma(acc_of(0), ahi_of(0), alo_of(0), w, U, l < L ? U : H, false, &mbar_mma[0]);
        }
        wait_mma(0);
        ++seq;
      }
      if (ch == 0) {
        ld_acc64(acc_of(0), v);
        if (valid) {
          float hres[64];
          if (primal || htan) load_row64(hB + ((size_t)(node0 + g) * ND + q) * H, hres);
          else {
#pragma unroll
            for (int c = 0; c < 64; ++c) hres[c] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] += hres[c] + (primal ? bp.bh[L][c] : 0.f);
          store_row64(hA + ((size_t)(node0 + g) * ND + q) * H, v);
        }
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }

  // =============================================================================================================
  // edge phase
  // =============================================================================================================
  struct Tile {
    int e0, ng, nrows;
    bool active;
  };

  // per-edge geometry + per-row metadata of one tile into the slot's shared arrays (all threads; ends with a barrier)
  __device__ __forceinline__ void tile_meta(int s, const Tile& t, int kind, int r) {
    float* gvs = gv + s * TC_GMAX * 3;
    if (t.active && tid < t.ng) {
      const int e = t.e0 + tid, i = e / (n - 1), jj = e - i * (n - 1);
      int j = i + 1 + jj; if (j >= n) j -= n;
      float sq = 0.f;
      for (int c = 0; c < dim; ++c) {
        const float vv = xs[i * dim + c] - xs[j * dim + c];
        gvs[tid * 3 + c] = vv;
        sq = fmaf(vv, vv, sq);
      }
      const int isz = (sq == 0.f);
      const float s1 = isz ? 1.f : sq;
      const float len = sqrtf(s1);
      gi[s * TC_GMAX + tid] = i; gj[s * TC_GMAX + tid] = j; giz[s * TC_GMAX + tid] = isz;
      gs1[s * TC_GMAX + tid] = s1; glen[s * TC_GMAX + tid] = len; ginv[s * TC_GMAX + tid] = 1.f / (m.C + len);
    }
    __syncthreads();
    if (t.active && tid < t.nrows) {
      const int g = tid / r, q = tid - g * r;
      rowgrp[s * 128 + tid] = g;
      if (q == 0) {
        rowslot[s * 128 + tid] = 0;
        rowsd[s * 128 + tid] = gs1[s * TC_GMAX + g];
      } else {
        const int i = gi[s * TC_GMAX + g], j = gj[s * TC_GMAX + g];
        const int k = dirmap(kind, q - 1, i, j, dim);
        float sd = 0.f;
        for (int c = 0; c < dim; ++c) sd = fmaf(gvs[g * 3 + c], xt[(i * dim + c) * D + k] - xt[(j * dim + c) * D + k], sd);
        rowslot[s * 128 + tid] = 1 + k;
        rowsd[s * 128 + tid] = giz[s * TC_GMAX + g] ? 0.f : 2.f * sd;
      }
    }
    __syncthreads();
  }

  // phi_e layer 0 by gather: z0 = P_s[j] + P_r[i] + (|v|^2 or its tangent) w_d.  Rows are fetched warp-per-row
  // (coalesced 256 B segments), transposed through the staging buffer to the thread-per-row ownership, then the
  // activation / tangent rule is applied and the A operand written.
  __device__ __forceinline__ void tile_build(int s, const Tile& t, int r, bool htan, const float* wd) {
    const bool valid = t.active && row < t.nrows;
    const int g = valid ? rowgrp[s * 128 + row] : 0;
    const bool primal = valid && (row == g * r);
    const int lane = tid & 31;
    float v[64];
    qbeg();
    for (int half = 0; half < 2; ++half) {
      const float2 w2 = *reinterpret_cast<const float2*>(wd + 64 * half + 2 * lane);
      // 8 rows per warp pass: all 16 global loads are issued before any of them is consumed
      for (int r0 = warp; r0 < t.nrows; r0 += 8 * (NTHREADS / 32)) {
        float2 pa[8], pb[8];
        float sd[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = r0 + u * (NTHREADS / 32);
          pa[u] = make_float2(0.f, 0.f);
          pb[u] = make_float2(0.f, 0.f);
          sd[u] = 0.f;
          if (rr < t.nrows) {
            const int gg = rowgrp[s * 128 + rr];
            sd[u] = rowsd[s * 128 + rr];
            if (rr == gg * r || htan) {
              const int slot = rowslot[s * 128 + rr];
              pa[u] = *reinterpret_cast<const float2*>(Ps + ((size_t)gj[s * TC_GMAX + gg] * ND + slot) * U + 64 * half + 2 * lane);
              pb[u] = *reinterpret_cast<const float2*>(Pr + ((size_t)gi[s * TC_GMAX + gg] * ND + slot) * U + 64 * half + 2 * lane);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int rr = r0 + u * (NTHREADS / 32);
          if (rr < t.nrows) {
            stage[rr * TC_SLD + 2 * lane] = fmaf(sd[u], w2.x, pa[u].x + pb[u].x);
            stage[rr * TC_SLD + 2 * lane + 1] = fmaf(sd[u], w2.y, pa[u].y + pb[u].y);
          }
        }
      }
      __syncthreads();
      if (ch == half) {
        if (valid) {
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] = stage[row * TC_SLD + c];
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) v[c] = 0.f;
        }
      }
      __syncthreads();
    }
    qend(P_BUILD_GATHER);
    act_rule(v, nullptr, valid, primal, g, t.ng);
    st_a64(ahi_of(s), alo_of(s), v);
    qend(P_BUILD_ACT);
  }

  // attention gate + message aggregation (egnn.py:99-104) from the fp32 phi_e outputs of this tile (all of whose edges
  // share one receiver).  v = this thread's 64 columns of m (primal) / m-dot (tangent).  All threads.
  __device__ __forceinline__ void tile_messages(int s, const Tile& t, int kind, int r, const float (&v)[64], const float* wa,
                                float bav) {
    const bool valid = row < t.nrows;
    const float inv_sqrt_nb = rsqrtf((float)(n - 1));
    const int i = gi[s * TC_GMAX];
    qbeg();
    rdot[(s * 2 + ch) * 128 + row] = valid ? dot64(v, wa) : 0.f;
    __syncthreads();
    qend(P_MSG_DOT);
    if (tid < t.ng)
      ge[s * TC_GMAX + tid] = ecnf_sigmoid(rdot[(s * 2) * 128 + tid * r] + rdot[(s * 2 + 1) * 128 + tid * r] + bav);
    // two passes over the column halves through the staging buffer [128 rows][65]
    for (int half = 0; half < 2; ++half) {
      __syncthreads();
      if (ch == half && valid) {
#pragma unroll
        for (int c = 0; c < 64; ++c) stage[row * TC_SLD + c] = v[c];
      }
      __syncthreads();
      qend(P_MSG_STAGE);
      // message rows  msg = m e,  msg-dot = m-dot e + m e(1-e)(m-dot . wa)  are formed on the fly inside the segmented
      // sums over the edges of this receiver; thread -> (column c, row-in-group q), lanes run over consecutive columns
      for (int idx = tid; idx < r * 64; idx += NTHREADS) {
        const int c = idx & 63, q = idx >> 6;
        const int col = 64 * half + c;
        const bool shared_q = (q == 0) || kind == KIND_MID || (q - 1 < dim);
        float acc = 0.f;
        for (int g = 0; g < t.ng; ++g) {
          const float e = ge[s * TC_GMAX + g];
          const float mp = stage[g * r * TC_SLD + c];
          float x;
          if (q == 0) {
            x = mp * e;
          } else {
            const int rr = g * r + q;
            const float ed = e * (1.f - e) * (rdot[(s * 2) * 128 + rr] + rdot[(s * 2 + 1) * 128 + rr]);
            x = fmaf(mp, ed, stage[rr * TC_SLD + c] * e);
          }
          if (shared_q) acc += x;
          else macc[(1 + gj[s * TC_GMAX + g] * dim + (q - 1 - dim)) * U + col] += x * inv_sqrt_nb;
        }
        if (shared_q) {
          const int slot = (q == 0) ? 0 : 1 + dirmap(kind, q - 1, i, 0, dim);
          macc[slot * U + col] += acc * inv_sqrt_nb;
        }
      }
    }
    __syncthreads();
    qend(P_MSG_SEG);
    // receiver complete -> flush its aggregate to global (coalesced) and clear the accumulator
    if (t.e0 + t.ng == (i + 1) * (n - 1)) {
      for (int idx = tid; idx < ND * U; idx += NTHREADS) {
        Mg[(size_t)i * ND * U + idx] = macc[idx];
        macc[idx] = 0.f;
      }
      __syncthreads();
    }
    qend(P_MSG_FLUSH);
  }

  // coordinate update (egnn.py:87-95) from the head outputs p (rdot) of one tile.  All threads; ends with a barrier.
  __device__ __forceinline__ void tile_coords(int s, const Tile& t, int kind, int r, float bpv) {
    if (t.active) {
      const float* gvs = gv + s * TC_GMAX * 3;
      const int i_first = gi[s * TC_GMAX], i_last = gi[s * TC_GMAX + t.ng - 1];
      const int nrec = i_last - i_first + 1;
      const int per = dim + (r - 1) * dim;   // primal + tangent work items per receiver
      for (int idx = tid; idx < nrec * per; idx += NTHREADS) {
        const int ri = idx / per, w = idx - ri * per, i = i_first + ri;
        const int ga = max(t.e0, i * (n - 1)) - t.e0, gb = min(t.e0 + t.ng, (i + 1) * (n - 1)) - t.e0;
        if (w < dim) {
          const int c = w;
          float acc = 0.f;
          for (int g = ga; g < gb; ++g) {
            const float pg = rdot[(s * 2) * 128 + g * r] + rdot[(s * 2 + 1) * 128 + g * r] + bpv;
            acc = fmaf(pg * gvs[g * 3 + c], ginv[s * TC_GMAX + g], acc);
          }
          xacc[i * dim + c] += acc;
        } else {
          const int t2 = w - dim, q = 1 + t2 / dim, c = t2 % dim;
          for (int g = ga; g < gb; ++g) {
            const int j = gj[s * TC_GMAX + g];
            const int k = dirmap(kind, q - 1, i, j, dim);
            if (kind == KIND_LAST && k != i * dim + c) continue;
            const int rr = g * r + q;
            const float pg = rdot[(s * 2) * 128 + g * r] + rdot[(s * 2 + 1) * 128 + g * r] + bpv;
            const float pd = rdot[(s * 2) * 128 + rr] + rdot[(s * 2 + 1) * 128 + rr];
            const float vc = gvs[g * 3 + c], inv = ginv[s * TC_GMAX + g];
            const float vd = xt[(i * dim + c) * D + k] - xt[(j * dim + c) * D + k];
            const float ld = giz[s * TC_GMAX + g] ? 0.f : rowsd[s * 128 + rr] / (2.f * glen[s * TC_GMAX + g]);
            const float cd = (pd * vc + pg * vd) * inv - pg * vc * ld * inv * inv;
            if (kind == KIND_LAST) dacc[i * dim + c] += cd;
            else xtacc[(i * dim + c) * D + k] += cd;
          }
        }
      }
    }
    __syncthreads();
  }

  __device__ __forceinline__ void edge_phase(int b, int kind, bool htan) {
    const EcnfBlockParams& bp = m.blk[b];
    const TcImgBlock& ib = img.blk[b];
    const int L = m.L;
    const int nact = kind == KIND_MID ? D : kind == KIND_LAST ? dim : 2 * dim;
    const int r = 1 + nact;
    // tiles hold whole edges; when messages are aggregated (not the last block) a tile never spans two receivers
    const int gpt = (kind == KIND_LAST) ? min(128 / r, TC_GMAX) : min(min(128 / r, TC_GMAX), n - 1);
    const int tpr = (n - 1 + gpt - 1) / gpt;                       // tiles per receiver (non-last blocks)
    const int ntiles = (kind == KIND_LAST) ? (E + gpt - 1) / gpt : n * tpr;
    const int NW = 2 * L - 1;                  // We[1..L-1], Wx[0..L-1]
    const float* wd = bp.We[0] + (size_t)2 * H * U;
    const float bpv = bp.bp[0], bav = bp.ba[0];
    pbeg();
    for (int i = tid; i < D; i += NTHREADS) { xacc[i] = 0.f; dacc[i] = 0.f; }
    if (kind != KIND_LAST) {
      for (int i = tid; i < D * D; i += NTHREADS) xtacc[i] = 0.f;
      for (int i = tid; i < ND * U; i += NTHREADS) macc[i] = 0.f;
    }
    auto wimg = [&](int w) { return w < L - 1 ? ib.We[w + 1] : ib.Wx[w - (L - 1)]; };
    // per-block vectors live in shared memory for the whole phase (the L1 is too small to keep them: 217 KB carve-out)
    for (int idx = tid; idx < (NW + 3) * U; idx += NTHREADS) {
      const int w = idx / U, c = idx - w * U;
      const float* src = w < L - 1 ? bp.be[w + 1] : w < NW ? bp.bx[w - (L - 1)] : w == NW ? wd : w == NW + 1 ? bp.wa : bp.wp;
      vecs[idx] = src[c];
    }
    auto wbias = [&](int w) { return vecs + w * U; };
    const float* wd_s = vecs + NW * U;
    const float* wa_s = vecs + (NW + 1) * U;
    const float* wp_s = vecs + (NW + 2) * U;
    const int npairs = (ntiles + 1) / 2;
    const uint32_t seq_base = wq_head;
    const uint32_t seq_end = seq_base + (uint32_t)(npairs * NW);
    wq_load(wimg(0), 2 * U * U * 2);
    if (seq_base + 1 < seq_end) wq_load(wimg(1 % NW), 2 * U * U * 2);
    __syncthreads();
    pend(P_EDGE_INIT);

    for (int p = 0; p < npairs; ++p) {
      Tile t[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int ti = 2 * p + s;
        t[s].active = ti < ntiles;
        if (kind == KIND_LAST) {
          t[s].e0 = ti * gpt;
          t[s].ng = t[s].active ? min(gpt, E - t[s].e0) : 0;
        } else {
          const int i = ti / tpr, k = ti - i * tpr;
          t[s].e0 = i * (n - 1) + k * gpt;
          t[s].ng = t[s].active ? min(gpt, (n - 1) - k * gpt) : 0;
        }
        t[s].nrows = t[s].ng * r;
      }
      const uint32_t seq0 = seq_base + (uint32_t)(p * NW);
      // stage 0: build both tiles and start their first GEMM
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        if (!t[s].active) continue;
        pbeg();
        tile_meta(s, t[s], kind, r);
        pend(P_META);
        tile_build(s, t[s], r, htan, wd_s);
        pend(P_BUILD);
        tc_sync();
        if (warp == 0) { const uint32_t w = wq_wait(seq0); issue_mma(acc_of(s), ahi_of(s), alo_of(s), w, U, U, false, &mbar_mma[s]); }
        pend(P_SYNC_ISSUE);
      }
      for (int w = 0; w < NW; ++w) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!t[s].active) continue;
          pbeg();
          wait_mma(s);
          pend(P_WAIT_MMA);
          // after the LAST active slot finished with weights `w`, their buffer is free: prefetch two images ahead
          const bool last_slot = (s == 1) || !t[1].active;
          if (last_slot && seq0 + w + 2 < seq_end) wq_load(wimg((w + 2) % NW), 2 * U * U * 2);
          const bool valid = row < t[s].nrows;
          const int g = valid ? rowgrp[s * 128 + row] : 0;
          const bool primal = valid && (row == g * r);
          float v[64];
          qbeg();
          ld_acc64(acc_of(s), v);
          qend(P_EPI_LD);
          act_rule(v, wbias(w), valid, primal, g, t[s].ng);
          qend(P_EPI_ACT);
          if (w < NW - 1) {
            st_a64(ahi_of(s), alo_of(s), v);
            qend(P_EPI_ST);
            pend(P_EPI);
            if (w == L - 2 && kind != KIND_LAST) tile_messages(s, t[s], kind, r, v, wa_s, bav);
            pend(P_MSG);
            tc_sync();
            if (warp == 0) {
              const uint32_t wsm = wq_wait(seq0 + w + 1);
              issue_mma(acc_of(s), ahi_of(s), alo_of(s), wsm, U, U, false, &mbar_mma[s]);
            }
            pend(P_SYNC_ISSUE);
          } else {
            rdot[(s * 2 + ch) * 128 + row] = valid ? dot64(v, wp_s) : 0.f;
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
            pend(P_EPI);
            tile_coords(s, t[s], kind, r, bpv);
            pend(P_COORDS);
          }
        }
      }
    }
  }

  // ---- one evaluation of (f, div f) at time t for the positions in xin (shared memory, D floats) ----
  __device__ __forceinline__ void eval(float t, const float* xin, const int32_t* feat, float* fout) {
    if (tid < dim) {
      float s = 0.f;
      for (int i = 0; i < n; ++i) s += xin[i * dim + tid];
      mu[tid] = s / (float)n;
    }
    if (tid >= 32 && tid < 32 + m.T / 2) {
      const int k = tid - 32;
      const float arg = (t * 1000.f) * m.freqs[k];
      tau[k] = sinf(arg);
      tau[k + m.T / 2] = cosf(arg);
    }
    __syncthreads();
    for (int i = tid; i < D; i += NTHREADS) {
      const float v = xin[i] - mu[i % dim];
      xs[i] = v;
      xs0[i] = v;
    }
    const float invn = 1.f / (float)n;
    for (int idx = tid; idx < D * D; idx += NTHREADS) {
      const int ra = idx / D, k = idx - ra * D;
      const int ia = ra / dim, ca = ra - ia * dim, ik = k / dim, ck = k - ik * dim;
      xt[idx] = (ca == ck) ? ((ia == ik ? 1.f : 0.f) - invn) : 0.f;
    }
    for (int idx = tid; idx < n * H; idx += NTHREADS) {
      const int node = idx / H, col = idx - node * H;
      int f = feat[node];
      f = max(0, min(m.nfeat - 1, f));
      hA[(size_t)node * ND * H + col] = m.embed[f * H + col];
    }
    __syncthreads();
    for (int b = 0; b < m.nblocks; ++b) {
      const bool last = (b == m.nblocks - 1);
      const int kind = last ? KIND_LAST : (b == 0 ? KIND_FIRST : KIND_MID);
      const bool htan = b > 0;
      pbeg();
      node_pre(b, htan);
      pend(P_NODE_PRE);
      edge_phase(b, kind, htan);
      pbeg();
      if (!last) node_post(b, htan);
      pend(P_NODE_POST);
      const float invnb = 1.f / (float)(n - 1);
      for (int i = tid; i < D; i += NTHREADS) xs[i] += xacc[i] * invnb;
      if (!last)
        for (int i = tid; i < D * D; i += NTHREADS) xt[i] += xtacc[i] * invnb;
      __syncthreads();
    }
    const float fs = m.final_scaling[0];
    for (int i = tid; i < D; i += NTHREADS) fout[i] = (xs[i] - xs0[i] - mu[i % dim]) * fs;
    if (tid == 0) {
      const float invnb = 1.f / (float)(n - 1);
      float s = 0.f;
      for (int d = 0; d < D; ++d) s += xt[d * D + d] + dacc[d] * invnb;
      fout[D] = fs * (s - (float)D);
    }
    __syncthreads();
  }
};

__global__ void __launch_bounds__(NTHREADS, 1) ecnf_solve_tc_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ long long s_traj;
  __shared__ float s_ctl[8];
  EngineTC eng(a);
  solve_body<EngineTC, true>(a, eng, s_traj, s_ctl);
  eng.finish(reinterpret_cast<long long*>(reinterpret_cast<char*>(a.counter) + 64));
}

// ---- weight images: fp32 [K][N] (flax kernel) -> bf16 hi / lo in the canonical layout with rows = N ----------------
struct TcPrepItem {
  int src_off;   // floats, into the parameter buffer
  int dst_off;   // bytes, into the image buffer
  int K, N;
};
struct TcPrepList {
  int count;
  TcPrepItem item[64];
};

__global__ void tc_prep_kernel(const float* __restrict__ params, unsigned char* __restrict__ image, const TcPrepList list) {
  const TcPrepItem it = list.item[blockIdx.y];
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(image + it.dst_off);
  __nv_bfloat16* lo = hi + it.K * it.N;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < it.K * it.N; idx += gridDim.x * blockDim.x) {
    const int k = idx / it.N, nn = idx - k * it.N;
    const float w = params[it.src_off + idx];
    const __nv_bfloat16 h = __float2bfloat16(w);
    const int d = canon_index(nn, k, it.N);
    hi[d] = h;
    lo[d] = __float2bfloat16(w - __bfloat162float(h));
  }
}

#undef stage
#undef G
#undef macc
#undef vecs
#undef xt
#undef xtacc
#undef dacc
#undef xs
#undef xs0
#undef xacc
#undef mu
#undef tau
#undef cvec
#undef rowsd
#undef rdot
#undef gv
#undef gs1
#undef glen
#undef ginv
#undef ge
#undef rowslot
#undef rowgrp
#undef gi
#undef gj
#undef giz
#undef mbar_mma
#undef mbar_w
#undef tmem_slot
#undef prof_s
#undef hA
#undef hB
#undef Ps
#undef Pr
#undef Mg

}  // namespace ecnf_solve_detail

// Engine A on the 5th-generation tensor cores (tcgen05 + TMEM), for mlp_units = 128, n_invariant_feat_hidden = 64
// (DW4 / LJ13 shapes) with the exact divergence.
//
// Same algorithm and tangent-row scheme as the fp32 SIMT engine (ecnf_solve_impl.cuh); the formulation is transposed
// ("feature on lane") so that everything the tangent rule needs is thread-local:
//
//     D^T[out feature (TMEM lane), row (TMEM column)] = W^T (A operand, TMEM) x Act^T (B operand, shared memory)
//
//   * a row tile is up to 128 rows = 128 accumulator COLUMNS; thread (f, hh) of the 256 epilogue threads owns output
//     feature f = lane f and the 64 columns [64 hh, +64) -- the natural tcgen05.ld 32x32b ownership;
//   * rows are (edge or node, slot) with slot 0 the primal value and slot 1+k the tangent in input direction k.  A tile
//     is two half-blocks of 64 columns made of segments "primal column, then its tangent columns" (a group split over
//     the two halves repeats its primal column), so the activation rule  a = silu(z + b) / a-dot = silu'(z) z-dot
//     is a running scalar per thread: no shared-memory exchange and no barrier between the layers of the MLP chain;
//   * activations: accumulator (TMEM, fp32) -> registers -> rule -> bf16 (hi, lo) split -> B operand in shared memory,
//     MN-major no-swizzle canonical layout (8 rows of one feature = one 16-byte store, a warp stores 512 contiguous B);
//   * weights: pre-split bf16 (hi, lo) images, loaded from L2 straight into TENSOR MEMORY (tcgen05.st) by the epilogue
//     threads, double buffered, so the MMA reads only B from shared memory;
//   * every Dense layer is 3 x (K/16) tcgen05.mma (hi*hi + lo*hi + hi*lo, fp32 accumulate: ~2^-17 relative per
//     product, measured 4e-6 by tools/probe_tc2.cu) -- single-pass bf16/tf32 misses the 1e-4 tolerance on log q;
//   * two tiles are in flight (two accumulators, two B buffers): while the epilogue threads work on one, the tensor
//     core multiplies the other.  A dedicated warp issues the MMAs (the issuing thread blocks while the tensor-core
//     queue is full, which must not hold up an epilogue warp); hand-over is by mbarriers only
//     (epilogue threads -> "ready[s]" -> issue warp -> tcgen05.commit -> "done[s]" -> epilogue threads);
//   * the only cross-lane work are the two Dense(1) heads (attention logit, coordinate head): a 62-shuffle transposing
//     butterfly per warp + one 4-way sum through shared memory;
//   * tile composition (which (group, slot) sits in which column) is precomputed per (n, dim, kind) into a table.
#pragma once
#include "ecnf_solve_impl.cuh"
#include "ecnf_tc.cuh"

namespace ecnf_solve_detail {

using namespace ecnf_tc;

constexpr int TCU = 128, TCH = 64;
constexpr int TC_NT = 288;          // 8 epilogue warps (thread (f, hh) owns feature / TMEM lane f and the column half hh) + 1 MMA-issue warp
constexpr int TC_EPI = 256;
#ifndef TC_WPREFETCH
#define TC_WPREFETCH 0
#endif
constexpr int TC_BOP = 65536;       // one B-operand buffer: hi image [K = 128][N = 128] + lo image
constexpr uint32_t TC_LBO = 128, TC_SBO = 2048;
constexpr int TC_WCOL = 256;        // first TMEM column of the weight buffers (accumulators: [0, 128) and [128, 256))

// column word of a tile table
constexpr uint32_t CW_VALID = 1u << 31, CW_PRIMAL = 1u << 30, CW_DUP = 1u << 29;
__host__ __device__ __forceinline__ int cw_gid(uint32_t w) { return (int)(w & 1023u); }
__host__ __device__ __forceinline__ int cw_q(uint32_t w) { return (int)((w >> 10) & 255u); }
__host__ __device__ __forceinline__ int cw_pc(uint32_t w) { return (int)((w >> 18) & 63u); }
enum { TH_NC0 = 0, TH_NC1, TH_N, TH_G0, TH_NG, TH_WIN, TH_FLUSH, TH_IFIRST, TH_ILAST, TH_SEGS };
// kinds with one row per group (TT_NODE1, TT_EDGE1) pack densely: column = local group index (group words only for the
// first 64 groups of a tile)

// sigmoid from ex2.approx / rcp.approx (5 instructions, ~1e-7 absolute: well inside the 3-pass GEMM error)
__device__ __forceinline__ float fast_sigmoid(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

__host__ __device__ inline int tc_macc_rows(int n, int dim) {
  const int ND = 1 + n * dim;
  const int rw = 40 / ND > 1 ? 40 / ND : 1;
  return (rw < n ? rw : n) * ND;
}

// Packs the groups (nodes or edges) of one table kind into tiles; returns the tile count, writes them when out != null.
// Tile = 128 column words | 64 group words (slot split + first columns of the group's segments) | 16 header words |
// (16 unused) | 4 primal masks.  Padding columns between segments are invalid (word 0).
__host__ __device__ inline int tc_pack(int kind, int n, int dim, uint32_t* out) {
  const int D = n * dim, ND = 1 + D;
  const bool edge = kind >= TT_FIRST;
  const int ngroups = edge ? n * (n - 1) : n;
  const int r = (kind == TT_NODE1 || kind == TT_EDGE1) ? 1 : kind == TT_NODE ? ND : kind == TT_FIRST ? 1 + 2 * dim : kind == TT_MID ? ND : 1 + dim;
  // message-accumulator window: tiles of the message-passing kinds never span two windows of receivers
  const int window = (kind == TT_FIRST || kind == TT_MID) ? (tc_macc_rows(n, dim) / ND) * (n - 1) : 0;
  // segments start on 8-column chunk boundaries (the activation rule works chunk-wise); all-primal node tiles pack densely
  const int al = r == 1 ? 1 : 8;
  int tile = 0, used0 = 0, used1 = 0, half = 0, ng = 0, g0 = 0;
  auto tp = [&](int t) { return out + (size_t)t * TC_TILE_WORDS; };
  auto close = [&](int gnext, int flush) {
    if (ng == 0) return;
    if (out) {
      uint32_t* h = tp(tile) + 192;
      const int n1 = (used1 + 15) & ~15, n0 = (used0 + 15) & ~15;
      h[TH_NC0] = (uint32_t)used0;
      h[TH_NC1] = (uint32_t)used1;
      h[TH_N] = (uint32_t)(used1 > 0 ? 64 + n1 : (n0 > 16 ? n0 : 16));
      h[TH_G0] = (uint32_t)g0;
      h[TH_NG] = (uint32_t)ng;
      const int wstart = window ? (g0 / window) * window : 0;
      h[TH_WIN] = (uint32_t)(edge ? wstart / (n - 1) : 0);
      h[TH_FLUSH] = (uint32_t)flush;
      h[TH_IFIRST] = (uint32_t)(edge ? g0 / (n - 1) : g0);
      h[TH_ILAST] = (uint32_t)(edge ? (g0 + ng - 1) / (n - 1) : g0 + ng - 1);
      // primal masks (4 x 32 columns) and, per half, which 8-column chunks start a segment
      uint32_t* t = tp(tile);
      uint32_t segs = 0u;
      for (int c = 0; c < 128; ++c)
        if (t[c] & CW_PRIMAL) {
          t[224 + (c >> 5)] |= 1u << (c & 31);
          if ((c & 7) == 0) segs |= 1u << (c >> 3);
        }
      h[TH_SEGS] = segs;
    }
    ++tile;
    used0 = used1 = 0; half = 0; ng = 0; g0 = gnext;
  };
  // one segment: [repeated primal] + slots q0..q1-1 from column col0 on
  auto seg = [&](int g, int col0, int q0, int q1, bool dup) {
    if (!out) return;
    uint32_t* cw = tp(tile);
    int c = col0;
    const uint32_t common = CW_VALID | (uint32_t)g | ((uint32_t)(col0 & 63) << 18);
    if (dup) cw[c++] = common | CW_PRIMAL | CW_DUP;
    for (int q = q0; q < q1; ++q) cw[c++] = common | (q == 0 ? CW_PRIMAL : 0u) | ((uint32_t)q << 10);
  };
  for (int g = 0; g < ngroups; ++g) {
    if (window && g > 0 && g % window == 0) close(g, 1);
    for (;;) {
      if (ng == 0 && used0 == 0 && out)
        for (int k = 0; k < TC_TILE_WORDS; ++k) tp(tile)[k] = 0;
      if (half == 0) {
        used0 = (used0 + al - 1) / al * al;
        const int rem = 64 - used0;
        if (r <= rem) {
          seg(g, used0, 0, r, false);
          if (out && ng < 64) tp(tile)[128 + ng] = (uint32_t)r | ((uint32_t)used0 << 8);
          used0 += r;
          break;
        }
        if (rem >= 2 && 1 + r - rem <= 64) {
          seg(g, used0, 0, rem, false);
          seg(g, 64, rem, r, true);
          if (out) tp(tile)[128 + ng] = (uint32_t)rem | ((uint32_t)used0 << 8) | (64u << 16);
          used0 = 64; used1 = 1 + r - rem; half = 1;
          break;
        }
        half = 1;
        if (r > 64) { close(g, 0); }
        continue;
      }
      used1 = (used1 + al - 1) / al * al;
      const int rem = 64 - used1;
      if (r <= rem) {
        seg(g, 64 + used1, 0, r, false);
        if (out && ng < 64) tp(tile)[128 + ng] = (uint32_t)r | ((uint32_t)(64 + used1) << 8);
        used1 += r;
        break;
      }
      close(g, 0);
    }
    ++ng;
  }
  close(ngroups, (window || kind == TT_EDGE1) ? 1 : 0);     // the primal-only edge kind aggregates over one window = all receivers
  return tile;
}

__host__ __device__ inline TcSmemLayout make_tc_layout(int n, int dim) {
  const int D = n * dim, S = D + 1, E = n * (n - 1);
  TcSmemLayout L;
  int o = 0;
  auto take = [&](int bytes) { int r = o; o += (bytes + 127) & ~127; return r; };
  L.bop = take(2 * TC_BOP);
  L.mrows = tc_macc_rows(n, dim);
  L.macc = take(2 * (L.mrows + 1) * TCU * 4);   // aggregated messages of the receiver window, one copy per column half (+ a dump row)
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
  L.colw = take(2 * 128 * 4);
  L.coloffS = take(2 * 128 * 4);
  L.coloffR = take(2 * 128 * 4);
  L.colsd = take(2 * 128 * 4);
  L.colmrow = take(2 * 128 * 4);
  L.colijk = take(2 * 128 * 4);
  L.pm = take(2 * 4 * 4);
  L.grpw = take(2 * 64 * 4);
  L.hdr = take(2 * 16 * 4);
  L.pdot = take(2 * 8 * 64 * 4);
  L.wA = take(8 * 64 * 4);
  L.wB = take(8 * 64 * 4);
  L.cdbuf = take(2 * 128 * 3 * 4);
  L.egv = take(E * 3 * 4);
  L.eglen = take(E * 4);
  L.eginv = take(E * 4);
  L.egs1 = take(E * 4);
  L.egiz = take(E * 4);
  L.bars = take(64);
  L.prof = take(32 * 8);
  L.total_bytes = o;
  return L;
}

extern __shared__ __align__(1024) unsigned char smem_tc[];

// Every buffer is addressed as (shared-memory base + offset from the kernel parameters): the offsets are constant-bank
// loads, so no pointer has to be kept alive in (or spilled from) registers across the very large inlined body.
#define TCF(field) (reinterpret_cast<float*>(smem_tc + a.lay.field))
#define TCI(field) (reinterpret_cast<int*>(smem_tc + a.lay.field))
#define TCW(field) (reinterpret_cast<uint32_t*>(smem_tc + a.lay.field))

// DIV = false: sample_cnf without a divergence -- primal rows only (tile kinds TT_NODE1 / TT_EDGE1, no tangent state)
template <bool DIV>
struct EngineTC {
  static constexpr int U = TCU, H = TCH;
  static constexpr int NT = TC_NT;
  const KernelArgs& a;     // the __grid_constant__ kernel parameter
  const EcnfModelDev& m;
  const TcImages& img;
  const int n, dim, D, ND, E;
  const int tid, f, hh, warp, lane;
  const bool is_epi;
  uint32_t tmem;         // TMEM base address
  uint32_t lane_addr;    // (32 * (warp & 3)) << 16
  uint32_t ph0, ph1;     // phase counters of the two slots (ready[] for the issue warp, done[] for the epilogue threads)

  enum { P_NODE_PRE, P_EDGE, P_NODE_POST, P_WAIT, P_BUILD, P_EPI, P_MSG, P_COORD, P_WLOAD, P_META, P_CTRL_WAIT, P_MISC,
         P_EPI_LD, P_EPI_ACT, P_EPI_ST, P_ARRIVE, P_GATHER, P_NCOUNT };
#ifdef ECNF_TC_PROFILE
  long long prof_t0, prof_t1;
  __device__ __forceinline__ long long* prof_s() const { return reinterpret_cast<long long*>(smem_tc + a.lay.prof); }
  __device__ __forceinline__ void pbeg() { prof_t0 = clock64(); }
  __device__ __forceinline__ void pend(int k) { const long long t1 = clock64(); if (tid == 0) prof_s()[k] += t1 - prof_t0; prof_t0 = t1; }
  __device__ __forceinline__ void qbeg() { prof_t1 = clock64(); }
  __device__ __forceinline__ void qend(int k) { const long long t1 = clock64(); if (tid == 0 || tid == TC_EPI) prof_s()[k] += t1 - prof_t1; prof_t1 = t1; }
#else
  __device__ __forceinline__ void pbeg() {}
  __device__ __forceinline__ void pend(int) {}
  __device__ __forceinline__ void qbeg() {}
  __device__ __forceinline__ void qend(int) {}
#endif

  __device__ __forceinline__ void set_eps(const float*) {}   // Hutchinson probes run on the SIMT engine
  __device__ __forceinline__ float* ode_ptr() const { return TCF(ode); }
  __device__ __forceinline__ float* red_ptr() const { return TCF(red); }
  __device__ __forceinline__ uint64_t* bar_ready(int s) const { return reinterpret_cast<uint64_t*>(smem_tc + a.lay.bars) + s; }
  __device__ __forceinline__ uint64_t* bar_done(int s) const { return reinterpret_cast<uint64_t*>(smem_tc + a.lay.bars) + 2 + s; }
  __device__ __forceinline__ uint32_t* tmem_slot() const { return reinterpret_cast<uint32_t*>(smem_tc + a.lay.bars + 32); }
  // per-CTA global scratch (L2 resident): h, h_in [n][ND][H]; P_s, P_r, aggregated messages, P_h [n][ND][U]
  __device__ __forceinline__ float* hA() const { return a.scratch + (size_t)blockIdx.x * a.scratch_stride; }
  __device__ __forceinline__ float* hB() const { return hA() + (size_t)n * ND * H; }
  __device__ __forceinline__ float* Ps() const { return hB() + (size_t)n * ND * H; }
  // P_s and P_r carry one extra, all-zero row (index n * ND): the gather of phi_e layer 0 points there for rows without a P part
  __device__ __forceinline__ float* Pr() const { return Ps() + ((size_t)n * ND + 1) * U; }
  __device__ __forceinline__ float* Mg() const { return Pr() + ((size_t)n * ND + 1) * U; }
  __device__ __forceinline__ float* Ph() const { return Mg() + (size_t)n * ND * U; }

  __device__ __forceinline__ EngineTC(const KernelArgs& a_)
      : a(a_), m(a_.m), img(a_.img), n(a_.m.n), dim(a_.m.dim), D(a_.m.n * a_.m.dim), ND(DIV ? 1 + a_.m.n * a_.m.dim : 1),
        E(a_.m.n * (a_.m.n - 1)), tid(threadIdx.x), f(threadIdx.x & 127), hh((threadIdx.x >> 7) & 1), warp(threadIdx.x >> 5),
        lane(threadIdx.x & 31), is_epi(threadIdx.x < TC_EPI) {
    if (tid == 0) {
      mbar_init(bar_ready(0), TC_EPI); mbar_init(bar_ready(1), TC_EPI);
      mbar_init(bar_done(0), 1); mbar_init(bar_done(1), 1);
#ifdef ECNF_TC_PROFILE
      for (int k = 0; k < P_NCOUNT; ++k) prof_s()[k] = 0;
#endif
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot(), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tmem = *tmem_slot();
    lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    ph0 = ph1 = 0;
    if (is_epi) {   // accumulators start finite (columns beyond a tile's N are read but never used)
      uint32_t z[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) z[c] = 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) tmem_st32(tmem + lane_addr + 128u * hh + 32u * q, z);
      tmem_wait_st();
      (hh ? Pr() : Ps())[(size_t)n * ND * U + f] = 0.f;      // the zero rows of P_s / P_r
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  __device__ __forceinline__ void finish(long long* out) {
#ifdef ECNF_TC_PROFILE
    if (out && tid == 0 && blockIdx.x == 0)
      for (int k = 0; k < P_NCOUNT; ++k) out[k] = prof_s()[k];
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
  }

  // ---- hand-over between the epilogue threads and the MMA-issue warp ----------------------------------------------
  __device__ __forceinline__ void epi_bar() const { named_bar_sync(1, TC_EPI); }
  // the four warps that own one column half (their partial dot products are summed by each of them)
  __device__ __forceinline__ void half_bar() const { named_bar_sync(2 + hh, TC_EPI / 2); }
  __device__ __forceinline__ void wait_slot(uint64_t* bar, int s) {
    const uint32_t par = (s ? ph1 : ph0) & 1u;
    mbar_wait_parked(bar, par, 20000u);
    ph0 += (s == 0); ph1 += (s != 0);
    tc_fence_after();
  }
  __device__ __forceinline__ void wait_done(int s) { qbeg(); wait_slot(bar_done(s), s); qend(P_WAIT); }
  // epilogue threads: my part of slot s's next operands (B in shared memory, weights / drained accumulator in TMEM) is complete
  __device__ __forceinline__ void arrive_ready(int s) {
    qbeg();
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar_ready(s));
    qend(P_ARRIVE);
  }
  // issue warp: wait for all epilogue threads, then acc[s] = W^T (TMEM columns a_col: hi [0, K/2), lo [K/2, K)) x B[s]
  // (K x N), 3-pass split; commit -> done[s]
  __device__ __forceinline__ void issue_mma(int s, uint32_t a_col, int K) {
    if (lane == 0) {
      wait_slot(bar_ready(s), s);
      const int N = hdr(s, TH_N);
      const uint32_t idesc = make_idesc_bf16(128, N) | IDESC_B_MN;
      const uint32_t bsm = smem_u32(smem_tc + a.lay.bop) + (uint32_t)s * TC_BOP;
      uint64_t bh = make_sdesc(bsm, TC_LBO, TC_SBO);
      uint64_t bl = make_sdesc(bsm + 32768u, TC_LBO, TC_SBO);
      const uint32_t acc = tmem + 128u * s;
      uint32_t a_hi = tmem + a_col, a_lo = tmem + a_col + (uint32_t)(K >> 1);
      const int nk = K >> 4;
      mma_ts(acc, a_hi, bh, idesc, 0u);
      mma_ts(acc, a_lo, bh, idesc, 1u);
      mma_ts(acc, a_hi, bl, idesc, 1u);
      for (int ks = 1; ks < nk; ++ks) {
        bh += (2u * TC_LBO) >> 4; bl += (2u * TC_LBO) >> 4; a_hi += 8; a_lo += 8;
        mma_ts(acc, a_hi, bh, idesc, 1u);
        mma_ts(acc, a_lo, bh, idesc, 1u);
        mma_ts(acc, a_hi, bl, idesc, 1u);
      }
      mma_commit(bar_done(s));
    }
  }

  // ---- per-thread column helpers (thread (f, hh): feature f, columns [64 hh, +64) of slot s) -----------------------
  __device__ __forceinline__ uint32_t my_acc(int s) const { return tmem + 128u * s + lane_addr + 64u * hh; }
  __device__ __forceinline__ void ld_acc(int s, float (&v)[64]) {
    uint32_t x[32], y[32];
    const uint32_t addr = my_acc(s);
    tmem_ld32(addr, x);
    tmem_ld32(addr + 32u, y);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 32; ++c) { v[c] = __uint_as_float(x[c]); v[32 + c] = __uint_as_float(y[c]); }
  }
  // B operand of slot s <- v (K row = my feature), bf16 hi / lo
  __device__ __forceinline__ void write_B(int s, const float (&v)[64]) {
    unsigned char* base = smem_tc + a.lay.bop + s * TC_BOP + (8 * hh) * (int)TC_SBO + f * 16;
#pragma unroll
    for (int g8 = 0; g8 < 8; ++g8) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) split_pack(v[8 * g8 + 2 * p], v[8 * g8 + 2 * p + 1], h[p], l[p]);
      *reinterpret_cast<uint4*>(base + g8 * (int)TC_SBO) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(base + 32768 + g8 * (int)TC_SBO) = make_uint4(l[0], l[1], l[2], l[3]);
    }
  }
  // Activation rule: a = silu(z + bias) on primal columns, a-dot = silu'(z_primal) z-dot on the tangent columns that
  // follow them.  Segments start on 8-column chunk boundaries, so the primal column of a segment is column 0 of a chunk
  // (a static register) and silu' is a running per-thread scalar; `segs` = which of my 8 chunks start a segment.
  __device__ __forceinline__ void act_rule(float (&v)[64], float bias, int s) {
    if constexpr (!DIV) {      // every column is a primal row; chunks beyond the used columns of my half are skipped
      const int nc = hdr(s, TH_NC0 + hh);   // (the sigmoid is MUFU bound and these tiles are mostly padding)
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        if (8 * ch < nc) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float z = v[8 * ch + u] + bias;
            v[8 * ch + u] = z * fast_sigmoid(z);
          }
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) v[8 * ch + u] = 0.f;
        }
      }
      return;
    }
    const uint32_t segs = (TCW(hdr)[s * 16 + TH_SEGS] >> (8 * hh)) & 0xffu;
    float cur = 0.f;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      if ((segs >> ch) & 1u) {
        const float z = v[8 * ch] + bias;
        const float sg = fast_sigmoid(z);
        cur = sg * (1.f + z * (1.f - sg));
        v[8 * ch] = z * sg;
      } else {
        v[8 * ch] *= cur;
      }
#pragma unroll
      for (int u = 1; u < 8; ++u) v[8 * ch + u] *= cur;
    }
  }
  template <int W>   // 2W partial sums -> W, exchanging with lane ^ (W/2)
  __device__ __forceinline__ void bfly_round(float (&x)[32]) const {
    const bool b = (lane & (W >> 1)) != 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const float keep = b ? x[W + i] : x[i], send = b ? x[i] : x[W + i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, W >> 1);
    }
  }
  // column sums over the 32 features of this warp of v[c] * wf (transposing butterfly): lane l ends with the columns
  // 2l and 2l+1 of its half, stored to pd[2l], pd[2l+1]
  __device__ __forceinline__ void warp_dot(const float (&v)[64], float wf, float* pd) {
    float x[32];
    {
      const bool b = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float lo = v[i] * wf, hi = v[32 + i] * wf;
        const float keep = b ? hi : lo, send = b ? lo : hi;
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    bfly_round<16>(x);
    bfly_round<8>(x);
    bfly_round<4>(x);
    bfly_round<2>(x);
    *reinterpret_cast<float2*>(pd + 2 * lane) = make_float2(x[0], x[1]);
  }
  // weight image (hi | lo, K x 128 lanes) -> registers -> TMEM columns [col, col + K)
  template <int K>
  __device__ __forceinline__ void fetch_w(int img_off, uint32_t (&wv)[2][K / 4]) {
    constexpr int NC = K / 16;     // uint4 chunks per thread and part
    const uint4* src = reinterpret_cast<const uint4*>(img.base + img_off);
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        const uint4 q = __ldg(src + (size_t)(p * (K / 8) + hh * NC + j) * 128 + f);
        wv[p][4 * j] = q.x; wv[p][4 * j + 1] = q.y; wv[p][4 * j + 2] = q.z; wv[p][4 * j + 3] = q.w;
      }
  }
  template <int K>
  __device__ __forceinline__ void store_w(const uint32_t (&wv)[2][K / 4], uint32_t col) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const uint32_t addr = tmem + col + lane_addr + (uint32_t)(p * (K / 2) + hh * (K / 4));
      if constexpr (K == 128) tmem_st32(addr, wv[p]);
      else tmem_st16(addr, wv[p]);
    }
    tmem_wait_st();
  }
  template <int K>
  __device__ __forceinline__ void load_w(int img_off, uint32_t col) {
    uint32_t wv[2][K / 4];
    fetch_w<K>(img_off, wv);
    store_w<K>(wv, col);
  }

  // ---- tile tables ---------------------------------------------------------------------------------------------------
  __device__ __forceinline__ const uint32_t* tile_ptr(int kind, int tile) const {
    return a.tabs.base + (size_t)(a.tabs.off[kind] + tile) * TC_TILE_WORDS;
  }
  __device__ __forceinline__ int hdr(int s, int k) const { return TCI(hdr)[s * 16 + k]; }
  // threads 128..227 copy the tile's group words, header, primal-column positions and primal masks
  __device__ __forceinline__ void meta_copy(int s, const uint32_t* tp) {
    if (tid >= 128 && tid < 228) {
      const uint32_t w = __ldg(tp + tid);
      if (tid < 192) TCW(grpw)[s * 64 + (tid - 128)] = w;
      else if (tid < 208) TCW(hdr)[s * 16 + (tid - 192)] = w;
      else if (tid >= 224) TCW(pm)[s * 4 + (tid - 224)] = w;
    }
  }

  // =============================================================================================================
  // The software pipeline shared by the three phases.  P provides
  //   ntiles, NL (layers per tile), kind (tile table), stream (weights streamed layer by layer through two buffers)
  //   prologue()                  epilogue threads, before the first tile
  //   build(s, tile)              epilogue threads: first B operand of the tile
  //   epi(s, tile, w)             epilogue threads: consume the accumulator of layer w (and write the next B operand)
  //   wimg(w)                     image offset of the weights of layer w (stream only)
  //   a_col(q, w), K(w)           MMA operand position / depth for layer w at stream position q
  // =============================================================================================================
  template <class P>
  __device__ __forceinline__ void run_phase(P& p) {
    const int ntiles = p.ntiles, NL = p.NL;
    const int npairs = (ntiles + 1) >> 1;
    const int total_q = npairs * NL;
    if (is_epi) {
      p.prologue();
      if (p.stream) {
        load_w<TCU>(p.wimg(0), TC_WCOL);
        if (total_q > 1) load_w<TCU>(p.wimg(1 % NL), TC_WCOL + 128);
      }
    }
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
      if (s >= ntiles) continue;
      if (is_epi) { p.build(s, s); arrive_ready(s); }
      else issue_mma(s, p.a_col(0, 0), p.K(0));
    }
    int q = 0;
#pragma unroll 1
    for (int pr = 0; pr < npairs; ++pr) {
#pragma unroll 1
      for (int w = 0; w < NL; ++w, ++q) {
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          const int tile = 2 * pr + s;
          if (tile >= ntiles) continue;
          const bool last_slot = (s == 1) || (tile + 1 >= ntiles);
          const int ntile = tile + 2;
          if (is_epi) {
            // the weights two layers ahead go into the buffer that layer q's MMAs (complete once done[last slot]
            // fires) have been reading: fetch them into registers before the wait, store after it
            const bool do_w = p.stream && last_slot && q + 2 < total_q;
#if TC_WPREFETCH
            uint32_t wv[2][TCU / 4];
            if (do_w) fetch_w<TCU>(p.wimg((w + 2) % NL), wv);
            wait_done(s);
            if (do_w) { qbeg(); store_w<TCU>(wv, TC_WCOL + 128 * (q & 1)); qend(P_WLOAD); }
#else
            wait_done(s);
            if (do_w) { qbeg(); load_w<TCU>(p.wimg((w + 2) % NL), TC_WCOL + 128 * (q & 1)); qend(P_WLOAD); }
#endif
            p.epi(s, tile, w);
            if (w < NL - 1) arrive_ready(s);
            else if (ntile < ntiles) { p.build(s, ntile); arrive_ready(s); }
          } else {
            if (w < NL - 1) issue_mma(s, p.a_col(q + 1, w + 1), p.K(w + 1));
            else if (ntile < ntiles) issue_mma(s, p.a_col(q + 1, 0), p.K(0));
          }
        }
      }
    }
  }

  // column metadata of a node tile (kinds TT_NODE1 / TT_NODE): row offset (node * ND + slot), primal masks, header
  __device__ __forceinline__ void node_meta(int s, int kind, int tile) {
    qbeg();
    epi_bar();     // every thread is done with the previous contents of slot s's metadata
    const uint32_t* tp = tile_ptr(kind, tile);
    if (tid < 128) {
      const uint32_t w = __ldg(tp + tid);
      TCW(colw)[s * 128 + tid] = w;
      TCI(coloffR)[s * 128 + tid] = (w & CW_VALID) ? cw_gid(w) * ND + cw_q(w) : -1;
    }
    meta_copy(s, tp);
    epi_bar();
    qend(P_META);
  }

  // =============================================================================================================
  // node phase 1: h_in = [h | tau] Wd + bd ; P_s = h_in We0[0:H] ; P_r = h_in We0[H:2H] + be0 ; P_h = h_in Wh0[U:U+H] + bh0
  // =============================================================================================================
  struct NodePre {
    EngineTC& e;
    int b, ntiles, NL, kind;
    static constexpr bool stream = false;
    __device__ __forceinline__ uint32_t a_col(int, int w) const { return TC_WCOL + 64u * w; }
    __device__ __forceinline__ int K(int) const { return TCH; }
    __device__ __forceinline__ int wimg(int) const { return 0; }
    __device__ __forceinline__ void prologue() {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      const TcImgBlock& ib = e.img.blk[b];
      if (e.tid < TCH) {
        float cv = bp.bd[e.tid];
        for (int k = 0; k < e.m.T; ++k) cv = fmaf(TCF(tau)[k], bp.Wd[(TCH + k) * TCH + e.tid], cv);
        TCF(cvec)[e.tid] = cv;
      }
      e.template load_w<TCH>(ib.Wd, TC_WCOL);
      e.template load_w<TCH>(ib.We0s, TC_WCOL + 64);
      e.template load_w<TCH>(ib.We0r, TC_WCOL + 128);
      if (NL > 3) e.template load_w<TCH>(ib.Wh0h, TC_WCOL + 192);
    }
    __device__ __forceinline__ void build(int s, int tile) {
      const KernelArgs& a = e.a;
      e.node_meta(s, kind, tile);
      float v[64];
      const float* src = e.hA() + e.f;
      const int* ro_ = TCI(coloffR) + s * 128 + 64 * e.hh;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const int ro = ro_[c];
        const float x = src[(size_t)(ro >= 0 && e.f < TCH ? ro : 0) * TCH];
        v[c] = (ro >= 0 && e.f < TCH) ? x : 0.f;
      }
      if (e.f < TCH) e.write_B(s, v);
    }
    __device__ __forceinline__ void epi(int s, int, int w) {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      float v[64];
      const float cv = TCF(cvec)[e.f & (TCH - 1)];
      const float bias = w == 0 ? cv : (w == 1 ? 0.f : (w == 2 ? bp.be[0][e.f] : bp.bh[0][e.f]));
      e.ld_acc(s, v);
      const uint32_t m0 = TCW(pm)[s * 4 + 2 * e.hh], m1 = TCW(pm)[s * 4 + 2 * e.hh + 1];
      const int* ro_ = TCI(coloffR) + s * 128 + 64 * e.hh;
      const bool act = (w > 0) || (e.f < TCH);
      const int ld = w == 0 ? TCH : TCU;
      float* dst = (w == 0 ? e.hB() : w == 1 ? e.Ps() : w == 2 ? e.Pr() : e.Ph()) + e.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const bool pr = (((c < 32 ? m0 : m1) >> (c & 31)) & 1u) != 0u;
        v[c] += pr ? bias : 0.f;
        const int ro = ro_[c];
        if (ro >= 0 && act) dst[(size_t)ro * ld] = v[c];    // a repeated primal column rewrites the same value
      }
      if (w == 0 && e.f < TCH) e.write_B(s, v);
    }
  };

  // =============================================================================================================
  // node phase 2: h <- phi_h([M | h_in]) + h_in        (the h_in part of layer 0 is P_h from phase 1)
  // =============================================================================================================
  struct NodePost {
    EngineTC& e;
    int b, ntiles, NL, kind;
    bool htan;
    static constexpr bool stream = true;
    __device__ __forceinline__ uint32_t a_col(int q, int) const { return TC_WCOL + 128u * (q & 1); }
    __device__ __forceinline__ int K(int) const { return TCU; }
    __device__ __forceinline__ void prologue() {}
    __device__ __forceinline__ int wimg(int w) const {
      const TcImgBlock& ib = e.img.blk[b];
      const int L = e.m.L;
      return w == 0 ? ib.Wh0m : (w < L ? ib.Wh[w] : ib.WhL);
    }
    __device__ __forceinline__ void build(int s, int tile) {
      const KernelArgs& a = e.a;
      e.node_meta(s, kind, tile);
      float v[64];
      const float* src = e.Mg() + e.f;
      const int* ro_ = TCI(coloffR) + s * 128 + 64 * e.hh;
#pragma unroll
      for (int c = 0; c < 64; ++c) {
        const int ro = ro_[c];
        const float x = src[(size_t)(ro >= 0 ? ro : 0) * TCU];
        v[c] = ro >= 0 ? x : 0.f;
      }
      e.write_B(s, v);
    }
    __device__ __forceinline__ void epi(int s, int, int w) {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      const int L = e.m.L;
      const int* ro_ = TCI(coloffR) + s * 128 + 64 * e.hh;
      float v[64];
      if (w == 0) {
        const uint32_t m0 = TCW(pm)[s * 4 + 2 * e.hh], m1 = TCW(pm)[s * 4 + 2 * e.hh + 1];
        const float* ph = e.Ph() + e.f;
        // + P_h (primal always; tangent columns where h_in carries tangents)
        e.ld_acc(s, v);
#pragma unroll
        for (int cb = 0; cb < 64; cb += 16) {
          float pv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int c = cb + u;
            const bool pr = (((c < 32 ? m0 : m1) >> (c & 31)) & 1u) != 0u;
            const int ro = ro_[c];
            const bool use = ro >= 0 && (pr || htan);
            const float x = ph[(size_t)(use ? ro : 0) * TCU];
            pv[u] = use ? x : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) v[cb + u] += pv[u];
        }
        e.act_rule(v, 0.f, s);
        e.write_B(s, v);
      } else if (w < L) {
        const float bias = bp.bh[w][e.f];
        e.ld_acc(s, v);
        e.act_rule(v, bias, s);
        e.write_B(s, v);
      } else {
        const uint32_t m0 = TCW(pm)[s * 4 + 2 * e.hh], m1 = TCW(pm)[s * 4 + 2 * e.hh + 1];
        const float bias = bp.bh[L][e.f & (TCH - 1)];
        const float* hin = e.hB() + (e.f & (TCH - 1));
        float* dst = e.hA() + e.f;
        e.ld_acc(s, v);
        if (e.f < TCH) {
#pragma unroll
          for (int cb = 0; cb < 64; cb += 16) {
            float hv[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int c = cb + u;
              const bool pr = (((c < 32 ? m0 : m1) >> (c & 31)) & 1u) != 0u;
              const int ro = ro_[c];
              const bool use = ro >= 0 && (pr || htan);
              const float x = hin[(size_t)(use ? ro : 0) * TCH];
              hv[u] = (use ? x : 0.f) + (pr ? bias : 0.f);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int ro = ro_[cb + u];
              if (ro >= 0) dst[(size_t)ro * TCH] = v[cb + u] + hv[u];
            }
          }
        }
      }
    }
  };

  // =============================================================================================================
  // edge phase: phi_e (layer 0 by gather) -> {attention gate + message aggregation, phi_x -> coordinate update}
  // =============================================================================================================
  struct EdgePh {
    EngineTC& e;
    int b, ntiles, NL, kind;   // kind = tile table (TT_FIRST / TT_MID / TT_LAST)
    bool htan, want_msg;       // want_msg: the aggregated messages feed phi_h (every block but the last)
    float wdf, waf, wpf;       // my feature's entry of w_d (|v|^2 column of phi_e layer 0), attention and head weights
    static constexpr bool stream = true;
    __device__ __forceinline__ uint32_t a_col(int q, int) const { return TC_WCOL + 128u * (q & 1); }
    __device__ __forceinline__ int K(int) const { return TCU; }
    __device__ __forceinline__ int ekind() const { return kind == TT_FIRST ? KIND_FIRST : kind == TT_MID ? KIND_MID : KIND_LAST; }   // TT_EDGE1 has no tangent columns: the value is never used
    __device__ __forceinline__ int wimg(int w) const {
      const TcImgBlock& ib = e.img.blk[b];
      const int L = e.m.L;
      return w < L - 1 ? ib.We[w + 1] : ib.Wx[w - (L - 1)];
    }
    __device__ __forceinline__ void prologue() {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      const int n = e.n, dim = e.dim, D = e.D;
      wdf = bp.We[0][(size_t)2 * TCH * TCU + e.f];
      waf = bp.wa[e.f];
      wpf = bp.wp[e.f];
      for (int i = e.tid; i < D; i += TC_EPI) { TCF(xacc)[i] = 0.f; TCF(dacc)[i] = 0.f; }
      if (DIV && kind != TT_LAST)
        for (int i = e.tid; i < D * D; i += TC_EPI) TCF(xtacc)[i] = 0.f;
      if (want_msg)
        for (int i = e.tid; i < 2 * (a.lay.mrows + 1) * TCU; i += TC_EPI) TCF(macc)[i] = 0.f;
      // per-edge geometry (egnn.py:73-76, numerical.py:7-10)
      for (int ed = e.tid; ed < e.E; ed += TC_EPI) {
        const int i = ed / (n - 1), jj = ed - i * (n - 1);
        int j = i + 1 + jj; if (j >= n) j -= n;
        float sq = 0.f;
        for (int c = 0; c < dim; ++c) {
          const float vv = TCF(xs)[i * dim + c] - TCF(xs)[j * dim + c];
          TCF(egv)[ed * 3 + c] = vv;
          sq = fmaf(vv, vv, sq);
        }
        const int isz = (sq == 0.f);
        const float s1 = isz ? 1.f : sq;
        const float len = sqrtf(s1);
        TCI(egiz)[ed] = isz; TCF(egs1)[ed] = s1; TCF(eglen)[ed] = len; TCF(eginv)[ed] = 1.f / (e.m.C + len);
      }
      // the first tile's metadata pass starts with a barrier
    }
    // per-column metadata of an edge tile
    __device__ __forceinline__ void meta(int s, int tile) {
      const KernelArgs& a = e.a;
      const int n = e.n, dim = e.dim, D = e.D, ND = e.ND;
      e.qbeg();
      e.epi_bar();
      const uint32_t* tp = e.tile_ptr(kind, tile);
      if (e.tid < 128) {
        const uint32_t w = __ldg(tp + e.tid);
        const int win = (int)__ldg(tp + 192 + TH_WIN);
        int oS = n * ND * TCU, oR = n * ND * TCU, mr = a.lay.mrows * TCU, ijk = 0;     // default: the zero rows
        float sd = 0.f;
        if (w & CW_VALID) {
          const int ed = cw_gid(w), q = cw_q(w);
          const int i = ed / (n - 1), jj = ed - i * (n - 1);
          int j = i + 1 + jj; if (j >= n) j -= n;
          int slot = 0, k = 0;
          if (q == 0) {
            sd = TCF(egs1)[ed];
          } else {
            k = dirmap(ekind(), q - 1, i, j, dim);
            slot = 1 + k;
            float acc = 0.f;
            for (int c = 0; c < dim; ++c)
              acc = fmaf(TCF(egv)[ed * 3 + c], TCF(xt)[(i * dim + c) * D + k] - TCF(xt)[(j * dim + c) * D + k], acc);
            sd = TCI(egiz)[ed] ? 0.f : 2.f * acc;
          }
          if (q == 0 || htan) { oS = (j * ND + slot) * TCU; oR = (i * ND + slot) * TCU; }
          if (!(w & CW_DUP) && want_msg) mr = ((i - win) * ND + slot) * TCU;
          ijk = i | (j << 5) | (k << 10);
        }
        TCW(colw)[s * 128 + e.tid] = w;
        TCI(coloffS)[s * 128 + e.tid] = oS;
        TCI(coloffR)[s * 128 + e.tid] = oR;
        TCF(colsd)[s * 128 + e.tid] = sd;
        TCI(colmrow)[s * 128 + e.tid] = mr;
        TCI(colijk)[s * 128 + e.tid] = ijk;
      }
      e.meta_copy(s, tp);
      e.epi_bar();
      e.qend(P_META);
    }
    // phi_e layer 0 by gather: z0 = P_s[j] + P_r[i] + (|v|^2 or its tangent) w_d, then the activation rule
    __device__ __forceinline__ void build(int s, int tile) {
      const KernelArgs& a = e.a;
      meta(s, tile);
      e.qbeg();
      float v[64];
      const float* ps = e.Ps() + e.f;
      const float* pr = e.Pr() + e.f;
#pragma unroll
      for (int cb = 0; cb < 64; cb += 16) {
        float pa[16], pb[16];
        const int c0 = s * 128 + 64 * e.hh + cb;
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const int4 oS = reinterpret_cast<const int4*>(TCI(coloffS) + c0)[u4], oR = reinterpret_cast<const int4*>(TCI(coloffR) + c0)[u4];
          pa[4 * u4] = ps[oS.x]; pa[4 * u4 + 1] = ps[oS.y]; pa[4 * u4 + 2] = ps[oS.z]; pa[4 * u4 + 3] = ps[oS.w];
          pb[4 * u4] = pr[oR.x]; pb[4 * u4 + 1] = pr[oR.y]; pb[4 * u4 + 2] = pr[oR.z]; pb[4 * u4 + 3] = pr[oR.w];
        }
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const float4 sd = reinterpret_cast<const float4*>(TCF(colsd) + c0)[u4];
          v[cb + 4 * u4] = fmaf(sd.x, wdf, pa[4 * u4] + pb[4 * u4]);
          v[cb + 4 * u4 + 1] = fmaf(sd.y, wdf, pa[4 * u4 + 1] + pb[4 * u4 + 1]);
          v[cb + 4 * u4 + 2] = fmaf(sd.z, wdf, pa[4 * u4 + 2] + pb[4 * u4 + 2]);
          v[cb + 4 * u4 + 3] = fmaf(sd.w, wdf, pa[4 * u4 + 3] + pb[4 * u4 + 3]);
        }
      }
      e.qend(P_GATHER);
      e.act_rule(v, 0.f, s);
      e.write_B(s, v);
      e.qend(P_BUILD);
    }
    __device__ __forceinline__ void epi(int s, int tile, int w) {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      const int L = e.m.L;
      const float bias = w < L - 1 ? bp.be[w + 1][e.f] : bp.bx[w - (L - 1)][e.f];
      float v[64];
      e.qbeg();
      e.ld_acc(s, v);
      e.qend(P_EPI_LD);
      e.act_rule(v, bias, s);
      e.qend(P_EPI_ACT);
      if (w < NL - 1) {
        e.write_B(s, v);
        e.qend(P_EPI_ST);
        if (w == L - 2 && want_msg) messages(s, v);
        e.qend(P_MSG);
      } else {
        e.qend(P_EPI);
        coords(s, v);
        e.qend(P_COORD);
      }
    }
    // attention gate + message aggregation (egnn.py:99-104) from the fp32 phi_e outputs in v
    __device__ __forceinline__ void messages(int s, const float (&v)[64]) {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      const int hh = e.hh, lane = e.lane, warp = e.warp;
      float* pd = TCF(pdot) + s * 512;
      e.warp_dot(v, waf, pd + warp * 64);
      e.half_bar();
      {
        const float bav = bp.ba[0];
        const float* p4 = pd + (4 * hh) * 64;
        float aco[2], bco[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = 2 * lane + u;
          const uint32_t cw = TCW(colw)[s * 128 + 64 * hh + c];
          const int pc = cw_pc(cw);
          const float raw = (p4[c] + p4[64 + c]) + (p4[128 + c] + p4[192 + c]);
          const float rawp = (p4[pc] + p4[64 + pc]) + (p4[128 + pc] + p4[192 + pc]);
          const float eg = ecnf_sigmoid(rawp + bav);
          aco[u] = eg;
          bco[u] = (cw & CW_PRIMAL) ? 0.f : eg * (1.f - eg) * raw;
        }
        *reinterpret_cast<float2*>(TCF(wA) + warp * 64 + 2 * lane) = make_float2(aco[0], aco[1]);
        *reinterpret_cast<float2*>(TCF(wB) + warp * 64 + 2 * lane) = make_float2(bco[0], bco[1]);
      }
      __syncwarp();
      if constexpr (!DIV) {
        // msg = m e per edge column; the columns of one receiver are consecutive, so a running sum is added to the
        // receiver's accumulator row whenever the row changes (padding columns go to the dump row)
        float* mac = TCF(macc) + (size_t)hh * (a.lay.mrows + 1) * TCU + e.f;
        const float* wa_ = TCF(wA) + warp * 64;
        const int* mr_ = TCI(colmrow) + s * 128 + 64 * hh;
        int cur = mr_[0];
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          const int mr = mr_[c];
          if (mr != cur) { mac[cur] += acc; acc = 0.f; cur = mr; }
          acc = fmaf(v[c], wa_[c], acc);
        }
        mac[cur] += acc;
      } else {
        // msg = m e,  msg-dot = m-dot e + m e (1 - e) (m-dot . wa)   accumulated per (receiver, slot) row; columns
        // that do not contribute (repeated primal, padding) go to the dump row
        float* mac = TCF(macc) + (size_t)hh * (a.lay.mrows + 1) * TCU + e.f;
        const float* wa_ = TCF(wA) + warp * 64;
        const float* wb_ = TCF(wB) + warp * 64;
        const int* mr_ = TCI(colmrow) + s * 128 + 64 * hh;
        const uint32_t segs = (TCW(hdr)[s * 16 + TH_SEGS] >> (8 * hh)) & 0xffu;
        float curm = 0.f, eg = 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const bool st = ((segs >> ch) & 1u) != 0u;     // a new edge starts here: its primal message and gate
          curm = st ? v[8 * ch] : curm;
          eg = st ? wa_[8 * ch] : eg;
          // the 8 columns of a chunk belong to one edge: distinct accumulator rows (padding shares the dump row), so
          // the loads need not wait for the stores
          const int4 ma = reinterpret_cast<const int4*>(mr_)[2 * ch], mb = reinterpret_cast<const int4*>(mr_)[2 * ch + 1];
          const float4 wa4 = reinterpret_cast<const float4*>(wb_)[2 * ch], wb4 = reinterpret_cast<const float4*>(wb_)[2 * ch + 1];
          const int mr[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
          const float wbv[8] = {wa4.x, wa4.y, wa4.z, wa4.w, wb4.x, wb4.y, wb4.z, wb4.w};
          float x[8], old[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            x[u] = fmaf(curm, wbv[u], v[8 * ch + u] * eg);
            old[u] = mac[mr[u]];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) mac[mr[u]] = old[u] + x[u];
        }
      }
      if (e.hdr(s, TH_FLUSH)) {
        // the receiver window is complete: write its aggregate to global (coalesced) and clear the accumulators
        e.epi_bar();
        const int win = e.hdr(s, TH_WIN);
        const int nrecv = min(a.lay.mrows / e.ND, e.n - win);
        const int rows = nrecv * e.ND;
        const float inv_sqrt_nb = rsqrtf((float)(e.n - 1));
        float* m0p = TCF(macc) + e.f;
        float* m1p = m0p + (size_t)(a.lay.mrows + 1) * TCU;
        float* dst = e.Mg() + (size_t)win * e.ND * TCU + e.f;
        for (int rr = hh; rr < rows; rr += 2) {
          dst[(size_t)rr * TCU] = (m0p[rr * TCU] + m1p[rr * TCU]) * inv_sqrt_nb;
          m0p[rr * TCU] = 0.f;
          m1p[rr * TCU] = 0.f;
        }
      }
    }
    // coordinate head + coordinate update (egnn.py:82-95) and its tangents
    __device__ __forceinline__ void coords(int s, const float (&v)[64]) {
      const KernelArgs& a = e.a;
      const EcnfBlockParams& bp = e.m.blk[b];
      const int n = e.n, dim = e.dim, D = e.D;
      const int ek = ekind();
      float* pd = TCF(pdot) + s * 512;
      e.warp_dot(v, wpf, pd + e.warp * 64);
      e.half_bar();
      const float bpv = bp.bp[0];
      float* cd = TCF(cdbuf) + s * 384;
      // stage 1: one (column, coordinate) contribution per thread
      for (int cc = 0; cc < dim; ++cc) {
        const int col = 64 * e.hh + (e.f & 63);     // a column of my own half (only its partial dots are complete)
        if ((e.f >> 6) != (cc & 1)) continue;       // the two 64-thread groups of a half alternate over the coordinates
        const uint32_t cw = TCW(colw)[s * 128 + col];
        float val = 0.f;
        if ((cw & CW_VALID) && !(cw & CW_DUP)) {
          const int ed = cw_gid(cw), q = cw_q(cw);
          const int ijk = TCI(colijk)[s * 128 + col];
          const int i = ijk & 31, j = (ijk >> 5) & 31, k = ijk >> 10;
          const float* p4 = pd + (4 * (col >> 6)) * 64;
          const int cl = col & 63, pc = cw_pc(cw);
          const float pg = (p4[pc] + p4[64 + pc]) + (p4[128 + pc] + p4[192 + pc]) + bpv;
          const float vc = TCF(egv)[ed * 3 + cc], inv = TCF(eginv)[ed];
          if (q == 0) {
            val = pg * vc * inv;
          } else if (ek != KIND_LAST || k == i * dim + cc) {
            const float pdv = (p4[cl] + p4[64 + cl]) + (p4[128 + cl] + p4[192 + cl]);
            const float vd = TCF(xt)[(i * dim + cc) * D + k] - TCF(xt)[(j * dim + cc) * D + k];
            const float ld = TCI(egiz)[ed] ? 0.f : TCF(colsd)[s * 128 + col] / (2.f * TCF(eglen)[ed]);
            val = (pdv * vc + pg * vd) * inv - pg * vc * ld * inv * inv;
          }
        }
        cd[col * 3 + cc] = val;
      }
      e.epi_bar();
      {
        // stage 2: fixed-order sums over the edges of a receiver
        const int r = !DIV ? 1 : ek == KIND_MID ? e.ND : ek == KIND_LAST ? 1 + dim : 1 + 2 * dim;
        const int g0 = e.hdr(s, TH_G0), ng = e.hdr(s, TH_NG), i_first = e.hdr(s, TH_IFIRST), i_last = e.hdr(s, TH_ILAST);
        const int per = r * dim;
        for (int idx = e.tid; idx < (i_last - i_first + 1) * per; idx += TC_EPI) {
          const int ri = idx / per, rem = idx - ri * per, q = rem / dim, cc = rem - q * dim;
          const int i = i_first + ri;
          if (ek == KIND_LAST && q != 0 && q - 1 != cc) continue;
          const int ga = max(g0, i * (n - 1)) - g0, gb = min(g0 + ng, (i + 1) * (n - 1)) - g0;
          const bool per_sender = (ek == KIND_FIRST && q > dim);
          float acc = 0.f;
          for (int lg = ga; lg < gb; ++lg) {
            int col = lg;       // primal-only tiles pack densely: column = local group index
            if constexpr (DIV) {
              const uint32_t gw = TCW(grpw)[s * 64 + lg];
              const int qsplit = (int)(gw & 255u), colA = (int)((gw >> 8) & 255u), colB = (int)((gw >> 16) & 255u);
              col = q < qsplit ? colA + q : colB + 1 + (q - qsplit);
            }
            const float val = cd[col * 3 + cc];
            if (per_sender) {
              const int ed = g0 + lg, jj = ed - i * (n - 1);
              int j = i + 1 + jj; if (j >= n) j -= n;
              TCF(xtacc)[(i * dim + cc) * D + j * dim + (q - 1 - dim)] += val;
            } else {
              acc += val;
            }
          }
          if (q == 0) TCF(xacc)[i * dim + cc] += acc;
          else if (ek == KIND_LAST) TCF(dacc)[i * dim + cc] += acc;
          else if (!per_sender) TCF(xtacc)[(i * dim + cc) * D + (ek == KIND_MID ? q - 1 : i * dim + q - 1)] += acc;
        }
      }
    }
  };

  // ---- one evaluation of (f, div f) at time t for the positions in xin (shared memory, D floats) ----
  __device__ __forceinline__ void eval(float t, const float* xin, const int32_t* feat, float* fout) {
    if (is_epi) {
      if (tid < dim) {
        float s = 0.f;
        for (int i = 0; i < n; ++i) s += xin[i * dim + tid];
        TCF(mu)[tid] = s / (float)n;
      }
      if (tid >= 32 && tid < 32 + m.T / 2) {
        const int k = tid - 32;
        const float arg = (t * 1000.f) * m.freqs[k];
        TCF(tau)[k] = sinf(arg);
        TCF(tau)[k + m.T / 2] = cosf(arg);
      }
      epi_bar();
      for (int i = tid; i < D; i += TC_EPI) {
        const float v = xin[i] - TCF(mu)[i % dim];
        TCF(xs)[i] = v;
        TCF(xs0)[i] = v;
      }
      const float invn = 1.f / (float)n;
      if constexpr (DIV)
      for (int idx = tid; idx < D * D; idx += TC_EPI) {
        const int ra = idx / D, k = idx - ra * D;
        const int ia = ra / dim, ca = ra - ia * dim, ik = k / dim, ck = k - ik * dim;
        TCF(xt)[idx] = (ca == ck) ? ((ia == ik ? 1.f : 0.f) - invn) : 0.f;
      }
      float* h0 = hA();
      for (int idx = tid; idx < n * H; idx += TC_EPI) {
        const int node = idx / H, col = idx - node * H;
        int ft = feat[node];
        ft = max(0, min(m.nfeat - 1, ft));
        h0[(size_t)node * ND * H + col] = m.embed[ft * H + col];
      }
      epi_bar();
    }
#pragma unroll 1
    for (int b = 0; b < m.nblocks; ++b) {
      const bool last = (b == m.nblocks - 1);
      const bool htan = DIV && b > 0;
      const int ekind = !DIV ? TT_EDGE1 : last ? TT_LAST : (b == 0 ? TT_FIRST : TT_MID);
      const int nkind = DIV ? TT_NODE : TT_NODE1;
      pbeg();
      {
        const int kind = htan ? TT_NODE : TT_NODE1;
        NodePre p{*this, b, a.tabs.cnt[kind], last ? 3 : 4, kind};
        run_phase(p);
      }
      if (is_epi) epi_bar();     // P_s / P_r / P_h / h_in of every node are in global memory
      pend(P_NODE_PRE);
      {
        EdgePh p{*this, b, a.tabs.cnt[ekind], 2 * m.L - 1, ekind, htan, !last, 0.f, 0.f, 0.f};
        run_phase(p);
      }
      if (is_epi) epi_bar();     // coordinate accumulators and aggregated messages complete
      pend(P_EDGE);
      if (!last) {
        NodePost p{*this, b, a.tabs.cnt[nkind], m.L + 1, nkind, htan};
        run_phase(p);
      }
      if (is_epi) {
        epi_bar();
        const float invnb = 1.f / (float)(n - 1);
        for (int i = tid; i < D; i += TC_EPI) TCF(xs)[i] += TCF(xacc)[i] * invnb;
        if (DIV && !last)
          for (int i = tid; i < D * D; i += TC_EPI) TCF(xt)[i] += TCF(xtacc)[i] * invnb;
        epi_bar();
      }
      pend(P_NODE_POST);
    }
    if (is_epi) {
      const float fs = m.final_scaling[0];
      for (int i = tid; i < D; i += TC_EPI) fout[i] = (TCF(xs)[i] - TCF(xs0)[i] - TCF(mu)[i % dim]) * fs;
      if (DIV && tid == 0) {
        const float invnb = 1.f / (float)(n - 1);
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += TCF(xt)[d * D + d] + TCF(dacc)[d] * invnb;
        fout[D] = fs * (s - (float)D);
      }
    }
    __syncthreads();
  }
};

template <bool DIV>
__global__ void __launch_bounds__(TC_NT, 1) ecnf_solve_tc_kernel(const __grid_constant__ KernelArgs a) {
  __shared__ long long s_traj;
  __shared__ float s_ctl[8];
  EngineTC<DIV> eng(a);
  solve_body<EngineTC<DIV>, DIV>(a, eng, s_traj, s_ctl);
  eng.finish(reinterpret_cast<long long*>(reinterpret_cast<char*>(a.counter) + 64));
}

// ---- weight images: fp32 [K][N] (flax kernel) -> bf16 hi | lo, 128 lanes (out features; zero rows beyond N), two
// consecutive k per 32-bit word, 4 words per 16-byte chunk:  uint4 index = (part * K/8 + chunk) * 128 + lane -----------
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
  uint32_t* dst = reinterpret_cast<uint32_t*>(image + it.dst_off);
  const int words = it.K / 2;                 // per lane and part
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 128 * words; idx += gridDim.x * blockDim.x) {
    const int wd = idx / 128, lane = idx - wd * 128;
    float x0 = 0.f, x1 = 0.f;
    if (lane < it.N) {
      x0 = params[it.src_off + (size_t)(2 * wd) * it.N + lane];
      x1 = params[it.src_off + (size_t)(2 * wd + 1) * it.N + lane];
    }
    uint32_t h, l;
    split_pack(x0, x1, h, l);
    const size_t o = ((size_t)(wd >> 2) * 128 + lane) * 4 + (wd & 3);
    dst[o] = h;
    dst[(size_t)(it.K / 8) * 128 * 4 + o] = l;
  }
}

__global__ void tc_tables_kernel(uint32_t* __restrict__ out, int n, int dim, TcTabs tabs) {
  const int k = threadIdx.x;
  if (k < TT_COUNT) tc_pack(k, n, dim, out + (size_t)tabs.off[k] * TC_TILE_WORDS);
}

#undef TCF
#undef TCI
#undef TCW

}  // namespace ecnf_solve_detail

// Tensor-core (tcgen05 / TMEM) instantiation of engine A + its weight-image preparation.
#include <vector>

#include "ecnf_solve_tc.cuh"

namespace ecnf_solve_detail {

bool tc_eligible(const ecnf_model* mdl, bool div) {
  const ecnf_config& c = mdl->cfg;
  if (c.mlp_units != TCU || c.n_hidden != TCH || c.n_layers < 2) return false;
  const int D = c.n_frames * c.dim;
  if (1 + D > 127) return false;   // a (node, slot) group must fit one tile: 64 + 1 + 63 columns
  const TcSmemLayout lay = make_tc_layout(c.n_frames, c.dim);
  if (!div && lay.mrows < c.n_frames) return false;   // primal-only edge tiles aggregate over all receivers at once
  return lay.total_bytes + 1024 <= 227 * 1024;
}

namespace {
// walks the images in a fixed order; fills the offsets and (optionally) the prep list
int64_t walk_images(const ecnf_model* mdl, TcImages* img, TcPrepList* list) {
  const ecnf_config& c = mdl->cfg;
  const int H = TCH, U = TCU, L = c.n_layers;
  int64_t off = 0;
  int cnt = 0;
  auto add = [&](int64_t src_off, int K, int N) {
    const int64_t o = off;
    if (list && cnt < 64) list->item[cnt] = TcPrepItem{(int)src_off, (int)o, K, N};
    ++cnt;
    off += (int64_t)K * 512;   // hi | lo, 128 lanes x K/2 words each
    return (int)o;
  };
  for (int b = 0; b < c.n_blocks; ++b) {
    const EcnfBlockOffsets& po = mdl->off[b];
    TcImgBlock ib{};
    ib.Wd = add(po.Wd, H, H);
    ib.We0s = add(po.We[0], H, U);
    ib.We0r = add(po.We[0] + (int64_t)H * U, H, U);
    for (int l = 1; l < L; ++l) ib.We[l] = add(po.We[l], U, U);
    for (int l = 0; l < L; ++l) ib.Wx[l] = add(po.Wx[l], U, U);
    ib.Wh0m = add(po.Wh[0], U, U);
    ib.Wh0h = add(po.Wh[0] + (int64_t)U * U, H, U);
    for (int l = 1; l < L; ++l) ib.Wh[l] = add(po.Wh[l], U, U);
    ib.WhL = add(po.Wh[L], U, H);
    if (img) img->blk[b] = ib;
  }
  if (list) list->count = cnt;
  return off;
}
}  // namespace

// tile counts / offsets of the five table kinds
TcTabs make_tabs(const ecnf_model* mdl) {
  TcTabs t{};
  int off = 0;
  for (int k = 0; k < TT_COUNT; ++k) {
    t.off[k] = off;
    t.cnt[k] = tc_pack(k, mdl->cfg.n_frames, mdl->cfg.dim, nullptr);
    off += t.cnt[k];
  }
  return t;
}
int64_t tables_bytes(const TcTabs& t) { return (int64_t)(t.off[TT_COUNT - 1] + t.cnt[TT_COUNT - 1]) * TC_TILE_WORDS * 4; }
int64_t images_bytes(const ecnf_model* mdl) { return (walk_images(mdl, nullptr, nullptr) + 255) & ~255LL; }

int64_t tc_image_bytes(const ecnf_model* mdl) { return images_bytes(mdl) + ((tables_bytes(make_tabs(mdl)) + 255) & ~255LL); }

int64_t tc_flops_per_eval(const ecnf_model* mdl) {
  const ecnf_config& c = mdl->cfg;
  int64_t sumN[TT_COUNT];
  for (int k = 0; k < TT_COUNT; ++k) {
    const int cnt = tc_pack(k, c.n_frames, c.dim, nullptr);
    std::vector<uint32_t> buf((size_t)cnt * TC_TILE_WORDS);
    tc_pack(k, c.n_frames, c.dim, buf.data());
    sumN[k] = 0;
    for (int t = 0; t < cnt; ++t) sumN[k] += buf[(size_t)t * TC_TILE_WORDS + 192 + TH_N];
  }
  auto layer = [](int64_t n_cols, int K) { return 3 * 2 * (int64_t)128 * n_cols * K; };   // hi*hi + lo*hi + hi*lo
  int64_t fl = 0;
  for (int b = 0; b < c.n_blocks; ++b) {
    const bool last = b == c.n_blocks - 1;
    fl += (last ? 3 : 4) * layer(sumN[b > 0 ? TT_NODE : TT_NODE1], TCH);
    fl += (2 * c.n_layers - 1) * layer(sumN[last ? TT_LAST : (b == 0 ? TT_FIRST : TT_MID)], TCU);
    if (!last) fl += (c.n_layers + 1) * layer(sumN[TT_NODE], TCU);
  }
  return fl;
}

int tc_tile_table(const ecnf_model* mdl, int kind, uint32_t* out, int64_t cap_words) {
  if (kind < 0 || kind >= TT_COUNT) return 0;
  const int cnt = tc_pack(kind, mdl->cfg.n_frames, mdl->cfg.dim, nullptr);
  if (out && (int64_t)cnt * TC_TILE_WORDS <= cap_words) tc_pack(kind, mdl->cfg.n_frames, mdl->cfg.dim, out);
  return cnt;
}

int launch_tc(const ecnf_model* mdl, KernelArgs& a, int grid, void* image_ws, bool div, cudaStream_t st) {
  TcPrepList local{};
  a.img.base = reinterpret_cast<const unsigned char*>(image_ws);
  walk_images(mdl, &a.img, &local);
  if (local.count > 64) {
    ecnf_set_error("tensor-core path: %d weight images exceed the prep list", local.count);
    return ECNF_ERR_UNSUPPORTED;
  }
  dim3 pgrid(16, local.count);
  tc_prep_kernel<<<pgrid, 256, 0, st>>>(mdl->d_params, reinterpret_cast<unsigned char*>(image_ws), local);
  ECNF_CHECK_CUDA(cudaGetLastError());
  a.tabs = make_tabs(mdl);
  uint32_t* tab_ws = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(image_ws) + images_bytes(mdl));
  a.tabs.base = tab_ws;
  tc_tables_kernel<<<1, 32, 0, st>>>(tab_ws, mdl->cfg.n_frames, mdl->cfg.dim, a.tabs);
  ECNF_CHECK_CUDA(cudaGetLastError());
  const TcSmemLayout L = make_tc_layout(mdl->cfg.n_frames, mdl->cfg.dim);
  a.lay = L;
  const size_t smem = (size_t)L.total_bytes;
  if (div) {
    ECNF_CHECK_CUDA(cudaFuncSetAttribute(ecnf_solve_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ecnf_solve_tc_kernel<true><<<grid, TC_NT, smem, st>>>(a);
  } else {
    ECNF_CHECK_CUDA(cudaFuncSetAttribute(ecnf_solve_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ecnf_solve_tc_kernel<false><<<grid, TC_NT, smem, st>>>(a);
  }
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

}  // namespace ecnf_solve_detail

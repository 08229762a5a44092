// Tensor-core (tcgen05 / TMEM) instantiation of engine A + its weight-image preparation.
#include "ecnf_solve_tc.cuh"

namespace ecnf_solve_detail {

bool tc_eligible(const ecnf_model* mdl, bool div) {
  const ecnf_config& c = mdl->cfg;
  if (!div || c.mlp_units != TCU || c.n_hidden != TCH || c.n_layers < 2) return false;
  const int D = c.n_frames * c.dim;
  if (1 + D > 128) return false;
  return make_tc_layout(c.n_frames, c.dim, c.n_layers).total_bytes + 1024 <= 227 * 1024;
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
    off += (int64_t)2 * K * N * 2;
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

int64_t tc_image_bytes(const ecnf_model* mdl) { return (walk_images(mdl, nullptr, nullptr) + 255) & ~255LL; }

int launch_tc(const ecnf_model* mdl, KernelArgs& a, int grid, void* image_ws, cudaStream_t st) {
  static TcPrepList list;   // filled per call below (host-side scratch; the call is not re-entrant across threads)
  TcPrepList local{};
  a.img.base = reinterpret_cast<const unsigned char*>(image_ws);
  walk_images(mdl, &a.img, &local);
  if (local.count > 64) {
    ecnf_set_error("tensor-core path: %d weight images exceed the prep list", local.count);
    return ECNF_ERR_UNSUPPORTED;
  }
  (void)list;
  dim3 pgrid(16, local.count);
  tc_prep_kernel<<<pgrid, 256, 0, st>>>(mdl->d_params, reinterpret_cast<unsigned char*>(image_ws), local);
  ECNF_CHECK_CUDA(cudaGetLastError());
  const TcSmemLayout L = make_tc_layout(mdl->cfg.n_frames, mdl->cfg.dim, mdl->cfg.n_layers);
  a.lay = L;
  const size_t smem = (size_t)L.total_bytes;
  ECNF_CHECK_CUDA(cudaFuncSetAttribute(ecnf_solve_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ecnf_solve_tc_kernel<<<grid, NTHREADS, smem, st>>>(a);
  ECNF_CHECK_CUDA(cudaGetLastError());
  return ECNF_OK;
}

}  // namespace ecnf_solve_detail

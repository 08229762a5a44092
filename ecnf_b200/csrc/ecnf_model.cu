// Model handle: hyper-parameters + the flat, 16-byte-aligned fp32 parameter buffer.
// Mirrors the flax variable dict of FlatEgnn (ecnf/cnf/build_cnf.py:65-97):
//   Embed_0/embedding, EGNN_0/Dense_b/{kernel,bias}, EGNN_0/b/{phi_e,phi_x_torso,phi_h}/Dense_l/{kernel,bias},
//   EGNN_0/b/Dense_0 (phi_x head), EGNN_0/b/Dense_1 (attention), EGNN_0/final_scaling.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ecnf_common.cuh"

namespace {
thread_local char g_err[512] = "";

struct TensorInfo {
  std::string name;
  int64_t offset, rows, cols;
};

// Walks the layout in flax order; fills offsets (aligned to 4 floats) and, optionally, the tensor list.
int64_t walk_layout(const ecnf_config& c, ecnf_model* m, std::vector<TensorInfo>* list) {
  int64_t off = 0;
  auto add = [&](const std::string& name, int64_t rows, int64_t cols) {
    const int64_t o = off;
    const int64_t cnt = rows * (cols > 0 ? cols : 1);
    if (list) list->push_back({name, o, rows, cols});
    off += (cnt + 3) & ~3LL;
    return o;
  };
  const int H = c.n_hidden, T = c.time_dim, U = c.mlp_units, L = c.n_layers;
  int64_t embed = add("Embed_0/embedding", c.n_features, H);
  if (m) m->embed_off = embed;
  for (int b = 0; b < c.n_blocks; ++b) {
    EcnfBlockOffsets o{};
    const std::string pre = "EGNN_0/" + std::to_string(b) + "/";
    o.Wd = add("EGNN_0/Dense_" + std::to_string(b) + "/kernel", H + T, H);
    o.bd = add("EGNN_0/Dense_" + std::to_string(b) + "/bias", H, 0);
    for (int l = 0; l < L; ++l) {
      o.We[l] = add(pre + "phi_e/Dense_" + std::to_string(l) + "/kernel", l == 0 ? 2 * H + 1 : U, U);
      o.be[l] = add(pre + "phi_e/Dense_" + std::to_string(l) + "/bias", U, 0);
    }
    for (int l = 0; l < L; ++l) {
      o.Wx[l] = add(pre + "phi_x_torso/Dense_" + std::to_string(l) + "/kernel", U, U);
      o.bx[l] = add(pre + "phi_x_torso/Dense_" + std::to_string(l) + "/bias", U, 0);
    }
    for (int l = 0; l <= L; ++l) {
      o.Wh[l] = add(pre + "phi_h/Dense_" + std::to_string(l) + "/kernel", l == 0 ? U + H : U, l == L ? H : U);
      o.bh[l] = add(pre + "phi_h/Dense_" + std::to_string(l) + "/bias", l == L ? H : U, 0);
    }
    o.wp = add(pre + "Dense_0/kernel", U, 1);
    o.bp = add(pre + "Dense_0/bias", 1, 0);
    o.wa = add(pre + "Dense_1/kernel", U, 1);
    o.ba = add(pre + "Dense_1/bias", 1, 0);
    if (m) m->off[b] = o;
  }
  int64_t fs = add("EGNN_0/final_scaling", 1, 0);
  if (m) m->final_scaling_off = fs;
  return off;
}

int validate(const ecnf_config* c) {
  if (!c) { ecnf_set_error("config is null"); return ECNF_ERR_INVALID; }
  if (c->n_frames < 2 || c->n_frames > ECNF_MAX_NODES) { ecnf_set_error("n_frames=%d outside [2,%d]", c->n_frames, ECNF_MAX_NODES); return ECNF_ERR_UNSUPPORTED; }
  if (c->dim != 2 && c->dim != 3) { ecnf_set_error("dim=%d: only 2 or 3", c->dim); return ECNF_ERR_UNSUPPORTED; }
  if (c->n_blocks < 1 || c->n_blocks > ECNF_MAX_BLOCKS) { ecnf_set_error("n_blocks=%d outside [1,%d]", c->n_blocks, ECNF_MAX_BLOCKS); return ECNF_ERR_UNSUPPORTED; }
  if (c->n_layers < 1 || c->n_layers > ECNF_MAX_LAYERS) { ecnf_set_error("n_layers=%d outside [1,%d]", c->n_layers, ECNF_MAX_LAYERS); return ECNF_ERR_UNSUPPORTED; }
  if (c->mlp_units != 64 && c->mlp_units != 128 && c->mlp_units != 256) { ecnf_set_error("mlp_units=%d: only 64/128/256", c->mlp_units); return ECNF_ERR_UNSUPPORTED; }
  if (c->n_hidden != 32 && c->n_hidden != 64) { ecnf_set_error("n_hidden=%d: only 32/64", c->n_hidden); return ECNF_ERR_UNSUPPORTED; }
  if (c->time_dim < 4 || c->time_dim > ECNF_MAX_T || (c->time_dim & 1)) { ecnf_set_error("time_dim=%d: even, 4..%d", c->time_dim, ECNF_MAX_T); return ECNF_ERR_UNSUPPORTED; }
  if (c->n_features < 1) { ecnf_set_error("n_features=%d < 1", c->n_features); return ECNF_ERR_INVALID; }
  if (!(c->base_scale > 0.f)) { ecnf_set_error("base_scale must be > 0"); return ECNF_ERR_INVALID; }
  return ECNF_OK;
}
}  // namespace

void ecnf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

EcnfModelDev ecnf_make_dev(const ecnf_model* m, const float* p) {
  EcnfModelDev d;
  memset(&d, 0, sizeof(d));
  const ecnf_config& c = m->cfg;
  d.n = c.n_frames; d.dim = c.dim; d.H = c.n_hidden; d.T = c.time_dim; d.U = c.mlp_units; d.L = c.n_layers;
  d.nblocks = c.n_blocks; d.nfeat = c.n_features; d.C = c.normalization_constant; d.base_scale = c.base_scale;
  d.sigma_min = c.sigma_min;
  for (int k = 0; k < ECNF_MAX_T / 2; ++k) d.freqs[k] = c.freqs[k];
  d.embed = p + m->embed_off;
  d.final_scaling = p + m->final_scaling_off;
  for (int b = 0; b < c.n_blocks; ++b) {
    const EcnfBlockOffsets& o = m->off[b];
    EcnfBlockParams& q = d.blk[b];
    q.Wd = p + o.Wd; q.bd = p + o.bd;
    for (int l = 0; l < c.n_layers; ++l) {
      q.We[l] = p + o.We[l]; q.be[l] = p + o.be[l];
      q.Wx[l] = p + o.Wx[l]; q.bx[l] = p + o.bx[l];
    }
    for (int l = 0; l <= c.n_layers; ++l) { q.Wh[l] = p + o.Wh[l]; q.bh[l] = p + o.bh[l]; }
    q.wp = p + o.wp; q.bp = p + o.bp; q.wa = p + o.wa; q.ba = p + o.ba;
  }
  return d;
}

extern "C" {

const char* ecnf_last_error(void) { return g_err; }
int ecnf_version(void) { return 100; }

int ecnf_model_create(const ecnf_config* cfg, const float* d_params, ecnf_model** out) {
  if (!out) { ecnf_set_error("out is null"); return ECNF_ERR_INVALID; }
  *out = nullptr;
  int rc = validate(cfg);
  if (rc != ECNF_OK) return rc;
  ecnf_model* m = new ecnf_model();
  memset(m, 0, sizeof(*m));
  m->cfg = *cfg;
  m->param_count = walk_layout(*cfg, m, nullptr);
  m->d_params = d_params;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    m->num_sms = 148;  // B200; lets the handle (layout queries) be used on a host without a GPU
    dev = -1;
  }
  m->device = dev;
  m->engine = 0;
  *out = m;
  return ECNF_OK;
}

void ecnf_model_destroy(ecnf_model* m) { delete m; }

int ecnf_model_clone(const ecnf_model* m, const float* d_params, ecnf_model** out) {
  if (!m || !out) { ecnf_set_error("ecnf_model_clone: null argument"); return ECNF_ERR_INVALID; }
  ecnf_model* c = new ecnf_model(*m);     // config, layout, engine choice, training chunk size: plain data
  if (d_params) c->d_params = d_params;
  *out = c;
  return ECNF_OK;
}

int ecnf_model_set_params(ecnf_model* m, const float* d_params) {
  if (!m) { ecnf_set_error("model is null"); return ECNF_ERR_INVALID; }
  m->d_params = d_params;
  return ECNF_OK;
}

int64_t ecnf_model_param_count(const ecnf_model* m) { return m ? m->param_count : 0; }

int ecnf_model_num_tensors(const ecnf_model* m) {
  if (!m) return 0;
  std::vector<TensorInfo> list;
  walk_layout(m->cfg, nullptr, &list);
  return (int)list.size();
}

int ecnf_model_param_layout(const ecnf_model* m, int idx, char* name, int name_cap, int64_t* offset, int64_t* rows,
                            int64_t* cols) {
  if (!m) { ecnf_set_error("model is null"); return ECNF_ERR_INVALID; }
  std::vector<TensorInfo> list;
  walk_layout(m->cfg, nullptr, &list);
  if (idx < 0 || idx >= (int)list.size()) { ecnf_set_error("tensor index %d out of range", idx); return ECNF_ERR_INVALID; }
  const TensorInfo& t = list[idx];
  if (name && name_cap > 0) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (offset) *offset = t.offset;
  if (rows) *rows = t.rows;
  if (cols) *cols = t.cols;
  return ECNF_OK;
}

float ecnf_warmup_cosine_lr(int64_t step, float init_value, float peak_value, int64_t warmup_steps, int64_t decay_steps,
                            float end_value) {
  // optax.warmup_cosine_decay_schedule (setup_training.py:100-106)
  if (step < warmup_steps) return init_value + (peak_value - init_value) * (float)step / (float)(warmup_steps > 0 ? warmup_steps : 1);
  int64_t n = decay_steps - warmup_steps;
  if (n < 1) n = 1;
  int64_t s = step - warmup_steps;
  if (s > n) s = n;
  const double c = 0.5 * (1.0 + cos(3.14159265358979323846 * (double)s / (double)n));
  return end_value + (peak_value - end_value) * (float)c;
}

}  // extern "C"

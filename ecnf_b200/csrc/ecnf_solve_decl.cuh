// Declarations shared by the engine-A dispatch (ecnf_solve.cu) and its per-(U,H) instantiation units.
#pragma once
#include "ecnf_common.cuh"

namespace ecnf_solve_detail {

// per-CTA global scratch (floats): h (current), h_in, P_s, P_r, aggregated messages
__host__ __device__ inline int64_t scratch_floats(int n, int dim, int H, int U, bool div) {
  const int64_t ND = div ? 1 + n * dim : 1;
  return n * ND * (2 * (int64_t)H + 3 * (int64_t)U);
}

struct KernelArgs {
  EcnfModelDev m;
  int mode;
  long long B;
  const float* x_init;   // [B, D]
  const float* t_in;     // [B] (VF modes)
  const int32_t* feat;   // [B, n]
  ecnf_solve_ctrl ctrl;
  float* out_x;          // [B, D]   (VF modes: f)
  float* out_logs;       // [B, 3]   (VF_DIV: out_div [B])
  int32_t* out_stats;    // [B, 4]
  float* scratch;        // per-CTA global scratch
  long long scratch_stride;
  unsigned int* counter;
};

template <int U, int H, bool DIV>
int launch_t(const ecnf_model* mdl, KernelArgs& a, int grid, cudaStream_t st);

}  // namespace ecnf_solve_detail

"""Coarse per-phase cycle breakdown of the tensor-core solve kernel (CTA 0), read back from the workspace header.
Usage (on a GPU box, library built with ECNF_TC_PROFILE=1): python tools/tc_profile.py [batch] [sample]   ("sample": the
primal-only sample_cnf path instead of sample + exact log q)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ecnf_b200 import lib as L
from ecnf_b200.cnf import build_cnf
from ecnf_b200.engine import PackedParams
from ecnf_b200.nets.egnn import init_flat_params

NAMES = ["node_pre", "edge", "node_post", "  wait_mma(epi)", "  build", "  epilogue", "  messages", "  coords", "  weight_load", "  tile_meta",
         "  issue_warp_wait", "misc", "  epi_ld", "  epi_act", "  epi_st", "  arrive", "  build_gather"]
NTOP = 3
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
MODE = L.MODE_SAMPLE if len(sys.argv) > 2 and sys.argv[2] == "sample" else L.MODE_SAMPLE_LOGQ
cnf = build_cnf(13, 3, 0.01, 1.0, 3, (128, 128, 128), 64, 8, 1)
eng = cnf.engine
params = PackedParams(torch.from_numpy(init_flat_params(eng, 0, 1.0)).cuda())
x0 = eng.base_sample(2, B)
for it in range(2):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    x1, logs, stats = eng.solve(params, MODE, x0, None, L.make_ctrl(use_fixed_step_size=True))
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
ws = eng._ws["solve"]
prof = ws[64:64 + 8 * len(NAMES)].view(torch.int64).cpu().numpy()
evals = int(stats[:, 2].sum()) / min(B, 148)
tot = prof[:NTOP].sum()
print(f"B={B} kernel {ms:.1f} ms; CTA0 ran ~{evals:.0f} evals; instrumented cycles {tot/1e6:.1f} M ({tot/evals/1e3:.0f} k/eval)")
for nm, v in zip(NAMES, prof):
    print(f"  {nm:12s} {v/1e6:10.2f} Mcyc  {100*v/tot:5.1f}%   {v/evals/1e3:8.1f} kcyc/eval")

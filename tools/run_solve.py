"""One launch of the solve kernel for profiling.  Usage: python tools/run_solve.py [lj13|aldp|dw4] [batch] [sample|logq] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ecnf_b200 import lib as L
from ecnf_b200.cnf import build_cnf
from ecnf_b200.engine import PackedParams
from ecnf_b200.nets.egnn import init_flat_params

name = sys.argv[1] if len(sys.argv) > 1 else "lj13"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148
mode = L.MODE_SAMPLE if len(sys.argv) > 3 and sys.argv[3] == "sample" else L.MODE_SAMPLE_LOGQ
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
cfg = bench.CFGS[name]
cnf = build_cnf(**cfg)
eng = cnf.engine
params = PackedParams(torch.from_numpy(init_flat_params(eng, 0, 1.0)).cuda())
x0 = eng.base_sample(2, B)
feat = (torch.arange(cfg["n_frames"], dtype=torch.int32, device="cuda") % cfg["n_features"]).repeat(B, 1).contiguous()
for it in range(reps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    x1, logs, stats = eng.solve(params, mode, x0, feat, L.make_ctrl(use_fixed_step_size=True))
    b.record()
    torch.cuda.synchronize()
    print(f"{name} B={B} {a.elapsed_time(b):.1f} ms  {B / a.elapsed_time(b) * 1e3:.1f} samples/s", flush=True)

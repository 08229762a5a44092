"""One small tensor-core solve (LJ13 and DW4, VF+div and a short fixed-step solve) for compute-sanitizer runs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ecnf_b200 import lib as L
from ecnf_b200.cnf import build_cnf
from ecnf_b200.engine import PackedParams
from ecnf_b200.nets.egnn import init_flat_params

for n, dim in ((13, 3), (4, 2)):
    cnf = build_cnf(n, dim, 0.01, 1.0, 3, (128, 128, 128), 64, 8, 1)
    eng = cnf.engine
    params = PackedParams(torch.from_numpy(init_flat_params(eng, 0, 1.0)).cuda())
    x0 = eng.base_sample(2, 3)
    t = torch.tensor([0.1, 0.5, 0.9], device="cuda")
    f, d = eng.apply_div(params, x0, t)
    x1, logs, stats = eng.solve(params, L.MODE_SAMPLE_LOGQ, x0, None, L.make_ctrl(use_fixed_step_size=True, step_size=0.5))
    torch.cuda.synchronize()
    print(n, dim, float(f.abs().max()), d.tolist(), logs[:, 0].tolist(), stats[:, 2].tolist())

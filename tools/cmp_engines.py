"""Compare the tensor-core engine with the fp32 SIMT engine on the adaptive 'stiffened' DW4 / LJ13 solves (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import ecnf_oracle as O
from ecnf_b200 import lib as L
from ecnf_b200.engine import Engine
from helpers import CASES, make_pair

for case, B in (("dw4", 8), ("lj13", 4)):
    n, dim, blocks, units, H, nfeat = CASES[case]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat, head_variance=1.0)
    eng = Engine(ecfg)
    rng = np.random.default_rng(7)
    eps = rng.standard_normal((B, n * dim)).astype(np.float32)
    feat = rng.integers(0, nfeat, (B, n)).astype(np.int32)
    x0 = O.base_sample_from_noise(ocfg, torch.tensor(eps)).numpy()
    res = {}
    for e in (1, 0):
        eng.set_engine(e)
        x1, logs, stats = eng.solve(tree, L.MODE_SAMPLE_LOGQ, x0, feat, L.make_ctrl())
        res[e] = (x1.cpu().numpy(), logs.cpu().numpy(), stats.cpu().numpy())
        print(case, "engine", e, "steps", res[e][2][:, 0], "acc", res[e][2][:, 1], "evals", res[e][2][:, 2], "status", res[e][2][:, 3])
    print(case, "max |dx|", np.abs(res[0][0] - res[1][0]).max(), "max |dlogq|", np.abs(res[0][1][:, 0] - res[1][1][:, 0]).max(),
          "logq", res[1][1][:, 0])
    # one evaluation: f and div differences
    t = rng.uniform(0, 1, B).astype(np.float32)
    out = {}
    for e in (1, 0):
        eng.set_engine(e)
        f, d = eng.apply_div(tree, x0, t, feat)
        out[e] = (f.cpu().numpy(), d.cpu().numpy())
    print(case, "vf rel diff", np.abs(out[0][0] - out[1][0]).max() / np.abs(out[1][0]).max(), "div diff", np.abs(out[0][1] - out[1][1]).max(), "div", out[1][1][:3])
eng.set_engine(0)

"""Throughput of the BASELINE.json configurations on one GPU (small batches; fixed dt = 0.05 so the work is exact).
Usage (GPU box): python tools/throughput.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ecnf_b200 import lib as L
from ecnf_b200.cnf import build_cnf
from ecnf_b200.engine import PackedParams
from ecnf_b200.nets.egnn import init_flat_params

CFGS = {
    "dw4 (configs[0])": (dict(n_frames=4, dim=2, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128,) * 3,
                              n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1), 148 * 8),
    "lj13 (configs[1])": (dict(n_frames=13, dim=3, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128,) * 3,
                               n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1), 148 * 2),
    "aldp (configs[3])": (dict(n_frames=22, dim=3, sigma_min=1e-6, base_scale=0.2, n_blocks_egnn=3, mlp_units=(64, 64),
                               n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=22), 148),
    "qm9pos (configs[2] net)": (dict(n_frames=19, dim=3, sigma_min=1e-6, base_scale=2.0, n_blocks_egnn=5, mlp_units=(256,) * 4,
                                     n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=1), 148),
}
for name, (cfg, B) in CFGS.items():
    cnf = build_cnf(**cfg)
    eng = cnf.engine
    params = PackedParams(torch.from_numpy(init_flat_params(eng, 0, 1.0)).cuda())
    x0 = eng.base_sample(2, B)
    feat = None
    if cfg["n_features"] > 1:
        feat = torch.arange(cfg["n_frames"], dtype=torch.int32, device="cuda").repeat(B, 1)
    ctrl = L.make_ctrl(use_fixed_step_size=True)
    for mode, label in ((L.MODE_SAMPLE_LOGQ, "sample + exact log q"), (L.MODE_SAMPLE, "sample only")):
        if name.startswith("qm9") and mode == L.MODE_SAMPLE_LOGQ and "--all" not in sys.argv:
            continue
        # the plain-sampling path is ~1 + D times cheaper per trajectory: use a larger batch for a stable number
        rep = 16 if mode == L.MODE_SAMPLE else 1
        xs = x0.repeat(rep, 1)
        fs = None if feat is None else feat.repeat(rep, 1)
        eng.solve(params, mode, x0[:8], None if feat is None else feat[:8], ctrl)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, _, st = eng.solve(params, mode, xs, fs, ctrl)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"{name:26s} {label:22s} B={B * rep:6d}  {ms:9.1f} ms  {B * rep / ms * 1e3:10.1f} samples/s  evals/sample {int(st[0, 2])}", flush=True)
    if name.startswith("qm9") and "--all" not in sys.argv:
        continue
    # the reference's approx=True branch (Hutchinson probe, one tangent): tensor-core engine vs fp32 SIMT
    rep = 4
    xs = x0.repeat(rep, 1)
    fs = None if feat is None else feat.repeat(rep, 1)
    eps = torch.randn_like(xs)
    for engine, label in ((0, "hutchinson (auto)"), (1, "hutchinson (fp32 SIMT)")):
        eng.set_engine(engine)
        eng.solve(params, L.MODE_SAMPLE_LOGQ, x0[:8], None if feat is None else feat[:8], ctrl, eps=eps[:8])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, _, st = eng.solve(params, L.MODE_SAMPLE_LOGQ, xs, fs, ctrl, eps=eps)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"{name:26s} {label:22s} B={B * rep:6d}  {ms:9.1f} ms  {B * rep / ms * 1e3:10.1f} samples/s  evals/sample {int(st[0, 2])}", flush=True)
    eng.set_engine(0)

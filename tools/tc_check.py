"""Quick numerical check of the tensor-core engine against the fp32 SIMT engine and the oracle, one case per subprocess
(a hang or fault in one case does not take the others down).  Usage (GPU box): python tools/tc_check.py [case ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    # name: (n, dim, blocks, units, H, n_features)
    "dw4": (4, 2, 3, (128, 128, 128), 64, 1),
    "lj13": (13, 3, 3, (128, 128, 128), 64, 1),
    "small_64_32": (5, 3, 2, (64, 64), 32, 3),
    "one_block": (6, 3, 1, (64, 64), 32, 1),
    "aldp": (22, 3, 3, (64, 64), 32, 22),
    "n2": (2, 2, 3, (128, 128), 64, 1),
}


def run_case(name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    from oracle import ecnf_oracle as O
    from ecnf_b200 import lib as L
    from ecnf_b200.engine import Engine
    from helpers import make_pair, rel_err
    if name in CASES:
        n, dim, blocks, units, H, nfeat = CASES[name]
    else:      # "n,dim,blocks,U,H,L"
        n, dim, blocks, U_, H, L_ = (int(v) for v in name.split(","))
        units, nfeat = (U_,) * L_, 1
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat, head_variance=0.5)
    eng = Engine(ecfg)
    B = 5
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, n * dim)).astype(np.float32) * 1.3 + 0.2
    t = rng.uniform(0, 1, B).astype(np.float32)
    feat = rng.integers(0, nfeat, (B, n)).astype(np.int32)
    p64 = O.to_torch(flat, torch.float64)
    f_ref, div_ref = O.vf_and_exact_div(p64, ocfg, torch.tensor(x, dtype=torch.float64), torch.tensor(t, dtype=torch.float64),
                                        torch.tensor(feat).long())
    for engine in (1, 0):
        eng.set_engine(engine)
        tag = "simt" if engine else "tc  "
        f = eng.apply(tree, x, t, feat)
        torch.cuda.synchronize()
        print(f"{name:12s} {tag} vf      rel err {rel_err(f.cpu().numpy(), f_ref.numpy()):.2e}", flush=True)
        f2, div = eng.apply_div(tree, x, t, feat)
        torch.cuda.synchronize()
        derr = np.abs(div.cpu().numpy() - div_ref.numpy()).max() / (np.abs(div_ref.numpy()).max() + 1.0)
        print(f"{name:12s} {tag} vf+div  rel err f {rel_err(f2.cpu().numpy(), f_ref.numpy()):.2e}  div {derr:.2e}", flush=True)
    if os.environ.get("TC_CHECK_QUICK"):
        return
    x0 = O.base_sample_from_noise(ocfg, torch.tensor(rng.standard_normal((B, n * dim)).astype(np.float32)))
    ctrl = L.make_ctrl(use_fixed_step_size=True, step_size=0.25)
    res = {}
    for engine in (1, 0):
        eng.set_engine(engine)
        res[engine] = eng.solve(tree, L.MODE_SAMPLE_LOGQ, x0.numpy(), feat, ctrl)
        torch.cuda.synchronize()
    a, c = res[0], res[1]
    print(f"{name:12s} solve tc vs simt: x {rel_err(a[0].cpu().numpy(), c[0].cpu().numpy()):.2e}  "
          f"logq {np.abs(a[1].cpu().numpy()[:, 0] - c[1].cpu().numpy()[:, 0]).max() / (np.abs(c[1].cpu().numpy()[:, 0]).max() + 1):.2e}", flush=True)
    # many trajectories: the work queue, two launches bit-identical
    Bm = 300
    x0m = O.base_sample_from_noise(ocfg, torch.tensor(rng.standard_normal((Bm, n * dim)).astype(np.float32))).numpy()
    featm = rng.integers(0, nfeat, (Bm, n)).astype(np.int32)
    eng.set_engine(0)
    r1 = eng.solve(tree, L.MODE_SAMPLE_LOGQ, x0m, featm, ctrl)
    r2 = eng.solve(tree, L.MODE_SAMPLE_LOGQ, x0m, featm, ctrl)
    eng.set_engine(1)
    r3 = eng.solve(tree, L.MODE_SAMPLE_LOGQ, x0m, featm, ctrl)
    torch.cuda.synchronize()
    print(f"{name:12s} B=300 deterministic {bool(torch.equal(r1[0], r2[0]) and torch.equal(r1[1], r2[1]))}  "
          f"tc vs simt x {rel_err(r1[0].cpu().numpy(), r3[0].cpu().numpy()):.2e} "
          f"logq {np.abs(r1[1].cpu().numpy()[:, 0] - r3[1].cpu().numpy()[:, 0]).max() / (np.abs(r3[1].cpu().numpy()[:, 0]).max() + 1):.2e}", flush=True)
    eng.set_engine(0)
    s1 = eng.solve(tree, L.MODE_SAMPLE, x0m, featm, ctrl)
    eng.set_engine(1)
    s3 = eng.solve(tree, L.MODE_SAMPLE, x0m, featm, ctrl)
    torch.cuda.synchronize()
    print(f"{name:12s} B=300 sample only tc vs simt x {rel_err(s1[0].cpu().numpy(), s3[0].cpu().numpy()):.2e}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", nm], timeout=int(os.environ.get("TC_CHECK_TIMEOUT", "120")), capture_output=True, text=True)
            print(r.stdout, end="")
            if r.returncode != 0:
                print(f"{nm}: exit {r.returncode}\n{r.stderr[-1500:]}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"{nm}: TIMEOUT (hang)\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else (e.stdout or '')}", flush=True)

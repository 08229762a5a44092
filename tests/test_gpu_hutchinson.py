"""GPU parity of the reference's `approx=True` branch (Hutchinson divergence, sample_and_log_prob.py:69-78, :123-133):
one forward-mode tangent in the probe direction (CUDA) vs the oracle's reverse-mode eps^T J eps."""
import numpy as np
import pytest
import torch

from oracle import ecnf_oracle as O
from ecnf_b200 import lib as L
from ecnf_b200.cnf import build_cnf, get_log_prob, sample_and_log_prob_cnf
from ecnf_b200.engine import Engine
from helpers import CASES, make_pair, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.mark.parametrize("case", list(CASES))
def test_hutchinson_div_matches_oracle(case, cuda_device):
    n, dim, blocks, units, H, nfeat = CASES[case]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat)
    eng = Engine(ecfg)
    B = 6
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, n * dim)).astype(np.float32) * 1.2 + 0.1
    t = rng.uniform(0, 1, B).astype(np.float32)
    eps = rng.standard_normal((B, n * dim)).astype(np.float32)
    feat = rng.integers(0, nfeat, (B, n)).astype(np.int32)
    p64 = O.to_torch(flat, torch.float64)
    f_ref, h_ref = O.vf_and_hutchinson_div(p64, ocfg, torch.tensor(x, dtype=torch.float64), torch.tensor(t, dtype=torch.float64),
                                           torch.tensor(feat).long(), torch.tensor(eps, dtype=torch.float64))
    f, h = eng.apply_hutchinson(tree, x, t, eps, feat)
    assert rel_err(f.cpu().numpy(), f_ref.numpy()) < 2e-5
    assert np.abs(h.cpu().numpy() - h_ref.numpy()).max() < 5e-5 * (np.abs(h_ref.numpy()).max() + 1.0)
    # a basis probe picks one diagonal entry of the Jacobian: consistent with the exact trace path
    e0 = np.zeros_like(eps); e0[:, 0] = 1.0
    _, d0 = eng.apply_hutchinson(tree, x, t, e0, feat)
    _, d0_ref = O.vf_and_hutchinson_div(p64, ocfg, torch.tensor(x, dtype=torch.float64), torch.tensor(t, dtype=torch.float64),
                                        torch.tensor(feat).long(), torch.tensor(e0, dtype=torch.float64))
    assert np.abs(d0.cpu().numpy() - d0_ref.numpy()).max() < 5e-5 * (np.abs(d0_ref.numpy()).max() + 1.0)


@pytest.mark.parametrize("case", ["small_64_32", "dw4"])
def test_fixed_step_approx_logq_matches_oracle(case, cuda_device):
    n, dim, blocks, units, H, nfeat = CASES[case]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat)
    cnf = build_cnf(n, dim, ocfg.sigma_min, ocfg.base_scale, blocks, tuple(units), H, 8, nfeat)
    B = 5
    rng = np.random.default_rng(12)
    x0 = O.base_sample_from_noise(ocfg, torch.tensor(rng.standard_normal((B, n * dim)), dtype=torch.float32))
    eps = torch.tensor(rng.standard_normal((B, n * dim)), dtype=torch.float32)
    feat = rng.integers(0, nfeat, (B, n)).astype(np.int32)
    ft = torch.tensor(feat).long()
    p32 = O.to_torch(flat, torch.float32)
    ctrl = O.SolveControl(fixed=True, step_size=0.05)
    x1_ref, lq_ref, _ = O.sample_and_log_prob_cnf(p32, ocfg, x0, ft, ctrl, eps=eps)
    x1, lq, stats = sample_and_log_prob_cnf(cnf, tree, 0, feat, approx=True, use_fixed_step_size=True, x0=x0, eps=eps,
                                           return_stats=True)
    assert (stats.cpu().numpy()[:, 2] == 121).all()
    assert rel_err(x1.cpu().numpy(), x1_ref.numpy()) < TOL
    assert np.abs(lq.cpu().numpy() - lq_ref.numpy()).max() < TOL * (np.abs(lq_ref.numpy()).max() + 1)
    lp_ref, lpb_ref, d_ref, _ = O.get_log_prob(p32, ocfg, x1_ref, ft, ctrl, eps=eps)
    lp, lpb, delta = get_log_prob(cnf, tree, x1_ref.numpy(), 0, feat, approx=True, use_fixed_step_size=True, eps=eps)
    assert np.abs(lp.cpu().numpy() - lp_ref.numpy()).max() < TOL * (np.abs(lp_ref.numpy()).max() + 1)
    assert np.abs(delta.cpu().numpy() - d_ref.numpy()).max() < TOL * (np.abs(d_ref.numpy()).max() + 1)


def test_probe_streams(cuda_device):
    """substream 0 is the raw noise underneath sample_base(key) (the reference reuses the base-sample key for the probe,
    SURVEY C#6); key-driven approx calls are deterministic and differ between keys."""
    n, dim, blocks, units, H, nfeat = CASES["dw4"]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H)
    cnf = build_cnf(n, dim, ocfg.sigma_min, ocfg.base_scale, blocks, tuple(units), H, 8, 1)
    eng = cnf.engine
    raw = eng.normal_noise(5, 7, 3, substream=0)
    assert torch.equal(eng.base_sample_from_noise(raw), eng.base_sample(5, 7, 3))
    assert abs(float(raw.mean())) < 0.5 and 0.5 < float(raw.std()) < 1.5
    a = sample_and_log_prob_cnf(cnf, tree, 5, None, approx=True, use_fixed_step_size=True, n_samples=4)
    b = sample_and_log_prob_cnf(cnf, tree, 5, None, approx=True, use_fixed_step_size=True, n_samples=4)
    c = sample_and_log_prob_cnf(cnf, tree, 6, None, approx=True, use_fixed_step_size=True, n_samples=4)
    assert torch.equal(a[1], b[1]) and not torch.equal(a[1], c[1])
    e = sample_and_log_prob_cnf(cnf, tree, 5, None, approx=False, use_fixed_step_size=True, n_samples=4)
    # same trajectories (the exact path may run on the tensor-core engine: equal to rounding), another divergence estimate
    assert float((a[0] - e[0]).abs().max()) < 1e-4 * float(e[0].abs().max())
    assert not torch.equal(a[1], e[1])


@pytest.mark.parametrize("shape", ["lj13", "aldp"])
def test_hutchinson_tensor_core_engine_matches_simt_through_the_ode_loop(shape, cuda_device):
    """The one-tangent (probe) mode of the tcgen05 engine against the fp32 SIMT engine, fixed-step sample + log q with
    injected probes, on the full LJ13 net and the full ALDP net (22 atoms: several message windows per block)."""
    n, dim, blocks, units, H, nfeat = CASES["lj13"] if shape == "lj13" else (22, 3, 3, (64, 64), 32, 22)
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat, head_variance=0.3)
    eng = Engine(ecfg)
    B = 5
    rng = np.random.default_rng(13)
    x0 = O.base_sample_from_noise(ocfg, torch.tensor(rng.standard_normal((B, n * dim)), dtype=torch.float32)).numpy()
    eps = rng.standard_normal((B, n * dim)).astype(np.float32)
    feat = np.tile(np.arange(n) % nfeat, (B, 1)).astype(np.int32)
    ctrl = L.make_ctrl(use_fixed_step_size=True)
    out = {}
    for engine in (0, 1):
        eng.set_engine(engine)
        x1, logs, stats = eng.solve(tree, L.MODE_SAMPLE_LOGQ, x0, feat, ctrl, eps=eps)
        out[engine] = (x1.cpu().numpy(), logs.cpu().numpy(), stats.cpu().numpy())
    eng.set_engine(0)
    assert (out[0][2][:, 2] == 121).all() and (out[0][2][:, 3] == 0).all()
    assert rel_err(out[0][0], out[1][0]) < TOL
    assert np.abs(out[0][1][:, 0] - out[1][1][:, 0]).max() < TOL * (np.abs(out[1][1][:, 0]).max() + 1)

"""GPU parity for the small kernels: base distribution, Philox noise, ESS statistics, target energies."""
import numpy as np
import pytest
import torch

from oracle import ecnf_oracle as O
from ecnf_b200.engine import Engine, ess_from_stats
from ecnf_b200 import lib as L
from helpers import CASES, make_pair

pytestmark = pytest.mark.gpu


def test_base_sample_and_log_prob(cuda_device):
    n, dim, blocks, units, H, nfeat = CASES["lj13"]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, base_scale=2.0)
    eng = Engine(ecfg)
    eps = np.random.default_rng(0).standard_normal((50, n * dim)).astype(np.float32)
    x0 = eng.base_sample_from_noise(eps).cpu().numpy()
    ref = O.base_sample_from_noise(ocfg, torch.tensor(eps)).numpy()
    assert np.abs(x0 - ref).max() < 1e-6
    lp = eng.base_log_prob(ref + 0.3).cpu().numpy()       # un-centred input: log_prob removes the mean first
    lp_ref = O.base_log_prob(ocfg, torch.tensor(ref + 0.3, dtype=torch.float64)).numpy()
    assert np.abs(lp - lp_ref).max() < 1e-4 * np.abs(lp_ref).max()


def test_philox_draws_are_shard_invariant_and_gaussian(cuda_device):
    n, dim, blocks, units, H, nfeat = CASES["lj13"]
    eng = Engine(make_pair(n, dim, blocks, units, H)[3])
    full = eng.base_sample(1234, 20000).cpu().numpy()
    a = eng.base_sample(1234, 12000, 0).cpu().numpy()
    b = eng.base_sample(1234, 8000, 12000).cpu().numpy()
    assert np.array_equal(full, np.concatenate([a, b]))           # keyed by global index
    assert not np.array_equal(full, eng.base_sample(1235, 20000).cpu().numpy())
    x = full.reshape(-1, n, dim)
    assert np.abs(x.mean(axis=1)).max() < 1e-5                    # zero centre of mass
    # remove_mean(N(0, I)) has per-coordinate variance (n-1)/n
    assert abs(x.var() - (n - 1) / n) < 0.01
    k = ((x - x.mean()) ** 4).mean() / x.var() ** 2
    assert abs(k - 3.0) < 0.1
    x0, t = eng.fm_draw_noise(7, 50000)
    t = t.cpu().numpy()
    assert t.min() >= 0 and t.max() < 1 and abs(t.mean() - 0.5) < 0.01 and abs(t.var() - 1 / 12) < 0.005


def test_ess_stats(cuda_device):
    eng = Engine(make_pair(*CASES["dw4"][:5])[3])
    rng = np.random.default_rng(3)
    lw = (rng.standard_normal(10001) * 2 - 40).astype(np.float32)
    st = eng.ess_stats(torch.tensor(lw)).cpu().numpy()
    rv, fw = ess_from_stats(st, lw.size)
    assert abs(rv - O.reverse_ess(lw.astype(np.float64))) < 1e-4 * rv
    assert abs(fw - O.forward_ess(lw.astype(np.float64), np.ones(lw.size, bool))) < 1e-4 * fw


def test_target_energies(cuda_device):
    rng = np.random.default_rng(4)
    eng = Engine(make_pair(*CASES["lj13"][:5])[3])
    x = (rng.standard_normal((64, 13, 3)) * 1.2).astype(np.float32)
    lp = eng.target_log_prob(L.TARGET_LJ, x.reshape(64, -1)).cpu().numpy()
    ref = -O.lj_energy(x.astype(np.float64))
    assert np.abs(lp - ref).max() < 1e-4 * np.abs(ref).max()
    eng4 = Engine(make_pair(*CASES["dw4"][:5])[3])
    x = (rng.standard_normal((64, 4, 2)) * 2).astype(np.float32)
    lp = eng4.target_log_prob(L.TARGET_DW, x.reshape(64, -1)).cpu().numpy()
    ref = -O.dw_energy(x.astype(np.float64))
    assert np.abs(lp - ref).max() < 1e-5 * np.abs(ref).max()

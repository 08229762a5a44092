"""GPU parity: vector field and exact divergence (forward-mode, CUDA) vs the CPU oracle (reverse-mode, fp64)."""
import numpy as np
import pytest
import torch

from oracle import ecnf_oracle as O
from ecnf_b200.engine import Engine
from helpers import CASES, make_pair, rel_err

pytestmark = pytest.mark.gpu

VF_TOL = 2e-5    # fp32 kernel vs fp64 oracle, relative to max |f|
DIV_TOL = 5e-5


@pytest.mark.parametrize("case", list(CASES))
def test_vf_and_div_match_oracle(case, cuda_device):
    n, dim, blocks, units, H, nfeat = CASES[case]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat)
    eng = Engine(ecfg)
    B = 5
    rng = np.random.default_rng(1)
    x = rng.standard_normal((B, n * dim)).astype(np.float32) * 1.3 + 0.2
    t = rng.uniform(0, 1, B).astype(np.float32)
    feat = rng.integers(0, nfeat, (B, n)).astype(np.int32)
    p64 = O.to_torch(flat, torch.float64)
    f_ref, div_ref = O.vf_and_exact_div(p64, ocfg, torch.tensor(x, dtype=torch.float64),
                                        torch.tensor(t, dtype=torch.float64), torch.tensor(feat).long())
    f = eng.apply(tree, x, t, feat).cpu().numpy()
    f2, div = eng.apply_div(tree, x, t, feat)
    assert rel_err(f, f_ref.numpy()) < VF_TOL
    assert rel_err(f2.cpu().numpy(), f_ref.numpy()) < VF_TOL
    err = np.abs(div.cpu().numpy() - div_ref.numpy()).max() / (np.abs(div_ref.numpy()).max() + 1.0)
    assert err < DIV_TOL, (div.cpu().numpy(), div_ref.numpy())


# the networks BASELINE.json's configs[2] and configs[3] run (examples/config/qm9.yaml:6-13, aldp.yaml:11-18), at full size
FULL_SIZE = {
    "aldp": (22, 3, 3, (64, 64), 32, 22),
    "qm9pos": (19, 3, 5, (256, 256, 256, 256), 32, 1),
}


@pytest.mark.parametrize("case", list(FULL_SIZE))
def test_full_size_networks_vf_and_div(case, cuda_device):
    n, dim, blocks, units, H, nfeat = FULL_SIZE[case]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat)
    eng = Engine(ecfg)
    B = 3
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, n * dim)).astype(np.float32) * 1.3 + 0.2
    t = rng.uniform(0, 1, B).astype(np.float32)
    feat = (np.tile(np.arange(n) % nfeat, (B, 1))).astype(np.int32)      # aldp: features = arange(22), targets/data.py:146
    p64 = O.to_torch(flat, torch.float64)
    f_ref, div_ref = O.vf_and_exact_div(p64, ocfg, torch.tensor(x, dtype=torch.float64),
                                        torch.tensor(t, dtype=torch.float64), torch.tensor(feat).long())
    f, div = eng.apply_div(tree, x, t, feat)
    assert rel_err(f.cpu().numpy(), f_ref.numpy()) < VF_TOL
    err = np.abs(div.cpu().numpy() - div_ref.numpy()).max() / (np.abs(div_ref.numpy()).max() + 1.0)
    assert err < DIV_TOL, (div.cpu().numpy(), div_ref.numpy())


def test_coincident_nodes_safe_norm(cuda_device):
    """safe_norm quirk (numerical.py:7-10, SURVEY C#7): |v|^2 fed to phi_e is 1 for coincident nodes."""
    n, dim, blocks, units, H, nfeat = CASES["small_64_32"]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat)
    eng = Engine(ecfg)
    x = np.random.default_rng(2).standard_normal((2, n * dim)).astype(np.float32)
    x[:, 3:6] = x[:, 0:3]
    t = np.asarray([0.3, 0.7], np.float32)
    feat = np.zeros((2, n), np.int32)
    p64 = O.to_torch(flat, torch.float64)
    f_ref, div_ref = O.vf_and_exact_div(p64, ocfg, torch.tensor(x, dtype=torch.float64),
                                        torch.tensor(t, dtype=torch.float64), torch.tensor(feat).long())
    f, div = eng.apply_div(tree, x, t, feat)
    assert rel_err(f.cpu().numpy(), f_ref.numpy()) < VF_TOL
    assert np.abs(div.cpu().numpy() - div_ref.numpy()).max() < DIV_TOL * (np.abs(div_ref.numpy()).max() + 1)


def test_rotation_equivariance_and_batch_sizes(cuda_device):
    """egnn_test.py:31 / utils/test.py:60-76 property on the CUDA path, plus ragged batch sizes (1, > #SMs)."""
    n, dim, blocks, units, H, nfeat = CASES["dw4"]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H)
    eng = Engine(ecfg)
    rng = np.random.default_rng(3)
    for B in (1, 333):
        x = rng.standard_normal((B, n, dim)).astype(np.float32)
        t = rng.uniform(0, 1, B).astype(np.float32)
        th = 0.7
        R = np.asarray([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], np.float32)
        f = eng.apply(tree, x.reshape(B, -1), t).cpu().numpy().reshape(B, n, dim)
        fr = eng.apply(tree, (x @ R.T).reshape(B, -1), t).cpu().numpy().reshape(B, n, dim)
        assert np.abs(f @ R.T - fr).max() < 1e-5 * (np.abs(f).max() + 1)
    assert eng.apply(tree, np.zeros((0, n * dim), np.float32), np.zeros(0, np.float32)).shape == (0, n * dim)

"""GPU: a jax-style key (uint32[2]) drives the façade with the reference's own threefry draws (utils/jax_random.py):
the results equal the ones obtained by injecting that noise explicitly, and the key bookkeeping follows
jax.random.split (gradient_step.py:30, loss.py:21-24, sample_and_log_prob.py:24)."""
import numpy as np
import pytest
import torch

from ecnf_b200.cnf import (build_cnf, flow_matching_loss_fn, flow_matching_update_fn, sample_cnf,
                           sample_and_log_prob_cnf, TrainingState)
from ecnf_b200.utils import jax_random as jr
from ecnf_b200.utils.optim import Adam, warmup_cosine_decay_schedule
from helpers import CASES, make_pair

pytestmark = pytest.mark.gpu


def test_jax_keys_reproduce_injected_noise(cuda_device):
    n, dim, blocks, units, H, nfeat = CASES["dw4"]
    ocfg, flat, tree, ecfg = make_pair(n, dim, blocks, units, H, n_features=nfeat, base_scale=1.5)
    cnf = build_cnf(n, dim, 0.01, 1.5, blocks, units, H, 8, nfeat)
    B = 6
    key = jr.PRNGKey(5)
    x_data = np.random.default_rng(0).standard_normal((B, n * dim)).astype(np.float32)
    feat = torch.zeros(B, n, dtype=torch.int32)
    # cnf.sample_base(key, n) (build_cnf.py:46-61): the reference's draw for a jax key
    xb = cnf.sample_base(key, 5)
    assert np.array_equal(xb.cpu().numpy(), jr.sample_base(key, 5, n, dim, 1.5))
    xs, lps = cnf.sample_and_log_prob_base(key, ())
    assert np.array_equal(xs.cpu().numpy(), jr.sample_base(key, 1, n, dim, 1.5)[0]) and lps.dim() == 0
    # loss.py:21-24
    x0, t = jr.fm_noise(key, B, n, dim, 1.5)
    l_key, _ = flow_matching_loss_fn(cnf, tree, x_data, key, feat)
    l_inj, _ = flow_matching_loss_fn(cnf, tree, x_data, 0, feat, x0=torch.from_numpy(x0), t=torch.from_numpy(t))
    assert abs(float(l_key) - float(l_inj)) < 1e-6 * abs(float(l_inj))      # the loss is summed with atomics
    # gradient_step.py:30: key, subkey = split(state.key); the loss is drawn from subkey
    opt = Adam(warmup_cosine_decay_schedule(1e-4, 1e-4, 10, 1000, 0.0))
    packed = cnf.engine.pack(tree)
    state = TrainingState(params=packed, opt_state=opt.init(packed), key=key, ema_params=None)
    new_state, info = flow_matching_update_fn(cnf, opt.update, state, x_data, feat)
    k_next, subkey = jr.split(key)
    assert np.array_equal(jr.as_key(new_state.key), k_next)
    l_sub, _ = flow_matching_loss_fn(cnf, tree, x_data, subkey, feat)
    assert abs(float(info["loss"]) - float(l_sub)) < 1e-6 * abs(float(l_sub))
    # sample_and_log_prob.py:24 under vmap over split(key, B) (setup_training.py:47)
    keys = jr.split(key, B)
    x0s = jr.sample_base_per_key(keys, n, dim, 1.5)
    a = sample_cnf(cnf, tree, keys, feat, use_fixed_step_size=True)
    b = sample_cnf(cnf, tree, 0, feat, use_fixed_step_size=True, x0=torch.from_numpy(x0s))
    c = sample_cnf(cnf, tree, key, feat, use_fixed_step_size=True)          # one key, B trajectories: split(key, B)
    assert torch.equal(a, b) and torch.equal(a, c)
    xa, lqa = sample_and_log_prob_cnf(cnf, tree, keys, feat, use_fixed_step_size=True)
    xb, lqb = sample_and_log_prob_cnf(cnf, tree, 0, feat, use_fixed_step_size=True, x0=torch.from_numpy(x0s))
    assert torch.equal(xa, xb) and torch.equal(lqa, lqb)

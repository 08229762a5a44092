"""Mirror of ecnf/cnf/build_cnf.py: build_cnf(...) -> FlowMatchingCNF backed by the CUDA engine."""
from functools import partial
from typing import Sequence

import numpy as np
import torch

from ..engine import CnfConfig, Engine, key_to_seed, timestep_frequencies
from ..utils import jax_random as jr
from ..nets.egnn import init_flat_params
from .core import FlowMatchingCNF, optimal_transport_conditional_vf


def get_timestep_embedding(timesteps, embedding_dim: int) -> torch.Tensor:
    """ecnf/cnf/build_cnf.py:18-32 (host-side convenience; the kernels compute it on device with sinf/cosf)."""
    t = torch.as_tensor(timesteps, dtype=torch.float32)
    assert t.dim() == 1
    freqs = torch.tensor(timestep_frequencies(embedding_dim), device=t.device)
    arg = (t * 1000)[:, None] * freqs[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def build_cnf(n_frames: int, dim: int, sigma_min: float, base_scale: float, n_blocks_egnn: int,
              mlp_units: Sequence[int], n_invariant_feat_hidden: int, time_embedding_dim: int,
              n_features: int) -> FlowMatchingCNF:
    """Same signature as ecnf/cnf/build_cnf.py:34-44."""
    cfg = CnfConfig(int(n_frames), int(dim), float(sigma_min), float(base_scale), int(n_blocks_egnn),
                    tuple(int(u) for u in mlp_units), int(n_invariant_feat_hidden), int(time_embedding_dim),
                    int(n_features))
    eng = Engine(cfg)

    def init(key, positions=None, time=None, node_features=None):
        """net.init(key, x, t, features) -> {'params': ...} pytree (flax naming)."""
        flat = init_flat_params(eng, key_to_seed(key))
        return eng.unpack(torch.from_numpy(flat).to(eng.device))

    def apply(params, positions, time, node_features=None):
        return eng.apply(params, positions, time, node_features)

    def _base(key, n: int):
        # a jax-style uint32[2] key: the reference's own draw (utils/jax_random.py); an integer seed: the device Philox stream
        if jr.is_key(key):
            return torch.from_numpy(jr.sample_base(key, n, cfg.n_frames, cfg.dim, cfg.base_scale)).to(eng.device)
        return eng.base_sample(key, n)

    def sample_base(key, n: int):
        return _base(key, int(n))

    def log_prob_base(x):
        return eng.base_log_prob(x)

    def sample_and_log_prob_base(seed, sample_shape=()):
        if isinstance(sample_shape, (int, np.integer)):      # distrax accepts an int for a 1-d sample shape
            sample_shape = (int(sample_shape),)
        n = int(np.prod(sample_shape)) if len(tuple(sample_shape)) else 1
        x = _base(seed, n)
        lp = eng.base_log_prob(x)
        if len(tuple(sample_shape)) == 0:
            return x[0], lp[0]
        return x.reshape(*sample_shape, cfg.D), lp.reshape(*sample_shape)

    return FlowMatchingCNF(init=init, apply=apply, sample_base=sample_base,
                           get_x_t_and_conditional_u_t=partial(optimal_transport_conditional_vf, sigma_min=sigma_min),
                           log_prob_base=log_prob_base, sample_and_log_prob_base=sample_and_log_prob_base, engine=eng)

"""Mirror of ecnf/nets/egnn.py at the API level: EGNN(...).init / .apply with the same parameter pytree.

The reference EGNN (egnn.py:117-190) takes (positions [B, n, dim], node_features [B, n, H], global_features
[B, T]) -- already-embedded features and a ready time embedding.  The CUDA engine fuses the Embed lookup and the
sinusoidal embedding of build_cnf.FlatEgnn (build_cnf.py:65-93), so the drop-in unit is FlatEgnn; this class
exposes it under the EGNN name with integer features and scalar times.
"""
import math
from typing import Optional, Sequence

import numpy as np
import torch

from ..engine import CnfConfig, Engine, key_to_seed


def init_param_tensors(layout, seed: int, head_variance: float = 0.001) -> dict:
    """{flax path: fp32 array} for an ordered [(path, shape)] layout (SURVEY Appendix D order).  Reference init
    statistics: Dense = lecun_normal (flax default), bias 0, Embed N(0, 1/H) (flax default variance_scaling(1, fan_in,
    normal, out_axis=0)), phi_x head variance_scaling(0.001, fan_avg, uniform) (egnn.py:84), final_scaling 1
    (egnn.py:188).  Pure numpy: usable without the CUDA library (bench.py's CPU arm draws the same parameters)."""
    rng = np.random.default_rng(seed)
    out = {}
    for path, shape in layout:
        shape = tuple(shape)
        leaf = path.split("/")[-1]
        cnt = int(np.prod(shape)) if shape else 1
        if leaf == "final_scaling":
            v = np.ones(1)
        elif leaf == "bias":
            v = np.zeros(cnt)
        elif leaf == "embedding":
            v = rng.standard_normal(cnt) / math.sqrt(shape[1])
        elif path.count("/") == 3 and path.endswith("Dense_0/kernel"):
            lim = math.sqrt(3.0 * head_variance / (0.5 * (shape[0] + shape[1])))
            v = rng.uniform(-lim, lim, cnt)
        else:
            # truncated normal (+-2 sigma) with variance 1/fan_in, like jax.nn.initializers.lecun_normal
            std = math.sqrt(1.0 / shape[0]) / 0.87962566103423978
            v = rng.standard_normal(cnt)
            bad = np.abs(v) > 2
            while bad.any():
                v[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(v) > 2
            v = v * std
        out[path] = v.astype(np.float32).reshape(shape)
    return out


def init_flat_params(eng: Engine, seed: int, head_variance: float = 0.001) -> np.ndarray:
    """init_param_tensors laid out as the flat buffer of the engine (the library's aligned parameter layout)."""
    tensors = init_param_tensors([(p, s) for p, _, s in eng.layout], seed, head_variance)
    flat = np.zeros(eng.param_count, np.float32)
    for path, off, shape in eng.layout:
        cnt = int(np.prod(shape)) if shape else 1
        flat[off:off + cnt] = tensors[path].ravel()
    return flat


class EGNN:
    """EGNN(n_blocks, mlp_units, n_invariant_feat_hidden, ...) -- egnn.py:117-128 defaults kept; `stable_mlp`,
    non-SiLU activations and residual switches other than the defaults are not built (dead under every shipped
    config)."""

    def __init__(self, n_blocks: int, mlp_units: Sequence[int], n_invariant_feat_hidden: int, n_nodes: int, dim: int,
                 time_embedding_dim: int = 8, n_features: int = 1, name: Optional[str] = None,
                 stable_mlp: bool = False, residual_h: bool = True, residual_x: bool = True,
                 normalization_constant: float = 1.0, variance_scaling_init: float = 0.001):
        if stable_mlp or not residual_h or not residual_x:
            raise NotImplementedError("only the reference defaults (stable_mlp=False, residual_h/x=True) are built")
        self.variance_scaling_init = variance_scaling_init
        self.cfg = CnfConfig(n_nodes, dim, 0.0, 1.0, n_blocks, tuple(mlp_units), n_invariant_feat_hidden,
                             time_embedding_dim, n_features, normalization_constant)
        self.engine = Engine(self.cfg)

    def init(self, key, *_):
        flat = init_flat_params(self.engine, key_to_seed(key), self.variance_scaling_init)
        return self.engine.unpack(torch.from_numpy(flat).to(self.engine.device))

    def apply(self, params, positions, node_feature_ids, time):
        """positions [B, n, dim] (or flat), integer node features [B, n], time [B] -> vectors [B, n, dim]."""
        pos = torch.as_tensor(positions, dtype=torch.float32)
        B = pos.shape[0]
        out = self.engine.apply(params, pos.reshape(B, -1), time, node_feature_ids)
        return out.reshape(B, self.cfg.n_frames, self.cfg.dim)

    __call__ = apply

"""Mirror of ecnf/cnf/loss.py."""
from typing import Optional, Tuple

import torch

from ..engine import split_key
from ..utils import jax_random as jr
from .core import FlowMatchingCNF


def flow_matching_loss_and_grad_fn(cnf: FlowMatchingCNF, params, x_data, key, features=None, *, x0=None, t=None,
                                   global_offset: int = 0, loss_denominator: Optional[float] = None):
    """loss and d loss / d params (flat buffer in the engine's parameter layout).  Noise as in loss.py:21-24:
    x0 ~ base, t ~ U[0,1): injected, or -- for a jax-style uint32[2] key -- the reference's own threefry draws restated on
    the host (utils/jax_random.py; `global_offset` is ignored, the caller shards the key as the reference would), or --
    for an integer seed -- the library's Philox stream on the device."""
    eng = cnf.engine
    x_data = torch.as_tensor(x_data, dtype=torch.float32, device=eng.device)
    if x_data.dim() != 2:
        raise ValueError("x_data must be rank 2 [B, n_frames*dim] (loss.py:17)")
    B = x_data.shape[0]
    if x0 is None or t is None:
        if jr.is_key(key):
            # a jax-style key (uint32[2]): the reference's own draws, restated on the host (utils/jax_random.py)
            c = eng.cfg
            x0_np, t_np = jr.fm_noise(key, B, c.n_frames, c.dim, c.base_scale)
            x0_, t_ = torch.from_numpy(x0_np).to(eng.device), torch.from_numpy(t_np).to(eng.device)
        else:
            key1, _ = split_key(key, 2)
            x0_, t_ = eng.fm_draw_noise(key1, B, global_offset)
        x0 = x0_ if x0 is None else x0
        t = t_ if t is None else t
    loss, grad = eng.fm_loss_grad(params, x_data, x0, t, features, loss_denominator)
    return loss[0], grad


def flow_matching_loss_fn(cnf: FlowMatchingCNF, params, x_data, key, features=None, *, x0=None, t=None) -> Tuple[torch.Tensor, dict]:
    """ecnf/cnf/loss.py:10-32: returns (loss, {'loss': loss})."""
    loss, _ = flow_matching_loss_and_grad_fn(cnf, params, x_data, key, features, x0=x0, t=t)
    return loss, {"loss": loss}

"""Mirror of ecnf/cnf/loss.py."""
from typing import Optional, Tuple

import torch

from ..engine import split_key
from .core import FlowMatchingCNF


def flow_matching_loss_and_grad_fn(cnf: FlowMatchingCNF, params, x_data, key, features=None, *, x0=None, t=None,
                                   global_offset: int = 0, loss_denominator: Optional[float] = None):
    """loss and d loss / d params (flat buffer in the engine's parameter layout).  Noise as in loss.py:21-24:
    x0 ~ base, t ~ U[0,1), drawn from the library's Philox stream unless injected."""
    eng = cnf.engine
    x_data = torch.as_tensor(x_data, dtype=torch.float32, device=eng.device)
    if x_data.dim() != 2:
        raise ValueError("x_data must be rank 2 [B, n_frames*dim] (loss.py:17)")
    B = x_data.shape[0]
    if x0 is None or t is None:
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

"""Mirror of ecnf/cnf/core.py: the FlowMatchingCNF bundle and the OT conditional path."""
from typing import Any, Callable, NamedTuple, Optional, Tuple

import torch


def optimal_transport_conditional_vf(x0, x1, t, sigma_min: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """ecnf/cnf/core.py:35-39.  Host-side convenience on device tensors (elementwise); the training path fuses
    this into ecnf_fm_loss_grad and never calls it."""
    t = torch.as_tensor(t, dtype=x0.dtype, device=x0.device)
    if t.dim() == 1 and x0.dim() == 2:
        t = t[:, None]
    x_t = (1 - (1 - sigma_min) * t) * x0 + t * x1
    u_t = x1 - (1 - sigma_min) * x0
    return x_t, u_t


class FlowMatchingCNF(NamedTuple):
    """Same six callables as ecnf/cnf/core.py:42-49, plus the engine that backs them."""
    init: Callable
    apply: Callable
    sample_base: Callable
    get_x_t_and_conditional_u_t: Callable
    log_prob_base: Callable
    sample_and_log_prob_base: Callable
    engine: Any = None

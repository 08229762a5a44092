"""Mirror of ecnf/cnf/gradient_step.py: TrainingState and the flow-matching update."""
from typing import Any, NamedTuple, Optional

import torch

from ..engine import PackedParams, split_key
from ..utils import jax_random as jr
from ..utils.optim import Adam, AdamState
from .core import FlowMatchingCNF
from .loss import flow_matching_loss_and_grad_fn


class TrainingState(NamedTuple):
    """ecnf/cnf/gradient_step.py:13-17.  `ema_params=None` (or any non-parameter sentinel) switches EMA off."""
    params: Any
    opt_state: Any
    key: Any
    ema_params: Optional[Any] = None


def flow_matching_update_fn(cnf: FlowMatchingCNF, opt_update, state: TrainingState, x_data, features=None,
                            ema_beta: float = 0.999, *, x0=None, t=None, grad_allreduce=None, global_offset: int = 0,
                            loss_denominator: Optional[float] = None, donate: bool = False):
    """ecnf/cnf/gradient_step.py:20-53.  `opt_update` must be the bound `update` of ecnf_b200.utils.optim.Adam
    (the stand-in for optax.adam(...).update).  Returns (new_state, info) with info = {loss, grad_norm,
    update_norm} as 0-d device tensors.  `grad_allreduce(flat_grad, loss)` is the data-parallel hook.
    `donate=True` is jax's donate_argnums for the state: parameters, moments and EMA are updated in place and the returned
    state aliases the buffers of the one passed in (which must not be used again); the default copies them, as a jitted
    function without donation returns fresh buffers."""
    opt = getattr(opt_update, "__self__", None)
    if not isinstance(opt, Adam):
        raise TypeError("opt_update must be ecnf_b200.utils.optim.Adam(...).update")
    eng = cnf.engine
    # gradient_step.py:30: key, subkey = jax.random.split(state.key) -- exact for a jax-style key, SplitMix64 for integer seeds
    key, subkey = jr.split(state.key) if jr.is_key(state.key) else split_key(state.key, 2)
    packed = eng.pack(state.params)
    loss, grad = flow_matching_loss_and_grad_fn(cnf, packed, x_data, subkey, features, x0=x0, t=t,
                                                global_offset=global_offset, loss_denominator=loss_denominator)
    if grad_allreduce is not None:
        loss = grad_allreduce(grad, loss)
    ost: AdamState = state.opt_state
    ema_on = isinstance(state.ema_params, (PackedParams, dict))
    if donate:
        new_flat, mu, nu = packed.flat, ost.mu, ost.nu
        ema = eng.pack(state.ema_params).flat if ema_on else None
        if ema is not None and ema.data_ptr() == new_flat.data_ptr():
            ema = ema.clone()          # the caller seeded the EMA with the parameter buffer itself
    else:
        new_flat = packed.flat.clone()
        mu, nu = ost.mu.clone(), ost.nu.clone()
        ema = eng.pack(state.ema_params).flat.clone() if ema_on else None
    lr = opt.learning_rate(ost.count)
    norms = eng.adam_step(new_flat, grad, mu, nu, ost.count, lr, ema, opt.b1, opt.b2, opt.eps, ema_beta)
    info = {"loss": loss, "grad_norm": norms[0], "update_norm": norms[1]}
    new_state = TrainingState(params=PackedParams(new_flat), opt_state=AdamState(ost.count + 1, mu, nu), key=key,
                              ema_params=PackedParams(ema) if ema_on else state.ema_params)
    return new_state, info

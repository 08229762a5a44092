"""optax stand-ins used by the reference trainer (ecnf/setup_training.py:96-109): adam + warmup-cosine schedule.
The arithmetic runs in the fused ecnf_adam_step kernel; this file only keeps step counts and the schedule."""
from typing import Callable, NamedTuple, Union

import torch

from .. import lib as L


class AdamState(NamedTuple):
    count: int
    mu: torch.Tensor
    nu: torch.Tensor


def warmup_cosine_decay_schedule(init_value: float, peak_value: float, warmup_steps: int, decay_steps: int,
                                 end_value: float = 0.0) -> Callable[[int], float]:
    """optax.warmup_cosine_decay_schedule, evaluated by the library (ecnf_warmup_cosine_lr)."""
    lib = L.load()
    return lambda step: float(lib.ecnf_warmup_cosine_lr(int(step), init_value, peak_value, int(warmup_steps),
                                                        int(decay_steps), end_value))


class Adam:
    """optax.adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8).  `.init(params)` / `.update` as in optax; `.update`
    is passed to flow_matching_update_fn, which recognises it and runs the fused kernel."""

    def __init__(self, learning_rate: Union[float, Callable[[int], float]], b1=0.9, b2=0.999, eps=1e-8):
        self._lr, self.b1, self.b2, self.eps = learning_rate, b1, b2, eps

    def learning_rate(self, count: int) -> float:
        return float(self._lr(count)) if callable(self._lr) else float(self._lr)

    def init(self, params, engine=None) -> AdamState:
        from ..engine import PackedParams
        if not isinstance(params, PackedParams):
            if engine is None:
                raise TypeError("Adam.init needs engine= to pack a pytree")
            params = engine.pack(params)
        return AdamState(0, torch.zeros_like(params.flat), torch.zeros_like(params.flat))

    def update(self, grads, state, params=None):  # optax signature; the fused path never calls it
        raise L.EcnfError("Adam.update is a marker for flow_matching_update_fn; the fused kernel does the update")

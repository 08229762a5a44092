"""optax stand-in: the two tree helpers flow_matching_update_fn calls; the optimiser itself is passed in by the caller."""
from typing import Any

import torch as _t

import jax as _jax

OptState = TransformUpdateFn = Any


def apply_updates(params, updates):
    return _jax.tree_map(lambda p, u: p + u, params, updates)


def global_norm(tree):
    return _t.sqrt(sum((leaf.double() ** 2).sum() for leaf in _jax.tree_util.tree_leaves(tree))).to(_t.get_default_dtype())

"""chex stand-in: the shape assertions the reference makes are really checked; the type aliases are `object`."""
from typing import Any

import torch as _t

Array = ArrayTree = PRNGKey = Shape = Any


def assert_rank(x, rank):
    ranks = rank if isinstance(rank, (set, tuple, list)) else (rank,)
    assert x.ndim in ranks, f"rank {x.ndim}, expected {rank}"


def assert_shape(x, shape):
    assert tuple(x.shape) == tuple(shape), f"shape {tuple(x.shape)}, expected {tuple(shape)}"


def assert_equal_shape(xs):
    shapes = [tuple(x.shape) for x in xs]
    assert all(s == shapes[0] for s in shapes), shapes


def assert_axis_dimension(x, axis, size):
    assert x.shape[axis] == size, (tuple(x.shape), axis, size)


def _leaves(tree):
    import jax
    return jax.tree_util.tree_leaves(tree)


def assert_tree_shape_suffix(tree, suffix):
    for leaf in _leaves(tree):
        assert tuple(leaf.shape[len(leaf.shape) - len(suffix):]) == tuple(suffix), (tuple(leaf.shape), suffix)


def assert_tree_shape_prefix(tree, prefix):
    for leaf in _leaves(tree):
        assert tuple(leaf.shape[:len(prefix)]) == tuple(prefix), (tuple(leaf.shape), prefix)


def assert_trees_all_close(a, b, rtol=1e-6, atol=0.0):
    assert _t.allclose(a, b, rtol=rtol, atol=atol)

"""Stand-in for the part of `jax` the reference's hot-path files use, on torch CPU tensors (see ../README.md)."""
import torch as _torch

from . import numpy  # noqa: F401  (jax.numpy)
from . import nn, random, tree_util  # noqa: F401
from .tree_util import tree_map  # jax.tree_map (the reference's era)

_vmap_hooks = []      # (snapshot, restore) pairs registered by the flax stand-in: auto-name counters restart per mapped sample


def _stack(outs):
    first = outs[0]
    if isinstance(first, tuple):
        return tuple(_stack([o[k] for o in outs]) for k in range(len(first)))
    return _torch.stack(list(outs))


def vmap(fun, in_axes=0, out_axes=0):
    """Leading-axis map as a Python loop (every use in the reference maps axis 0 of all arguments)."""
    assert in_axes == 0 and out_axes == 0

    def mapped(*args):
        n = args[0].shape[0]
        snaps = [snap() for snap, _ in _vmap_hooks]
        outs = []
        for i in range(n):
            for (_, restore), s in zip(_vmap_hooks, snaps):
                restore(s)
            outs.append(fun(*[a[i] for a in args]))
        return _stack(outs)
    return mapped


def jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def vjp(fun, x):
    """(fun(x), vjp_fn) with vjp_fn(ct) -> (ct^T dfun/dx,) by reverse mode, like jax.vjp for one argument."""
    xg = x.detach().clone().requires_grad_(True)
    with _torch.enable_grad():
        y = fun(xg)

    def vjp_fn(ct):
        (g,) = _torch.autograd.grad(y, xg, ct.to(y.dtype), retain_graph=True)
        return (g,)
    return y.detach(), vjp_fn


def _leaf_copy(tree):
    return tree_map(lambda a: a.detach().clone().requires_grad_(True), tree)


def value_and_grad(fun, argnums=0, has_aux=False):
    def wrapped(*args, **kwargs):
        args = list(args)
        args[argnums] = _leaf_copy(args[argnums])
        with _torch.enable_grad():
            out = fun(*args, **kwargs)
        val, aux = out if has_aux else (out, None)
        leaves, treedef = tree_util.tree_flatten(args[argnums])
        grads = _torch.autograd.grad(val, leaves, allow_unused=True)
        grads = [(_torch.zeros_like(l) if g is None else g) for g, l in zip(grads, leaves)]
        gtree = tree_util.tree_unflatten(treedef, grads)
        val = val.detach()
        aux = tree_map(lambda a: a.detach() if isinstance(a, _torch.Tensor) else a, aux) if aux is not None else None
        return ((val, aux), gtree) if has_aux else (val, gtree)
    return wrapped


def grad(fun, argnums=0, has_aux=False):
    vg = value_and_grad(fun, argnums=argnums, has_aux=has_aux)

    def wrapped(*args, **kwargs):
        out, g = vg(*args, **kwargs)
        return (g, out[1]) if has_aux else g
    return wrapped


def jacrev(fun):
    def wrapped(x, *rest):
        y, vjp_fn = vjp(lambda z: fun(z, *rest), x)
        return _torch.stack([vjp_fn(row)[0] for row in _torch.eye(y.shape[0], dtype=y.dtype)])
    return wrapped


jacfwd = jacrev      # same matrix; the reference never relies on the mode

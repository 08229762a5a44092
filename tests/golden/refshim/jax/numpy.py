"""jax.numpy stand-in: the functions the reference's hot path calls, one-to-one onto torch (float dtype = torch's default:
float32 like jax with x64 off, or float64 for the ground-truth run)."""
import math as _math

import torch as _t

pi = _math.pi
ndarray = _t.Tensor
float32, float64, int32, int64 = _t.float32, _t.float64, _t.int32, _t.int64


def _a(x):
    return x if isinstance(x, _t.Tensor) else _t.as_tensor(x)


def _f(x):        # python numbers are weakly typed in jax: floats take the default float dtype
    if isinstance(x, _t.Tensor):
        return x
    return _t.tensor(float(x), dtype=_t.get_default_dtype())


def array(x, dtype=None):
    if dtype is int:
        dtype = _t.long
    if isinstance(x, _t.Tensor):
        return x.to(dtype) if dtype is not None else x
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], (list, tuple, _t.Tensor)):
        rows = [array(r) for r in x]
        out = _t.stack(rows)
        return out.to(dtype) if dtype is not None else out
    if isinstance(x, (list, tuple)) and len(x) and any(isinstance(e, _t.Tensor) for e in x):
        out = _t.stack([_f(e) for e in x])
        return out.to(dtype) if dtype is not None else out
    out = _t.tensor(x)
    if out.dtype == _t.float64 and dtype is None:
        out = out.to(_t.get_default_dtype())
    return out.to(dtype) if dtype is not None else out


asarray = array


def concatenate(xs, axis=0): return _t.cat(list(xs), dim=axis)
def stack(xs, axis=0): return _t.stack(list(xs), dim=axis)
def reshape(x, shape): return _t.reshape(x, tuple(shape))
def squeeze(x, axis=None): return _t.squeeze(x) if axis is None else _t.squeeze(x, dim=axis)
def repeat(x, repeats, axis=None): return _t.repeat_interleave(x, repeats, dim=axis)
def sum(x, axis=None, keepdims=False): return _t.sum(x) if axis is None else _t.sum(x, dim=axis, keepdim=keepdims)
def mean(x, axis=None, keepdims=False): return _t.mean(x) if axis is None else _t.mean(x, dim=axis, keepdim=keepdims)
def where(c, a, b): return _t.where(c if c.dtype == _t.bool else c != 0, a, b)      # jnp accepts a non-boolean condition
def sqrt(x): return _t.sqrt(_f(x))
def log(x): return _t.log(_f(x))
def exp(x): return _t.exp(_f(x))
def sin(x): return _t.sin(_f(x))
def cos(x): return _t.cos(_f(x))
def abs(x): return _t.abs(_a(x))
def arange(*args): return _t.arange(*args)
def zeros(shape, dtype=None): return _t.zeros(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=dtype)
def ones(shape, dtype=None): return _t.ones(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=dtype)
def zeros_like(x): return _t.zeros_like(x)
def ones_like(x): return _t.ones_like(x)
def eye(n): return _t.eye(n)
def trace(x): return _t.trace(x)
def dot(a, b): return _t.dot(a, b)
def matmul(a, b): return _t.matmul(a, b)

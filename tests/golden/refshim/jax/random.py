"""jax.random stand-in: the repo's numpy restatement of threefry2x32 / uniform / normal (ecnf_b200/utils/jax_random.py,
pinned by the Random123 known-answer vectors and values printed in the JAX documentation), returned as torch tensors."""
import importlib.util as _ilu
import os as _os

import numpy as _np
import torch as _t

_p = _os.path.join(_os.path.dirname(__file__), "..", "..", "..", "..", "ecnf_b200", "utils", "jax_random.py")
_spec = _ilu.spec_from_file_location("_ecnf_jax_random", _os.path.abspath(_p))
_jr = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_jr)


class _Key(_t.Tensor):
    pass


def _key(k):
    return _np.asarray(k.numpy() if isinstance(k, _t.Tensor) else k).astype(_np.uint32)


def _wrap_key(a):
    return _t.from_numpy(_np.asarray(a).astype(_np.int64))      # uint32 values carried in int64 tensors


def PRNGKey(seed): return _wrap_key(_jr.PRNGKey(seed))
def split(key, num=2): return _wrap_key(_jr.split(_key(key), num))


def _shape(shape):
    return tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, _t.Size)) else (shape,)))


def normal(key, shape=(), dtype=None):
    return _t.from_numpy(_np.asarray(_jr.normal(_key(key), _shape(shape)))).to(dtype or _t.get_default_dtype())


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return _t.from_numpy(_np.asarray(_jr.uniform(_key(key), _shape(shape), minval, maxval))).to(dtype or _t.get_default_dtype())

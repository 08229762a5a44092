"""Reader for the reference's training checkpoints (`state_%08i.pkl`, written by `ecnf/utils/loop.py:144-153` as
`pickle.dump(TrainingState)` with `TrainingState(params, opt_state, key, ema_params)` from
`ecnf/cnf/gradient_step.py:13-17`) -- SURVEY.md 8(f) rank 4.

The pickles reference classes of jax, flax, optax and ecnf, none of which is importable here, so a *restricted*
unpickler rebuilds them from stubs: jax arrays become numpy arrays (jax pickles an array as
`_reconstruct_array(numpy_reduce_fun, numpy_reduce_args, ndarray_state, aval_state)`), flax `FrozenDict`s become dicts,
optax optimiser states become plain tuples that remember their class name.  Anything outside that white list raises
`pickle.UnpicklingError`: a checkpoint is foreign data and `pickle.load` on it would execute arbitrary code.

The result feeds `Engine.pack(ckpt.params)` (same flax parameter pytree).

STATUS: the jax / flax / optax pickle layouts are restated from their published sources (jax >= 0.4 `jax/_src/array.py`,
jax < 0.4 `jax/_src/device_array.py`, flax `FrozenDict.__reduce__`); no reference checkpoint is available in this image
(they live in a private wandb project), so the reader is exercised with synthetic pickles only
(tests/test_checkpoint_reader.py).
"""
from __future__ import annotations

import io
import pickle
from typing import Any, NamedTuple, Optional

import numpy as np


class ReferenceCheckpoint(NamedTuple):
    params: Any                 # {'params': {...}} nested dicts of numpy arrays
    opt_state: Any              # optax state as (class-tagged) tuples of numpy arrays
    key: Optional[np.ndarray]   # raw uint32[2] threefry key
    ema_params: Any             # same tree as params, or None (the reference stores jnp.array(None) as "no EMA")


class _TaggedTuple(tuple):
    """Stand-in for a NamedTuple class that cannot be imported (optax states): keeps the values and the class name."""
    _tag = "?"

    def __new__(cls, *args, **kwargs):
        return tuple.__new__(cls, tuple(args) + tuple(kwargs.values()))

    def __repr__(self):
        return f"{self._tag}{tuple.__repr__(self)}"


class _FrozenDict(dict):
    """flax.core.frozen_dict.FrozenDict -> dict (its __reduce__ passes the unfrozen dict; older versions pickle `_dict`)."""

    def __init__(self, *args, **kwargs):
        args = tuple(a for a in args if a is not None)
        super().__init__(*args, **kwargs)

    def __setstate__(self, state):
        if isinstance(state, dict) and "_dict" in state:
            self.update(state["_dict"])


def _reconstruct_array(fun, args, arr_state, aval_state=None):
    """jax._src.array._reconstruct_array / jax._src.device_array.reconstruct_device_array, without the device_put."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


class _TrainingState(_TaggedTuple):
    _tag = "TrainingState"


_ARRAY_RECONSTRUCTORS = {
    ("jax._src.array", "_reconstruct_array"),
    ("jax._src.device_array", "reconstruct_device_array"),
    ("jax.interpreters.xla", "reconstruct_device_array"),
    ("jaxlib.xla_extension", "_reconstruct_array"),
}
_NUMPY_OK = {"_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer"}
_BUILTINS_OK = {"dict", "list", "tuple", "set", "frozenset", "int", "float", "complex", "bool", "bytes", "bytearray", "str",
                "slice", "range", "object"}


class _RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if (module, name) in _ARRAY_RECONSTRUCTORS:
            return _reconstruct_array
        if module.startswith("numpy") and name in _NUMPY_OK:
            if name in ("ndarray", "dtype"):
                return getattr(np, name)
            import importlib
            for mod in (module, "numpy.core.multiarray", "numpy._core.multiarray", "numpy.core.numeric", "numpy._core.numeric"):
                try:
                    return getattr(importlib.import_module(mod), name)
                except (ImportError, AttributeError):
                    continue
        if module == "builtins" and name in _BUILTINS_OK:
            return getattr(__import__("builtins"), name)
        if module == "collections" and name == "OrderedDict":
            import collections
            return collections.OrderedDict
        if module == "copyreg" and name in ("_reconstructor", "__newobj__"):
            import copyreg
            return getattr(copyreg, name)
        if module.startswith("flax.") and name == "FrozenDict":
            return _FrozenDict
        if module.startswith("ecnf.") and name == "TrainingState":
            return _TrainingState
        if module.startswith("optax.") or module.startswith("ecnf."):
            return type(name, (_TaggedTuple,), {"_tag": f"{module}.{name}"})
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a checkpoint (not on the white list)")


def _to_numpy_tree(x):
    if isinstance(x, dict):
        return {k: _to_numpy_tree(v) for k, v in x.items()}
    if isinstance(x, _TaggedTuple):
        return x
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy_tree(v) for v in x)
    return x


def _is_none_sentinel(x) -> bool:
    # gradient_step.py:46-50 / setup_training.py:137: "no EMA" is stored as jnp.array(None), a 0-d NaN (or object) array
    if x is None:
        return True
    if isinstance(x, np.ndarray) and x.ndim == 0:
        return x.dtype == object or (np.issubdtype(x.dtype, np.floating) and bool(np.isnan(x)))
    return False


def loads_reference_checkpoint(data: bytes) -> ReferenceCheckpoint:
    state = _RestrictedUnpickler(io.BytesIO(data)).load()
    if not isinstance(state, tuple) or len(state) < 3:
        raise ValueError(f"not a TrainingState pickle: got {type(state).__name__} with {len(state) if hasattr(state, '__len__') else '?'} fields")
    params, opt_state, key = state[0], state[1], state[2]
    ema = state[3] if len(state) > 3 else None
    params = _to_numpy_tree(params)
    if not (isinstance(params, dict) and "params" in params):
        raise ValueError("TrainingState.params is not a flax {'params': ...} pytree")
    return ReferenceCheckpoint(params=params, opt_state=opt_state, key=None if key is None else np.asarray(key),
                               ema_params=None if _is_none_sentinel(ema) else _to_numpy_tree(ema))


def load_reference_checkpoint(path) -> ReferenceCheckpoint:
    """Read `state_%08i.pkl`; `Engine.pack(ckpt.params)` (or `ckpt.ema_params`) gives the device parameter buffer."""
    with open(path, "rb") as f:
        return loads_reference_checkpoint(f.read())

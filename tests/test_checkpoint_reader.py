"""Host logic: the restricted reader for the reference's `state_%08i.pkl` checkpoints (utils/loop.py:144-153).
No reference checkpoint exists in this image, so the pickles are synthesised with stand-in modules that reproduce the
pickle layouts of jax arrays (`_reconstruct_array`), flax `FrozenDict`, optax states and `TrainingState`."""
import collections
import os
import pickle
import sys
import types

import numpy as np
import pytest

from ecnf_b200.utils.checkpoint import loads_reference_checkpoint
from oracle import ecnf_oracle as O


def _fake_modules():
    mods = {}

    def mod(name):
        m = types.ModuleType(name)
        mods[name] = m
        return m

    for n in ("jax", "jax._src", "flax", "flax.core", "optax", "optax._src", "ecnf", "ecnf.cnf"):
        mod(n)
    ja = mod("jax._src.array")

    def _reconstruct_array(fun, args, arr_state, aval_state):
        raise AssertionError("the reader must not call into jax")
    _reconstruct_array.__module__ = "jax._src.array"
    _reconstruct_array.__qualname__ = "_reconstruct_array"
    ja._reconstruct_array = _reconstruct_array

    class ArrayImpl:
        def __init__(self, v):
            self._value = np.asarray(v)

        def __reduce__(self):
            fun, args, arr_state = self._value.__reduce__()
            return (_reconstruct_array, (fun, args, arr_state, {"weak_type": False, "named_shape": {}}))
    ja.ArrayImpl = ArrayImpl

    fd = mod("flax.core.frozen_dict")

    class FrozenDict(dict):
        def __reduce__(self):
            return (FrozenDict, (dict(self),))
    FrozenDict.__module__ = "flax.core.frozen_dict"
    FrozenDict.__qualname__ = "FrozenDict"
    fd.FrozenDict = FrozenDict

    ot = mod("optax._src.transform")
    ot.ScaleByAdamState = collections.namedtuple("ScaleByAdamState", "count mu nu")
    ot.ScaleByAdamState.__module__ = "optax._src.transform"
    ot.ScaleByScheduleState = collections.namedtuple("ScaleByScheduleState", "count")
    ot.ScaleByScheduleState.__module__ = "optax._src.transform"
    gs = mod("ecnf.cnf.gradient_step")
    gs.TrainingState = collections.namedtuple("TrainingState", "params opt_state key ema_params")
    gs.TrainingState.__module__ = "ecnf.cnf.gradient_step"
    return mods


def _dump_state(with_ema):
    cfg = O.CnfConfig(n_frames=4, dim=2, n_blocks_egnn=2, mlp_units=(64, 64), n_invariant_feat_hidden=32)
    tree = O.flat_to_nested(O.init_params(cfg, seed=3))
    mods = _fake_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        A = mods["jax._src.array"].ArrayImpl
        FD = mods["flax.core.frozen_dict"].FrozenDict

        def wrap(t):
            return FD({k: wrap(v) for k, v in t.items()}) if isinstance(t, dict) else A(t)
        params = wrap(tree)
        ot, gs = mods["optax._src.transform"], mods["ecnf.cnf.gradient_step"]
        opt_state = (ot.ScaleByAdamState(A(np.int32(7)), wrap(tree), wrap(tree)), ot.ScaleByScheduleState(A(np.int32(7))))
        ema = wrap(tree) if with_ema else A(np.float32("nan"))       # jnp.array(None) is a 0-d NaN
        state = gs.TrainingState(params, opt_state, A(np.asarray([0, 42], np.uint32)), ema)
        data = pickle.dumps(state)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return tree, data


def _assert_tree_equal(a, b):
    assert set(a) == set(b)
    for k in a:
        if isinstance(a[k], dict):
            _assert_tree_equal(a[k], b[k])
        else:
            assert isinstance(b[k], np.ndarray) and np.array_equal(np.asarray(a[k]), b[k]), k


@pytest.mark.parametrize("with_ema", [True, False])
def test_reads_training_state_without_jax(with_ema):
    assert "jax" not in sys.modules or not hasattr(sys.modules["jax"], "numpy")
    tree, data = _dump_state(with_ema)
    ck = loads_reference_checkpoint(data)
    _assert_tree_equal(tree, ck.params)
    assert ck.key.dtype == np.uint32 and ck.key.tolist() == [0, 42]
    if with_ema:
        _assert_tree_equal(tree, ck.ema_params)
    else:
        assert ck.ema_params is None
    adam = ck.opt_state[0]
    assert "ScaleByAdamState" in repr(type(adam)._tag) and int(adam[0]) == 7
    _assert_tree_equal(tree, adam[1])


def test_parameters_pack_into_the_library_layout():
    """The unpickled pytree is the flax layout the engine packs."""
    from ecnf_b200.engine import CnfConfig, Engine
    tree, data = _dump_state(False)
    ck = loads_reference_checkpoint(data)
    eng = Engine(CnfConfig(4, 2, 0.01, 1.0, 2, (64, 64), 32, 8, 1))
    paths = {p for p, _, _ in eng.layout}
    flat_ref = O.nested_to_flat(tree) if hasattr(O, "nested_to_flat") else None
    def walk(t, pre=""):
        for k, v in t.items():
            if isinstance(v, dict):
                yield from walk(v, pre + k + "/")
            else:
                yield pre + k
    got = {p[len("params/"):] for p in walk(ck.params)}
    assert got == paths, (sorted(got ^ paths)[:5])


def test_refuses_foreign_classes():
    class Evil:
        def __reduce__(self):
            return (os.system, ("echo pwned",))
    with pytest.raises(pickle.UnpicklingError):
        loads_reference_checkpoint(pickle.dumps(Evil()))
    with pytest.raises(pickle.UnpicklingError):
        loads_reference_checkpoint(pickle.dumps(collections.Counter("abc")))


def test_module_names_are_compared_never_imported():
    """A module string that merely starts with 'numpy' / 'optax' must not be imported or trusted."""
    import pickletools  # noqa: F401  (documentation of the opcodes used below)
    for mod in ("numpyro.evil", "numpy_foo", "optaxx.mod", "flaxy.core"):
        payload = b"\x80\x02c" + mod.encode() + b"\n_reconstruct\n."
        with pytest.raises(pickle.UnpicklingError):
            loads_reference_checkpoint(payload)
        assert mod not in sys.modules
    # memory-exhaustion constructors are not on the white list
    for name in ("bytearray", "range", "bytes"):
        with pytest.raises(pickle.UnpicklingError):
            loads_reference_checkpoint(b"\x80\x02cbuiltins\n" + name.encode() + b"\n.")


def test_optax_state_adapter_finds_adam_moments():
    from ecnf_b200.utils.checkpoint import find_adam_state
    tree, data = _dump_state(True)
    ck = loads_reference_checkpoint(data)
    count, mu, nu = find_adam_state(ck.opt_state)
    assert int(count) == 7
    _assert_tree_equal(tree, mu)
    _assert_tree_equal(tree, nu)
    assert find_adam_state(("nothing", 3)) is None

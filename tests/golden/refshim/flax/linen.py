"""flax.linen stand-in: a dataclass-like Module with `setup`, `@compact`, `param`, `init` / `apply`, flax's sub-module naming
rule (explicit `name=`, the attribute name for modules assigned in `setup`, `ClassName_<k>` in construction order inside a
compact method), and `Dense` / `Embed`.  Parameters are torch tensors in a nested dict {"params": {...}}."""
import collections
import math

import numpy as np
import torch

import jax as _jax

_stack = []                       # modules whose method is executing (innermost last)
_run = {"init": False, "root": None, "rng": None}


def compact(fn):
    fn._compact = True
    return fn


# ---- initialisers: (key, shape, dtype=None) -> tensor; the VALUES are irrelevant for the fixtures (they are overwritten
# by seeded parameters), the SHAPES are what the reference's init pins --------------------------------------------------
def _fans(shape):
    if len(shape) < 2:
        return (shape[0] if shape else 1), (shape[0] if shape else 1)
    return shape[-2], shape[-1]


class _Initializers:
    @staticmethod
    def variance_scaling(scale, mode, distribution):
        def init(key, shape, dtype=None):
            fan_in, fan_out = _fans(tuple(shape))
            denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": 0.5 * (fan_in + fan_out)}[mode]
            var = scale / max(denom, 1)
            g = _run["rng"]
            if distribution == "uniform":
                a = math.sqrt(3.0 * var)
                x = g.uniform(-a, a, size=tuple(shape))
            else:
                x = g.standard_normal(size=tuple(shape)) * math.sqrt(var)
            return torch.as_tensor(x, dtype=torch.get_default_dtype())
        return init

    @staticmethod
    def lecun_normal():
        return _Initializers.variance_scaling(1.0, "fan_in", "truncated_normal")

    @staticmethod
    def zeros_init():
        return lambda key, shape, dtype=None: torch.zeros(tuple(shape))

    @staticmethod
    def ones_init():
        return lambda key, shape, dtype=None: torch.ones(tuple(shape))

    @staticmethod
    def normal(stddev=1e-2):
        return lambda key, shape, dtype=None: torch.as_tensor(_run["rng"].standard_normal(size=tuple(shape)) * stddev,
                                                              dtype=torch.get_default_dtype())


initializers = _Initializers()


class _Linear:
    default_kernel_init = staticmethod(_Initializers.lecun_normal())


linear = _Linear()


def _wrap_call(fn):
    def wrapped(self, *args, **kwargs):
        self._ensure_setup()
        saved = self._counters
        self._counters = collections.defaultdict(int)
        _stack.append(self)
        try:
            return fn(self, *args, **kwargs)
        finally:
            _stack.pop()
            self._counters = saved
    wrapped.__wrapped__ = fn
    return wrapped


class Module:
    _fields = ("name",)

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = []
        for klass in reversed(cls.__mro__):
            if klass in (object, Module):
                continue
            for f in klass.__dict__.get("__annotations__", {}):
                if f not in fields and f != "parent":
                    fields.append(f)
        if "name" not in fields:
            fields.append("name")                      # flax: `parent` / `name` are the trailing keyword fields
        cls._fields = tuple(fields)
        if "__call__" in cls.__dict__:
            cls.__call__ = _wrap_call(cls.__dict__["__call__"])

    def __init__(self, *args, **kwargs):
        object.__setattr__(self, "_in_setup", False)
        object.__setattr__(self, "_setup_done", False)
        object.__setattr__(self, "_counters", collections.defaultdict(int))
        object.__setattr__(self, "_parent", _stack[-1] if _stack else None)
        names = list(self._fields)
        if len(args) > len(names):
            raise TypeError(f"{type(self).__name__}: too many positional arguments")
        given = dict(zip(names, args))
        for k, v in kwargs.items():
            if k not in names:
                raise TypeError(f"{type(self).__name__}: unexpected field {k!r}")
            if k in given:
                raise TypeError(f"{type(self).__name__}: field {k!r} given twice")
            given[k] = v
        for f in names:
            if f in given:
                object.__setattr__(self, f, given[f])
            elif f == "name":
                object.__setattr__(self, f, None)
            elif hasattr(type(self), f):
                object.__setattr__(self, f, getattr(type(self), f))
            else:
                raise TypeError(f"{type(self).__name__}: missing field {f!r}")
        parent = self._parent
        if parent is not None and self.name is None and not parent._in_setup:
            k = parent._counters[type(self).__name__]          # compact: ClassName_<k> in construction order
            parent._counters[type(self).__name__] = k + 1
            object.__setattr__(self, "name", f"{type(self).__name__}_{k}")

    def __setattr__(self, key, value):
        if isinstance(value, Module) and self._in_setup and value.name is None:
            object.__setattr__(value, "name", key)             # setup: the attribute name
            object.__setattr__(value, "_parent", self)
        object.__setattr__(self, key, value)

    def setup(self):
        pass

    def _ensure_setup(self):
        if self._setup_done:
            return
        object.__setattr__(self, "_setup_done", True)
        object.__setattr__(self, "_in_setup", True)
        _stack.append(self)
        try:
            self.setup()
        finally:
            _stack.pop()
            object.__setattr__(self, "_in_setup", False)

    def _scope(self):
        if self._parent is None:
            return _run["root"]
        parent_scope = self._parent._scope()
        if self.name not in parent_scope:
            if not _run["init"]:
                raise KeyError(f"no parameters for sub-module {self.name!r} ({type(self).__name__})")
            parent_scope[self.name] = {}
        return parent_scope[self.name]

    def param(self, name, init_fn, *init_args):
        scope = self._scope()
        if name not in scope:
            if not _run["init"]:
                raise KeyError(f"parameter {name!r} of {type(self).__name__} {self.name!r} is missing")
            scope[name] = init_fn(None, *init_args)
        return scope[name]

    # ---- top-level entry points (root module only) ----
    def init(self, rngs, *args, **kwargs):
        root = {}
        saved = dict(_run)
        _run.update(init=True, root=root, rng=np.random.default_rng(0))
        try:
            self(*args, **kwargs)
        finally:
            _run.update(saved)
        return {"params": root}

    def apply(self, variables, *args, **kwargs):
        saved = dict(_run)
        _run.update(init=False, root=variables["params"], rng=None)
        try:
            return self(*args, **kwargs)
        finally:
            _run.update(saved)


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = linear.default_kernel_init
    bias_init: object = _Initializers.zeros_init()

    @compact
    def __call__(self, x):
        kernel = self.param("kernel", self.kernel_init, (x.shape[-1], self.features))
        y = torch.matmul(x, kernel)
        if self.use_bias:
            y = y + self.param("bias", self.bias_init, (self.features,))
        return y


class Embed(Module):
    num_embeddings: int
    features: int
    embedding_init: object = _Initializers.normal(1.0)

    @compact
    def __call__(self, ids):
        table = self.param("embedding", self.embedding_init, (self.num_embeddings, self.features))
        return table[ids.long()]


# jax.vmap over a method that builds sub-modules: every mapped sample sees the same auto-name counters (one set of parameters)
def _snap():
    return dict(_stack[-1]._counters) if _stack else None


def _restore(s):
    if _stack and s is not None:
        object.__setattr__(_stack[-1], "_counters", collections.defaultdict(int, s))


_jax._vmap_hooks.append((_snap, _restore))

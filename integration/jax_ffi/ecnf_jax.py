"""JAX binding of libecnf_b200.so through `jax.ffi` (XLA typed FFI): the host side BASELINE.json's north_star asks for.

NOT IMPORTABLE IN THIS REPOSITORY'S IMAGE -- JAX is not installed here, so this module and
`ecnf_jax_ffi.cc` are the reference-side binding a maintainer adds, written against the published jax.ffi API
(`jax.ffi.register_ffi_target`, `jax.ffi.pycapsule`, `jax.ffi.ffi_call`) and exercised by no test here.  The product
path that IS built and tested in this repository is the ctypes host layer (`ecnf_b200/`), which calls the same C entry
points with the same arguments.

What stays on the JAX side (so that results are key-for-key those of the reference): every random draw
(`jax.random.split / normal / uniform`, exactly where the reference draws: sample_and_log_prob.py:24,55,130,
loss.py:21-24, gradient_step.py:30), the parameter pytree (ravelled into the library's flat layout by `pack`), and
the ESS arithmetic on the five sufficient statistics.

    import ecnf_jax as E
    cnf = E.build_cnf(13, 3, 0.01, 1.0, 3, (128, 128, 128), 64, 8, 1)          # same nine arguments as the reference
    x1, log_q = E.sample_and_log_prob_cnf(cnf, params, keys, features)          # keys [B, 2], features [B, n]
    loss, grads = E.flow_matching_loss_and_grad(cnf, params, x_data, key, features)
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np
import jax
import jax.numpy as jnp

ROOT = Path(__file__).resolve().parents[2]
# the ctypes table of the C-ABI (ecnf_b200/lib.py) loaded on its own: the package's torch-based host layer is not needed
_spec = importlib.util.spec_from_file_location("ecnf_b200_lib", ROOT / "ecnf_b200" / "lib.py")
L = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(L)

_HANDLERS = {
    "ecnf_vf_forward": "EcnfVfForward", "ecnf_vf_forward_div": "EcnfVfForwardDiv",
    "ecnf_vf_forward_hutchinson": "EcnfVfForwardHutchinson", "ecnf_solve": "EcnfSolve",
    "ecnf_solve_hutchinson": "EcnfSolveHutchinson", "ecnf_base_sample_from_noise": "EcnfBaseSampleFromNoise",
    "ecnf_base_log_prob": "EcnfBaseLogProb", "ecnf_fm_loss_grad": "EcnfFmLossGrad", "ecnf_adam_step": "EcnfAdamStep",
    "ecnf_ess_stats": "EcnfEssStats", "ecnf_target_log_prob": "EcnfTargetLogProb",
}
_registered = False


def register(shim_path: os.PathLike | None = None) -> None:
    """dlopen libecnf_b200.so (RTLD_GLOBAL, so that the shim resolves the C-ABI) and the shim, and register one FFI target
    per handler."""
    global _registered
    if _registered:
        return
    C.CDLL(os.fspath(L.LIB_PATH), mode=C.RTLD_GLOBAL)
    shim = C.CDLL(os.fspath(shim_path or Path(__file__).with_name("libecnf_jax_ffi.so")))
    for target, symbol in _HANDLERS.items():
        jax.ffi.register_ffi_target(target, jax.ffi.pycapsule(getattr(shim, symbol)), platform="CUDA")
    _registered = True


@dataclass(frozen=True)
class Cnf:
    """The template handle (hyper-parameters, engine choice, training chunk size) + the parameter layout."""
    handle: int
    n_frames: int
    dim: int
    param_count: int
    layout: tuple          # ((flax path, float offset, shape), ...) in ecnf_model_param_layout order

    @property
    def D(self) -> int:
        return self.n_frames * self.dim


def build_cnf(n_frames, dim, sigma_min, base_scale, n_blocks_egnn, mlp_units, n_invariant_feat_hidden,
              time_embedding_dim, n_features) -> Cnf:
    """ecnf/cnf/build_cnf.py:34-102 (same argument order)."""
    register()
    lib = L.load()
    cfg = L.Config()
    cfg.n_frames, cfg.dim, cfg.n_blocks, cfg.n_layers = n_frames, dim, n_blocks_egnn, len(mlp_units)
    cfg.mlp_units, cfg.n_hidden, cfg.time_dim, cfg.n_features = mlp_units[0], n_invariant_feat_hidden, time_embedding_dim, n_features
    cfg.sigma_min, cfg.base_scale, cfg.normalization_constant = sigma_min, base_scale, 1.0
    half = time_embedding_dim // 2                      # build_cnf.py:25-27, computed in float32 like the reference
    freqs = np.exp(np.arange(half, dtype=np.float32) * np.float32(-np.log(10000.0) / (half - 1)))
    for k in range(half):
        cfg.freqs[k] = float(freqs[k])
    h = C.c_void_p()
    L.check(lib.ecnf_model_create(C.byref(cfg), None, C.byref(h)), "ecnf_model_create")
    layout = []
    name, off, rows, cols = C.create_string_buffer(256), C.c_int64(), C.c_int64(), C.c_int64()
    for i in range(lib.ecnf_model_num_tensors(h)):
        L.check(lib.ecnf_model_param_layout(h, i, name, 256, C.byref(off), C.byref(rows), C.byref(cols)), "param_layout")
        shape = () if rows.value == 0 else ((rows.value,) if cols.value == 0 else (rows.value, cols.value))
        layout.append((name.value.decode(), off.value, shape))
    return Cnf(h.value, n_frames, dim, int(lib.ecnf_model_param_count(h)), tuple(layout))


def pack(cnf: Cnf, params) -> jax.Array:
    """flax variable dict {"params": {...}} -> the library's flat fp32 buffer (16-byte aligned tensors, zero padding)."""
    flat = jnp.zeros((cnf.param_count,), jnp.float32)
    for path, off, shape in cnf.layout:
        node = params["params"]
        for part in path.split("/"):
            node = node[part]
        flat = jax.lax.dynamic_update_slice(flat, jnp.ravel(node).astype(jnp.float32), (off,))
    return flat


def unpack(cnf: Cnf, flat):
    """the inverse of `pack` (for the flat gradient)."""
    tree: dict = {}
    for path, off, shape in cnf.layout:
        parts = path.split("/")
        node = tree
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        size = int(np.prod(shape)) if shape else 1
        node[parts[-1]] = jnp.reshape(flat[off:off + size], shape)
    return {"params": tree}


def _ws(nbytes: int) -> jax.Array:
    return jnp.zeros((int(nbytes),), jnp.uint8)         # XLA owns the scratch; the library never allocates


def _ctrl(use_fixed_step_size, rtol, atol, step_size):
    return dict(fixed=np.int32(bool(use_fixed_step_size)), step_size=np.float32(step_size), rtol=np.float32(rtol),
                atol=np.float32(atol), dtmin=np.float32(1e-5), max_steps=np.int32(4096), err_scale=np.float32(1.0))


def _solve(cnf: Cnf, mode: int, flat, x_init, features, ctrl, eps=None):
    B = x_init.shape[0]
    ws = _ws(L.load().ecnf_solve_workspace_bytes(cnf.handle, mode, B))
    out = (jax.ShapeDtypeStruct((B, cnf.D), jnp.float32), jax.ShapeDtypeStruct((B, 3), jnp.float32),
           jax.ShapeDtypeStruct((B, 4), jnp.int32))
    attrs = dict(model=np.int64(cnf.handle), mode=np.int32(mode), **ctrl)
    feat = features.astype(jnp.int32)
    if eps is None:
        return jax.ffi.ffi_call("ecnf_solve", out)(flat, x_init, feat, ws, **attrs)
    return jax.ffi.ffi_call("ecnf_solve_hutchinson", out)(flat, x_init, feat, eps, ws, **attrs)


def apply(cnf: Cnf, params, x, t, features):
    """cnf.apply(params, x, t, features)  (build_cnf.py:68-93)."""
    B = x.shape[0]
    ws = _ws(L.load().ecnf_solve_workspace_bytes(cnf.handle, L.MODE_VF, B))
    return jax.ffi.ffi_call("ecnf_vf_forward", jax.ShapeDtypeStruct((B, cnf.D), jnp.float32))(
        pack(cnf, params), x, t, features.astype(jnp.int32), ws, model=np.int64(cnf.handle))


def _from_noise(cnf: Cnf, eps):
    """x0 = base_scale * remove_mean(eps) for a batch of raw normal draws [B, D] (zero_com_base.py:44-47, 88-93)."""
    return jax.ffi.ffi_call("ecnf_base_sample_from_noise", jax.ShapeDtypeStruct(eps.shape, jnp.float32))(
        eps, model=np.int64(cnf.handle))


def sample_base(cnf: Cnf, key, n: int):
    """cnf.sample_base(key, n)  (build_cnf.py:46-48): the draw stays jax.random's; returns (x0 [n, D], raw noise)."""
    eps = jax.random.normal(key, (n, cnf.n_frames, cnf.dim), jnp.float32).reshape(n, cnf.D)
    return _from_noise(cnf, eps), eps


def _per_key_base(cnf: Cnf, keys):
    """vmap over keys of cnf.sample_base(key, 1)[0]: the draws are vmapped (pure jax), the library is called once."""
    eps = jax.vmap(lambda k: jax.random.normal(k, (1, cnf.n_frames, cnf.dim), jnp.float32).reshape(cnf.D))(keys)
    return _from_noise(cnf, eps), eps


def log_prob_base(cnf: Cnf, x):
    return jax.ffi.ffi_call("ecnf_base_log_prob", jax.ShapeDtypeStruct((x.shape[0],), jnp.float32))(
        x, model=np.int64(cnf.handle))


def sample_cnf(cnf: Cnf, params, keys, features, use_fixed_step_size=False, rtol=1e-5, atol=1e-5, step_size=0.05):
    """sample_and_log_prob.py:11-38, vmapped over the leading axis of `keys` [B, 2] / `features` [B, n]."""
    x0, _ = _per_key_base(cnf, keys)
    x1, _, stats = _solve(cnf, L.MODE_SAMPLE, pack(cnf, params), x0, features, _ctrl(use_fixed_step_size, rtol, atol, step_size))
    return x1


def get_log_prob(cnf: Cnf, params, x, keys, features, approx=False, use_fixed_step_size=False, rtol=1e-5, atol=1e-5,
                 step_size=0.05):
    """sample_and_log_prob.py:41-94: (log_p, log_prob_base, delta); approx=True draws eps = normal(key, x.shape) (:55)."""
    eps = jax.vmap(lambda k: jax.random.normal(k, (cnf.D,), jnp.float32))(keys) if approx else None
    _, logs, _ = _solve(cnf, L.MODE_LOGPROB, pack(cnf, params), x, features, _ctrl(use_fixed_step_size, rtol, atol, step_size), eps)
    return logs[:, 0], logs[:, 1], logs[:, 2]


def sample_and_log_prob_cnf(cnf: Cnf, params, keys, features, approx=False, use_fixed_step_size=False, rtol=1e-5,
                            atol=1e-5, step_size=0.05):
    """sample_and_log_prob.py:97-149: (x1, log_q).  approx=True reuses the base-sample key for the probe like the reference
    (:130,137: the same key gives the same normal draw, so eps is the raw noise underneath x0)."""
    x0, raw = _per_key_base(cnf, keys)
    x1, logs, _ = _solve(cnf, L.MODE_SAMPLE_LOGQ, pack(cnf, params), x0, features,
                         _ctrl(use_fixed_step_size, rtol, atol, step_size), raw if approx else None)
    return x1, logs[:, 0]


def flow_matching_loss_and_grad(cnf: Cnf, params, x_data, key, features):
    """loss.py:10-32 + jax.grad (gradient_step.py:31-37) as ONE call: returns (loss, grads pytree)."""
    B = x_data.shape[0]
    key1, key2 = jax.random.split(key)                                   # loss.py:21
    x0, _ = sample_base(cnf, key1, B)                                    # loss.py:22
    t = jax.random.uniform(key2, (B,), jnp.float32)                      # loss.py:24
    ws = _ws(L.load().ecnf_fm_workspace_bytes(cnf.handle, B))
    loss, grad = jax.ffi.ffi_call("ecnf_fm_loss_grad", (jax.ShapeDtypeStruct((1,), jnp.float32),
                                                        jax.ShapeDtypeStruct((cnf.param_count,), jnp.float32)))(
        pack(cnf, params), x_data, x0, t, features.astype(jnp.int32), ws, model=np.int64(cnf.handle),
        loss_denominator=np.float32(B * cnf.D))
    return loss[0], unpack(cnf, grad)


def adam_step(flat, grad, mu, nu, ema, count: int, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8, ema_beta=0.999):
    """optax.adam + EMA + global norms on the flat buffers, in place (gradient_step.py:39-50)."""
    shapes = tuple(jax.ShapeDtypeStruct(a.shape, jnp.float32) for a in (flat, mu, nu, ema)) + (jax.ShapeDtypeStruct((2,), jnp.float32),)
    return jax.ffi.ffi_call("ecnf_adam_step", shapes, input_output_aliases={0: 0, 2: 1, 3: 2, 4: 3})(
        flat, grad, mu, nu, ema, count=np.int64(count), step=np.int64(step), lr=np.float32(lr), b1=np.float32(b1),
        b2=np.float32(b2), eps=np.float32(eps), ema_beta=np.float32(ema_beta))


def ess(log_w):
    """reverse / forward ESS from the five mergeable statistics (setup_training.py:175-182, utils/evaluation.py:10-22)."""
    s = jax.ffi.ffi_call("ecnf_ess_stats", jax.ShapeDtypeStruct((5,), jnp.float32))(log_w)
    n = log_w.shape[0]
    reverse = s[1] ** 2 / (n * s[2])
    forward = jnp.exp(2.0 * jnp.log(float(n)) - (s[0] + jnp.log(s[1])) - (s[3] + jnp.log(s[4])))
    return reverse, forward

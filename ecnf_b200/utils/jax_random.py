"""Host-side restatement of the part of `jax.random` the reference's hot path draws its noise from, so that a jax-style
key (uint32[2]) gives the SAME noise here as in the reference:

  * `flow_matching_loss_fn` (ecnf/cnf/loss.py:21-24): `key1, key2 = jax.random.split(key)`;
    `x0 = cnf.sample_base(key1, B)`; `t = jax.random.uniform(key2, shape=(B,))`
  * `cnf.sample_base(key, n)` (ecnf/cnf/build_cnf.py:46-61, zero_com_base.py:15-18,40-42,88-93): distrax passes the key
    through unchanged, so it is `base_scale * remove_mean(jax.random.normal(key, (n, n_nodes, dim)))`
  * `flow_matching_update_fn` (gradient_step.py:30): `key, subkey = jax.random.split(state.key)`
  * `sample_cnf` / `sample_and_log_prob_cnf` (sample_and_log_prob.py:24,112): `x0 = cnf.sample_base(key, 1)[0]` per
    trajectory, the caller vmaps over `jax.random.split(key, n)` (setup_training.py:47)

jax is not installed here, so this follows the published algorithm of jax's default PRNG of the reference's era
(`jax._src.prng`: threefry2x32, 20 rounds, non-partitionable bit generation; `jax._src.random`: `_uniform`,
`_normal_real` = sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1)); XLA's float32 `ErfInv` polynomial, M. Giles 2010).
It is PINNED by known answers (tests/test_jax_random.py): the three Random123 / `jax/tests/random_test.py` vectors for
threefry2x32, `split(PRNGKey(0)) = [[4146024105, 967050713], [2718843009, 1272950319]]`,
`uniform(PRNGKey(0)) = 0.41845703`, `normal(PRNGKey(0), (1,)) = -0.20584226`, `normal(subkey, (1,)) = -1.2515389`,
the ten values of the quickstart's `normal(PRNGKey(0), (10,))` and `normal(PRNGKey(42)) = -0.18471177` (values printed
in the JAX documentation: quickstart, "Sharp Bits", "Pseudo random numbers").  The last ulp of `normal` depends on
the libm `log`; XLA:CPU's may differ from numpy's there.

This is the drop-in (key-parity) noise path: numpy on the host, then one H2D copy.  The fast path for large batches
stays the device Philox stream (`ecnf_base_sample`, integer seeds).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

_U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey for a non-negative Python int (threefry key = [high word, low word])."""
    seed = int(seed)
    if seed < 0:
        raise ValueError("negative seeds are not supported")
    return np.asarray([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=_U32)


def is_key(key) -> bool:
    """True for a jax-style raw key (or a batch of keys): an array of uint32 with a trailing axis of 2."""
    if isinstance(key, (int, np.integer)) or key is None:
        return False
    try:
        import torch
        if isinstance(key, torch.Tensor):
            return key.dtype == torch.uint32 and key.dim() >= 1 and key.shape[-1] == 2
    except ImportError:      # pragma: no cover
        pass
    a = np.asarray(key)
    return a.dtype == _U32 and a.ndim >= 1 and a.shape[-1] == 2


def as_key(key) -> np.ndarray:
    try:
        import torch
        if isinstance(key, torch.Tensor):
            key = key.cpu().numpy()
    except ImportError:      # pragma: no cover
        pass
    a = np.asarray(key, dtype=_U32)
    if a.shape[-1] != 2:
        raise ValueError("a threefry key has a trailing axis of 2")
    return a


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(key: np.ndarray, x0: np.ndarray, x1: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """The Threefry-2x32 block function, 20 rounds (Salmon et al. 2011; jax._src.prng._threefry2x32_lowering)."""
    with np.errstate(over="ignore"):
        k0, k1 = np.asarray(key[0], _U32), np.asarray(key[1], _U32)      # scalars, or arrays broadcast against x0 / x1
        ks = (k0, k1, (k0 ^ k1 ^ _U32(0x1BD11BDA)).astype(_U32))
        x0 = x0.astype(_U32) + ks[0]
        x1 = x1.astype(_U32) + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r) ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
    return x0.astype(_U32), x1.astype(_U32)


def threefry_2x32(key: np.ndarray, count: np.ndarray) -> np.ndarray:
    """jax._src.prng.threefry_2x32: the flattened counts are split into two halves (odd sizes padded with one zero)."""
    count = np.asarray(count, dtype=_U32).ravel()
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, _U32)])
    h = count.size // 2
    a, b = threefry2x32(as_key(key), count[:h], count[h:])
    out = np.concatenate([a, b])
    return out[:-1] if odd else out


def split(key, num: int = 2) -> np.ndarray:
    """jax.random.split -> [num, 2] keys."""
    return threefry_2x32(key, np.arange(2 * num, dtype=_U32)).reshape(num, 2)


def random_bits(key, shape) -> np.ndarray:
    """32 random bits per element (jax._src.prng.threefry_random_bits, non-partitionable order)."""
    size = int(np.prod(shape, dtype=np.int64)) if len(tuple(shape)) else 1
    if size >= 2 ** 32:
        raise ValueError("more than 2^32 - 1 draws from one key")
    return threefry_2x32(key, np.arange(size, dtype=_U32)).reshape(tuple(shape))


def uniform(key, shape=(), minval: float = 0.0, maxval: float = 1.0) -> np.ndarray:
    """jax.random.uniform, float32: mantissa bits -> [1, 2) -> [minval, maxval)."""
    bits = random_bits(key, shape)
    floats = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, (floats * (hi - lo) + lo).astype(np.float32))


_ERFINV_CENTRAL = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
                   -0.00417768164, 0.246640727, 1.50140941)
_ERFINV_TAIL = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
                0.00943887047, 1.00167406, 2.83297682)


def erf_inv(x: np.ndarray) -> np.ndarray:
    """XLA's float32 ErfInv (Giles, "Approximating the erfinv function"): w = -log((1-x)(1+x)), two polynomials."""
    x = np.asarray(x, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log(((np.float32(1) - x) * (np.float32(1) + x)).astype(np.float32))).astype(np.float32)
        central = w < np.float32(5)
        wc = np.where(central, w - np.float32(2.5), np.sqrt(w) - np.float32(3)).astype(np.float32)
        p = np.where(central, np.float32(_ERFINV_CENTRAL[0]), np.float32(_ERFINV_TAIL[0])).astype(np.float32)
        for a, b in zip(_ERFINV_CENTRAL[1:], _ERFINV_TAIL[1:]):
            p = (np.where(central, np.float32(a), np.float32(b)) + p * wc).astype(np.float32)
        out = (p * x).astype(np.float32)
    with np.errstate(invalid="ignore"):
        return np.where(np.abs(x) == 1, np.float32(np.inf) * x, out).astype(np.float32)


def normal(key, shape=()) -> np.ndarray:
    """jax.random.normal, float32 (jax._src.random._normal_real)."""
    lo = np.nextafter(np.float32(-1), np.float32(0))
    u = uniform(key, shape, lo, 1.0)
    return (np.float32(np.sqrt(2)) * erf_inv(u)).astype(np.float32)


# ---- the reference's draws -----------------------------------------------------------------------------------------
def sample_base(key, n: int, n_nodes: int, dim: int, base_scale: float) -> np.ndarray:
    """cnf.sample_base(key, n) -> [n, n_nodes * dim] (build_cnf.py:46-61; zero_com_base.py:88-93)."""
    x = normal(key, (n, n_nodes, dim))
    x = x - x.mean(axis=-2, keepdims=True, dtype=np.float32)
    return (x.reshape(n, n_nodes * dim) * np.float32(base_scale)).astype(np.float32)


def normal_per_key(keys, size: int) -> np.ndarray:
    """vmap over keys of `jax.random.normal(key, (size,))` -> [B, size], all keys in one vectorised threefry pass."""
    keys = as_key(keys).reshape(-1, 2)
    count = np.arange(size + (size % 2), dtype=_U32)
    count[size:] = 0                                       # odd sizes: one ZERO padding counter, last output dropped
    h = count.size // 2
    kb = (keys[:, 0:1], keys[:, 1:2])                      # [B, 1] key words broadcast against the counter halves
    a, b = threefry2x32(kb, count[None, :h], count[None, h:])
    bits = np.concatenate([a, b], axis=1)[:, :size]
    lo = np.nextafter(np.float32(-1), np.float32(0))
    floats = ((bits >> _U32(9)) | _U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    u = np.maximum(lo, (floats * (np.float32(1.0) - lo) + lo).astype(np.float32))
    return (np.float32(np.sqrt(2)) * erf_inv(u)).astype(np.float32)


def sample_base_per_key(keys, n_nodes: int, dim: int, base_scale: float) -> np.ndarray:
    """vmap over keys of `cnf.sample_base(key, 1)[0]` (sample_and_log_prob.py:24 under setup_training.py:47) -> [B, D]."""
    x = normal_per_key(keys, n_nodes * dim).reshape(-1, n_nodes, dim)
    x = x - x.mean(axis=-2, keepdims=True, dtype=np.float32)
    return (x.reshape(-1, n_nodes * dim) * np.float32(base_scale)).astype(np.float32)


def fm_noise(key, batch: int, n_nodes: int, dim: int, base_scale: float) -> Tuple[np.ndarray, np.ndarray]:
    """(x0 [B, D], t [B]) of flow_matching_loss_fn (loss.py:21-24)."""
    key1, key2 = split(key)
    return sample_base(key1, batch, n_nodes, dim, base_scale), uniform(key2, (batch,))

"""Host logic: the numpy restatement of jax.random (ecnf_b200/utils/jax_random.py) against known answers.
jax is not installed, so the pins are published values: the Random123 known-answer vectors for Threefry-2x32 (the ones
jax/tests/random_test.py::testThreefry2x32 checks) and values printed in the JAX documentation for PRNGKey(0)."""
import numpy as np

from ecnf_b200.utils import jax_random as jr


def test_threefry2x32_known_answers():
    kat = [((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, want in kat:
        got = jr.threefry_2x32(np.asarray(key, np.uint32), np.asarray(ctr, np.uint32))
        assert tuple(int(v) for v in got) == want


def test_split_uniform_normal_documented_values():
    key = jr.PRNGKey(0)
    assert key.tolist() == [0, 0] and jr.PRNGKey(42).tolist() == [0, 42] and jr.PRNGKey((7 << 32) + 5).tolist() == [7, 5]
    k, sub = jr.split(key)
    assert k.tolist() == [4146024105, 967050713] and sub.tolist() == [2718843009, 1272950319]
    assert abs(float(jr.uniform(key)) - 0.41845703) < 1e-8
    assert abs(float(jr.normal(key, (1,))[0]) - (-0.20584226)) < 2e-7
    assert abs(float(jr.normal(sub, (1,))[0]) - (-1.2515389)) < 2e-7
    # JAX quickstart: x = random.normal(random.PRNGKey(0), (10,)); JAX-101 "Pseudo random numbers": normal(PRNGKey(42))
    quick = [-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377, -0.1521442, -0.67135346, -0.5908641,
             0.73168886, 0.5673026]
    assert np.abs(jr.normal(key, (10,)) - np.asarray(quick, np.float32)).max() < 2e-7
    assert abs(float(jr.normal(jr.PRNGKey(42))) - (-0.18471177)) < 2e-7


def test_bits_layout_and_ranges():
    key = jr.PRNGKey(3)
    # an odd number of draws pads the counter array with one zero and drops the last output
    assert np.array_equal(jr.random_bits(key, (5,)), jr.threefry_2x32(key, np.arange(5, dtype=np.uint32)))
    assert np.array_equal(jr.random_bits(key, (2, 3)).ravel(), jr.random_bits(key, (6,)))
    u = jr.uniform(key, (20000,))
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    z = jr.normal(key, (20000,))
    assert np.isfinite(z).all() and abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    # erf_inv against the definition: erf(erf_inv(x)) = x
    from math import erf
    xs = np.asarray([-0.999, -0.9, -0.3, 0.0, 0.5, 0.99, 0.99999], np.float32)
    back = np.asarray([erf(float(v)) for v in jr.erf_inv(xs)])
    assert np.abs(back - xs).max() < 2e-6


def test_reference_draws_compose_as_in_loss_py():
    """flow_matching_loss_fn (loss.py:21-24): key1, key2 = split(key); x0 = sample_base(key1, B); t = uniform(key2, (B,))."""
    key = jr.PRNGKey(11)
    n, dim, B, scale = 13, 3, 6, 2.0
    x0, t = jr.fm_noise(key, B, n, dim, scale)
    k1, k2 = jr.split(key)
    z = jr.normal(k1, (B, n, dim))
    want = ((z - z.mean(axis=1, keepdims=True)) * np.float32(scale)).reshape(B, n * dim)
    assert np.allclose(x0, want, atol=1e-6) and np.array_equal(t, jr.uniform(k2, (B,)))
    assert np.abs(x0.reshape(B, n, dim).mean(axis=1)).max() < 1e-6          # zero centre of mass
    # per-trajectory keys (sample_cnf under vmap over split(key, B)): each row is its own sample_base(key_i, 1)[0]
    keys = jr.split(key, B)
    rows = jr.sample_base_per_key(keys, n, dim, scale)
    assert rows.shape == (B, n * dim) and np.array_equal(rows[2], jr.sample_base(keys[2], 1, n, dim, scale)[0])
    # the vectorised per-key pass equals one call per key (odd and even sizes)
    for size in (39, 8):
        assert np.array_equal(jr.normal_per_key(keys, size), np.stack([jr.normal(k, (size,)) for k in keys]))
    assert jr.is_key(keys) and jr.is_key(key) and not jr.is_key(7) and not jr.is_key(np.zeros(2, np.float32))

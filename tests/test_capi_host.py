"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/ecnf_b200.h declares, the
layout queries and error paths work without a GPU, and the host-side helpers (keys, sharding, ESS merging) behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from oracle import ecnf_oracle as O
from ecnf_b200 import lib as L
from ecnf_b200.engine import CnfConfig, Engine, ess_from_stats, key_to_seed, split_key
from ecnf_b200.distributed import merge_ess_stats_list, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ecnf_b200.h")).read()
    declared = set(re.findall(r"\b(ecnf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert lib.ecnf_version() == 100


@pytest.mark.parametrize("name", ["dw4", "lj13", "qm9", "aldp"])
def test_param_layout_matches_flax_pytree(name):
    ocfg = O.CONFIGS[name]
    eng = Engine(CnfConfig(ocfg.n_frames, ocfg.dim, ocfg.sigma_min, ocfg.base_scale, ocfg.n_blocks_egnn, ocfg.mlp_units,
                           ocfg.n_invariant_feat_hidden, ocfg.time_embedding_dim, ocfg.n_features))
    ref = O.param_layout(ocfg)
    assert [(p, tuple(s)) for p, _, s in eng.layout] == [(p, tuple(s)) for p, s in ref]
    end = 0
    for path, off, shape in eng.layout:
        assert off % 4 == 0 and off >= end           # 16-byte aligned, non-overlapping, in flax order
        end = off + (int(np.prod(shape)) if shape else 1)
    assert end <= eng.param_count < end + 4


def test_error_paths_return_codes_and_messages():
    lib = L.load()
    cfg = L.Config()
    cfg.n_frames, cfg.dim, cfg.n_blocks, cfg.n_layers, cfg.mlp_units, cfg.n_hidden, cfg.time_dim, cfg.n_features = 13, 3, 3, 3, 100, 64, 8, 1
    cfg.base_scale = 1.0
    h = C.c_void_p()
    rc = lib.ecnf_model_create(C.byref(cfg), None, C.byref(h))
    assert rc == -3 and b"mlp_units" in lib.ecnf_last_error() and not h.value
    with pytest.raises(L.EcnfError):
        Engine(CnfConfig(13, 3, 0.01, 1.0, 3, (128, 64, 128), 64, 8, 1))        # non-uniform widths
    with pytest.raises(L.EcnfError):
        Engine(CnfConfig(13, 4, 0.01, 1.0, 3, (128,) * 3, 64, 8, 1))            # dim 4
    eng = Engine(CnfConfig(13, 3, 0.01, 1.0, 3, (128,) * 3, 64, 8, 1))
    assert lib.ecnf_solve_workspace_bytes(eng.handle, L.MODE_SAMPLE_LOGQ, 10_000) > 100e6
    rc = lib.ecnf_solve(eng.handle, 99, None, None, 1, None, None, None, None, None, 0, None)
    assert rc == -1
    rc = lib.ecnf_adam_step(None, None, None, None, None, 0, 0, 1e-3, 0.9, 0.999, 1e-8, 0.999, None, None)
    assert rc == -1 and b"ecnf_adam_step" in lib.ecnf_last_error()
    if not torch.cuda.is_available():
        with pytest.raises(L.EcnfError):       # no silent CPU fallback
            eng.apply(O.flat_to_nested(O.init_params(O.CONFIGS["lj13"])), np.zeros((1, 39), np.float32), np.zeros(1, np.float32))


def test_schedule_entry_point_matches_oracle():
    lib = L.load()
    for step in (0, 3, 10, 11, 500, 1000, 5000):
        assert abs(lib.ecnf_warmup_cosine_lr(step, 1e-4, 1e-3, 10, 1000, 0.0) - O.warmup_cosine_lr(step, 1e-4, 1e-3, 10, 1000)) < 1e-9


def test_keys_shards_and_ess_merge():
    assert key_to_seed(7) == 7 and key_to_seed(np.asarray([1, 2], np.uint32)) == (1 << 32) | 2
    a, b = split_key(0)
    assert a != b and split_key(0) == [a, b] and split_key(1) != [a, b]
    for B, W in ((10_000, 8), (100_000, 8), (10, 4), (7, 8)):
        ranges = [shard_range(B, r, W) for r in range(W)]
        assert ranges[0][0] == 0 and ranges[-1][1] == B
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(W - 1))
        assert max(e - s for s, e in ranges) - min(e - s for s, e in ranges) <= 1
    rng = np.random.default_rng(0)
    lw = rng.standard_normal(4000) * 3 - 100

    def stats(w):
        mx, nmx = w.max(), (-w).max()
        return [mx, np.exp(w - mx).sum(), np.exp(2 * (w - mx)).sum(), nmx, np.exp(-w - nmx).sum()]
    merged = merge_ess_stats_list(torch.tensor([stats(lw[:1000]), stats(lw[1000:1500]), stats(lw[1500:])], dtype=torch.float64))
    rv, fw = ess_from_stats(merged.tolist(), lw.size)
    assert abs(rv - O.reverse_ess(lw)) < 1e-10 and abs(fw - O.forward_ess(lw, np.ones(lw.size, bool))) < 1e-10


def test_per_handle_attributes_fm_chunk_and_engine():
    """ecnf_model_set_fm_chunk / ecnf_model_set_engine are attributes of ONE handle: the training workspace follows the
    chunk size (a chunk's activations only), bad values are refused, and a second handle is untouched."""
    lib = L.load()
    qm9 = CnfConfig(19, 3, 1e-6, 2.0, 5, (256,) * 4, 32, 8, 1)
    a, b = Engine(qm9), Engine(qm9)
    full = lib.ecnf_fm_workspace_bytes(a.handle, 512)
    assert full > 7e9                                    # 2 x 5 x 4 [175 104 x 256] fp32 pre-activations (7.2 GB) and more
    a.set_fm_chunk(128)
    quarter = lib.ecnf_fm_workspace_bytes(a.handle, 512)
    assert quarter < 0.3 * full
    assert lib.ecnf_fm_workspace_bytes(a.handle, 100) == lib.ecnf_fm_workspace_bytes(b.handle, 100)   # B <= chunk: one chunk
    assert lib.ecnf_fm_workspace_bytes(b.handle, 512) == full
    a.set_fm_chunk(0)
    assert lib.ecnf_fm_workspace_bytes(a.handle, 512) == full
    assert lib.ecnf_model_set_fm_chunk(a.handle, -1) == -1 and b"fm_chunk" in lib.ecnf_last_error()
    assert lib.ecnf_model_set_engine(a.handle, 7) == -1
    with pytest.raises(L.EcnfError):
        a.set_engine(2)


def test_model_clone_binds_other_parameters_without_touching_the_original():
    """ecnf_model_clone: the per-call parameter binding an XLA FFI handler uses (clone -> launch -> destroy)."""
    lib = L.load()
    eng = Engine(CnfConfig(22, 3, 1e-6, 0.2, 3, (64, 64), 32, 8, 22))
    eng.set_engine(1)
    eng.set_fm_chunk(7)
    h = C.c_void_p()
    assert lib.ecnf_model_clone(eng.handle, None, C.byref(h)) == 0 and h.value and h.value != eng.handle
    try:
        assert lib.ecnf_model_param_count(h) == eng.param_count
        assert lib.ecnf_fm_workspace_bytes(h, 100) == lib.ecnf_fm_workspace_bytes(eng.handle, 100)     # same chunk size
        assert lib.ecnf_solve_tensor_flops_per_eval(h) == 0                                             # same engine choice (SIMT)
        assert lib.ecnf_model_set_engine(h, 0) == 0                                                     # the clone's own attribute
        assert lib.ecnf_solve_tensor_flops_per_eval(h) > 0 and lib.ecnf_solve_tensor_flops_per_eval(eng.handle) == 0
    finally:
        lib.ecnf_model_destroy(h)
    assert lib.ecnf_model_clone(None, None, C.byref(h)) == -1


def test_jax_ffi_shim_typechecks_against_the_c_abi():
    """integration/jax_ffi/ecnf_jax_ffi.cc cannot be built here (no jaxlib headers); against a minimal stand-in of
    xla/ffi/api/ffi.h every call it makes into include/ecnf_b200.h must still type-check, and every handler the Python
    side registers must be defined."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = os.path.join(ROOT, "integration", "jax_ffi", "ecnf_jax_ffi.cc")
    cuda_inc = "/usr/local/cuda/include"
    r = subprocess.run([gxx, "-fsyntax-only", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "tests", "mock_xla"),
                        "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    py = open(os.path.join(ROOT, "integration", "jax_ffi", "ecnf_jax.py")).read()
    symbols = set(re.findall(r'"(Ecnf[A-Za-z]+)"', py))
    assert len(symbols) == 11
    for s in symbols:
        assert f"XLA_FFI_DEFINE_HANDLER_SYMBOL({s}," in text, s

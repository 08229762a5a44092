"""Parity against the reference's OWN source files, executed over stand-ins of their dependencies.

`tests/golden/make_refsrc_golden.py` imports `/root/reference`'s modules unmodified on top of `tests/golden/refshim/`
(torch-backed stand-ins for the few dozen jax / flax / distrax / chex / e3nn_jax functions they call), runs them on seeded
inputs and writes `tests/golden/refsrc_*.npz` + `refsrc_layouts.json`; nothing of this repository's oracle or CUDA path
takes part in producing them.  Here the CPU oracle (not gpu-marked) and the CUDA path (gpu-marked) are checked against
those vectors.  What this pins and what it does not is stated in `tests/golden/refshim/README.md` and DESIGN.md section 2
(not pinned: the stand-ins' own semantics, XLA's float32 rounding, diffrax, optax's Adam)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import ecnf_oracle as O
from ecnf_b200.utils import jax_random as jr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(glob.glob(os.path.join(GOLDEN, "refsrc_*.npz")))
assert FILES, "tests/golden/refsrc_*.npz are committed fixtures"


def _load(path):
    g = np.load(path, allow_pickle=False)
    kw = json.loads(str(g["config"]))
    kw["mlp_units"] = tuple(kw["mlp_units"])
    flat = {k[len("param:"):]: g[k].astype(np.float64) for k in g.files if k.startswith("param:")}     # float32-representable values
    grads = {k[len("grad:"):]: g[k] for k in g.files if k.startswith("grad:")}
    return g, kw, O.CnfConfig(**kw), flat, grads


@pytest.fixture
def f64_frequencies(monkeypatch):
    """The oracle evaluates the timestep frequency table in float32 whatever the compute dtype (that is what jnp does with
    x64 off, build_cnf.py:25-27).  The float64 run of the reference's source evaluates it in float64, so for the float64
    comparisons the oracle gets the same table; the float32 run (the reference's real setting) is compared separately."""
    def table(T):
        half = T // 2
        return np.exp(np.arange(half, dtype=np.float64) * -(np.log(10_000.0) / (half - 1)))
    monkeypatch.setattr(O, "timestep_frequencies", table)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


def test_parameter_tree_is_the_one_the_reference_builds():
    """flax names and shapes for the four BASELINE networks: oracle layout == library layout == the tree
    the reference's own `cnf.init` produced."""
    from ecnf_b200.engine import CnfConfig, Engine
    layouts = json.load(open(os.path.join(GOLDEN, "refsrc_layouts.json")))
    assert set(layouts) >= {"dw4", "lj13", "qm9", "aldp"}
    for name, ref in layouts.items():
        ref = {p: tuple(s) for p, s in ref}
        ocfg = O.CONFIGS.get(name) or O.CnfConfig(n_frames=5, dim=3, sigma_min=0.01, base_scale=0.7, n_blocks_egnn=2, mlp_units=(64, 64),
                                                  n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=3)
        mine = {p: tuple(s) for p, s in O.param_layout(ocfg)}
        assert mine == ref, (name, set(mine) ^ set(ref))
        eng = Engine(CnfConfig(ocfg.n_frames, ocfg.dim, ocfg.sigma_min, ocfg.base_scale, ocfg.n_blocks_egnn, ocfg.mlp_units,
                               ocfg.n_invariant_feat_hidden, ocfg.time_embedding_dim, ocfg.n_features))
        assert {p: tuple(s) for p, _, s in eng.layout} == ref, name


@pytest.mark.parametrize("path", FILES)
def test_oracle_vector_field_divergence_and_joint_fields(path, f64_frequencies, monkeypatch):
    g, kw, cfg, flat, _ = _load(path)
    p64, p32 = O.to_torch(flat, torch.float64), O.to_torch(flat, torch.float32)
    x, t, feat = torch.tensor(g["x"]), torch.tensor(g["t"]), torch.tensor(g["feat"]).long()
    f, div = O.vf_and_exact_div(p64, cfg, x, t, feat)
    assert _rel(f.numpy(), g["f"]) < 1e-10 and np.abs(div.numpy() - g["div"]).max() < 1e-9
    _, h = O.vf_and_hutchinson_div(p64, cfg, x, t, feat, torch.tensor(g["lp_eps"]))
    assert np.abs(h.numpy() - g["lp_hutch_div"]).max() < 1e-9 * (1 + np.abs(g["lp_hutch_div"]).max())
    monkeypatch.undo()                        # float32 like-for-like: the oracle's own float32 frequency table
    f32, div32 = O.vf_and_exact_div(p32, cfg, x.float(), t.float(), feat)
    assert _rel(f32.numpy(), g["f_f32"]) < 2e-5 and np.abs(div32.numpy() - g["div_f32"]).max() < 2e-4
    # the reference's joint vector field IS (apply, +trace of the Jacobian) -- sample_and_log_prob.py:58-67
    assert _rel(g["lp_f"], g["f"]) < 1e-12 and np.abs(g["lp_exact_div"] - g["div"]).max() < 1e-10
    # Hutchinson variant with eps = normal(key, x.shape) (sample_and_log_prob.py:55, 69-78)
    D = cfg.n_frames * cfg.dim
    assert np.abs(jr.normal_per_key(g["keys"], D) - g["lp_eps"]).max() < 1e-7


@pytest.mark.parametrize("path", FILES)
def test_oracle_base_distribution_and_key_reuse(path, f64_frequencies):
    g, kw, cfg, flat, _ = _load(path)
    n, dim, s = cfg.n_frames, cfg.dim, cfg.base_scale
    # cnf.sample_base(key, n) and cnf.log_prob_base(x) (build_cnf.py:46-61: the ildj scaled by (n-1)/n)
    assert np.abs(jr.sample_base(g["base_key"], 5, n, dim, s) - g["base_samples_f32"]).max() < 1e-6
    assert np.abs(O.base_log_prob(cfg, torch.tensor(g["x"])).numpy() - g["base_logp"]).max() < 1e-9
    # sample_and_log_prob_cnf: x0, log_prob_base = cnf.sample_and_log_prob_base(seed=key) per trajectory (:112)
    x0 = jr.sample_base_per_key(g["keys"], n, dim, s)
    assert np.abs(x0 - g["sl_x0_f32"]).max() < 1e-6
    assert np.abs(O.base_log_prob(cfg, torch.tensor(g["sl_x0"])).numpy() - g["sl_logp_base"]).max() < 1e-9
    # ... and its Hutchinson probe re-uses the key: eps is the raw noise underneath x0 (:130, SURVEY C#6)
    eps = torch.tensor(g["sl_eps"])
    assert np.abs(O.base_sample_from_noise(cfg, eps).numpy() - g["sl_x0"]).max() < 1e-6
    p64 = O.to_torch(flat, torch.float64)
    f, h = O.vf_and_hutchinson_div(p64, cfg, torch.tensor(g["sl_x0"]), torch.tensor(g["t"]), torch.tensor(g["feat"]).long(), eps)
    assert _rel(f.numpy(), g["sl_f"]) < 1e-10
    assert np.abs(h.numpy() - g["sl_hutch_div"]).max() < 1e-9 * (1 + np.abs(g["sl_hutch_div"]).max())
    # sample_cnf: x0 = cnf.sample_base(key, 1)[0] (:24)
    assert np.abs(x0[0] - g["sc_x0"]).max() < 1e-6


@pytest.mark.parametrize("path", FILES)
def test_oracle_target_energies_and_forward_ess(path):
    """leonard_jones.py:10-27, double_well.py:9-19 (log p = -E) and utils/evaluation.py:10-22."""
    g = np.load(path, allow_pickle=False)
    x = g["energy_x"]
    assert np.abs(-O.lj_energy(x) - g["lj_logp"]).max() < 1e-9 * np.abs(g["lj_logp"]).max()
    assert np.abs(-O.dw_energy(x) - g["dw_logp"]).max() < 1e-9 * np.abs(g["dw_logp"]).max()
    fe = O.forward_ess(g["ess_log_w"], g["ess_mask"].astype(bool))
    assert abs(fe - float(g["forward_ess"])) < 1e-10 * float(g["forward_ess"])


@pytest.mark.parametrize("path", FILES[:1])
def test_solver_call_sites_of_the_reference(path):
    """The arguments the reference passes to diffeqsolve (captured from its own calls): what the on-device loop mirrors."""
    g = np.load(path, allow_pickle=False)
    s = json.loads(str(g["call_sites"]))
    ctrl = {"rtol": 1e-5, "atol": 1e-5, "dtmin": 1e-5}
    assert all(v["solver"] == "Dopri5" for v in s.values())
    assert s["sample_cnf_adaptive"] == dict(t0=0.0, t1=1.0, dt0=None, y0_is_tuple=False, controller=ctrl, solver="Dopri5")
    assert s["sample_cnf_fixed"]["dt0"] == 0.05 and s["sample_cnf_fixed"]["controller"] is None
    assert s["get_log_prob_adaptive"] == dict(t0=1.0, t1=0.0, dt0=None, y0_is_tuple=True, controller=ctrl, solver="Dopri5")
    assert s["get_log_prob_fixed"]["dt0"] == -0.05 and s["get_log_prob_fixed"]["y0_is_tuple"]
    assert s["sample_and_log_prob_adaptive"] == dict(t0=0.0, t1=1.0, dt0=None, y0_is_tuple=True, controller=ctrl, solver="Dopri5")
    assert s["sample_and_log_prob_fixed"]["y0_is_tuple"] is False      # quirk C#2: the fixed branch passes y0 = x0 and cannot run
    d = O.SolveControl()
    assert (d.rtol, d.atol, d.dtmin, d.step_size) == (1e-5, 1e-5, 1e-5, 0.05)


@pytest.mark.parametrize("path", FILES)
def test_oracle_flow_matching_loss_gradient_and_update_bookkeeping(path, f64_frequencies):
    g, kw, cfg, flat, grads = _load(path)
    n, dim, s = cfg.n_frames, cfg.dim, cfg.base_scale
    B = g["x_data"].shape[0]
    feat = torch.tensor(g["feat"]).long()
    # the loss's own draws: key1, key2 = split(key); x0 = sample_base(key1, B); t = uniform(key2, (B,))  (loss.py:21-24)
    x0, t = jr.fm_noise(g["fm_key"], B, n, dim, s)
    assert np.abs(x0 - g["fm_x0_f32"]).max() < 1e-6 and np.abs(t - g["fm_t_f32"]).max() < 1e-7
    loss, og = O.fm_loss_and_grad(flat, cfg, torch.tensor(g["x_data"]), torch.tensor(g["fm_x0"]), torch.tensor(g["fm_t"]), feat,
                                  dtype=torch.float64)
    assert abs(float(loss) - float(g["fm_loss"])) < 1e-11 * abs(float(g["fm_loss"]))
    assert set(grads) <= set(og) and len(grads) >= 25
    for k, v in grads.items():                                  # stored in float32: 2^-24 relative rounding of the stored value
        assert np.abs(og[k].numpy() - v).max() <= 2e-7 * (np.abs(v).max() + 1e-30) + 1e-300, k
    # flow_matching_update_fn: key, subkey = split(state.key); the loss uses subkey (gradient_step.py:30-37)
    key_out, subkey = jr.split(g["upd_key_in"])
    assert (key_out == g["upd_key_out"]).all()
    ux0, ut = jr.fm_noise(subkey, B, n, dim, s)
    uloss, ug = O.fm_loss_and_grad(flat, cfg, torch.tensor(g["x_data"]), torch.tensor(ux0).double(), torch.tensor(ut).double(), feat,
                                   dtype=torch.float64)
    assert abs(float(uloss) - float(g["upd_loss"])) < 1e-6 * abs(float(g["upd_loss"]))       # (float32 draws fed in float64)
    gn = float(np.sqrt(sum((v.numpy() ** 2).sum() for v in ug.values())))
    assert abs(gn - float(g["upd_grad_norm"])) < 1e-6 * gn and abs(0.05 * gn - float(g["upd_update_norm"])) < 1e-6 * gn
    assert float(g["upd_opt_state"]) == 1.0 and bool(g["upd_sentinel_kept"])
    for leaf in ("EGNN_0/final_scaling", "EGNN_0/1/phi_e/Dense_1/kernel"):
        new = flat[leaf] - 0.05 * ug[leaf].numpy()                                          # apply_updates(params, updates)
        assert np.abs(new - g["upd_param:" + leaf]).max() < 1e-6 * (np.abs(new).max() + 1e-30)
        ema = 0.999 * flat[leaf] + 0.001 * g["upd_param:" + leaf]                           # EMA of the NEW params (:46-50)
        assert np.abs(ema - g["upd_ema:" + leaf]).max() < 1e-12 * (np.abs(ema).max() + 1e-30) + 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES)
def test_cuda_matches_the_reference_source(path, cuda_device):
    from ecnf_b200.engine import CnfConfig, Engine
    g, kw, cfg, flat, grads = _load(path)
    eng = Engine(CnfConfig(cfg.n_frames, cfg.dim, cfg.sigma_min, cfg.base_scale, cfg.n_blocks_egnn, cfg.mlp_units,
                           cfg.n_invariant_feat_hidden, cfg.time_embedding_dim, cfg.n_features))
    tree = O.flat_to_nested({k: v.astype(np.float32) for k, v in flat.items()})
    x, t, feat = g["x"].astype(np.float32), g["t"].astype(np.float32), g["feat"].astype(np.int32)
    for engine in (0, 1):
        eng.set_engine(engine)
        f, div = eng.apply_div(tree, x, t, feat)
        assert _rel(f.cpu().numpy(), g["f"]) < 2e-5, engine
        assert np.abs(div.cpu().numpy() - g["div"]).max() < 1e-4 * (1 + np.abs(g["div"]).max()), engine
        _, h = eng.apply_hutchinson(tree, x, t, g["lp_eps"].astype(np.float32), feat)
        assert np.abs(h.cpu().numpy() - g["lp_hutch_div"]).max() < 1e-4 * (1 + np.abs(g["lp_hutch_div"]).max()), engine
    eng.set_engine(0)
    assert np.abs(eng.base_log_prob(torch.tensor(x)).cpu().numpy() - g["base_logp"]).max() < 1e-4 * (1 + np.abs(g["base_logp"]).max())
    eps = torch.tensor(g["sl_eps"].astype(np.float32))
    assert np.abs(eng.base_sample_from_noise(eps).cpu().numpy() - g["sl_x0"]).max() < 1e-6
    from ecnf_b200 import lib as L
    xe = torch.tensor(g["energy_x"].astype(np.float32)).reshape(g["energy_x"].shape[0], -1)
    for kind, key, tol in ((L.TARGET_LJ, "lj_logp", 1e-4), (L.TARGET_DW, "dw_logp", 2e-5)):
        got = eng.target_log_prob(kind, xe).cpu().numpy()
        assert np.abs(got - g[key]).max() < tol * np.abs(g[key]).max(), key
    loss, grad = eng.fm_loss_grad(tree, g["x_data"].astype(np.float32), g["fm_x0"].astype(np.float32), g["fm_t"].astype(np.float32), feat)
    assert abs(float(loss[0]) - float(g["fm_loss"])) < 1e-5 * abs(float(g["fm_loss"]))
    mine = O.nested_to_flat(eng.unpack(grad, to_numpy=True)["params"])
    for k, v in grads.items():
        scale = np.abs(v).max()
        if scale == 0.0:
            assert np.abs(mine[k]).max() == 0.0, k          # dead parameters of the last block: exact zeros on both sides
        else:
            assert np.abs(mine[k] - v).max() < 1e-4 * scale, k


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES)
def test_facade_is_key_for_key_the_reference(path, cuda_device):
    """The reference-shaped Python API given the SAME jax-style keys as the reference's own code: same base draws, same base
    log-densities, same flow-matching loss (loss.py:21-24 draws included), same update bookkeeping."""
    from ecnf_b200.cnf import build_cnf, flow_matching_loss_fn, flow_matching_update_fn, TrainingState
    from ecnf_b200.utils.optim import Adam
    g, kw, cfg, flat, grads = _load(path)
    cnf = build_cnf(**kw)
    tree = O.flat_to_nested({k: v.astype(np.float32) for k, v in flat.items()})
    feat = g["feat"].astype(np.int32)
    assert np.abs(cnf.sample_base(g["base_key"], 5).cpu().numpy() - g["base_samples"]).max() < 1e-6
    assert np.abs(cnf.log_prob_base(torch.tensor(g["x"].astype(np.float32))).cpu().numpy() - g["base_logp"]).max() \
        < 1e-4 * (1 + np.abs(g["base_logp"]).max())
    x0, lp0 = cnf.sample_and_log_prob_base(g["keys"][0], ())
    assert np.abs(x0.cpu().numpy() - g["sl_x0"][0]).max() < 1e-6
    assert abs(float(lp0) - float(g["sl_logp_base"][0])) < 1e-4 * (1 + abs(float(g["sl_logp_base"][0])))
    loss, info = flow_matching_loss_fn(cnf, tree, g["x_data"].astype(np.float32), g["fm_key"], feat)
    assert abs(float(loss) - float(g["fm_loss"])) < 1e-5 * abs(float(g["fm_loss"]))
    # one update step: the new key and the loss (drawn from the sub-key) are the reference's
    opt = Adam(1e-4)
    params = cnf.engine.pack(tree)
    state = TrainingState(params=params, opt_state=opt.init(params), key=g["upd_key_in"], ema_params=params)
    new_state, uinfo = flow_matching_update_fn(cnf, opt.update, state, g["x_data"].astype(np.float32), feat)
    assert (np.asarray(new_state.key) == g["upd_key_out"]).all()
    assert abs(float(uinfo["loss"]) - float(g["upd_loss"])) < 1e-5 * abs(float(g["upd_loss"]))
    assert abs(float(uinfo["grad_norm"]) - float(g["upd_grad_norm"])) < 1e-4 * float(g["upd_grad_norm"])

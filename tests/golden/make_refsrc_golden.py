#!/usr/bin/env python
"""Golden vectors from the reference's OWN source files, executed here over stand-ins of their dependencies.

The reference (`/root/reference`) cannot run as shipped (no jax / flax / distrax / chex / e3nn_jax / diffrax / optax in
this image).  `tests/golden/refshim/` holds minimal torch-backed stand-ins of exactly the functions its hot-path files
call; this script imports the reference's modules UNMODIFIED on top of them, runs them on seeded inputs in float64 (the
ground truth) and float32, and writes `tests/golden/refsrc_<case>.npz`.  Nothing of this repository's oracle or CUDA path
is involved in producing the vectors.  `tests/test_refsrc_golden.py` then checks the CPU oracle (`-m "not gpu"`) and the
CUDA path (`-m gpu`) against them.

Run HERE (the container that has /root/reference):   python tests/golden/make_refsrc_golden.py
The `.npz` files travel to the GPU box; this script and the stand-ins do not need to.

Recorded per case (suffix _f32 = the float32 run, otherwise float64):
  param:<flax path>     seeded parameters written into the tree the reference's own `cnf.init` built (names + shapes are its)
  x, t, feat            inputs of the vector field
  f, div                cnf.apply (build_cnf.py:68-93 -> egnn.py:131-190) and the trace of its Jacobian per sample
  lp_*                  get_log_prob's joint vector fields (sample_and_log_prob.py:58-78) evaluated at (t, x): exact and
                        Hutchinson (with the probe eps = normal(key, x.shape) of :55), and the arguments of its diffeqsolve call
  sl_*                  sample_and_log_prob_cnf (:97-149): x0 / log_prob_base of cnf.sample_and_log_prob_base(seed=key), the
                        Hutchinson field whose probe re-uses `key` (:130), the arguments of the diffeqsolve calls (incl. the
                        fixed-step branch's y0, quirk C#2)
  sc_*                  sample_cnf (:11-38): x0 = cnf.sample_base(key, 1)[0] and the call-site arguments
  base_*                cnf.sample_base(key, n), cnf.log_prob_base(x)   (build_cnf.py:46-61, zero_com_base.py)
  lj_logp, dw_logp      the LJ / DW target log-densities of `energy_x` (leonard_jones.py:10-27, double_well.py:9-19)
  forward_ess           calculate_forward_ess(ess_log_w, ess_mask)      (utils/evaluation.py:10-22)
  fm_*                  flow_matching_loss_fn (loss.py:10-32) through jax.value_and_grad: x0, t (its own draws), loss, every gradient
                        (stored in float32; LJ13 keeps block 1 and the top-level tensors only)
  upd_*                 flow_matching_update_fn (gradient_step.py:20-53) with a caller-supplied linear `opt_update`: key
                        bookkeeping, info, new params / EMA of two leaves
  layout:<config>       the parameter tree (paths + shapes) the reference's init builds for the four BASELINE networks
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ECNF_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(1, REFERENCE)

import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402
import diffrax  # noqa: E402
from ecnf.cnf.build_cnf import build_cnf  # noqa: E402  (the reference)
from ecnf.cnf.loss import flow_matching_loss_fn  # noqa: E402
from ecnf.cnf.gradient_step import flow_matching_update_fn, TrainingState  # noqa: E402
from ecnf.cnf.sample_and_log_prob import sample_cnf, get_log_prob, sample_and_log_prob_cnf  # noqa: E402
from ecnf.targets.target_energy import leonard_jones, double_well  # noqa: E402
from ecnf.utils.evaluation import calculate_forward_ess  # noqa: E402

CASES = {
    "dw4": dict(n_frames=4, dim=2, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128, 128, 128),
                n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1),
    "small_64_32": dict(n_frames=5, dim=3, sigma_min=0.01, base_scale=0.7, n_blocks_egnn=2, mlp_units=(64, 64),
                        n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=3),
    "lj13": dict(n_frames=13, dim=3, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128, 128, 128),
                 n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1),
}
LAYOUTS = dict(CASES, **{
    "qm9": dict(n_frames=19, dim=3, sigma_min=1e-6, base_scale=2.0, n_blocks_egnn=5, mlp_units=(256, 256, 256, 256),
                n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=1),
    "aldp": dict(n_frames=22, dim=3, sigma_min=1e-6, base_scale=0.2, n_blocks_egnn=3, mlp_units=(64, 64),
                 n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=22),
})
B = 3


def flatten(tree, pre=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(flatten(v, pre + k + "/"))
        else:
            out[pre + k] = v
    return out


def set_leaf(tree, path, value):
    parts = path.split("/")
    node = tree
    for p in parts[:-1]:
        node = node[p]
    node[parts[-1]] = value


def seeded_values(shapes, seed):
    """Deterministic O(1)-field parameters for the tree the reference built (float64 masters)."""
    rng = np.random.default_rng(seed)
    vals = {}
    for path in sorted(shapes):
        shape = shapes[path]
        if path.endswith("final_scaling"):
            v = np.asarray(1.25)
        elif path.endswith("embedding"):
            v = rng.standard_normal(shape) / np.sqrt(shape[1])
        elif path.endswith("bias"):
            v = rng.standard_normal(shape) * 0.1
        else:
            v = rng.standard_normal(shape) / np.sqrt(shape[0])
        vals[path] = np.asarray(v, np.float32).astype(np.float64)      # float32-representable: stored as float32, exact in both runs
    return vals


def to_np(a):
    return np.asarray(a.detach().cpu().numpy()) if isinstance(a, torch.Tensor) else np.asarray(a)


def capture(fn, *args, **kwargs):
    """Runs a reference function until its diffeqsolve call; returns the captured call (term + keyword arguments)."""
    try:
        fn(*args, **kwargs)
    except diffrax.Captured:
        return dict(diffrax.LAST)
    raise AssertionError("diffeqsolve was not reached")


def call_site(cap):
    kw = cap["kwargs"]
    ctrl = kw.get("stepsize_controller")
    y0 = kw["y0"]
    return dict(t0=float(kw["t0"]), t1=float(kw["t1"]), dt0=None if kw["dt0"] is None else float(kw["dt0"]),
                y0_is_tuple=isinstance(y0, tuple), controller=None if ctrl is None else {k: float(v) for k, v in ctrl.kwargs.items()},
                solver=type(cap["solver"]).__name__)


def run_case(name, kw, dtype, masters):
    torch.set_default_dtype(dtype)
    n, dim = kw["n_frames"], kw["dim"]
    D = n * dim
    cnf = build_cnf(**kw)
    rng = np.random.default_rng(7)
    x = torch.tensor(rng.standard_normal((B, D)) * 1.2 + 0.1, dtype=dtype)
    t = torch.tensor(rng.uniform(0, 1, B), dtype=dtype)
    feat = torch.tensor(rng.integers(0, kw["n_features"], (B, n)))
    params = cnf.init(jax.random.PRNGKey(42), x, t, feat)
    shapes = {k: tuple(v.shape) for k, v in flatten(params["params"]).items()}
    if masters is None:
        masters = seeded_values(shapes, seed=1)
    for path, v in masters.items():
        assert tuple(v.shape) == shapes[path], path
        set_leaf(params["params"], path, torch.tensor(v, dtype=dtype))
    out = {"x": to_np(x), "t": to_np(t), "feat": to_np(feat)}

    # ---- vector field and the trace of its Jacobian
    out["f"] = to_np(cnf.apply(params, x, t, feat))
    divs = []
    for i in range(B):
        jac = jax.jacrev(lambda xi: cnf.apply(params, xi[None], t[i][None], feat[i][None])[0])(x[i])
        divs.append(torch.trace(jac))
    out["div"] = to_np(torch.stack(divs))

    # ---- get_log_prob: joint vector fields at (t_i, x_i) and the call site
    keys = jax.random.split(jax.random.PRNGKey(5), B)
    out["keys"] = to_np(keys).astype(np.uint32)
    lp_exact, lp_hutch, lp_eps, lp_f = [], [], [], []
    for i in range(B):
        cap = capture(get_log_prob, cnf, params, x[i], keys[i], feat[i], False, False)
        vf, dv = cap["term"].vector_field(t[i], (x[i], jnp.zeros(())), None)
        lp_f.append(vf); lp_exact.append(dv)
        cap_h = capture(get_log_prob, cnf, params, x[i], keys[i], feat[i], True, False)
        _, dvh = cap_h["term"].vector_field(t[i], (x[i], jnp.zeros(())), None)
        lp_hutch.append(dvh)
        lp_eps.append(jax.random.normal(keys[i], x[i].shape))        # sample_and_log_prob.py:55
    out.update(lp_f=to_np(torch.stack(lp_f)), lp_exact_div=to_np(torch.stack(lp_exact)), lp_hutch_div=to_np(torch.stack(lp_hutch)),
               lp_eps=to_np(torch.stack(lp_eps)))
    sites = {"get_log_prob_adaptive": call_site(cap),
             "get_log_prob_fixed": call_site(capture(get_log_prob, cnf, params, x[0], keys[0], feat[0], False, True, 1e-5, 1e-5, 0.05))}

    # ---- sample_and_log_prob_cnf: base draw, Hutchinson field with the key re-used for the probe, call sites
    sl_x0, sl_lpb, sl_f, sl_hutch, sl_eps = [], [], [], [], []
    for i in range(B):
        cap = capture(sample_and_log_prob_cnf, cnf, params, keys[i], feat[i], True, False)
        x0 = cap["kwargs"]["y0"][0]
        vf, dvh = cap["term"].vector_field(t[i], (x0, jnp.zeros(())), None)
        sl_x0.append(x0); sl_f.append(vf); sl_hutch.append(dvh)
        _, lpb = cnf.sample_and_log_prob_base(seed=keys[i], sample_shape=())
        sl_lpb.append(lpb)
        sl_eps.append(jax.random.normal(keys[i], x0.shape))          # sample_and_log_prob.py:130
    out.update(sl_x0=to_np(torch.stack(sl_x0)), sl_logp_base=to_np(torch.stack(sl_lpb)), sl_f=to_np(torch.stack(sl_f)),
               sl_hutch_div=to_np(torch.stack(sl_hutch)), sl_eps=to_np(torch.stack(sl_eps)))
    sites["sample_and_log_prob_adaptive"] = call_site(cap)
    sites["sample_and_log_prob_fixed"] = call_site(capture(sample_and_log_prob_cnf, cnf, params, keys[0], feat[0], False, True,
                                                           1e-5, 1e-5, 0.05))
    cap = capture(sample_cnf, cnf, params, keys[0], feat[0], False)
    sites["sample_cnf_adaptive"] = call_site(cap)
    out["sc_x0"] = to_np(cap["kwargs"]["y0"])
    sites["sample_cnf_fixed"] = call_site(capture(sample_cnf, cnf, params, keys[0], feat[0], True, 1e-5, 1e-5, 0.05))
    out["call_sites"] = np.asarray(json.dumps(sites))

    # ---- base distribution
    bkey = jax.random.PRNGKey(9)
    out["base_key"] = to_np(bkey).astype(np.uint32)
    out["base_samples"] = to_np(cnf.sample_base(bkey, 5))
    out["base_logp"] = to_np(cnf.log_prob_base(x))

    # ---- target log-densities and the forward ESS (leonard_jones.py:10-27, double_well.py:9-19, utils/evaluation.py:10-22)
    xe = torch.tensor(rng.standard_normal((B + 2, n, dim)) * 0.9 + 0.3, dtype=dtype)
    out.update(energy_x=to_np(xe), lj_logp=to_np(leonard_jones.log_prob_fn(xe)), dw_logp=to_np(double_well.log_prob_fn(xe)))
    log_w = torch.tensor(rng.standard_normal(64) * 2.0 - 30.0, dtype=dtype)
    mask = torch.tensor((rng.uniform(0, 1, 64) < 0.8).astype(np.int64))
    out.update(ess_log_w=to_np(log_w), ess_mask=to_np(mask), forward_ess=to_np(calculate_forward_ess(log_w, mask)["forward_ess"]))

    # ---- flow-matching loss and its gradient
    x_data = rng.standard_normal((B, n, dim))
    x_data = torch.tensor((x_data - x_data.mean(axis=1, keepdims=True)).reshape(B, D), dtype=dtype)
    fm_key = jax.random.PRNGKey(11)
    k1, k2 = jax.random.split(fm_key)
    out.update(x_data=to_np(x_data), fm_key=to_np(fm_key).astype(np.uint32), fm_x0=to_np(cnf.sample_base(k1, B)),
               fm_t=to_np(jax.random.uniform(k2, shape=(B,))))
    (loss, info), grads = jax.value_and_grad(flow_matching_loss_fn, has_aux=True, argnums=1)(cnf, params, x_data, fm_key, feat)
    out["fm_loss"] = to_np(loss)
    assert float(info["loss"]) == float(loss)
    for path, g in flatten(grads["params"]).items():
        out["grad:" + path] = to_np(g)

    # ---- update function with a linear stand-in optimiser (optax.adam itself is not part of the reference's tree)
    def opt_update(g, opt_state, params=None):
        return jax.tree_map(lambda a: -0.05 * a, g), opt_state + 1
    ukey = jax.random.PRNGKey(3)
    state = TrainingState(params=params, opt_state=torch.zeros(()), key=ukey, ema_params=params)
    new_state, uinfo = flow_matching_update_fn(cnf, opt_update, state, x_data, feat)
    out.update(upd_key_in=to_np(ukey).astype(np.uint32), upd_key_out=to_np(new_state.key).astype(np.uint32),
               upd_loss=to_np(uinfo["loss"]), upd_grad_norm=to_np(uinfo["grad_norm"]), upd_update_norm=to_np(uinfo["update_norm"]),
               upd_opt_state=to_np(new_state.opt_state))
    for leaf in ("EGNN_0/final_scaling", "EGNN_0/1/phi_e/Dense_1/kernel"):
        out["upd_param:" + leaf] = to_np(flatten(new_state.params["params"])[leaf])
        out["upd_ema:" + leaf] = to_np(flatten(new_state.ema_params["params"])[leaf])
    sentinel = torch.tensor(float("nan"))                     # setup_training.py:137 passes an array as "no EMA"
    st2, _ = flow_matching_update_fn(cnf, opt_update, state._replace(ema_params=sentinel), x_data, feat)
    out["upd_sentinel_kept"] = np.asarray(st2.ema_params is sentinel)
    return out, masters


def main():
    os.makedirs(HERE, exist_ok=True)
    for name, kw in CASES.items():
        r64, masters = run_case(name, kw, torch.float64, None)
        r32, _ = run_case(name, kw, torch.float32, masters)
        out = dict(r64)
        for k in ("f", "div", "lp_exact_div", "lp_hutch_div", "sl_x0", "sl_logp_base", "sl_hutch_div", "base_samples", "base_logp",
                  "fm_x0", "fm_t", "fm_loss", "upd_loss", "upd_grad_norm", "lj_logp", "dw_logp"):
            out[k + "_f32"] = r32[k]
        out.update({"param:" + p: v.astype(np.float32) for p, v in masters.items()})
        # gradients are stored in float32 (the fixtures stay a few MB); the big LJ13 case keeps block 1 and the top-level tensors
        for k in [k for k in out if k.startswith("grad:")]:
            if name == "lj13" and "/1/" not in k and k.count("/") > 2:
                del out[k]
            else:
                out[k] = out[k].astype(np.float32)
        out["config"] = np.asarray(json.dumps(kw))
        np.savez_compressed(os.path.join(HERE, f"refsrc_{name}.npz"), **out)
        print(f"{name}: f max {np.abs(out['f']).max():.3f}, div {out['div']}, loss {float(out['fm_loss']):.5f}, "
              f"{sum(k.startswith('grad:') for k in out)} gradient tensors")
    layouts = {}
    torch.set_default_dtype(torch.float32)
    for name, kw in LAYOUTS.items():
        cnf = build_cnf(**kw)
        D = kw["n_frames"] * kw["dim"]
        params = cnf.init(jax.random.PRNGKey(0), torch.zeros(1, D), torch.zeros(1), torch.zeros(1, kw["n_frames"], dtype=torch.long))
        layouts[name] = [[p, list(v.shape)] for p, v in flatten(params["params"]).items()]
    with open(os.path.join(HERE, "refsrc_layouts.json"), "w") as fh:
        json.dump(layouts, fh)
    print("layouts:", {k: len(v) for k, v in layouts.items()})


if __name__ == "__main__":
    main()

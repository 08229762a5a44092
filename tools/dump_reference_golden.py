#!/usr/bin/env python
"""Pins parity to the REAL reference the day a JAX environment exists ("first run the real reference to dump
golden vectors").

Needs `jax, flax, diffrax, distrax, optax, chex` importable and the reference tree on PYTHONPATH (default
/root/reference).  Runs the reference's own functions -- nothing of this repo's oracle or CUDA path is involved -- on small
seeded cases and writes `tests/golden/ref_<case>.npz` (+ one `ref_state_<case>.pkl` written exactly like
ecnf/utils/loop.py:144-153 does).  `tests/test_reference_golden.py` picks the files up when present and checks BOTH the
CPU oracle and the CUDA path against them; until then it skips and parity stays "unpinned" (DESIGN.md section 2).

What is recorded per case (all float32, the reference runs with x64 off):
  params (flax pytree, flattened to {"path": array}), x, t, feat
  f            cnf.apply(params, x, t, feat)                                       build_cnf.py:68-93
  div          trace of jax.jacfwd(apply wrt x) per sample                          sample_and_log_prob.py:58-67
  keys, x0     per-trajectory keys and cnf.sample_base(key, 1)[0]                   sample_and_log_prob.py:24
  x1_fixed     sample_cnf(..., use_fixed_step_size=True)                            sample_and_log_prob.py:11-38
  x1, logq     sample_and_log_prob_cnf(...) adaptive (the fixed branch is broken)   sample_and_log_prob.py:97-149
  logp3_fixed  get_log_prob(x1, fixed step) -> (log_p, log_p_base, delta)           sample_and_log_prob.py:41-94
  logp3        get_log_prob(x1) adaptive
  fm_key, fm_x0, fm_t, loss, grads     flow_matching_loss_fn + jax.grad             loss.py:10-32
  upd_*        one flow_matching_update_fn step with optax.adam(1e-4)               gradient_step.py:20-53
"""
import argparse
import os
import pickle
import sys

import numpy as np

CASES = {
    # name: build_cnf kwargs (small enough that the reference's D reverse passes per stage finish in minutes on a CPU)
    "dw4": dict(n_frames=4, dim=2, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128, 128, 128),
                n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1),
    "small_64_32": dict(n_frames=5, dim=3, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=2, mlp_units=(64, 64),
                        n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=3),
    "lj13": dict(n_frames=13, dim=3, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128, 128, 128),
                 n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1),
}


def flatten(tree, pre=""):
    out = {}
    for k, v in tree.items():
        if hasattr(v, "items"):
            out.update(flatten(v, pre + k + "/"))
        else:
            out[pre + k] = np.asarray(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    ap.add_argument("--cases", default="dw4,small_64_32,lj13")
    ap.add_argument("--batch", type=int, default=3)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    try:
        import jax
        import jax.numpy as jnp
        import optax
        import flax  # noqa: F401
        import diffrax  # noqa: F401
        import distrax  # noqa: F401
    except ImportError as e:
        raise SystemExit(f"the reference's dependencies are not importable here ({e}); parity stays unpinned")
    from ecnf.cnf.build_cnf import build_cnf
    from ecnf.cnf.sample_and_log_prob import sample_cnf, get_log_prob, sample_and_log_prob_cnf
    from ecnf.cnf.loss import flow_matching_loss_fn
    from ecnf.cnf.gradient_step import flow_matching_update_fn, TrainingState

    versions = {m.__name__: getattr(m, "__version__", "?") for m in (jax, flax, diffrax, distrax, optax)}
    for name in args.cases.split(","):
        kw = CASES[name]
        n, dim, B = kw["n_frames"], kw["dim"], args.batch
        D = n * dim
        cnf = build_cnf(**kw)
        rng = np.random.default_rng(7)
        x = rng.standard_normal((B, D)).astype(np.float32)
        t = rng.uniform(0, 1, B).astype(np.float32)
        feat = rng.integers(0, kw["n_features"], (B, n)).astype(np.int32)      # rank 2, as build_cnf.py:73-75 asserts
        params = cnf.init(jax.random.PRNGKey(42), jnp.asarray(x), jnp.asarray(t), jnp.asarray(feat))
        # the flax init makes the coordinate head tiny (variance 0.001): scale it up so the field is O(1) and the
        # adaptive controller does real work, as the "stiffened" synthetic parameters of the benches do
        p = flax.core.unfreeze(params) if hasattr(flax.core, "unfreeze") else params
        for b in range(kw["n_blocks_egnn"]):
            blk = p["params"]["EGNN_0"][str(b)]
            blk["Dense_0"]["kernel"] = blk["Dense_0"]["kernel"] * 30.0
        params = p

        f = cnf.apply(params, jnp.asarray(x), jnp.asarray(t), jnp.asarray(feat))

        def single(xi, ti, fi):
            return cnf.apply(params, xi[None], ti[None], fi[None])[0]
        div = jax.vmap(lambda xi, ti, fi: jnp.trace(jax.jacfwd(single)(xi, ti, fi)))(jnp.asarray(x), jnp.asarray(t), jnp.asarray(feat))

        keys = jax.random.split(jax.random.PRNGKey(5), B)
        x0 = jnp.stack([cnf.sample_base(k, 1)[0] for k in keys])
        feat1 = jnp.asarray(feat)
        x1_fixed = jnp.stack([sample_cnf(cnf, params, keys[i], feat1[i], True, 1e-5, 1e-5, 0.05) for i in range(B)])
        x1, logq = [], []
        for i in range(B):
            a, b_ = sample_and_log_prob_cnf(cnf, params, keys[i], feat1[i], False, False)
            x1.append(a); logq.append(b_)
        x1, logq = jnp.stack(x1), jnp.stack(logq)
        lp_fixed = jnp.stack([jnp.stack(get_log_prob(cnf, params, x1[i], keys[i], feat1[i], False, True, 1e-5, 1e-5, 0.05))
                              for i in range(B)])
        lp = jnp.stack([jnp.stack(get_log_prob(cnf, params, x1[i], keys[i], feat1[i], False, False)) for i in range(B)])

        fm_key = jax.random.PRNGKey(11)
        k1, k2 = jax.random.split(fm_key)
        fm_x0 = cnf.sample_base(k1, B)
        fm_t = jax.random.uniform(k2, shape=(B,))
        x_data = rng.standard_normal((B, n, dim)).astype(np.float32)
        x_data = (x_data - x_data.mean(axis=1, keepdims=True)).reshape(B, D)
        (loss, _), grads = jax.value_and_grad(flow_matching_loss_fn, has_aux=True, argnums=1)(
            cnf, params, jnp.asarray(x_data), fm_key, feat1)

        opt = optax.adam(1e-4)
        state = TrainingState(params=params, opt_state=opt.init(params), key=jax.random.PRNGKey(3), ema_params=params)
        new_state, info = flow_matching_update_fn(cnf, opt.update, state, jnp.asarray(x_data), feat1)

        out = dict(x=x, t=t, feat=feat, f=np.asarray(f), div=np.asarray(div), keys=np.asarray(keys), x0=np.asarray(x0),
                   x1_fixed=np.asarray(x1_fixed), x1=np.asarray(x1), logq=np.asarray(logq), logp3_fixed=np.asarray(lp_fixed),
                   logp3=np.asarray(lp), fm_key=np.asarray(fm_key), fm_x0=np.asarray(fm_x0), fm_t=np.asarray(fm_t),
                   x_data=x_data, loss=np.asarray(loss), upd_key_in=np.asarray(state.key), upd_key_out=np.asarray(new_state.key),
                   upd_loss=np.asarray(info["loss"]), upd_grad_norm=np.asarray(info["grad_norm"]),
                   upd_update_norm=np.asarray(info["update_norm"]), versions=np.asarray(repr(versions)))
        out.update({"param:" + k: v for k, v in flatten(params).items()})
        out.update({"grad:" + k: v for k, v in flatten(grads).items()})
        out.update({"upd_param:" + k: v for k, v in flatten(new_state.params).items()})
        out.update({"upd_ema:" + k: v for k, v in flatten(new_state.ema_params).items()})
        np.savez_compressed(os.path.join(args.out, f"ref_{name}.npz"), **out)
        with open(os.path.join(args.out, f"ref_state_{name}.pkl"), "wb") as fh:      # as ecnf/utils/loop.py:144-153
            pickle.dump(new_state, fh)
        print(name, "written;", versions)


if __name__ == "__main__":
    main()

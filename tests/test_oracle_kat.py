"""CPU tests that pin the oracle (oracle/ecnf_oracle.py) before anything is compared against it.

The reference has no golden vectors and cannot run here, so the oracle is validated by analytic
known-answer tests derived from the reference's own smoke scripts (core_test.py:23 linear field; egnn_test.py:31
equivariance), by independent implementations (scipy RK45, torch.optim.Adam, finite differences) and by the
committed float64 fixtures under tests/golden/.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import ecnf_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def small_cfg(**kw):
    base = dict(n_frames=5, dim=3, n_blocks_egnn=2, mlp_units=(16, 16), n_invariant_feat_hidden=32)   # egnn_test.py:17-22
    base.update(kw)
    return O.CnfConfig(**base)


def test_param_layout_counts_match_survey_appendix_d():
    for name, total in (("lj13", 510_407), ("dw4", 510_407), ("qm9", 3_794_283), ("aldp", 92_487)):
        cfg = O.CONFIGS[name]
        n = sum(int(np.prod(s)) if s else 1 for _, s in O.param_layout(cfg))
        assert n == total, (name, n)
    flat = O.init_params(O.CONFIGS["dw4"], 0)
    assert O.nested_to_flat(O.flat_to_nested(flat)).keys() == flat.keys()


def test_edge_order_and_safe_norm():
    send, recv = O.fully_connected_edges(4)          # utils/graph.py:6-14
    assert recv.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    assert send.tolist() == [1, 2, 3, 2, 3, 0, 3, 0, 1, 0, 1, 2]


def test_rotation_translation_permutation_properties():
    cfg = small_cfg()
    p = O.to_torch(O.init_params(cfg, 0, head_variance=1.0, bias_std=0.1), torch.float64)
    B, n, dim = 3, cfg.n_frames, cfg.dim
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, n, dim, dtype=torch.float64, generator=g)
    t = torch.tensor([0.1, 0.5, 0.9], dtype=torch.float64)
    feat = torch.zeros(B, n, dtype=torch.long)
    f = O.egnn_apply(p, cfg, x.reshape(B, -1), t, feat).reshape(B, n, dim)
    Q, _ = np.linalg.qr(np.random.default_rng(0).standard_normal((3, 3)))
    Q = torch.tensor(Q)
    fr = O.egnn_apply(p, cfg, (x @ Q.T).reshape(B, -1), t, feat).reshape(B, n, dim)
    assert (f @ Q.T - fr).abs().max() < 1e-12                        # egnn_test.py:31 (atol 1e-6 there)
    # translation: the field only sees x - mean(x), except the explicit "- mean(x)" of egnn.py:186 (SURVEY C#1)
    shift = torch.tensor([0.3, -1.0, 2.0], dtype=torch.float64)
    ft = O.egnn_apply(p, cfg, (x + shift).reshape(B, -1), t, feat).reshape(B, n, dim)
    assert (ft - (f - shift * p["EGNN_0/final_scaling"])).abs().max() < 1e-12
    perm = torch.tensor([2, 0, 4, 1, 3])
    fp = O.egnn_apply(p, cfg, x[:, perm].reshape(B, -1), t, feat).reshape(B, n, dim)
    assert (fp - f[:, perm]).abs().max() < 1e-12


def test_exact_divergence_loop_vs_batched_vs_finite_differences():
    cfg = small_cfg()
    p = O.to_torch(O.init_params(cfg, 1, head_variance=1.0, bias_std=0.1), torch.float64)
    B = 2
    x = torch.randn(B, cfg.D, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    t = torch.tensor([0.2, 0.8], dtype=torch.float64)
    feat = torch.zeros(B, cfg.n_frames, dtype=torch.long)
    f1, d1 = O.vf_and_exact_div(p, cfg, x, t, feat, batched=False)
    f2, d2 = O.vf_and_exact_div(p, cfg, x, t, feat, batched=True)
    assert (d1 - d2).abs().max() < 1e-12 and (f1 - f2).abs().max() == 0
    eps, fd = 1e-6, torch.zeros(B, dtype=torch.float64)
    for d in range(cfg.D):
        e = torch.zeros_like(x)
        e[:, d] = eps
        fd += (O.egnn_apply(p, cfg, x + e, t, feat)[:, d] - O.egnn_apply(p, cfg, x - e, t, feat)[:, d]) / (2 * eps)
    assert (d1 - fd).abs().max() < 1e-7
    # the "- mean(positions)" quirk contributes -dim * final_scaling ... to the trace of a zero network
    p0 = {k: torch.zeros_like(v) for k, v in p.items()}
    p0["EGNN_0/final_scaling"] = torch.tensor(1.0, dtype=torch.float64)
    _, d0 = O.vf_and_exact_div(p0, cfg, x, t, feat)
    assert (d0 + cfg.dim).abs().max() < 1e-12


def test_linear_field_known_answer_core_test():
    """core_test.py:23: apply = x*2 + x on a Gaussian base with scale 100: x(1) = e^3 x0, log q = log p0(x0) - 3 dim,
    and get_log_prob of x(1) returns the same number."""
    dim, scale = 3, 100.0
    x0 = torch.randn(11, dim, dtype=torch.float64, generator=torch.Generator().manual_seed(0)) * scale
    joint = lambda t, y: torch.cat([3 * y[:, :-1], torch.full((y.shape[0], 1), 3.0 * dim, dtype=y.dtype)], dim=1)
    y0 = torch.cat([x0, torch.zeros(11, 1, dtype=torch.float64)], dim=1)
    for ctrl, tol in ((O.SolveControl(), 3e-5), (O.SolveControl(fixed=True), 1e-6), (O.SolveControl(rtol=1e-10, atol=1e-10), 1e-9)):
        y1, st = O.dopri5(joint, y0, 0.0, 1.0, ctrl)
        assert ((y1[:, :-1] - math.exp(3) * x0).abs().max() / (math.exp(3) * scale)) < tol
        assert (y1[:, -1] - 3 * dim).abs().max() < 1e-5
        logp0 = lambda x: (-0.5 * (x / scale) ** 2 - math.log(scale) - 0.5 * math.log(2 * math.pi)).sum(-1)
        log_q = logp0(x0) - y1[:, -1]
        yb, _ = O.dopri5(joint, torch.cat([y1[:, :-1], torch.zeros(11, 1, dtype=torch.float64)], dim=1), 1.0, 0.0, ctrl)
        log_q_back = logp0(yb[:, :-1]) + yb[:, -1]
        assert (log_q - log_q_back).abs().max() < 1e-3
        if ctrl.fixed:
            assert (st.n_steps == 20).all() and (st.n_evals == 121).all()


def test_dopri5_against_scipy_rk45():
    from scipy.integrate import solve_ivp
    A = np.asarray([[-0.5, 2.0, 0.0], [-2.0, -0.5, 0.3], [0.1, 0.0, -1.0]])
    rhs = lambda t, y: A @ y + np.asarray([math.sin(3 * t), 0.0, math.cos(2 * t)])
    y0 = np.asarray([1.0, -0.5, 0.25])
    ref = solve_ivp(rhs, (0, 1), y0, method="RK45", rtol=1e-12, atol=1e-12).y[:, -1]
    At = torch.tensor(A)

    def func(t, y):
        forcing = torch.stack([torch.sin(3 * t), torch.zeros_like(t), torch.cos(2 * t)], dim=1)
        return y @ At.T + forcing
    y1, st = O.dopri5(func, torch.tensor(y0)[None], 0.0, 1.0, O.SolveControl(rtol=1e-9, atol=1e-9))
    assert np.abs(y1.numpy()[0] - ref).max() < 1e-7
    y1c, stc = O.dopri5(func, torch.tensor(y0)[None], 0.0, 1.0, O.SolveControl(rtol=1e-5, atol=1e-5))
    assert np.abs(y1c.numpy()[0] - ref).max() < 1e-4 and stc.n_steps[0] < st.n_steps[0]
    # per-trajectory control: a batch gives the same answer as each row alone
    ys = torch.tensor(np.stack([y0, 3 * y0]))
    yb, sb = O.dopri5(func, ys, 0.0, 1.0, O.SolveControl())
    y_single, s_single = O.dopri5(func, ys[1:2], 0.0, 1.0, O.SolveControl())
    assert (yb[1] - y_single[0]).abs().max() < 1e-14 and sb.n_steps[1] == s_single.n_steps[0]
    # max_steps is reported, not raised
    _, sm = O.dopri5(func, ys, 0.0, 1.0, O.SolveControl(max_steps=2))
    assert (sm.status == 1).all() and (sm.n_steps == 2).all()


def test_base_distribution_and_ot_path():
    cfg = O.CnfConfig(n_frames=13, dim=3, base_scale=2.0)
    eps = torch.randn(5, 39, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    x0 = O.base_sample_from_noise(cfg, eps)
    assert x0.reshape(5, 13, 3).mean(dim=1).abs().max() < 1e-14
    lp = O.base_log_prob(cfg, x0)
    z = x0 / 2.0
    ref = -0.5 * (z * z).sum(-1) - 0.5 * 36 * math.log(2 * math.pi) - 36 * math.log(2.0)
    assert (lp - ref).abs().max() < 1e-12
    assert (O.base_log_prob(cfg, x0 + 5.0) - lp).abs().max() < 1e-9     # log_prob removes the mean first
    x1 = torch.randn(5, 39, dtype=torch.float64)
    t = torch.rand(5, dtype=torch.float64)
    xt, ut = O.ot_conditional_vf(x0, x1, t, 0.01)
    assert (xt - ((1 - 0.99 * t[:, None]) * x0 + t[:, None] * x1)).abs().max() == 0 and (ut - (x1 - 0.99 * x0)).abs().max() == 0


def test_adam_and_schedule_against_independent_implementations():
    rng = np.random.default_rng(0)
    w = torch.tensor(rng.standard_normal(50), requires_grad=True)
    opt = torch.optim.Adam([w], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p, m, v = w.detach().numpy().copy(), np.zeros(50), np.zeros(50)
    for step in range(5):
        g = rng.standard_normal(50)
        w.grad = torch.tensor(g)
        opt.step()
        p, m, v, _ = O.adam_step(p, g, m, v, step, 1e-3)
        # torch divides sqrt(v)/sqrt(bc2) + eps, optax sqrt(v/bc2) + eps: identical up to rounding
        assert np.abs(p - w.detach().numpy()).max() < 1e-12
    assert O.warmup_cosine_lr(0, 1e-4, 1e-3, 10, 100) == 1e-4
    assert abs(O.warmup_cosine_lr(5, 1e-4, 1e-3, 10, 100) - 5.5e-4) < 1e-12
    assert O.warmup_cosine_lr(10, 1e-4, 1e-3, 10, 100) == 1e-3
    assert abs(O.warmup_cosine_lr(55, 1e-4, 1e-3, 10, 100) - 5e-4) < 1e-12
    assert abs(O.warmup_cosine_lr(100, 1e-4, 1e-3, 10, 100)) < 1e-15 and abs(O.warmup_cosine_lr(500, 1e-4, 1e-3, 10, 100)) < 1e-15


def test_ess_and_target_energies():
    lw = np.random.default_rng(2).standard_normal(1000) * 1.5 + 3
    w = np.exp(lw)
    assert abs(O.reverse_ess(lw) - (w.sum() ** 2 / (w * w).sum()) / 1000) < 1e-12
    assert abs(O.forward_ess(lw, np.ones(1000, bool)) - 1.0 / ((1 / w).mean() * w.mean())) < 1e-12
    assert O.reverse_ess(np.zeros(10)) == pytest.approx(1.0)
    x = np.random.default_rng(3).standard_normal((2, 13, 3))
    e = O.lj_energy(x)
    ref = 0.0
    for i in range(13):
        for j in range(13):
            if i != j:
                d = np.linalg.norm(x[0, i] - x[0, j])
                ref += 0.5 * (d ** -12 - 2 * d ** -6)
    ref += 0.5 * ((x[0] - x[0].mean(0)) ** 2).sum()
    assert abs(e[0] - ref) < 1e-9 * abs(ref)
    xd = np.random.default_rng(4).standard_normal((1, 4, 2)) * 2
    ref = sum(0.5 * (-4 * (np.linalg.norm(xd[0, i] - xd[0, j]) - 4) ** 2 + 0.9 * (np.linalg.norm(xd[0, i] - xd[0, j]) - 4) ** 4)
              for i in range(4) for j in range(4) if i != j)
    assert abs(O.dw_energy(xd)[0] - ref) < 1e-10 * abs(ref)


@pytest.mark.parametrize("name", ["dw4", "small_64_32"])
def test_oracle_fp32_reproduces_golden_fp64(name):
    from golden.make_golden import CASES
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    n, dim, blocks, units, H, nfeat = CASES[name]
    cfg = O.CnfConfig(n_frames=n, dim=dim, n_blocks_egnn=blocks, mlp_units=units, n_invariant_feat_hidden=H, n_features=nfeat)
    flat = O.init_params(cfg, seed=42, head_variance=1.0, bias_std=0.1)
    p32 = O.to_torch(flat, torch.float32)
    feat = torch.tensor(g["feat"])
    f, div = O.vf_and_exact_div(p32, cfg, torch.tensor(g["x"], dtype=torch.float32), torch.tensor(g["t"], dtype=torch.float32), feat)
    assert np.abs(f.numpy() - g["f"]).max() < 2e-5 * np.abs(g["f"]).max()
    assert np.abs(div.numpy() - g["div"]).max() < 5e-5 * (np.abs(g["div"]).max() + 1)
    x1, logq, _ = O.sample_and_log_prob_cnf(p32, cfg, torch.tensor(g["x0"], dtype=torch.float32), feat, O.SolveControl(fixed=True))
    assert np.abs(x1.numpy() - g["x1"]).max() < 1e-4 * np.abs(g["x1"]).max()
    assert np.abs(logq.numpy() - g["logq"]).max() < 1e-4 * (np.abs(g["logq"]).max() + 1)
    loss, grads = O.fm_loss_and_grad(flat, cfg, torch.tensor(g["x_data"]), torch.tensor(g["x0"]), torch.tensor(g["t"]), feat)
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    assert np.abs(grads["EGNN_0/0/phi_e/Dense_1/kernel"].numpy() - g["grad_phi_e"]).max() < 1e-4 * np.abs(g["grad_phi_e"]).max()


def test_hutchinson_estimator_is_unbiased_and_matches_jvp():
    """approx branch (sample_and_log_prob.py:69-78): eps^T J eps from one reverse pass equals eps . (J eps) by finite
    differences, and its mean over probes is the exact trace."""
    cfg = O.CnfConfig(n_frames=4, dim=2, n_blocks_egnn=2, mlp_units=(64, 64), n_invariant_feat_hidden=32)
    flat = O.init_params(cfg, seed=0, head_variance=1.0, bias_std=0.1)
    p = O.to_torch(flat, torch.float64)
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.standard_normal((1, 8)))
    t = torch.tensor([0.3], dtype=torch.float64)
    feat = torch.zeros(1, 4, dtype=torch.long)
    _, div = O.vf_and_exact_div(p, cfg, x, t, feat)
    eps = torch.tensor(rng.standard_normal((1, 8)))
    _, h = O.vf_and_hutchinson_div(p, cfg, x, t, feat, eps)
    d = 1e-6
    jv = (O.egnn_apply(p, cfg, x + d * eps, t, feat) - O.egnn_apply(p, cfg, x - d * eps, t, feat)) / (2 * d)
    assert abs(float((jv * eps).sum()) - float(h)) < 1e-6 * (abs(float(h)) + 1)
    K = 3000
    xs, ts, fs = x.expand(K, 8), t.expand(K), feat.expand(K, 4)
    _, hs = O.vf_and_hutchinson_div(p, cfg, xs, ts, fs, torch.tensor(rng.standard_normal((K, 8))))
    assert abs(float(hs.mean()) - float(div)) < 4 * float(hs.std()) / np.sqrt(K)

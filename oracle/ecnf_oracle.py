"""CPU oracle for the ecnf hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module.  The product path (``ecnf_b200``) never does: it fails loudly when the CUDA library is missing.

PARITY: the reference (``/root/reference``) is pure Python/JAX and needs jax, flax, diffrax, distrax, e3nn_jax and
optax, none of which exist in this image, and its own tests pin no values (``ecnf/cnf/core_test.py:42-43`` asserts
nothing, ``ecnf/nets/egnn_test.py:31`` checks equivariance only).  This file is a *restatement* of the reference written
from its sources.  It is PINNED to vectors produced by the reference's own source files executed over minimal stand-ins
of their dependencies (``tests/golden/refshim``, ``tests/golden/make_refsrc_golden.py`` -> ``tests/golden/refsrc_*.npz``,
checked by ``tests/test_refsrc_golden.py``): parameter tree, vector field, exact and Hutchinson divergences, base
distribution, flow-matching loss, every gradient, update bookkeeping, solver call-site arguments.  It stays UNPINNED for
the diffrax solver (``dopri5`` below: restated from the published algorithm) and optax's Adam (``adam_step``), which are
validated by analytic known-answer tests only (``tests/test_oracle_kat.py``).

Everything is written in torch on the CPU with an explicit dtype (float32 = like-for-like with the
reference, float64 = ground truth) so that autograd can supply the reverse-mode Jacobian exactly as the
reference builds it (``ecnf/cnf/sample_and_log_prob.py:64-66``) and the flow-matching gradients
(``ecnf/cnf/gradient_step.py:31``).

Third-party arithmetic restated here (none vendored or pinned by the reference, ``requirements.txt:1-16``):
  * flax.linen.Dense / Embed                  -> ``x @ kernel + bias`` / row gather
  * e3nn_jax.scatter_sum                      -> ``index_add`` by receiver
  * distrax Transformed(ScalarAffine)         -> ``base_log_prob``
  * diffrax Dopri5 + PIDController + diffeqsolve (restated from the published algorithm, see ``dopri5``)
  * optax.adam + warmup_cosine_decay_schedule -> ``adam_step`` / ``warmup_cosine_lr``
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# --------------------------------------------------------------------------------------------------
# configuration + parameter pytree (flax naming)
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CnfConfig:
    """Arguments of ``build_cnf`` (``ecnf/cnf/build_cnf.py:34-44``)."""
    n_frames: int
    dim: int
    sigma_min: float = 0.01
    base_scale: float = 1.0
    n_blocks_egnn: int = 3
    mlp_units: Tuple[int, ...] = (128, 128, 128)
    n_invariant_feat_hidden: int = 64
    time_embedding_dim: int = 8
    n_features: int = 1
    normalization_constant: float = 1.0  # ecnf/nets/egnn.py:127

    @property
    def D(self) -> int:
        return self.n_frames * self.dim


CONFIGS = {
    # examples/config/{dw4,lj13,qm9,aldp}.yaml
    "dw4": CnfConfig(n_frames=4, dim=2, sigma_min=0.01, base_scale=1.0),
    "lj13": CnfConfig(n_frames=13, dim=3, sigma_min=0.01, base_scale=1.0),
    "qm9": CnfConfig(n_frames=19, dim=3, sigma_min=1e-6, base_scale=2.0, n_blocks_egnn=5,
                     mlp_units=(256, 256, 256, 256), n_invariant_feat_hidden=32),
    "aldp": CnfConfig(n_frames=22, dim=3, sigma_min=1e-6, base_scale=0.2, n_blocks_egnn=3,
                      mlp_units=(64, 64), n_invariant_feat_hidden=32, n_features=22),
}


def param_layout(cfg: CnfConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (path, shape) list of the flax pytree created by ``FlatEgnn`` (build_cnf.py:65-93,
    egnn.py:38-47,83-85,99,166-169,188).  The order is the flat-buffer order of the C library."""
    H, T, U = cfg.n_invariant_feat_hidden, cfg.time_embedding_dim, list(cfg.mlp_units)
    out: List[Tuple[str, Tuple[int, ...]]] = [("Embed_0/embedding", (cfg.n_features, H))]
    for b in range(cfg.n_blocks_egnn):
        out.append((f"EGNN_0/Dense_{b}/kernel", (H + T, H)))
        out.append((f"EGNN_0/Dense_{b}/bias", (H,)))
        for name, first_in, units in (("phi_e", 2 * H + 1, U), ("phi_x_torso", U[-1], U),
                                      ("phi_h", U[-1] + H, U + [H])):
            fan_in = first_in
            for l, u in enumerate(units):
                out.append((f"EGNN_0/{b}/{name}/Dense_{l}/kernel", (fan_in, u)))
                out.append((f"EGNN_0/{b}/{name}/Dense_{l}/bias", (u,)))
                fan_in = u
        out.append((f"EGNN_0/{b}/Dense_0/kernel", (U[-1], 1)))   # phi_x head, egnn.py:83-85
        out.append((f"EGNN_0/{b}/Dense_0/bias", (1,)))
        out.append((f"EGNN_0/{b}/Dense_1/kernel", (U[-1], 1)))   # attention logit, egnn.py:99
        out.append((f"EGNN_0/{b}/Dense_1/bias", (1,)))
    out.append(("EGNN_0/final_scaling", ()))
    return out


def init_params(cfg: CnfConfig, seed: int = 0, head_variance: float = 0.001,
                bias_std: float = 0.0) -> Dict[str, np.ndarray]:
    """Synthetic parameters with the reference's init statistics: Dense kernels
    N(0, 1/fan_in) (flax lecun_normal up to truncation), zero bias, Embed N(0, 1/H), phi_x head
    U(+-sqrt(3 v / fan_avg)) (egnn.py:84).  ``head_variance=1`` gives the 'stiffened' field used by the
    benchmarks; ``bias_std>0`` makes biases non-trivial for parity tests."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for path, shape in param_layout(cfg):
        leaf = path.split("/")[-1]
        if leaf == "final_scaling":
            v = np.ones((), np.float64)
        elif leaf == "embedding":
            v = rng.standard_normal(shape) / math.sqrt(shape[1])
        elif leaf == "bias":
            v = rng.standard_normal(shape) * bias_std
        elif path.endswith("/Dense_0/kernel") and path.count("/") == 3:  # EGNN_0/b/Dense_0/kernel
            lim = math.sqrt(3.0 * head_variance / (0.5 * (shape[0] + shape[1])))
            v = rng.uniform(-lim, lim, shape)
        else:
            v = rng.standard_normal(shape) / math.sqrt(shape[0])
        out[path] = np.asarray(v, np.float32)
    return out


def flat_to_nested(flat: Dict[str, np.ndarray]) -> dict:
    """{'a/b/c': v} -> {'params': {'a': {'b': {'c': v}}}} (the flax variable dict)."""
    root: dict = {}
    for path, v in flat.items():
        d = root
        parts = path.split("/")
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return {"params": root}


def nested_to_flat(tree: dict) -> Dict[str, np.ndarray]:
    tree = tree.get("params", tree)
    out: Dict[str, np.ndarray] = {}

    def rec(prefix, d):
        for k, v in d.items():
            p = f"{prefix}/{k}" if prefix else str(k)
            if isinstance(v, dict):
                rec(p, v)
            else:
                out[p] = v
    rec("", tree)
    return out


def to_torch(flat: Dict[str, np.ndarray], dtype=torch.float32, requires_grad=False) -> Dict[str, torch.Tensor]:
    return {k: torch.tensor(np.asarray(v), dtype=dtype).requires_grad_(requires_grad) for k, v in flat.items()}


# --------------------------------------------------------------------------------------------------
# vector field: FlatEgnn -> EGNN -> EGCL
# --------------------------------------------------------------------------------------------------
def fully_connected_edges(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """ecnf/utils/graph.py:6-14: receiver-major, sender (i+1+j) % n."""
    recv, send = [], []
    for i in range(n):
        for j in range(n - 1):
            recv.append(i)
            send.append((i + 1 + j) % n)
    return np.asarray(send, np.int64), np.asarray(recv, np.int64)


def timestep_frequencies(T: int) -> np.ndarray:
    """fp32 frequency table exactly as jnp evaluates it with x64 off (build_cnf.py:25-27)."""
    half = T // 2
    emb = np.float32(np.log(np.float32(10_000.0)) / np.float32(half - 1))
    return np.exp(np.arange(half, dtype=np.float32) * -emb).astype(np.float32)


def timestep_embedding(t: torch.Tensor, T: int) -> torch.Tensor:
    """ecnf/cnf/build_cnf.py:18-32.  ``t`` [B] -> [B, T] = [sin | cos]."""
    freqs = torch.tensor(timestep_frequencies(T)).to(t.dtype)
    arg = (t * 1000)[:, None] * freqs[None, :]
    return torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)


def _silu(x):
    return x * torch.sigmoid(x)


def _mlp(p, prefix: str, n_layers: int, x, activate_final: bool):
    """ecnf/nets/mlp.py:7-19."""
    for l in range(n_layers):
        x = x @ p[f"{prefix}/Dense_{l}/kernel"] + p[f"{prefix}/Dense_{l}/bias"]
        if l < n_layers - 1 or activate_final:
            x = _silu(x)
    return x


def egnn_apply(p: Dict[str, torch.Tensor], cfg: CnfConfig, x: torch.Tensor, t: torch.Tensor,
               feat: torch.Tensor) -> torch.Tensor:
    """``cnf.apply(params, x[B, n*dim], t[B], features[B, n]) -> [B, n*dim]``.

    build_cnf.py:68-93 (reshape, Embed, time embedding), egnn.py:144-190 (torso), egnn.py:49-114 (EGCL).
    """
    n, dim, H = cfg.n_frames, cfg.dim, cfg.n_invariant_feat_hidden
    L = len(cfg.mlp_units)
    B = x.shape[0]
    send, recv = fully_connected_edges(n)
    send_t, recv_t = torch.as_tensor(send), torch.as_tensor(recv)
    pos = x.reshape(B, n, dim)
    h = p["Embed_0/embedding"][feat.reshape(B, n).long()]                    # build_cnf.py:79-80
    tau = timestep_embedding(t, cfg.time_embedding_dim)                     # build_cnf.py:83
    mean = pos.mean(dim=1, keepdim=True)
    vec = pos - mean                                                        # egnn.py:160
    vec0 = vec
    for b in range(cfg.n_blocks_egnn):
        h = torch.cat([h, tau[:, None, :].expand(B, n, tau.shape[1])], dim=2)   # egnn.py:166
        h = h @ p[f"EGNN_0/Dense_{b}/kernel"] + p[f"EGNN_0/Dense_{b}/bias"]      # egnn.py:167
        pre = f"EGNN_0/{b}"
        v = vec[:, recv_t] - vec[:, send_t]                                 # egnn.py:73
        s = (v * v).sum(dim=-1, keepdim=True)
        length = torch.where(s == 0, torch.ones_like(s), s) ** 0.5          # numerical.py:7-10
        edge_in = torch.cat([h[:, send_t], h[:, recv_t], length ** 2], dim=-1)  # egnn.py:76
        m = _mlp(p, f"{pre}/phi_e", L, edge_in, True)                       # egnn.py:79
        px = _mlp(p, f"{pre}/phi_x_torso", L, m, True)                      # egnn.py:82
        px = px @ p[f"{pre}/Dense_0/kernel"] + p[f"{pre}/Dense_0/bias"]     # egnn.py:83-85
        shifts = px * v / (cfg.normalization_constant + length)            # egnn.py:87-91
        shift_i = torch.zeros_like(vec).index_add(1, recv_t, shifts) / (n - 1)  # egnn.py:92-95
        e = torch.sigmoid(m @ p[f"{pre}/Dense_1/kernel"] + p[f"{pre}/Dense_1/bias"])  # egnn.py:99-101
        m_i = torch.zeros(B, n, m.shape[-1], dtype=m.dtype).index_add(1, recv_t, m * e) / math.sqrt(n - 1)
        h_out = _mlp(p, f"{pre}/phi_h", L + 1, torch.cat([m_i, h], dim=-1), False)  # egnn.py:105-106
        h = h_out + h                                                       # egnn.py:110-111
        vec = vec + shift_i                                                 # egnn.py:112-113
    out = (vec - vec0 - mean) * p["EGNN_0/final_scaling"]                   # egnn.py:183-188
    return out.reshape(B, n * dim)


def vf_and_exact_div(p, cfg: CnfConfig, x: torch.Tensor, t: torch.Tensor, feat: torch.Tensor, batched: bool = False):
    """(f, tr df/dx) with the full Jacobian built in REVERSE mode from the D one-hot cotangents, like
    ``jax.vmap(vjp_fn)(jnp.eye(D))`` + ``jnp.trace`` (sample_and_log_prob.py:64-66).  Independent of the CUDA
    kernel's forward-mode formulation on purpose.  ``batched=True`` pushes the D cotangents through one
    vectorised backward pass (what vmap does); ``batched=False`` loops over them, which torch runs faster on
    the CPU for B >= 8 and is therefore the default (and what the CPU baseline times)."""
    xg = x.detach().clone().requires_grad_(True)
    f = egnn_apply(p, cfg, xg, t, feat)
    B, D = f.shape
    if batched:
        eye = torch.eye(D, dtype=f.dtype)[:, None, :].expand(D, B, D)
        (jac,) = torch.autograd.grad(f, xg, grad_outputs=eye, is_grads_batched=True)   # [D(out), B, D(in)]
        div = torch.diagonal(jac.permute(1, 0, 2), dim1=1, dim2=2).sum(dim=1)
        return f.detach(), div
    div = torch.zeros(B, dtype=x.dtype)
    for d in range(D):
        (g,) = torch.autograd.grad(f[:, d].sum(), xg, retain_graph=d + 1 < D)
        div = div + g[:, d]
    return f.detach(), div


def vf_and_hutchinson_div(p, cfg: CnfConfig, x: torch.Tensor, t: torch.Tensor, feat: torch.Tensor, eps: torch.Tensor):
    """(f, eps^T J eps): the Hutchinson estimate of tr df/dx with ONE fixed probe per trajectory, in reverse mode exactly
    like the reference's ``approx`` branch: ``(eps_dfdy,) = vjp_fn(eps); approx_div = sum(eps_dfdy * eps)``
    (sample_and_log_prob.py:69-78, :123-133).  The CUDA kernel computes the same scalar in forward mode (eps . J eps)."""
    xg = x.detach().clone().requires_grad_(True)
    f = egnn_apply(p, cfg, xg, t, feat)
    (eps_dfdy,) = torch.autograd.grad(f, xg, grad_outputs=eps.to(f.dtype))
    return f.detach(), (eps_dfdy * eps.to(f.dtype)).sum(dim=1)


# --------------------------------------------------------------------------------------------------
# base distribution, OT path, flow-matching loss, optimiser, ESS
# --------------------------------------------------------------------------------------------------
def remove_mean(x: torch.Tensor, n: int, dim: int) -> torch.Tensor:
    """zero_com_base.py:59-62 on the flat layout."""
    x3 = x.reshape(-1, n, dim)
    return (x3 - x3.mean(dim=1, keepdim=True)).reshape(x.shape)


def base_sample_from_noise(cfg: CnfConfig, eps: torch.Tensor) -> torch.Tensor:
    """``x0 = s * remove_mean(eps)`` (zero_com_base.py:88-93, build_cnf.py:46-61)."""
    return cfg.base_scale * remove_mean(eps, cfg.n_frames, cfg.dim)


def base_log_prob(cfg: CnfConfig, x: torch.Tensor) -> torch.Tensor:
    """distrax.Transformed log-prob: zero_com_base.py:44-47,64-84 + ildj scaled by (n-1)/n
    (build_cnf.py:53-54)."""
    n, dim, s = cfg.n_frames, cfg.dim, cfg.base_scale
    z = remove_mean(x / s, n, dim)
    dof = (n - 1) * dim
    return -0.5 * (z * z).sum(dim=-1) - 0.5 * dof * math.log(2 * math.pi) - dof * math.log(s)


def ot_conditional_vf(x0, x1, t, sigma_min: float):
    """ecnf/cnf/core.py:35-39 (t broadcast over the event axis as jax.vmap does in loss.py:25)."""
    t = t[:, None]
    return (1 - (1 - sigma_min) * t) * x0 + t * x1, x1 - (1 - sigma_min) * x0


def fm_loss(p, cfg: CnfConfig, x_data, x0, t, feat):
    """ecnf/cnf/loss.py:21-29 with the noise (x0, t) injected instead of drawn from a jax key."""
    x_t, u_t = ot_conditional_vf(x0, x_data, t, cfg.sigma_min)
    v = egnn_apply(p, cfg, x_t, t, feat)
    return ((v - u_t) ** 2).mean()


def fm_loss_and_grad(flat_params: Dict[str, np.ndarray], cfg: CnfConfig, x_data, x0, t, feat,
                     dtype=torch.float32):
    p = to_torch(flat_params, dtype, requires_grad=True)
    loss = fm_loss(p, cfg, x_data.to(dtype), x0.to(dtype), t.to(dtype), feat)
    names = [k for k, _ in param_layout(cfg)]
    # the last block's phi_h / attention never reach the output (egnn.py:180-190): jax.grad gives zeros there
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(p[k])) for k, g in zip(names, grads)}


def warmup_cosine_lr(step: int, init_value: float, peak_value: float, warmup_steps: int,
                     decay_steps: int, end_value: float = 0.0) -> float:
    """optax.warmup_cosine_decay_schedule (setup_training.py:100-106): linear init->peak over
    ``warmup_steps``, then cosine peak->end over ``decay_steps - warmup_steps``."""
    if step < warmup_steps:
        return init_value + (peak_value - init_value) * step / max(warmup_steps, 1)
    n = max(decay_steps - warmup_steps, 1)
    c = min(step - warmup_steps, n) / n
    cosine = 0.5 * (1 + math.cos(math.pi * c))
    return end_value + (peak_value - end_value) * cosine


def adam_step(theta, g, m, v, step: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam (scale_by_adam + scale(-lr)), ``step`` = count BEFORE this update (0-based)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mh = m / (1 - b1 ** (step + 1))
    vh = v / (1 - b2 ** (step + 1))
    upd = -lr * mh / (np.sqrt(vh) + eps)
    return theta + upd, m, v, upd


def reverse_ess(log_w: np.ndarray) -> float:
    """setup_training.py:182: 1 / sum(softmax(log_w)^2) / N."""
    w = np.exp(log_w - log_w.max())
    w = w / w.sum()
    return float(1.0 / np.sum(w * w) / log_w.shape[0])


def forward_ess(log_w: np.ndarray, mask: np.ndarray) -> float:
    """utils/evaluation.py:10-22."""
    from scipy.special import logsumexp
    lw = np.where(mask, log_w, 0.0)
    nm = np.log(mask.sum())
    log_z_inv = logsumexp(-lw, b=mask) - nm
    log_z = logsumexp(lw, b=mask) - nm
    return float(np.exp(-log_z_inv - log_z))


def lj_energy(x: np.ndarray, eps=1.0, tau=1.0, r=1.0, harmonic=0.5) -> np.ndarray:
    """targets/target_energy/leonard_jones.py:10-27, batched [B, n, dim] -> [B]."""
    n = x.shape[1]
    send, recv = fully_connected_edges(n)
    v = x[:, send] - x[:, recv]
    s = (v * v).sum(-1)
    d = np.sqrt(np.where(s == 0, 1.0, s))
    term = (r / d) ** 12 - 2 * (r / d) ** 6
    e = eps / (2 * tau) * term.sum(-1)
    com = x.mean(axis=1, keepdims=True)
    return e + harmonic * ((x - com) ** 2).sum(axis=(1, 2))


def dw_energy(x: np.ndarray, a=0.0, b=-4.0, c=0.9, d0=4.0, tau=1.0) -> np.ndarray:
    """targets/target_energy/double_well.py:9-19, batched."""
    n = x.shape[1]
    send, recv = fully_connected_edges(n)
    v = x[:, send] - x[:, recv]
    s = (v * v).sum(-1)
    d = np.sqrt(np.where(s == 0, 1.0, s)) - d0
    return (a * d + b * d ** 2 + c * d ** 4).sum(-1) / tau / 2


# --------------------------------------------------------------------------------------------------
# diffrax restated: Dopri5 + PIDController(I-only) + diffeqsolve, batched with per-trajectory control
# --------------------------------------------------------------------------------------------------
# Dormand-Prince 5(4) tableau.  b_err = b - b_hat with the embedded weights diffrax/torchdiffeq use
# (1951/21600, 0, 22642/50085, 451/720, -12231/42400, 649/6300, 1/60); the classic Hairer weights give an
# estimate exactly 1.5x larger.  diffrax is not on disk: this is a recollection, exposed as SolveControl.err_scale
# (and ecnf_solve_ctrl.err_scale in the C-ABI) so that a mismatch is a configuration change.
DP_C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
DP_A = (
    (),
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
DP_B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0)
DP_BERR = (35 / 384 - 1951 / 21600, 0.0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
           -2187 / 6784 + 12231 / 42400, 11 / 84 - 649 / 6300, -1 / 60)


@dataclass
class SolveControl:
    """diffeqsolve / PIDController attributes used by the reference
    (sample_and_log_prob.py:33-37,85-89,140-144) with diffrax defaults for the rest."""
    fixed: bool = False
    step_size: float = 0.05
    rtol: float = 1e-5
    atol: float = 1e-5
    dtmin: float = 1e-5
    max_steps: int = 4096
    safety: float = 0.9
    factormin: float = 0.2
    factormax: float = 10.0
    error_order: float = 5.0
    err_scale: float = 1.0   # multiplies the embedded error estimate (1.5 = classic Hairer-Wanner b_hat)


@dataclass
class SolveStats:
    n_steps: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    n_accepted: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    n_evals: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    status: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))  # 0 ok, 1 max_steps hit


def _rms(z: torch.Tensor) -> torch.Tensor:
    return torch.sqrt((z * z).mean(dim=1))


def dopri5(func: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], y0: torch.Tensor, t0: float, t1: float,
           ctrl: SolveControl) -> Tuple[torch.Tensor, SolveStats]:
    """Integrate ``dy/dt = func(t[B], y[B, S])`` from t0 to t1 (either direction) independently per row.

    Time reversal as diffrax does it: internal time tau = direction * t increases, the field is
    ``direction * func(direction * tau, y)``.  Per-trajectory (t, dt, accept) so each row follows the
    step sequence the un-vmapped reference solve would; finished rows are frozen.
    """
    B = y0.shape[0]
    dt_ = y0.dtype
    direction = 1.0 if t0 < t1 else -1.0
    T0, T1 = t0 * direction, t1 * direction
    tol_clip = 1e-10 if dt_ == torch.float64 else 1e-6
    n_evals = np.zeros(B, np.int64)

    def f(tau, y, mask=None):
        out = direction * func(direction * tau, y)
        n_evals[(np.ones(B, bool) if mask is None else mask.numpy())] += 1
        return out

    def clip_to_end(tprev, tnext, keep):
        clip = tnext > T1 - tol_clip
        tclip = torch.where(keep, torch.full_like(tnext, T1), tprev + 0.5 * (T1 - tprev))
        return torch.where(clip, tclip, tnext)

    tprev = torch.full((B,), T0, dtype=dt_)
    y = y0.clone()
    f0 = f(tprev, y)                                            # FSAL initialisation
    if ctrl.fixed:
        dt0 = torch.full((B,), abs(ctrl.step_size), dtype=dt_)
    else:
        # Hairer-Wanner initial step (diffrax _select_initial_step).  diffrax evaluates f(t0, y0) a
        # second time here; it is the identical computation (XLA CSEs it), so it is reused, not recounted.
        scale = ctrl.atol + y.abs() * ctrl.rtol
        d0, d1 = _rms(y / scale), _rms(f0 / scale)
        cond = (d0 < 1e-5) | (d1 < 1e-5)
        h0 = torch.where(cond, torch.full_like(d0, 1e-6), 0.01 * d0 / torch.where(cond, torch.ones_like(d1), d1))
        f1 = f(tprev + h0, y + h0[:, None] * f0)
        d2 = _rms((f1 - f0) / scale) / h0
        md = torch.maximum(d1, d2)
        h1 = torch.where(md <= 1e-15, torch.maximum(torch.full_like(h0, 1e-6), h0 * 1e-3),
                         (0.01 / md) ** (1.0 / ctrl.error_order))
        dt0 = torch.minimum(100 * h0, h1)
        dt0 = torch.clamp(dt0, min=ctrl.dtmin)
    tnext = clip_to_end(tprev, tprev + dt0, torch.ones(B, dtype=torch.bool))
    at_dtmin = torch.zeros(B, dtype=torch.bool)
    n_steps = np.zeros(B, np.int32)
    n_acc = np.zeros(B, np.int32)
    status = np.zeros(B, np.int32)
    active = tprev < T1
    while bool(active.any()):
        dt = tnext - tprev
        k = [f0 * dt[:, None]]
        for s in range(1, 7):
            ys = y.clone()
            for j, a in enumerate(DP_A[s]):
                if a != 0.0:
                    ys = ys + a * k[j]
            fs = f(tprev + DP_C[s] * dt, ys, active)
            k.append(fs * dt[:, None])
        y1 = ys                                                  # stage 7 input IS the 5th-order solution
        yerr = sum(c * kk for c, kk in zip(DP_BERR, k) if c != 0.0)
        if ctrl.fixed:
            keep = torch.ones(B, dtype=torch.bool)
            new_prev, new_next = tnext, tnext + dt
        else:
            sc = ctrl.atol + torch.maximum(y.abs(), y1.abs()) * ctrl.rtol
            err = _rms(ctrl.err_scale * yerr / sc)
            keep = (err < 1) | at_dtmin
            inv = torch.where(err == 0, torch.full_like(err, float("inf")), 1.0 / err)
            factor = ctrl.safety * inv ** (1.0 / ctrl.error_order)
            fmin = torch.where(keep, torch.ones_like(err), torch.full_like(err, ctrl.factormin))
            factor = torch.minimum(torch.maximum(factor, fmin), torch.full_like(err, ctrl.factormax))
            ndt = dt * factor
            new_at = ndt <= ctrl.dtmin
            ndt = torch.clamp(ndt, min=ctrl.dtmin)
            new_prev = torch.where(keep, tnext, tprev)
            new_next = new_prev + ndt
            at_dtmin = torch.where(active, new_at, at_dtmin)
        new_prev = torch.minimum(new_prev, torch.full_like(new_prev, T1))
        new_next = clip_to_end(new_prev, new_next, keep)
        upd = active & keep
        y = torch.where(upd[:, None], y1, y)
        f0 = torch.where(upd[:, None], fs, f0)                   # FSAL: k7 / dt of an accepted step
        tprev = torch.where(active, new_prev, tprev)
        tnext = torch.where(active, new_next, tnext)
        n_steps += active.numpy().astype(np.int32)
        n_acc += upd.numpy().astype(np.int32)
        hit = active.numpy() & (n_steps >= ctrl.max_steps) & (tprev < T1).numpy()
        status[hit] = 1
        active = active & (tprev < T1) & torch.as_tensor(n_steps < ctrl.max_steps)
    return y, SolveStats(n_steps, n_acc, n_evals.astype(np.int32), status)


# --------------------------------------------------------------------------------------------------
# sample / log-prob drivers (sample_and_log_prob.py), batched, noise injected
# --------------------------------------------------------------------------------------------------
def sample_cnf(p, cfg: CnfConfig, x0: torch.Tensor, feat: torch.Tensor, ctrl: SolveControl):
    """sample_and_log_prob.py:11-38 given the base draw ``x0`` (no divergence)."""
    def func(t, y):
        with torch.no_grad():
            return egnn_apply(p, cfg, y, t, feat)
    return dopri5(func, x0, 0.0, 1.0, ctrl)


def _joint(p, cfg, feat, eps=None):
    """exact trace (eps is None) or the Hutchinson estimate with the fixed per-trajectory probe ``eps`` [B, D]"""
    def func(t, y):
        if eps is None:
            fx, div = vf_and_exact_div(p, cfg, y[:, :-1], t, feat)
        else:
            fx, div = vf_and_hutchinson_div(p, cfg, y[:, :-1], t, feat, eps)
        return torch.cat([fx, div[:, None]], dim=1)
    return func


def sample_and_log_prob_cnf(p, cfg: CnfConfig, x0: torch.Tensor, feat: torch.Tensor, ctrl: SolveControl, eps=None):
    """sample_and_log_prob.py:97-149; ``eps`` = None: exact branch, else the ``approx`` (Hutchinson) branch with that
    probe.  Fixed-step uses the evident intent y0=(x0, 0) (the reference passes y0=x0 there and cannot run, SURVEY
    Appendix C#2)."""
    y0 = torch.cat([x0, torch.zeros(x0.shape[0], 1, dtype=x0.dtype)], dim=1)
    y1, st = dopri5(_joint(p, cfg, feat, eps), y0, 0.0, 1.0, ctrl)
    log_q = base_log_prob(cfg, x0) - y1[:, -1]
    return y1[:, :-1], log_q, st


def get_log_prob(p, cfg: CnfConfig, x: torch.Tensor, feat: torch.Tensor, ctrl: SolveControl, eps=None):
    """sample_and_log_prob.py:41-94 (``eps``: see sample_and_log_prob_cnf): t 1 -> 0; returns (log_p, log_prob_base, delta)."""
    y0 = torch.cat([x, torch.zeros(x.shape[0], 1, dtype=x.dtype)], dim=1)
    y1, st = dopri5(_joint(p, cfg, feat, eps), y0, 1.0, 0.0, ctrl)
    lpb = base_log_prob(cfg, y1[:, :-1])
    delta = y1[:, -1]
    return lpb + delta, lpb, delta, st

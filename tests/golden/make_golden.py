"""Generates tests/golden/*.npz from the CPU oracle in float64.

The reference itself cannot run here (no jax/flax/diffrax in the image), so these vectors pin the
*restatement*, not the reference: they freeze today's oracle outputs so that neither the oracle nor the CUDA path can
drift silently.  Re-run with `python tests/golden/make_golden.py` only when the oracle is deliberately changed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ecnf_oracle as O  # noqa: E402

CASES = {
    # name: (n, dim, blocks, units, H, n_features)
    "dw4": (4, 2, 3, (128, 128, 128), 64, 1),
    "small_64_32": (5, 3, 2, (64, 64), 32, 3),
}


def main():
    for name, (n, dim, blocks, units, H, nfeat) in CASES.items():
        cfg = O.CnfConfig(n_frames=n, dim=dim, n_blocks_egnn=blocks, mlp_units=units, n_invariant_feat_hidden=H,
                          n_features=nfeat)
        flat = O.init_params(cfg, seed=42, head_variance=1.0, bias_std=0.1)
        rng = np.random.default_rng(7)
        B = 4
        x = rng.standard_normal((B, n * dim))
        t = rng.uniform(0, 1, B)
        feat = rng.integers(0, nfeat, (B, n))
        p64 = O.to_torch(flat, torch.float64)
        f, div = O.vf_and_exact_div(p64, cfg, torch.tensor(x), torch.tensor(t), torch.tensor(feat))
        x0 = O.base_sample_from_noise(cfg, torch.tensor(rng.standard_normal((B, n * dim))))
        x1, logq, st = O.sample_and_log_prob_cnf(p64, cfg, x0, torch.tensor(feat), O.SolveControl(fixed=True))
        x_data = O.remove_mean(torch.tensor(rng.standard_normal((B, n * dim))), n, dim)
        loss, grads = O.fm_loss_and_grad(flat, cfg, x_data, x0, torch.tensor(t), torch.tensor(feat), dtype=torch.float64)
        out = dict(x=x, t=t, feat=feat, f=f.numpy(), div=div.numpy(), x0=x0.numpy(), x1=x1.numpy(), logq=logq.numpy(),
                   x_data=x_data.numpy(), loss=np.asarray(float(loss)),
                   grad_phi_e=grads["EGNN_0/0/phi_e/Dense_1/kernel"].numpy(),
                   grad_head=grads["EGNN_0/0/Dense_0/kernel"].numpy(),
                   grad_embed=grads["Embed_0/embedding"].numpy(),
                   base_logp=O.base_log_prob(cfg, x0).numpy())
        np.savez_compressed(os.path.join(os.path.dirname(__file__), f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

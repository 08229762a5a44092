"""World-size-2 gloo tests (CPU) of the multi-process host logic: sharding by global sample index, the gradient
all-reduce hook and the ESS-statistics all-gather give the single-process answer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ecnf_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ecnf_b200.distributed import gather_log_weights, make_grad_allreduce, merge_ess_stats, shard_range
    torch.set_num_threads(1)
    cfg = O.CnfConfig(n_frames=4, dim=2, n_blocks_egnn=2, mlp_units=(16, 16), n_invariant_feat_hidden=16)
    flat = O.init_params(cfg, 0, head_variance=1.0, bias_std=0.1)
    rng = np.random.default_rng(5)
    Bg = 6
    x_data = torch.tensor(rng.standard_normal((Bg, cfg.D)))
    x0 = torch.tensor(rng.standard_normal((Bg, cfg.D)))
    t = torch.tensor(rng.uniform(0, 1, Bg))
    feat = torch.zeros(Bg, cfg.n_frames, dtype=torch.long)
    lo, hi = shard_range(Bg, rank, world)
    # each rank: sum over its rows / (global_B * D)  ==  local mean * (local_B / global_B)
    loss, grads = O.fm_loss_and_grad(flat, cfg, x_data[lo:hi], x0[lo:hi], t[lo:hi], feat[lo:hi], dtype=torch.float64)
    w = (hi - lo) / Bg
    flat_grad = torch.cat([g.reshape(-1) for g in grads.values()]) * w
    loss = make_grad_allreduce(world)(flat_grad, (loss * w).reshape(1))
    lw_all = torch.tensor(rng.standard_normal(8) * 2)
    lw = lw_all[rank * 4:(rank + 1) * 4]
    mx, nmx = lw.max(), (-lw).max()
    st = torch.stack([mx, torch.exp(lw - mx).sum(), torch.exp(2 * (lw - mx)).sum(), nmx, torch.exp(-lw - nmx).sum()])
    merged = merge_ess_stats(st)
    gathered = gather_log_weights(lw)
    if rank == 0:
        out["loss"], out["grad"], out["ess"], out["lw"] = loss.item(), flat_grad.numpy(), merged.numpy(), gathered.numpy()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_ess_merge():
    from ecnf_b200.engine import ess_from_stats
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    cfg = O.CnfConfig(n_frames=4, dim=2, n_blocks_egnn=2, mlp_units=(16, 16), n_invariant_feat_hidden=16)
    flat = O.init_params(cfg, 0, head_variance=1.0, bias_std=0.1)
    rng = np.random.default_rng(5)
    x_data = torch.tensor(rng.standard_normal((6, cfg.D)))
    x0 = torch.tensor(rng.standard_normal((6, cfg.D)))
    t = torch.tensor(rng.uniform(0, 1, 6))
    feat = torch.zeros(6, cfg.n_frames, dtype=torch.long)
    loss, grads = O.fm_loss_and_grad(flat, cfg, x_data, x0, t, feat, dtype=torch.float64)
    ref = torch.cat([g.reshape(-1) for g in grads.values()]).numpy()
    assert abs(out["loss"] - float(loss)) < 1e-12
    assert np.abs(out["grad"] - ref).max() < 1e-12 * (np.abs(ref).max() + 1)
    lw_all = rng.standard_normal(8) * 2
    assert np.array_equal(out["lw"], lw_all)
    rv, fw = ess_from_stats(out["ess"].tolist(), 8)
    assert abs(rv - O.reverse_ess(lw_all)) < 1e-12 and abs(fw - O.forward_ess(lw_all, np.ones(8, bool))) < 1e-12

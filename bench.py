#!/usr/bin/env python
"""Benchmark of the ecnf hot path on B200 (contract in the task statement; metric from BASELINE.json).

Headline workload (BASELINE.json configs[1]): LJ13 (13 particles, 3-D) EGNN CNF, `sample_and_log_prob_cnf` with the
exact divergence, Dopri5, batch 10 000 per GPU, followed by the LJ target log-density, log-weights and the ESS
sufficient statistics (setup_training.py:166-185).  One "step" = one such batch.  The solver runs the reference's
fixed-step branch (use_fixed_step_size=True, dt=0.05 -> exactly 121 vector-field evaluations per trajectory) so the
work per sample is deterministic and the roofline numerator is exact; parameters are synthetic ('stiffened' init,
SURVEY 8(d)) because no trained checkpoint exists offline.

Also measured in the same run (reported under "extra"): the QM9-positional flow-matching training step (batch 512,
BASELINE.json configs[2]) in steps/s.

`--impl reference` times the CPU restatement of the reference (oracle/, torch fp32, reverse-mode Jacobian exactly
like sample_and_log_prob.py:64-66) on the host cores: JAX is not installable in this image, so the reference itself
cannot run ("reference arm").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LJ13 = dict(n_frames=13, dim=3, sigma_min=0.01, base_scale=1.0, n_blocks_egnn=3, mlp_units=(128, 128, 128),
            n_invariant_feat_hidden=64, time_embedding_dim=8, n_features=1)
QM9 = dict(n_frames=19, dim=3, sigma_min=1e-6, base_scale=2.0, n_blocks_egnn=5, mlp_units=(256, 256, 256, 256),
           n_invariant_feat_hidden=32, time_embedding_dim=8, n_features=1)
N_EVALS_FIXED = 121          # 1 FSAL init + 6 stages x 20 steps (dt = 0.05)
# one ncu --set full capture of ecnf_solve_tc_kernel (profiles/r1_solve_tc_full.txt): dram read + write bytes per trajectory
NCU_DRAM_BYTES_PER_TRAJ = (19.957129e9 + 43.069807e9) / 148   # ncu --set full, 148 trajectories (profiles/r1_solve_tc_full.txt)
METRIC = "LJ13 samples/s with exact log-q (Dopri5)"
UNIT = "samples/s"


def fwd_flops(c) -> float:
    """2 * blocks * (E * M_e + n * M_n), dense matmuls exactly as the reference's op graph."""
    n, H, T, U, L = c["n_frames"], c["n_invariant_feat_hidden"], c["time_embedding_dim"], c["mlp_units"][0], len(c["mlp_units"])
    E = n * (n - 1)
    m_e = (2 * H + 1) * U + (L - 1) * U * U + L * U * U + 2 * U
    m_n = (H + T) * H + (U + H) * U + (L - 1) * U * U + U * H
    return 2.0 * c["n_blocks_egnn"] * (E * m_e + n * m_n)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        # the busiest half of the samples = "under load"
        sm_sorted = sorted(sm)
        return {"sm_mhz": statistics.median(sm_sorted[len(sm_sorted) // 2:]) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------
def cpu_sample_logq(flat_params: dict, n_traj: int, threads: int, seed: int = 2):
    """Time `n_traj` LJ13 trajectories of sample_and_log_prob (exact, fixed dt=0.05) on the host cores."""
    from oracle import ecnf_oracle as O
    torch.set_num_threads(threads)
    torch.set_flush_denormal(True)
    ocfg = O.CnfConfig(**LJ13)
    p = O.to_torch(flat_params, torch.float32)
    rng = np.random.default_rng(seed)
    x0 = O.base_sample_from_noise(ocfg, torch.tensor(rng.standard_normal((n_traj, ocfg.D)), dtype=torch.float32))
    feat = torch.zeros(n_traj, ocfg.n_frames, dtype=torch.long)
    t0 = time.perf_counter()
    x1, logq, st = O.sample_and_log_prob_cnf(p, ocfg, x0, feat, O.SolveControl(fixed=True, step_size=0.05))
    lw = -O.lj_energy(x1.numpy().reshape(n_traj, 13, 3).astype(np.float64)) - logq.numpy()
    O.reverse_ess(lw)
    dt = time.perf_counter() - t0
    assert int(st.n_evals[0]) == N_EVALS_FIXED
    return n_traj / dt, dt


def synthetic_params_numpy(cfg: dict, seed: int = 0, head_variance: float = 1.0) -> dict:
    """Same synthetic parameters as the GPU arm, as a {flax path: array} dict, without touching CUDA."""
    from ecnf_b200.engine import CnfConfig, Engine
    from ecnf_b200.nets.egnn import init_flat_params
    eng = Engine(CnfConfig(**cfg))
    flat = init_flat_params(eng, seed, head_variance)
    out = {}
    for path, off, shape in eng.layout:
        cnt = int(np.prod(shape)) if shape else 1
        out[path] = flat[off:off + cnt].reshape(shape).copy()
    return out, eng, flat


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    params, _, _ = synthetic_params_numpy(LJ13)
    budget = 200.0 / max(1, args.steps + args.warmup)        # seconds per step
    # calibrate: one small solve tells the per-trajectory cost
    _, t2 = cpu_sample_logq(params, 2, threads)
    n_traj = int(max(1, min(64, (budget / (t2 / 2)) * 0.7)))
    for _ in range(args.warmup):
        cpu_sample_logq(params, n_traj, threads)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_sample_logq(params, n_traj, threads)
        times.append(dt)
    value = n_traj * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "lj13_sample_and_log_prob_exact_dopri5_fixed_dt0.05+lj_log_weights+ess", "batch_per_step": n_traj,
                   "n_evals_per_sample": N_EVALS_FIXED},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_traj} LJ13 trajectories per step (121 evals each), torch-CPU fp32 restatement "
                                   "of the reference with reverse-mode Jacobian; the JAX reference cannot be installed "
                                   "in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=10_000, help="trajectories per GPU per step")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary QM9 training-step measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the bounded CPU baseline")
    ap.add_argument("--adaptive", action="store_true", help="PID-controlled steps (rtol=atol=1e-5) instead of dt=0.05")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from ecnf_b200 import lib as L
    from ecnf_b200.cnf import build_cnf
    from ecnf_b200.engine import PackedParams, ess_from_stats
    from ecnf_b200.nets.egnn import init_flat_params
    from ecnf_b200.distributed import merge_ess_stats

    cnf = build_cnf(**LJ13)
    eng = cnf.engine
    flat_host = init_flat_params(eng, 0, head_variance=1.0)
    params = PackedParams(torch.from_numpy(flat_host).to(dev))
    B = args.batch
    goff = rank * B                                      # noise keyed by GLOBAL sample index
    feat = torch.zeros(B, 13, dtype=torch.int32, device=dev)
    ctrl = L.make_ctrl(use_fixed_step_size=not args.adaptive)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    kernel_events = []

    def step_resident(x0, timed=False):
        if timed:
            ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ka.record()
        x1, logs, stats = eng.solve(params, L.MODE_SAMPLE_LOGQ, x0, feat, ctrl)
        if timed:
            kb.record()
            kernel_events.append((ka, kb))
        log_w = eng.target_log_prob(L.TARGET_LJ, x1) - logs[:, 0]
        st = eng.ess_stats(log_w)
        if world > 1:
            st = merge_ess_stats(st)
        return x1, logs, stats, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    x0_res = eng.base_sample(2, B, goff)
    for _ in range(args.warmup):
        step_resident(x0_res)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    n_evals_total = 0
    for i in range(args.steps):
        flush.zero_()
        barrier()
        ev[i][0].record()
        x1, logs, stats, st = step_resident(x0_res, timed=True)
        ev[i][1].record()
        barrier()
        n_evals_total += int(stats[:, 2].sum().item())
    t_steps = [a.elapsed_time(b) for a, b in ev]                       # ms, per step
    t_total = torch.tensor([sum(t_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
    value = world * B * args.steps / (t_total.item() * 1e-3)
    ms_per_step = t_total.item() / args.steps

    # ---- dominant kernel (the persistent solve kernel), CUDA events on its launch stream inside the timed steps
    k_times = [a.elapsed_time(b) for a, b in kernel_events]
    kstats = stats
    clocks = sampler.stop()
    k_ms = sum(k_times) / len(k_times)
    evals_per_launch = int(kstats[:, 2].sum().item())
    f_fwd = fwd_flops(LJ13)
    alg_flops = evals_per_launch * (1 + 39) * f_fwd                    # (1 + D) * F_fwd per eval
    pk, pk_kind = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])      # the kernel runs for seconds: sustained figure
    achieved_tf = alg_flops / (k_ms * 1e-3) / 1e12

    # ---- end to end through the public API with host buffers (H2D of the noise + features, D2H of the results)
    from ecnf_b200.cnf import sample_and_log_prob_cnf
    rng = np.random.default_rng(1000 + rank)
    eps_host = torch.from_numpy(rng.standard_normal((B, 39)).astype(np.float32)).pin_memory()
    feat_host = torch.zeros(B, 13, dtype=torch.int32).pin_memory()
    out_x = torch.empty(B, 39, dtype=torch.float32).pin_memory()
    out_lq = torch.empty(B, dtype=torch.float32).pin_memory()
    out_ess = torch.empty(5, dtype=torch.float32).pin_memory()

    def step_e2e():
        eps = eps_host.to(dev, non_blocking=True)
        f = feat_host.to(dev, non_blocking=True)
        x0 = eng.base_sample_from_noise(eps)
        x1, log_q = sample_and_log_prob_cnf(cnf, params, None, f, use_fixed_step_size=not args.adaptive, x0=x0)
        log_w = eng.target_log_prob(L.TARGET_LJ, x1) - log_q
        st = eng.ess_stats(log_w)
        if world > 1:
            st = merge_ess_stats(st)
        out_x.copy_(x1, non_blocking=True)
        out_lq.copy_(log_q, non_blocking=True)
        out_ess.copy_(st, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(1, min(args.steps, 2))
    barrier()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        step_e2e()
    b.record()
    barrier()
    e2e_ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (e2e_ms.item() * 1e-3)
    rv_ess, fw_ess = ess_from_stats(out_ess.tolist(), world * B)

    # ---- secondary: QM9-positional flow-matching training step, batch 512 per GPU (weak scaling)
    extra = {}
    launches = args.steps * 3 + e2e_steps * 4
    if not args.no_train:
        # ---- secondary: plain sampling without a divergence (sample_cnf; BASELINE.json configs[4], the reference's
        #      load_checkpoint_measure_sampling_time.py path), LJ13, 4736 trajectories per GPU, device-resident noise
        Bs = min(B, 148 * 32)
        eng.solve(params, L.MODE_SAMPLE, x0_res[:Bs], feat[:Bs], ctrl)
        barrier()
        sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sa.record()
        eng.solve(params, L.MODE_SAMPLE, x0_res[:Bs], feat[:Bs], ctrl)
        sb.record()
        barrier()
        s_ms = torch.tensor([sa.elapsed_time(sb)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(s_ms, op=dist.ReduceOp.MAX)
        extra["sample_only"] = {"metric": "LJ13 sample_cnf samples/s (no divergence, Dopri5 " + ("adaptive" if args.adaptive else "dt=0.05") + ")",
                                "value": world * Bs / (s_ms.item() * 1e-3), "unit": "samples/s", "batch_per_gpu": Bs,
                                "ms": s_ms.item()}
        launches += 2 * 3
        extra["fm_train"] = bench_train(args, dev, rank, world)
        launches += extra["fm_train"].pop("_launches")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        p_np, _, _ = synthetic_params_numpy(LJ13)
        threads = os.cpu_count() or 1
        v, dt = cpu_sample_logq(p_np, 4, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"4 LJ13 trajectories x 121 evals (fixed dt=0.05) in {dt:.1f} s, torch-CPU fp32 restatement of "
                         "the reference (reverse-mode Jacobian); JAX reference not installable here"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "lj13_sample_and_log_prob_exact_dopri5_" + ("adaptive_rtol1e-5" if args.adaptive else "fixed_dt0.05")
                                   + "+lj_log_weights+ess", "batch_per_gpu": B, "global_batch": world * B,
                       "n_evals_per_sample": evals_per_launch / B, "params": "synthetic stiffened init (seed 0)",
                       "l2": "256 MiB flush buffer written between timed steps", "parallelism": f"dp{world} (independent trajectories; ESS statistics all-gathered)"},
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf, "traffic": NCU_DRAM_BYTES_PER_TRAJ * B if NCU_DRAM_BYTES_PER_TRAJ else None,
                         "kernel": "ecnf_solve_tc_kernel (tcgen05 + TMEM, 3-pass bf16 split, fp32 accumulate)",
                         "kernel_ms": k_ms, "algorithmic_flops_per_launch": alg_flops,
                         "executed_tensor_flops_per_launch": evals_per_launch * int(eng.lib.ecnf_solve_tensor_flops_per_eval(eng.handle)),
                         "peak_source": f"{pk_kind} bf16 dense sustained (MEASURED_PEAKS.json); numerator = algorithmic "
                                        "(1+D)*F_fwd per evaluation (SURVEY 8(d)); the kernel executes 3 bf16 passes over "
                                        "a structurally reduced tangent set, see executed_tensor_flops_per_launch",
                         "traffic_source": "profiles/r1_solve_tc_full.txt: dram bytes of one ncu --set full capture, per trajectory x batch"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 39 * 4 + B * 13 * 4,
                    "d2h_bytes_per_step": B * 39 * 4 + B * 4 + 20},
            "gpu_launches": launches,
            "clocks": clocks,
            "extra": {**extra, "reverse_ess": rv_ess, "forward_ess": fw_ess, "step_ms": t_steps},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_train(args, dev, rank, world):
    """QM9-positional FM training step (loss + grad + all-reduce + Adam/EMA), batch 512 per GPU."""
    import torch.distributed as dist
    from ecnf_b200.cnf import build_cnf, flow_matching_update_fn, TrainingState
    from ecnf_b200.engine import PackedParams
    from ecnf_b200.nets.egnn import init_flat_params
    from ecnf_b200.utils.optim import Adam, warmup_cosine_decay_schedule
    from ecnf_b200.distributed import make_grad_allreduce
    cnf = build_cnf(**QM9)
    eng = cnf.engine
    B = 512
    params = PackedParams(torch.from_numpy(init_flat_params(eng, 0)).to(dev))
    opt = Adam(warmup_cosine_decay_schedule(1e-4, 1e-4, 10, 100_000, 0.0))
    state = TrainingState(params=params, opt_state=opt.init(params), key=rank, ema_params=params)
    rng = np.random.default_rng(3 + rank)
    x = rng.standard_normal((B, 19, 3)).astype(np.float32) * 1.5
    x = (x - x.mean(axis=1, keepdims=True)).reshape(B, 57)
    x_host = torch.from_numpy(x).pin_memory()
    feat_host = torch.zeros(B, 19, dtype=torch.int32).pin_memory()
    hook = make_grad_allreduce(world) if world > 1 else None
    denom = float(world * B * 57)

    def step(st):
        xd = x_host.to(dev, non_blocking=True)
        fd = feat_host.to(dev, non_blocking=True)
        st, info = flow_matching_update_fn(cnf, opt.update, st, xd, fd, grad_allreduce=hook,
                                           global_offset=rank * B, loss_denominator=denom)
        return st, info

    for _ in range(3):
        state, info = step(state)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = 10
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(K):
        state, info = step(state)
    loss = float(info["loss"])          # D2H read of the step's metric
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / K
    flops = 3.0 * fwd_flops(QM9) * B                     # per GPU (weak scaling: every rank steps its own 512 graphs)
    pk, _ = peaks()
    tf = flops / (ms_step * 1e-3) / 1e12
    return {"metric": "QM9-positional FM train steps/s (batch 512 per GPU, loss+grad+Adam+EMA, H2D of the batch inside)",
            "value": 1e3 / ms_step, "unit": "steps/s", "ms_per_step": ms_step, "loss": loss,
            "global_batch": B * world, "graphs_per_s": B * world * 1e3 / ms_step,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                         "unit": "TFLOP/s", "frac": tf / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                         "algorithmic_flops_per_step": flops, "per": "GPU"},
            "_launches": 13 * 360}


if __name__ == "__main__":
    main()

"""Host logic of bench.py: the algorithmic-flop figures are SURVEY.md 8(d)'s, and the reference arm prints a contract line."""
import json
import os
import subprocess
import sys

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_flops_match_the_survey_table():
    # F_fwd per sample and evaluation: LJ13 97.843 MFLOP, QM9-positional 1.6808 GFLOP
    assert abs(bench.fwd_flops(bench.LJ13) - 97.8432e6) < 1e3
    assert abs(bench.fwd_flops(bench.QM9) - 1.6807552e9) < 1e3
    # exact-divergence evaluation = (1 + D) F_fwd = 3.9137 GFLOP; fixed dt = 0.05 -> 6 * 20 + 1 evaluations
    assert abs((1 + 39) * bench.fwd_flops(bench.LJ13) - 3.9137e9) < 1e6 and bench.N_EVALS_FIXED == 121
    # flow-matching step = 3 F_fwd per graph: 2.582 TFLOP at batch 512
    assert abs(3 * bench.fwd_flops(bench.QM9) * 512 - 2.5816e12) < 1e9


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun (N > 1) rank 0 alone runs the CPU reference; the other ranks exit 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cpu_arm_builds_its_parameters_without_the_cuda_library():
    """The reference arm must not load libecnf_b200.so (it did, through Engine)."""
    code = ("import bench; p = bench.host_params('lj13'); "
            "maps = open('/proc/self/maps').read(); "
            "print(len(p), 'libecnf_b200' in maps)")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr
    n, loaded = r.stdout.split()
    assert int(n) > 50 and loaded == "False"

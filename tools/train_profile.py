"""A few QM9-positional flow-matching steps (batch 512) for per-kernel timing under ncu (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench
ap = argparse.Namespace(steps=2, warmup=3)
r = bench.bench_train(ap, torch.device("cuda", 0), 0, 1)
print(r["value"], "steps/s", r["ms_per_step"], "ms")

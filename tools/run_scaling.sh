#!/bin/bash
# Multi-GPU measurement session (run on a GPU box with N GPUs):  tools/run_scaling.sh N [aldp]
# Writes one contract line per workload to gpurun_out/r2_<workload>_n<N>.json
N=$1; shift
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ "$N" = 1 ] && RUN="python"
PORT=29600
run() { # name, args...
  name=$1; shift
  PORT=$((PORT+1))
  if [ "$N" = 1 ]; then cmd="python bench.py --gpus 1 $*"; else cmd="$RUN --master-port $PORT bench.py --gpus $N $*"; fi
  timeout 900 $cmd > gpurun_out/r2_${name}_n${N}.json 2> gpurun_out/r2_${name}_n${N}.err
  echo "== $name n=$N rc=$? $(head -c 300 gpurun_out/r2_${name}_n${N}.json)"
}
for w in "$@"; do
  case $w in
    aldp)   run aldp_strong --workload aldp --scaling strong --steps 1 --warmup 1 --no-cpu --no-count ;;
    lj13s)  run lj13_strong --workload lj13 --scaling strong --steps 3 --warmup 2 --no-cpu --no-extra --no-count ;;
    lj13w)  run lj13_weak --workload lj13 --steps 2 --warmup 2 --no-cpu --no-extra --no-count ;;
    sweep)  run sweep --workload sweep --steps 2 --no-cpu --no-count --sweep-max ${SWEEP_MAX:-1000000} ;;
    fm)     run fm --workload fm --no-cpu ;;
    dflt)   run default --steps 1 --warmup 1 --no-cpu ;;
  esac
done

#!/bin/bash
# One-GPU measurement session (run on the GPU box through gpurun):  tools/gpu_session.sh <tag> [steps...]
# Every step writes into gpurun_out/<tag>_*; steps: tests tests2 thr thrq bench benchq launches ncu_lj13 ncu_aldp ncu_sample prof fm fmsweep fmlaunch aldp sweep
TAG=$1; shift
O=gpurun_out
mkdir -p $O
for s in "$@"; do
  case $s in
    tests)    timeout 1200 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "== tests rc=$? $(tail -1 $O/${TAG}_tests.log)" ;;
    thr)      timeout 600 python tools/throughput.py --all > $O/${TAG}_throughput.txt 2>&1; echo "== thr rc=$?"; cat $O/${TAG}_throughput.txt ;;
    bench)    timeout 1500 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "== bench rc=$? $(head -c 600 $O/${TAG}_bench.json)" ;;
    benchq)   timeout 900 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > $O/${TAG}_benchq.json 2> $O/${TAG}_benchq.err; echo "== benchq rc=$? $(head -c 400 $O/${TAG}_benchq.json)" ;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches.csv \
                python bench.py --steps 1 --warmup 3 --batch 1184 --no-cpu --no-count > $O/${TAG}_launches.log 2>&1; echo "== launches rc=$?" ;;
    ncu_lj13) timeout 600 ncu --set full --clock-control none --import-source on -k regex:ecnf_solve_tc_kernel --launch-skip 1 --launch-count 1 \
                -o $O/${TAG}_solve_lj13 -f python tools/run_solve.py lj13 148 logq 2 > $O/${TAG}_ncu_lj13.log 2>&1; echo "== ncu_lj13 rc=$?" ;;
    ncu_aldp) timeout 600 ncu --set full --clock-control none --import-source on -k regex:ecnf_solve_tc_kernel --launch-skip 1 --launch-count 1 \
                -o $O/${TAG}_solve_aldp -f python tools/run_solve.py aldp 148 logq 2 > $O/${TAG}_ncu_aldp.log 2>&1; echo "== ncu_aldp rc=$?" ;;
    ncu_sample) timeout 600 ncu --set full --clock-control none --import-source on -k regex:ecnf_solve_tc_kernel --launch-skip 1 --launch-count 1 \
                -o $O/${TAG}_solve_sample -f python tools/run_solve.py lj13 4736 sample 2 > $O/${TAG}_ncu_sample.log 2>&1; echo "== ncu_sample rc=$?" ;;
    prof)     for w in "lj13 148" "aldp 148" "lj13 1184 sample"; do
                ECNF_B200_LIB=ecnf_b200/libecnf_b200_prof.so timeout 300 python tools/tc_profile.py $w >> $O/${TAG}_phase_cycles.txt 2>&1; done
              echo "== prof"; cat $O/${TAG}_phase_cycles.txt ;;
    tests2)   timeout 900 python -m pytest tests/test_gpu_hutchinson.py tests/test_gpu_train.py tests/test_gpu_round2.py -m gpu -x -q > $O/${TAG}_tests2.log 2>&1; echo "== tests2 rc=$? $(tail -3 $O/${TAG}_tests2.log)" ;;
    thrq)     timeout 600 python tools/throughput.py > $O/${TAG}_throughput.txt 2>&1; echo "== thr rc=$?"; cat $O/${TAG}_throughput.txt ;;
    fmsweep)  for ch in 512 256 171 128 64; do
                timeout 300 python bench.py --workload fm --no-cpu --fm-chunk $ch > $O/${TAG}_fm_chunk$ch.json 2> $O/${TAG}_fm_chunk$ch.err
                echo "== fm chunk $ch rc=$? $(head -c 330 $O/${TAG}_fm_chunk$ch.json)"; done ;;
    fmlaunch) for ch in 512 128; do
                timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_fm_launches_chunk$ch.csv \
                  python tools/train_profile.py $ch > $O/${TAG}_fm_launches_chunk$ch.log 2>&1; echo "== fmlaunch $ch rc=$?"; done ;;
    fm)       timeout 600 python bench.py --workload fm > $O/${TAG}_fm.json 2> $O/${TAG}_fm.err; echo "== fm rc=$? $(head -c 600 $O/${TAG}_fm.json)" ;;
    aldp)     timeout 900 python bench.py --workload aldp --steps 1 --warmup 1 > $O/${TAG}_aldp.json 2> $O/${TAG}_aldp.err; echo "== aldp rc=$? $(head -c 600 $O/${TAG}_aldp.json)" ;;
    sweep)    timeout 900 python bench.py --workload sweep --steps 2 > $O/${TAG}_sweep.json 2> $O/${TAG}_sweep.err; echo "== sweep rc=$? $(head -c 900 $O/${TAG}_sweep.json)" ;;
  esac
done

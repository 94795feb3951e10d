#!/usr/bin/env bash
# Everything a round needs from ONE gpurun call (getting a box is charged per call):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_suite.sh r02'
#
# GPU parity tests, smoke(), both bench arms, the MSG / multiview lines, the in-situ timeline and
# the ncu launch list of an eager step.  Every leg runs under its own timeout and writes to
# gpurun_out/<tag>_*; a failing leg does not stop the others.  Numbers printed under ncu are
# never bench values (the launch list is for the kernels' SHARES only).
set -u
tag=${1:-suite}
out=gpurun_out
mkdir -p "$out"
leg() {  # leg <name> <seconds> <command...>
  local name=$1 secs=$2
  shift 2
  local t0=$SECONDS
  timeout "$secs" "$@" > "$out/${tag}_${name}.out" 2> "$out/${tag}_${name}.err"
  echo "$name rc=$? $((SECONDS - t0))s" | tee -a "$out/${tag}_summary.txt"
}
: > "$out/${tag}_summary.txt"
leg pytest 900 python -m pytest tests -m gpu -x -q
leg smoke 300 python -c 'import __graft_entry__ as g; g.smoke()'
leg bench_ref 300 python bench.py --impl reference
CUDA_LAUNCH_BLOCKING=1 leg bench_ref_blocking 300 python bench.py --impl reference
leg bench 400 python bench.py
leg bench_msg 300 python bench.py --msg --no-cpu-baseline --no-per-op
leg bench_msg_mv 300 python bench.py --msg --multiview --no-cpu-baseline --no-per-op
leg timeline 300 python tools/timeline.py "$tag"
leg launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
  --log-file "$out/${tag}_launches.csv" python bench.py --steps 1 --warmup 1 --settle 1 --no-graph --no-prefetch \
  --no-cpu-baseline --no-per-op
tail -n 3 "$out/${tag}_pytest.out"
cat "$out/${tag}_summary.txt"

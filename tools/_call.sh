set -u
out=gpurun_out; tag=s4k
leg() { local name=$1 secs=$2; shift 2; local t0=$SECONDS; timeout "$secs" "$@" > "$out/${tag}_${name}.out" 2> "$out/${tag}_${name}.err"; echo "$name rc=$? $((SECONDS - t0))s" | tee -a "$out/${tag}_summary.txt"; }
: > "$out/${tag}_summary.txt"
leg pytest 900 python -m pytest tests -m gpu -q
leg bench 400 python bench.py --no-cpu-baseline --no-per-op
leg bench_msg 300 python bench.py --msg --no-cpu-baseline --no-per-op
leg trace 200 python tools/trace_step.py s4k
tail -n 6 "$out/${tag}_pytest.out"; for f in bench bench_msg; do cat "$out/${tag}_$f.out" | tail -1 | cut -c1-230; done; cat "$out/${tag}_summary.txt"

set -u
out=gpurun_out; tag=s4h
leg() { local name=$1 secs=$2; shift 2; local t0=$SECONDS; timeout "$secs" "$@" > "$out/${tag}_${name}.out" 2> "$out/${tag}_${name}.err"; echo "$name rc=$? $((SECONDS - t0))s" | tee -a "$out/${tag}_summary.txt"; }
: > "$out/${tag}_summary.txt"
leg bench_pdl 400 python bench.py --no-cpu-baseline --no-per-op
PN2_PDL=0 leg bench_nopdl 400 python bench.py --no-cpu-baseline --no-per-op
leg bench_pdl2 400 python bench.py --no-cpu-baseline --no-per-op
leg pytest 900 python -m pytest tests -m gpu -q -x
leg trace 200 python tools/trace_step.py s4h
tail -n 4 "$out/${tag}_pytest.out"; for f in bench_pdl bench_nopdl bench_pdl2; do cat "$out/${tag}_$f.out" | tail -1 | cut -c1-230; tail -2 "$out/${tag}_$f.err"; done; cat "$out/${tag}_summary.txt"

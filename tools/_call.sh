set -u
out=gpurun_out; tag=s4i
leg() { local name=$1 secs=$2; shift 2; local t0=$SECONDS; timeout "$secs" "$@" > "$out/${tag}_${name}.out" 2> "$out/${tag}_${name}.err"; echo "$name rc=$? $((SECONDS - t0))s" | tee -a "$out/${tag}_summary.txt"; }
: > "$out/${tag}_summary.txt"
leg bench 400 python bench.py
leg trace 200 python tools/trace_step.py s4i
leg lead 200 python -m pytest tests/test_lead_gpu.py -q
cat "$out/${tag}_bench.out" | tail -1; grep "lead::" $out/trace_s4i.txt; tail -2 $out/${tag}_lead.out; cat "$out/${tag}_summary.txt"

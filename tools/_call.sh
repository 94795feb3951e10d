set -u
out=gpurun_out; tag=s4l
: > $out/${tag}.txt
for kb in 200 100 64 200 100; do
  echo "== PN2_WGRAD_SMEM_KB=$kb" >> $out/${tag}.txt
  PN2_WGRAD_SMEM_KB=$kb timeout 200 python bench.py --no-cpu-baseline --no-per-op 2>/dev/null | cut -c1-200 >> $out/${tag}.txt
done
cat $out/${tag}.txt

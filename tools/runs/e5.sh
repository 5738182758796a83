mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e5_pytest.log
for w in cfg2 cfg3; do
  for pre in 1 0; do
   echo "== $w pre_agg=$pre"; DRGNN_PRE_AGG=$pre timeout 120 python tools/step_breakdown.py $w 200 2>&1 | grep -v "graph=False\|structure kernel\|edges \|emit split\|CTA of graph 0: 0\|stage 0 " | tail -9
   DRGNN_PRE_AGG=$pre timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e5_bench_${w}_pre$pre.json 2> gpurun_out/e5_bench_${w}_pre$pre.err
  done
done > gpurun_out/e5_breakdown.log

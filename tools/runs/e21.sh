mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e21_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
   timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e21_bench_${w}.json 2> gpurun_out/e21_bench_${w}.err
done
DRGNN_PRE_AGG=1 timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload cfg2 > gpurun_out/e21_bench_cfg2_pre1.json 2> gpurun_out/e21_bench_cfg2_pre1.err
for w in cfg2 cfg4; do echo "== $w"; timeout 120 python tools/step_breakdown.py $w 200 2>&1 | grep -v "graph=False\|edges \|emit split\|CTA of graph 0: 0\|stage 0 " | tail -9; done > gpurun_out/e21_breakdown.log

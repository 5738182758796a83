mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/e24_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
   timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e24_bench_${w}.json 2> gpurun_out/e24_bench_${w}.err
done
DRGNN_PRE_AGG=1 timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload cfg2 > gpurun_out/e24_bench_cfg2_pre1.json 2> gpurun_out/e24_bench_cfg2_pre1.err
for w in cfg3 cfg4 cfg5; do echo "== $w"; timeout 120 python tools/step_breakdown.py $w 100 2>&1 | grep "blob structure\|graph=True" -A2 | grep -v "^--"; done > gpurun_out/e24_breakdown.log

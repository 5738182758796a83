mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/e22_pytest.log
for pre in 0 1; do
DRGNN_PRE_AGG=$pre timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload cfg2 > gpurun_out/e22_bench_cfg2_pre$pre.json 2> gpurun_out/e22_bench_cfg2_pre$pre.err
done
for w in cfg4 cfg5; do DRGNN_PRE_AGG=0 timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e22_bench_${w}_pre0.json 2> gpurun_out/e22_bench_${w}_pre0.err; done
echo "== cfg2 pre0"; DRGNN_PRE_AGG=0 timeout 120 python tools/step_breakdown.py cfg2 200 2>&1 | grep "blob structure" -A1 > gpurun_out/e22_breakdown.log

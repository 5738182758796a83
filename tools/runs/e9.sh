mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e9_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
   timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e9_bench_${w}.json 2> gpurun_out/e9_bench_${w}.err
done
timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload cfg3 --layers 3 > gpurun_out/e9_bench_cfg3_l3.json 2> gpurun_out/e9_bench_cfg3_l3.err

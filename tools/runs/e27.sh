mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/e27_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
   timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e27_bench_${w}.json 2> gpurun_out/e27_bench_${w}.err
done

mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e14_pytest.log
for w in cfg2 cfg3 cfg5; do
   timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e14_bench_${w}.json 2> gpurun_out/e14_bench_${w}.err
done

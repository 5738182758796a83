mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_step3_gpu.py tests/test_pinned_path_gpu.py -m gpu -x -q 2>&1 | tail -4) > gpurun_out/e13_pytest.log
for w in cfg2 cfg3; do
  echo "== $w"; timeout 120 python tools/step_breakdown.py $w 200 2>&1 | grep -v "graph=False\|structure kernel\|edges \|emit split\|CTA of graph 0: 0\|stage 0 \|load+minmax" | tail -7
  timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e13_bench_${w}.json 2> gpurun_out/e13_bench_${w}.err
done > gpurun_out/e13_breakdown.log

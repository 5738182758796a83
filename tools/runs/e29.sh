mkdir -p gpurun_out
timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload cfg3 > gpurun_out/e29_bench_cfg3.json 2> gpurun_out/e29_bench_cfg3.err
echo "== cfg3"; timeout 120 python tools/step_breakdown.py cfg3 100 2>&1 | grep "CTAs\|graph=True" > gpurun_out/e29_breakdown.log
(timeout 300 python -m pytest tests/test_step3_gpu.py -m gpu -x -q 2>&1 | tail -3) > gpurun_out/e29_pytest.log

mkdir -p gpurun_out
for cs in 2 1; do
  DRGNN_FEED_COPY_STREAMS=$cs timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 > gpurun_out/e10_bench_cs$cs.json 2> gpurun_out/e10_bench_cs$cs.err
  echo "== copy streams $cs"; DRGNN_FEED_COPY_STREAMS=$cs timeout 100 python tools/e2e_bound.py 400 2>&1 | tail -5
done > gpurun_out/e10_bound.log
(timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_modules_gpu.py -m gpu -x -q 2>&1 | tail -3) > gpurun_out/e10_pytest.log

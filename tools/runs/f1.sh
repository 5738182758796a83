mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/f1_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5) > gpurun_out/f1_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f1_bench_default.json 2> gpurun_out/f1_bench_default.err

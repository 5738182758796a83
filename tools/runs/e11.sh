mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/e11_pytest.log
for w in cfg4 cfg5; do
 for lo in 1 0; do
   DRGNN_LOCAL_STRUCTURE=$lo timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e11_bench_${w}_lo$lo.json 2> gpurun_out/e11_bench_${w}_lo$lo.err
 done
done

mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e1_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
  for tc in 1 0; do
    DRGNN_FUSED_TC=$tc timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e1_bench_${w}_tc$tc.json 2> gpurun_out/e1_bench_${w}_tc$tc.err
  done
done
for w in cfg2 cfg3 cfg4 cfg5; do
  for tc in 1 0; do
   echo "== $w tc=$tc"; DRGNN_FUSED_TC=$tc timeout 120 python tools/step_breakdown.py $w 200 2>&1 | tail -8
  done
done > gpurun_out/e1_breakdown.log

mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e3_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
  for tc in 0 1; do
   echo "== $w tc=$tc"; DRGNN_FUSED_TC=$tc timeout 120 python tools/step_breakdown.py $w 200 2>&1 | grep -v "graph=False\|structure kernel\|edges \|load+minmax\|emit split" | tail -9
  done
done > gpurun_out/e3_breakdown.log
for lpt in 1 0; do
  DRGNN_STEP3_LPT=$lpt timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload cfg5 > gpurun_out/e3_bench_cfg5_lpt$lpt.json 2> gpurun_out/e3_bench_cfg5_lpt$lpt.err
done

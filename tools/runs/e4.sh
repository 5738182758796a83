mkdir -p gpurun_out
(DRGNN_FUSED_TC=1 timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_step3_gpu.py tests/test_pinned_path_gpu.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/e4_pytest.log
for w in cfg2 cfg3 cfg4 cfg5; do
  for tc in 0 1; do
   echo "== $w tc=$tc"; DRGNN_FUSED_TC=$tc timeout 120 python tools/step_breakdown.py $w 200 2>&1 | grep -v "graph=False\|structure kernel\|edges \|load+minmax\|emit split\|CTA of graph 0: 0\|stage 0 " | tail -9
  done
done > gpurun_out/e4_breakdown.log

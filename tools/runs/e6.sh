mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/e6_pytest.log
for w in cfg2 cfg3; do
 for pdl in 1 0; do
  for pre in 1 0; do
   DRGNN_PDL_PREP=$pdl DRGNN_PRE_AGG=$pre timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e6_bench_${w}_pdl${pdl}_pre$pre.json 2> gpurun_out/e6_bench_${w}_pdl${pdl}_pre$pre.err
  done
 done
done

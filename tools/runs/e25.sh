mkdir -p gpurun_out
for rep in 1 2; do
for w in cfg2 cfg4 cfg5; do
 for pre in 1 0; do
   DRGNN_PRE_AGG=$pre timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 --workload $w > gpurun_out/e25_bench_${w}_pre${pre}_$rep.json 2> gpurun_out/e25_bench_${w}_pre${pre}_$rep.err
 done
done
done

mkdir -p gpurun_out
numactl -H > gpurun_out/e20_numa.log 2>&1 || (ls /sys/devices/system/node/ > gpurun_out/e20_numa.log; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) >> gpurun_out/e20_numa.log; done)
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d numa $(cat $f)"; fi; done >> gpurun_out/e20_numa.log
for rep in 1 2; do
 for nb in 1 0; do
  DRGNN_BENCH_NUMA=$nb timeout 200 python bench.py --no-cpu --no-roofline --steps 20 --warmup 5 > gpurun_out/e20_bench_numa${nb}_$rep.json 2> gpurun_out/e20_bench_numa${nb}_$rep.err
 done
done

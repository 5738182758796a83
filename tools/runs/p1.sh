# round-2 profiles: ncu launch lists of the bench command (graphs off so kernels are listed individually) and
# full captures of the step / structure kernels per configuration
mkdir -p gpurun_out
for w in cfg2 cfg3 cfg4 cfg5; do
  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_${w}.csv \
     python bench.py --workload $w --steps 4 --warmup 1 --repeats 1 --no-graph --no-roofline --no-cpu --pool 4 > gpurun_out/p1_${w}_launch.log 2>&1
done
for w in cfg2 cfg3 cfg4 cfg5; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:'graph_step2|graph_step3|graph_blob|graph_local|graph_finalize|net_step_reduce' --launch-skip 6 -c 4 \
     -o gpurun_out/r2_step_${w} -f python tools/one_step.py $w 6 > gpurun_out/p1_${w}_full.log 2>&1
done
ls -la gpurun_out/*.ncu-rep

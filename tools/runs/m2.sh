mkdir -p gpurun_out
run() { # N workload
  N=$1; W=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 20 --warmup 5 --workload $W --no-cpu --no-roofline "$@" > gpurun_out/m2_bench_${W}_n$N.json 2> gpurun_out/m2_bench_${W}_n$N.err
}
run 8 cfg2
run 8 cfg4
run 8 cfg5
run 8 cfg3
run 2 cfg2
run 4 cfg2

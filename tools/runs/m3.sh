mkdir -p gpurun_out
run() { N=$1; W=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 20 --warmup 5 --workload $W --no-cpu --no-roofline "$@" > gpurun_out/m3_bench_${W}_n$N.json 2> gpurun_out/m3_bench_${W}_n$N.err
}
run 2 cfg2
run 2 cfg4
run 2 cfg5

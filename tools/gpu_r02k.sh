mkdir -p gpurun_out
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02k_bench_n$N.json 2> gpurun_out/r02k_bench_n$N.err; echo "rc=$?"; tail -c 800 gpurun_out/r02k_bench_n$N.err; tail -c 300 gpurun_out/r02k_bench_n$N.json

# multi-GPU validation: parity scripts at N ranks (slab and RCB partitions), then the bench line
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 tests/dist_check_rcb.py > gpurun_out/r02_dist_rcb_n$N.log 2>&1; echo "rcb rc=$?"; grep -E "OK|PASSED" gpurun_out/r02_dist_rcb_n$N.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 tests/dist_check.py > gpurun_out/r02_dist_slab_n$N.log 2>&1; echo "slab rc=$?"; grep -E "OK|PASSED" gpurun_out/r02_dist_slab_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n$N.err | tail -3; tail -c 200 gpurun_out/r02_bench_n$N.json

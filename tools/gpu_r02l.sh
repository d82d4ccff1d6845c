mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scale_parity.py -x -q -k "torchrun" > gpurun_out/r02l_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02l_dist.log; tail -5 gpurun_out/r02l_dist.log
bash tools/gpu_r02k.sh 4 2>&1 | tail -3
bash tools/gpu_r02k.sh 2 2>&1 | tail -3
FFCUDA_HALO_OVERLAP=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 4 --steps 3 --warmup 3 --no-parity --no-strong > gpurun_out/r02l_bench_n4_nooverlap.json 2>/dev/null; echo rc=$?

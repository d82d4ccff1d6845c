#!/bin/bash
# round 2, session 2, second 2-GPU run: the plugin driving two GPUs (FFCUDA_NGPU=2), with peer mailboxes shared by pointer inside
# the process, and with NCCL only (FFCUDA_P2P=0); the multi-process checks again (comm.cu changed)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_plugin.py -m gpu -q -k "two_gpus" --durations=5 > gpurun_out/r03d_pytest_ngpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r03d_pytest_ngpu.log
FFCUDA_P2P=0 timeout 300 python -m pytest tests/test_plugin.py -m gpu -q -k "two_gpus" > gpurun_out/r03d_pytest_ngpu_nccl.log 2>&1; echo "pytest (NCCL only) rc=$?"; tail -5 gpurun_out/r03d_pytest_ngpu_nccl.log
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 $2 > gpurun_out/$3 2>&1; echo "$2 rc=$?"; tail -2 gpurun_out/$3; }
run 29711 tests/dist_check_rcb.py r03d_dist_rcb.log
run 29712 tests/dist_check_p2.py r03d_dist_p2.log

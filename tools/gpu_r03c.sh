#!/bin/bash
# round 2, session 2, 2-GPU run: distributed GMRES (in dist_check_rcb), P2 spaces on a distributed mesh (dist_check_p2), the slab
# check as a regression of the halo refactoring, and the single-GPU GMRES tests (kernels refactored)
mkdir -p gpurun_out
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 $2 > gpurun_out/$3 2>&1; echo "$2 rc=$?"; tail -4 gpurun_out/$3; }
run 29701 tests/dist_check_p2.py r03c_dist_p2.log
run 29702 tests/dist_check_rcb.py r03c_dist_rcb.log
run 29703 tests/dist_check.py r03c_dist_slab.log
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gmres" > gpurun_out/r03c_pytest_gmres.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r03c_pytest_gmres.log

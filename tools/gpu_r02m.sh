mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_scale_parity.py -x -q -k "torchrun and not 3 and not 4" > gpurun_out/r02m_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_dist.log; tail -3 gpurun_out/r02m_dist.log
bash tools/gpu_r02k.sh 2 2>&1 | tail -2

mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests/test_gpu_scale_parity.py -x -q -k "torchrun" > gpurun_out/r02i_dist.log 2>&1; echo "rc=$?" >> gpurun_out/r02i_dist.log; tail -30 gpurun_out/r02i_dist.log

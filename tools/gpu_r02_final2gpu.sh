mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_scale_parity.py -x -q -k "torchrun and 2" > gpurun_out/r02_dist2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_dist2.log; tail -3 gpurun_out/r02_dist2.log
bash tools/gpu_r02_benchN.sh 2

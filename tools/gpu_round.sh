# One GPU-box call: parity tests, bench (both arms), launch list and full ncu capture of the top kernels.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python tools/prof_driver.py 128 5 > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_asm_p1_lean|k_spmv_sell|k_sym_p1|k_rhs_p1_lean' -s 0 -c 8 -o gpurun_out/prof_${TAG} -f python tools/prof_driver.py 128 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log; tail -15 gpurun_out/r02h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; tail -2 gpurun_out/r02h_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02h_bench.err; head -c 6000 gpurun_out/r02h_bench.json

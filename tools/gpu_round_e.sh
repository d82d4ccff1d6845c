# Round-1e GPU call: full parity suite, smoke, bench (both arms), launch list of bench.py, timings of the widened rows.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.json
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>&1; tail -c 400 gpurun_out/bench_ref.json
timeout 300 python tools/widen_run.py 96 > gpurun_out/widen_r01e.jsonl 2> gpurun_out/widen.err; cat gpurun_out/widen_r01e.jsonl; tail -3 gpurun_out/widen.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 3 > gpurun_out/launches.log 2>&1
ls -la gpurun_out | tail -12

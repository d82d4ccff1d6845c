mkdir -p gpurun_out
timeout 300 python tools/cg_graph_check.py 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "cg or golden or config" 2>&1 | tail -3

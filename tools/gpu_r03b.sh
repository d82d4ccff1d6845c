#!/bin/bash
# round 2, session 2, second GPU run: boundary integrals with derivatives (new golden cases + plugin cases) and the boundary
# regressions they share code with
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_plugin.py -m gpu -q --durations=8 \
  -k "bnd_grad or robin or neumann or traction or bnd_g or device_mesh" > gpurun_out/r03b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r03b_pytest.log
tail -25 gpurun_out/r03b_pytest.log

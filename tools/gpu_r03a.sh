#!/bin/bash
# round 2, session 2, first GPU run: whole GPU suite (the plugin tests now go through the device cube) + script-level timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r03a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r03a_pytest.log
tail -5 gpurun_out/r03a_pytest.log
timeout 200 python tools/plugin_e2e.py 128 > gpurun_out/r03a_plugin_e2e.log 2>&1; echo "e2e rc=$?"
tail -12 gpurun_out/r03a_plugin_e2e.log

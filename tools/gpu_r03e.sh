#!/bin/bash
# the bench's .edp at cube(128) with the plugin on 1 and on 2 GPUs: script-level wall clock of the three statements
mkdir -p gpurun_out
timeout 200 python tools/plugin_e2e.py 128 > gpurun_out/r03e_plugin_e2e_1gpu.log 2>&1; echo "1 GPU rc=$?"; head -3 gpurun_out/r03e_plugin_e2e_1gpu.log
FFCUDA_NGPU=2 timeout 200 python tools/plugin_e2e.py 128 > gpurun_out/r03e_plugin_e2e_2gpu.log 2>&1; echo "2 GPUs rc=$?"; grep -v "^ *[0-9]* :" gpurun_out/r03e_plugin_e2e_2gpu.log | head -20

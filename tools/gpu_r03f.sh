#!/bin/bash
# the bench's .edp at cube(128) with the plugin on 2 GPUs (gang created at load time), time marks of the distributed solve
mkdir -p gpurun_out
FFCUDA_NGPU=2 timeout 200 python tools/plugin_e2e.py 128 > gpurun_out/r03f_plugin_e2e_2gpu.log 2>&1; echo "2 GPUs rc=$?"; grep -v "^ *[0-9]* :" gpurun_out/r03f_plugin_e2e_2gpu.log | head -20

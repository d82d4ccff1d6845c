mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck python tools/fan_check.py 16 small > gpurun_out/r02j_sanitize.log 2>&1; tail -4 gpurun_out/r02j_sanitize.log
FFCUDA_VERBOSE=1 timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|rhs_rows|round-1|fans:|Error|error|first" | tail -6

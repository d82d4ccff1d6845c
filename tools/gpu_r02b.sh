set -x
mkdir -p gpurun_out
export FFCUDA_VERBOSE=1
timeout 300 compute-sanitizer --tool memcheck python tools/fan_check.py 16 small > gpurun_out/r02b_sanitize.log 2>&1; tail -25 gpurun_out/r02b_sanitize.log
timeout 300 python tools/fan_check.py 128 > gpurun_out/r02b_fan.log 2>&1; tail -30 gpurun_out/r02b_fan.log
for thr in 128 512; do FFCUDA_FAN_THREADS=$thr timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|round-1|fans:" ; done
for rows in 64 128 192; do ROWS=$rows timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|round-1|fans:|tiles:" ; done

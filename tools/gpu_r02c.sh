set -x
mkdir -p gpurun_out
export FFCUDA_VERBOSE=1
timeout 300 compute-sanitizer --tool memcheck python tools/fan_check.py 16 small > gpurun_out/r02c_sanitize.log 2>&1; tail -3 gpurun_out/r02c_sanitize.log
timeout 300 python tools/fan_check.py 128 > gpurun_out/r02c_fan.log 2>&1; grep -v "^ffcuda" gpurun_out/r02c_fan.log | tail -8; grep "^ffcuda" gpurun_out/r02c_fan.log | tail -2
for thr in 128 512; do FFCUDA_FAN_THREADS=$thr timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows" ; done
for rows in 64 128; do ROWS=$rows timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|round-1|fans:|tiles:" | tail -4; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm_fans" -s 2 -c 1 -o gpurun_out/r02c_fans python tools/fan_check.py 128 > gpurun_out/r02c_ncu.log 2>&1; tail -3 gpurun_out/r02c_ncu.log
ncu -i gpurun_out/r02c_fans.ncu-rep --page raw --csv > gpurun_out/r02c_raw.csv 2>/dev/null
ncu -i gpurun_out/r02c_fans.ncu-rep --page source --csv > gpurun_out/r02c_src.csv 2>/dev/null
ls -la gpurun_out/r02c*

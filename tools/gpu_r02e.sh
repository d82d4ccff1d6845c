mkdir -p gpurun_out
export FFCUDA_VERBOSE=1
timeout 300 compute-sanitizer --tool memcheck python tools/fan_check.py 16 small > gpurun_out/r02e_sanitize.log 2>&1; tail -2 gpurun_out/r02e_sanitize.log
for rows in 96 64 48; do for thr in 128 64 256; do echo "== rows $rows threads $thr"; ROWS=$rows FFCUDA_FAN_THREADS=$thr timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|round-1|fans:|Error|error" | tail -3; done; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm_fans" -s 33 -c 1 -o gpurun_out/r02e_fans python tools/fan_check.py 128 > gpurun_out/r02e_ncu.log 2>&1; tail -2 gpurun_out/r02e_ncu.log
ncu -i gpurun_out/r02e_fans.ncu-rep --page raw --csv > gpurun_out/r02e_raw.csv 2>/dev/null
ncu -i gpurun_out/r02e_fans.ncu-rep --page source --csv > gpurun_out/r02e_src.csv 2>/dev/null

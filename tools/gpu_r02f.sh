mkdir -p gpurun_out
export FFCUDA_VERBOSE=1
timeout 300 python tools/fan_check.py 16 small 2>&1 | grep -c "rel err"
for rows in 96 64; do echo "== rows $rows"; ROWS=$rows timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|round-1|fans:|Error|error" | tail -3; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm_fans" -s 33 -c 1 -o gpurun_out/r02f_fans python tools/fan_check.py 128 > gpurun_out/r02f_ncu.log 2>&1; tail -2 gpurun_out/r02f_ncu.log
ncu -i gpurun_out/r02f_fans.ncu-rep --page raw --csv > gpurun_out/r02f_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_fans.ncu-rep --page source --csv > gpurun_out/r02f_src.csv 2>/dev/null

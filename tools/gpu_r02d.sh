set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_asm_fans" -s 33 -c 1 -o gpurun_out/r02d_fans python tools/fan_check.py 128 > gpurun_out/r02d_ncu.log 2>&1; tail -3 gpurun_out/r02d_ncu.log
ncu -i gpurun_out/r02d_fans.ncu-rep --page raw --csv > gpurun_out/r02d_raw.csv 2>/dev/null
ncu -i gpurun_out/r02d_fans.ncu-rep --page source --csv > gpurun_out/r02d_src.csv 2>/dev/null
ls -la gpurun_out/r02d*

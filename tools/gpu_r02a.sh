set -x
mkdir -p gpurun_out
nvidia-smi -L
./tools/micro/pipes > gpurun_out/r02a_pipes.txt 2>&1; cat gpurun_out/r02a_pipes.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sym_p1_fused|k_asm_tiles|k_rhs_tiles" -c 9 -o gpurun_out/r02a_base python tools/prof_final.py 128 > gpurun_out/r02a_ncu.log 2>&1; tail -3 gpurun_out/r02a_ncu.log
ncu -i gpurun_out/r02a_base.ncu-rep --page raw --csv > gpurun_out/r02a_raw.csv 2>/dev/null; ls -la gpurun_out/r02a*

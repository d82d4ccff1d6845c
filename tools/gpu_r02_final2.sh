# final single-GPU measurements of round 2 on the committed build
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n1.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-strong > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_asm_fans|k_rhs_fans|k_sym_p1_fused|k_spmv_sell" -s 10 -c 5 -o gpurun_out/r02_final python tools/prof_final.py 128 > gpurun_out/r02_ncu.log 2>&1; tail -2 gpurun_out/r02_ncu.log
ncu -i gpurun_out/r02_final.ncu-rep --page raw --csv > gpurun_out/r02_raw.csv 2>/dev/null
timeout 600 python tools/configs_run.py 1 3 4 --it3 0 > gpurun_out/r02_configs_1_3_4.jsonl 2> gpurun_out/r02_configs.err; echo "configs rc=$?"

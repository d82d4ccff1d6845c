mkdir -p gpurun_out
timeout 300 python tools/fan_check.py 16 small 2>&1 | grep -c "rel err"
for rows in 96 64; do echo "== rows $rows"; ROWS=$rows timeout 200 python tools/fan_check.py 128 2>&1 | grep -E "asm_rows|round-1|Error|error" | tail -3; done

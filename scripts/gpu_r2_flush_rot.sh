#!/bin/bash
# First GPU call of the next round: validate and measure the opt-in rotated flush reads of the Hex8 tile kernel
# (fb200_set_tuning "hex8_flush_rot" / FB200_HEX8_FLUSH_ROT=1; profiles/r01/README.md, "Flush bank-conflict model").
mkdir -p gpurun_out
FB200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_hex8_tile.py -q -k rotated > gpurun_out/r2_rot_pytest.log 2>&1; tail -n 3 gpurun_out/r2_rot_pytest.log
for rot in 0 1 0 1; do
  FB200_HEX8_FLUSH_ROT=$rot timeout 300 python bench.py --no-e2e --no-cpu --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rot=$rot', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])" | tee -a gpurun_out/r2_rot_bench.log
done
FB200_HEX8_FLUSH_ROT=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:hex8_tile -s 3 -c 1 -o gpurun_out/prof_tile_rot python bench.py --no-e2e --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_rot.log 2>&1

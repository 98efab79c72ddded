#!/bin/bash
# session 3: Hex8 tile kernel - parity + tile-size / debug-knob sweep on C3 (short)
mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 300 python -m pytest tests/test_hex8_tile.py -m gpu -q --maxfail=8 > gpurun_out/pytest_tile.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_tile.log
tail -n 4 gpurun_out/pytest_tile.log
for tile in 64; do
FB200_HEX8_TILE=$tile timeout 150 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_tile${tile}.log 2>&1
for dbg in 1 2 3 4; do
FB200_HEX8_TILE=$tile FB200_DEBUG=$dbg timeout 150 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_tile${tile}_dbg${dbg}.log 2>&1
done; done
for f in gpurun_out/b_tile*.log; do echo -n "$f "; tail -n 1 $f | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3))
except Exception as e: print('ERR', e)"; done

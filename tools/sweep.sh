#!/bin/bash
# BASELINE.json configs 3 and 4 on one GPU: W4A8 / W8A8 x codebook size 64 / 128 / 256, max fusion, 2 / 4 agents.
# usage: tools/sweep.sh > gpurun_out/sweep.jsonl
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>/dev/null | grep '^{'; }
for wb in 8 4; do for k in 64 128 256; do run --w-bits $wb --dict-size $k; done; done
run --fusion max
run --agents 4
run --agents 2

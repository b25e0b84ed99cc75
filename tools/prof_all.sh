#!/bin/bash
# Time the five representative conv layers (CUDA-graph replay, not under a profiler).  usage: tools/prof_all.sh [tag]
for l in shrink1 shrink0 s0 s1 s2 d0 d1 d2; do python tools/prof_layer.py $l 50 --graph; done

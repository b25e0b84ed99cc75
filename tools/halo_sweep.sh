#!/bin/bash
# Layer times of the HALO mainloop for (BLOCK_N, weights-resident) choices vs the tap-per-stage mainloop.
for l in s0 s1 s2 shrink1 shrink0; do
  echo "== $l old mainloop"; QV2X_HALO=0 python tools/prof_layer.py $l 50 --graph
  echo "== $l halo (cost model)"; python tools/prof_layer.py $l 50 --graph
  for bn in 64 128 256; do for res in 0 1; do
    echo "== $l halo bn=$bn res=$res"; QV2X_HALO_BN=$bn QV2X_HALO_RES=$res python tools/prof_layer.py $l 50 --graph
  done; done
done

export QV2X_LIB=quantv2x_b200/libqv2x_dbg.so
for l in shrink1 s1 shrink0; do for d in 0 1 16 17; do python tools/prof_layer.py $l 50 --graph --debug=$d; done; done

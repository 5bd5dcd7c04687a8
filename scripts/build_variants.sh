#!/bin/bash
# builds tuning variants of libgelcu.so: gel_b200/libgelcu_t<threads>_g<grab>.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off"
for v in "$@"; do
  t=${v%%:*}; g=${v##*:}
  nvcc $FLAGS -DGEL_RASTER_THREADS=$t -DGEL_GRAB=$g -shared gel_b200/csrc/gelcu.cu -o gel_b200/libgelcu_t${t}_g${g}.so &
done
wait
ls gel_b200/*.so

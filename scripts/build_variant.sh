#!/bin/bash
# build_variant.sh <name> [nvcc -D flags ...]  ->  gel_b200/libgelcu_<name>.so (travels to the GPU box; use with GELCU_LIB=<path>)
# prints registers / spills of the kernels named in $KERNELS (regex on the mangled name)
name=$1; shift
out=gel_b200/libgelcu_$name.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off -shared \
  gel_b200/csrc/gelcu.cu "$@" -Xptxas -v -o $out 2> build/ptxas_$name.log || { tail -20 build/ptxas_$name.log; exit 1; }
python - "$name" "${KERNELS:-direct_raster_kernelILi[01]ELb1|direct_resolve_kernelILb0ELb1ELb0|raster_band_kernelILb0|transform_kernel}" <<'PY'
import re, sys
name, pat = sys.argv[1], sys.argv[2]
txt = open(f"build/ptxas_{name}.log").read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \S+\n\s+(.*?)\n.*?Used (\d+) registers", txt, re.S):
    if re.search(pat, m.group(1)): print(f"  {name}: {m.group(1)[:70]:70s} regs {m.group(3)}  {m.group(2)}")
PY

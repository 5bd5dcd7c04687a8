#!/bin/bash
# session 3, call I: parity of the final code, then the round's evidence (scripts/gpu_r2_evidence.sh: bench + reference lines, launch list, full captures)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ) > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_i.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/gpu_r2_evidence.sh

#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 300 python tools/fused_flaky.py 2>&1 | tail -13
echo SELL=0; WB_SPMV_SELL=0 timeout -k 10 300 python tools/fused_flaky.py 2>&1 | tail -12

#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for t in 128 256; do
  WB_SELL_CTA=$t timeout -k 10 300 python tools/microbench.py --skip-pcs --its 50 2>&1 | grep -E "spmv_us"
done

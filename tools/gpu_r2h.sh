#!/bin/bash
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 200 python tools/fused_probe.py --dims 50 50 50 2>&1 | tail -1

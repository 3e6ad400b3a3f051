#!/bin/bash
# robustness of the persistent kernel's synchronisation: random per-thread delays at every phase boundary
# (WB_FUSED_JITTER) must not change a single bit of the result
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
run() {
  env "$@" timeout -k 10 300 python tools/microbench.py --skip-pcs --fused-only $DIMS > gpurun_out/r2n.json 2> gpurun_out/r2n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2n.json"))["gmres30_full_fused"]
print("$* $DIMS:", d["its"], "its", round(d["us_per_it"],2), "us/it", d["rnorm"])
PY
}
run WB_FUSED_JITTER=0
run WB_FUSED_JITTER=1024
run WB_FUSED_JITTER=4096
run WB_FUSED_JITTER=16384

#!/bin/bash
# norm of the new Krylov vector from the dot-product pass (WB_FUSED_NORM): parity tests and the effect on iterations / time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout -k 10 600 python -m pytest tests/test_gpu_fused.py -q 2>&1 | tail -8 | cut -c1-300
for nm in 2; do
  WB_FUSED_NORM=$nm timeout -k 10 300 python tools/microbench.py --skip-pcs --fused-only > gpurun_out/r2p_norm$nm.json 2> gpurun_out/r2p_norm$nm.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2p_norm$nm.json"))["gmres30_full_fused"]
print("norm mode $nm:", d["its"], "its", round(d["us_per_it"],2), "us/it", d["rnorm"], d["breakdown_us"])
PY
done

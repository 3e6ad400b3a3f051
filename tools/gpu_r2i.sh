#!/bin/bash
# round-2 GPU job I (N GPUs): bench lines of config 2 (strong scaling) and config 5 (weak scaling) with the persistent kernel
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
run() {  # label, bench args..., env via WB_*
  label=$1; shift
  WB_FUSED_VERBOSE=1 timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2i_${label}_n$N.json 2> gpurun_out/r2i_${label}_n$N.err
  grep -a "wb_fused" gpurun_out/r2i_${label}_n$N.err | head -1
  python - <<PY
import json
try:
    txt=open("gpurun_out/r2i_${label}_n$N.json").read()
    d=json.loads([l for l in txt.split("\n") if l.startswith("{")][-1])
    print("$label N=$N", round(d["value"],3), "steps/s", round(d["ms_per_step"],2), "ms", d["config"]["ksp_iterations_per_step"], "its", d["config"]["us_per_ksp_iteration"], "us/it", "launches", d["gpu_launches"], "parity", (d.get("parity") or {}).get("residual_relerr"), "reason", d["config"]["ksp_reason"], d["config"]["newton_reason"])
    print("   ", d.get("ksp_breakdown_min_mean_max_over_ctas"))
    print("   ", d["phases_ms"])
except Exception as e:
    print("$label N=$N FAILED", e); print(open("gpurun_out/r2i_${label}_n$N.err").read()[-1500:])
PY
}
run c2
run c5 --config 5

#!/bin/bash
# round-2 GPU job D (N GPUs): multi-GPU parity with the persistent kernel, then strong-scaling bench lines with and without it
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
if [ "$2" != "notest" ]; then
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2d_multi_tests_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_multi_tests_n$N.log
tail -12 gpurun_out/r2d_multi_tests_n$N.log
fi
run() {  # label, extra env, extra args
  label=$1; shift
  env "$@" WB_FUSED_VERBOSE=1 timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/r2d_bench_${label}_n$N.json 2> gpurun_out/r2d_bench_${label}_n$N.err
  grep -a "wb_fused" gpurun_out/r2d_bench_${label}_n$N.err | head -2
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_bench_${label}_n$N.json"))
    print("$label N=$N", round(d["value"],3), "steps/s", round(d["ms_per_step"],2), "ms", d["config"]["ksp_iterations_per_step"], "its", d["config"]["us_per_ksp_iteration"], "us/it", d.get("ksp_breakdown_us_per_iteration"), "launches", d["gpu_launches"], "parity", (d.get("parity") or {}).get("residual_relerr"), "reason", d["config"]["ksp_reason"])
except Exception as e:
    print("$label N=$N FAILED", e); print(open("gpurun_out/r2d_bench_${label}_n$N.err").read()[-1500:])
PY
}
run fused WB_FUSED=1
run unfused WB_FUSED=0

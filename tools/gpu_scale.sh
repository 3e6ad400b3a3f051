#!/bin/bash
# strong-scaling bench line of config 2 at N GPUs (and optionally the weak-scaling line of config 5)
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
run() {
  label=$1; shift
  if [ "$N" = "1" ]; then
    timeout -k 10 500 python bench.py --gpus 1 --steps 5 --warmup 3 "$@" > gpurun_out/r2s_${label}_n$N.json 2> gpurun_out/r2s_${label}_n$N.err
  else
    timeout -k 10 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r2s_${label}_n$N.json 2> gpurun_out/r2s_${label}_n$N.err
  fi
  python - <<PY
import json
try:
    txt=open("gpurun_out/r2s_${label}_n$N.json").read()
    d=json.loads([l for l in txt.split("\n") if l.startswith("{")][-1])
    print("$label N=$N", round(d["value"],3), "steps/s", round(d["ms_per_step"],2), "ms", d["config"]["ksp_iterations_per_step"], "its", d["config"]["us_per_ksp_iteration"], "us/it", "launches", d["gpu_launches"], "parity", (d.get("parity") or {}).get("residual_relerr"), "reason", d["config"]["ksp_reason"], d["config"]["newton_reason"], "e2e", round(d["e2e"]["value"],3))
    print("   ", d.get("ksp_breakdown_min_mean_max_over_ctas"))
except Exception as e:
    print("$label N=$N FAILED", e); print(open("gpurun_out/r2s_${label}_n$N.err").read()[-1500:])
PY
}
for cfg in $2; do
  if [ "$cfg" = "2" ]; then run c2 --no-cpu-baseline; fi
  if [ "$cfg" = "5" ]; then run c5 --config 5 --no-cpu-baseline; fi
  if [ "$cfg" = "tests" ]; then
    timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2s_multi_tests_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r2s_multi_tests_n$N.log
    tail -3 gpurun_out/r2s_multi_tests_n$N.log
  fi
  if [ "$cfg" = "ref" ]; then
    # the CPU arm launched the way the driver launches it at N > 1: rank 0 alone runs, on all host threads
    timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2s_ref_n$N.json 2> gpurun_out/r2s_ref_n$N.err
    grep -a '^{' gpurun_out/r2s_ref_n$N.json | tail -1 | cut -c1-300
  fi
done

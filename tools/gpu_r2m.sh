#!/bin/bash
# round-2 GPU job M (1 GPU): L2 prefetch distance of the Gram-Schmidt passes of the persistent kernel, and a source-level
# stall profile of that kernel (CSV of the source page only: the report itself is too large to travel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for pf in 0 2 3 5; do
  WB_FUSED_PF=$pf timeout -k 10 300 python tools/microbench.py --skip-pcs --fused-only > gpurun_out/r2m_pf$pf.json 2> gpurun_out/r2m_pf$pf.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2m_pf$pf.json"))["gmres30_full_fused"]
print("pf=$pf", d["its"], "its", round(d["us_per_it"],2), "us/it", {k: v for k, v in d["breakdown_us"].items()})
PY
done
timeout -k 10 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
WB_FUSED_PF=3 timeout -k 10 600 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -2
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gmres_fused' -s 1 -c 1 \
    -f -o gpurun_out/r2m_fused python bench.py --steps 1 --warmup 1 --ksp-maxit 90 --no-cpu-baseline --no-parity --spmv-launches 2 > gpurun_out/r2m_fused_bench.log 2>&1
echo "fused capture rc=$?"
ncu -i gpurun_out/r2m_fused.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2m_fused_source.csv 2>gpurun_out/r2m_source.err
ls -la gpurun_out/r2m_fused.ncu-rep
rm -f gpurun_out/r2m_fused.ncu-rep

#!/bin/bash
# Round-2 evidence run (1 GPU, under gpurun): launch list of the bench command, `ncu --set full` of the kernels on
# the Newton-step path, then the bench lines themselves (never under ncu).  Outputs under gpurun_out/r2f_*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader > gpurun_out/r2f_env.txt
# 1. launch list of the bench command (one pass per kernel, serialised: compare shares)
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --spmv-launches 5 > gpurun_out/r2f_launches_bench.log 2>&1
echo "launch list rc=$?"
# 2. full captures: the stand-alone SpMV, the assembly kernels, the factorisation; then the persistent solver kernel
#    (one launch is a whole KSPSolve: capped at 90 iterations so that the ~40 replays stay short)
timeout -k 10 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_sell_spmv|k_sell_fill|k_jacobian|k_residual|k_eos|k_ilu0_factor|k_ilu_repack' -s 12 -c 12 \
    -f -o gpurun_out/r2f_full python bench.py --steps 1 --warmup 1 --ksp-maxit 40 --no-cpu-baseline --no-parity --spmv-launches 5 > gpurun_out/r2f_full_bench.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/r2f_full.ncu-rep --page raw --csv > gpurun_out/r2f_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2f_full.ncu-rep   # tens of MB: only gpurun_out/ <= 64 MiB travels back
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gmres_fused' -s 1 -c 1 \
    -f -o gpurun_out/r2f_fused python bench.py --steps 1 --warmup 1 --ksp-maxit 90 --no-cpu-baseline --no-parity --spmv-launches 2 > gpurun_out/r2f_fused_bench.log 2>&1
echo "fused capture rc=$?"
ncu -i gpurun_out/r2f_fused.ncu-rep --page raw --csv > gpurun_out/r2f_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/r2f_fused.ncu-rep --page details > gpurun_out/r2f_fused_details.txt 2>/dev/null
rm -f gpurun_out/r2f_fused.ncu-rep
# the additive Schwarz experiment and its parity test
timeout -k 10 600 python -m pytest tests/test_gpu_linalg.py -x -q -k asm > gpurun_out/r2f_asm_tests.log 2>&1; tail -3 gpurun_out/r2f_asm_tests.log
timeout -k 10 600 python tools/microbench.py --asm --skip-pcs > gpurun_out/r2f_micro_asm.json 2> gpurun_out/r2f_micro_asm.err
echo "asm microbench rc=$?"; grep -a "asm1\|gmres30_full_fused" -A3 gpurun_out/r2f_micro_asm.json | cut -c1-200 | head -40
# 3. bench lines
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err
echo "bench c2 rc=$?"
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_ref_c2.json 2> gpurun_out/r2f_ref_c2.err
echo "reference arm rc=$?"
timeout -k 10 600 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c4.json 2> gpurun_out/r2f_bench_c4.err
echo "bench c4 rc=$?"
timeout -k 10 600 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2f_bench_c5.json 2> gpurun_out/r2f_bench_c5.err
echo "bench c5 rc=$?"
timeout -k 10 600 python bench.py --ksp bcgs --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_c2_bcgs.json 2> gpurun_out/r2f_bench_c2_bcgs.err
echo "bench c2 bcgs rc=$?"
for f in c2 c4 c5 c2_bcgs; do grep -a '^{' gpurun_out/r2f_bench_$f.json | tail -1 | cut -c1-400; done
grep -a '^{' gpurun_out/r2f_ref_c2.json | tail -1 | cut -c1-600

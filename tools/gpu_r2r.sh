#!/bin/bash
# `ncu --set full` of the stand-alone SpMV (k_sell_spmv) for configs 2 and 4: DRAM traffic per launch -> profiles/spmv_traffic.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for cfg in 2 4; do
  timeout -k 10 600 ncu --set full --clock-control none -k 'regex:k_sell_spmv' -s 3 -c 2 -f -o gpurun_out/r2r_spmv_c$cfg \
      python bench.py --config $cfg --steps 1 --warmup 1 --ksp-maxit 10 --no-cpu-baseline --no-parity --spmv-launches 6 > gpurun_out/r2r_spmv_c$cfg.log 2>&1
  echo "capture config $cfg rc=$?"
  ncu -i gpurun_out/r2r_spmv_c$cfg.ncu-rep --page raw --csv > gpurun_out/r2r_spmv_c${cfg}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2r_spmv_c$cfg.ncu-rep --page details > gpurun_out/r2r_spmv_c${cfg}_details.txt 2>/dev/null
  rm -f gpurun_out/r2r_spmv_c$cfg.ncu-rep
done
cuobjdump -sass -fun '_Z11k_sell_spmvILi2ELi128EEv8SellArgs' waiwera_b200/libwaiwera_b200.so 2>/dev/null | grep -c LDG

#!/bin/bash
# round-2 run J (1 GPU): launch list of the bench command and full ncu captures of the dominant kernels (after the plain run exited 0)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --skip strong,config5,parity,aw,e2e,cpu"
$B > gpurun_out/r2j_plain.json 2> gpurun_out/r2j_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv $B > gpurun_out/r2j_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'grid_dense_kernel|degrid_reg_kernel|bin_scatter_kernel|bin_hist_kernel' -s 8 -c 4 -o gpurun_out/r2j_prof -f $B > gpurun_out/r2j_ncu2.log 2>&1
ls -la gpurun_out | grep r2j

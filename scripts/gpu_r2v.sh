#!/bin/bash
# round-2 run V (1 GPU): ncu --set full of the S=31 kernels (config 5's dominant pair) in the dense regime, after the plain run exited 0
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --support 31 --nw 16 --vis 5e7 --skip strong,config5,parity,aw,e2e,cpu"
$B > gpurun_out/r2v_plain.json 2> gpurun_out/r2v_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:'grid_dense_kernel|degrid_tile_kernel' -s 6 -c 2 -o gpurun_out/r2v_prof -f $B > gpurun_out/r2v_ncu.log 2>&1
ls -la gpurun_out | grep r2v

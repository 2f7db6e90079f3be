#!/bin/bash
# round-2 run K (1 GPU): how the tap loads treat L1 (ncu: the dense gridder sits at 96 % of the L1TEX pipe), scatter occupancy
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --skip strong,config5,parity,aw,e2e,cpu"
$B > gpurun_out/r2k_ld0.json 2> gpurun_out/r2k_ld0.err
SKAGRID_TAP_LOAD=1 $B > gpurun_out/r2k_ld1.json 2> gpurun_out/r2k_ld1.err
SKAGRID_TAP_LOAD=2 $B > gpurun_out/r2k_ld2.json 2> gpurun_out/r2k_ld2.err
SKAGRID_SCATTER_OCC=6 $B > gpurun_out/r2k_occ6.json 2> gpurun_out/r2k_occ6.err
SKAGRID_SCATTER_OCC=8 $B > gpurun_out/r2k_occ8.json 2> gpurun_out/r2k_occ8.err
SKAGRID_TAP_LOAD=1 $B --support 31 --nw 16 --vis 5e7 > gpurun_out/r2k_s31_ld1.json 2> gpurun_out/r2k_s31_ld1.err
$B --support 31 --nw 16 --vis 5e7 > gpurun_out/r2k_s31_ld0.json 2> gpurun_out/r2k_s31_ld0.err
ls gpurun_out | grep -c r2k

#!/bin/bash
# round-2 run F (1 GPU): the per-GPU share of the strong-scaling step (1.25e7 visibilities) under the plan's tuning switches
mkdir -p gpurun_out
B="python bench.py --steps 10 --warmup 3 --vis 1.25e7 --skip strong,config5,parity,aw,e2e,cpu"
$B > gpurun_out/r2f_default.json 2> gpurun_out/r2f_default.err
SKAGRID_TILE=16 $B > gpurun_out/r2f_tile16.json 2> gpurun_out/r2f_tile16.err
SKAGRID_CELLSORT=0 $B > gpurun_out/r2f_nocell.json 2> gpurun_out/r2f_nocell.err
SKAGRID_TILE=16 SKAGRID_CELLSORT=0 $B > gpurun_out/r2f_tile16_nocell.json 2> gpurun_out/r2f_tile16_nocell.err
B="python bench.py --steps 10 --warmup 3 --vis 2.5e7 --skip strong,config5,parity,aw,e2e,cpu"
$B > gpurun_out/r2f_25_default.json 2> gpurun_out/r2f_25_default.err
SKAGRID_TILE=16 $B > gpurun_out/r2f_25_tile16.json 2> gpurun_out/r2f_25_tile16.err
ls gpurun_out | grep r2f

#!/bin/bash
# round-2 run G (2 GPUs): active-row plans, traced config-5 exchange stages, gridder variant 6 at the strong-scaling share
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip aw,e2e > gpurun_out/r2g_n2.json 2> gpurun_out/r2g_n2.err
echo "bench rc=$?"
B="python bench.py --steps 10 --warmup 3 --skip strong,config5,parity,aw,e2e,cpu"
$B > gpurun_out/r2g_n1.json 2> gpurun_out/r2g_n1.err
$B --vis 1.25e7 > gpurun_out/r2g_n1_small.json 2> gpurun_out/r2g_n1_small.err
$B --vis 1.25e7 --variant 6 > gpurun_out/r2g_n1_small_v6.json 2> gpurun_out/r2g_n1_small_v6.err
$B --variant 6 > gpurun_out/r2g_n1_v6.json 2> gpurun_out/r2g_n1_v6.err
tail -3 gpurun_out/r2g_n2.err

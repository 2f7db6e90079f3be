#!/bin/bash
# round-2 run H (2 GPUs): atomic-free routing kernels + time-balanced slabs (config 5), multi-rank parity again
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2h_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2h_mgpu_check.log
python -m pytest tests -m gpu -x -q -k "route or tile_sharded" 2>&1 | tail -3 >> gpurun_out/r2h_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --skip aw,e2e,strong,parity > gpurun_out/r2h_n2.json 2> gpurun_out/r2h_n2.err
echo "bench rc=$?" >> gpurun_out/r2h_mgpu_check.log
grep -v "^\*\*\*\|^$\|Warning\|warn\|OMP" gpurun_out/r2h_mgpu_check.log | tail -8; tail -3 gpurun_out/r2h_n2.err

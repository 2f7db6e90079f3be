#!/bin/bash
# round-2 run C (2 GPUs): multi-rank parity, the full N=2 bench line, A/B of the reduction forms
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2c_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2c_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c_n2.json 2> gpurun_out/r2c_n2.err
echo "bench rc=$?" >> gpurun_out/r2c_mgpu_check.log
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --skip config5,aw,e2e,parity --one-group > gpurun_out/r2c_n2_onegroup.json 2> gpurun_out/r2c_n2_onegroup.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --skip config5,aw,e2e,parity --allreduce > gpurun_out/r2c_n2_allreduce.json 2> gpurun_out/r2c_n2_allreduce.err
tail -4 gpurun_out/r2c_mgpu_check.log; tail -3 gpurun_out/r2c_n2.err

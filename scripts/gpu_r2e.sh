#!/bin/bash
# round-2 run E (2 GPUs): peer-memory exchange (CUDA IPC over NVLink) -- parity, then the bench line against the NCCL form
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2e_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2e_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip aw,e2e > gpurun_out/r2e_n2_peer.json 2> gpurun_out/r2e_n2_peer.err
echo "bench peer rc=$?" >> gpurun_out/r2e_mgpu_check.log
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --skip aw,e2e,parity --nccl > gpurun_out/r2e_n2_nccl.json 2> gpurun_out/r2e_n2_nccl.err
echo "bench nccl rc=$?" >> gpurun_out/r2e_mgpu_check.log
grep -v "^\*\*\*\|^$\|Warning\|warn" gpurun_out/r2e_mgpu_check.log | tail -8; tail -5 gpurun_out/r2e_n2_peer.err

#!/bin/bash
# round-2 run M (8 GPUs): parity + bench after the stream-pool / SM-gather / balance changes (no e2e)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2m_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2m_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --skip e2e,aw > gpurun_out/r2m_n8.json 2> gpurun_out/r2m_n8.err
echo "bench rc=$?" >> gpurun_out/r2m_mgpu_check.log
grep "world=\|rc=" gpurun_out/r2m_mgpu_check.log; tail -3 gpurun_out/r2m_n8.err

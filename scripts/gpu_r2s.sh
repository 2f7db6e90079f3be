#!/bin/bash
# round-2 run S (8 GPUs): config 4 weak + strong with the three all-gather forms
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for m in sm fused; do
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --skip aw,e2e,config5,parity --allgather $m > gpurun_out/r2s_n8_$m.json 2> gpurun_out/r2s_n8_$m.err
done
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2s_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2s_mgpu_check.log
grep "world=\|rc=" gpurun_out/r2s_mgpu_check.log

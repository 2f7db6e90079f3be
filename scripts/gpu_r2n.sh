#!/bin/bash
# round-2 run N (4 GPUs): parity at world 4, config-4 weak + strong with the SM transpose-gather
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2n_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2n_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --skip e2e,aw,config5 > gpurun_out/r2n_n4.json 2> gpurun_out/r2n_n4.err
echo "bench rc=$?" >> gpurun_out/r2n_mgpu_check.log
grep "world=\|rc=" gpurun_out/r2n_mgpu_check.log; tail -3 gpurun_out/r2n_n4.err

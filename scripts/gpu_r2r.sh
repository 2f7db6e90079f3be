#!/bin/bash
# round-2 run R (8 GPUs): exchange primitives with the all-peers-at-once gather kernels, parity, config 5 (traced) with SM pulls
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29510 scripts/bench_peer.py --mb 64 > gpurun_out/r2r_peer_n8_64.json 2> gpurun_out/r2r_peer_n8_64.err
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2r_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2r_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 --skip aw,e2e,strong,parity > gpurun_out/r2r_n8_c5.json 2> gpurun_out/r2r_n8_c5.err
echo "bench rc=$?" >> gpurun_out/r2r_mgpu_check.log
grep "world=\|rc=" gpurun_out/r2r_mgpu_check.log

#!/bin/bash
# round-2 run P (2 GPUs): exchange primitives at N=2, parity with the fused sum-and-broadcast, N=2 bench (config 4 only), config-5 pulls by CE vs SM
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29510 scripts/bench_peer.py --mb 256 > gpurun_out/r2p_peer_n2_256.json 2> gpurun_out/r2p_peer_n2_256.err
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2p_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2p_mgpu_check.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --skip aw,e2e,config5 > gpurun_out/r2p_n2.json 2> gpurun_out/r2p_n2.err
echo "bench rc=$?" >> gpurun_out/r2p_mgpu_check.log
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --skip aw,e2e,config5,parity --allgather ce > gpurun_out/r2p_n2_ce.json 2> gpurun_out/r2p_n2_ce.err
SKAGRID_PEER_PULL=sm timeout 900 $TR --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --skip aw,e2e,strong,parity > gpurun_out/r2p_n2_c5_sm.json 2> gpurun_out/r2p_n2_c5_sm.err
grep "world=\|rc=" gpurun_out/r2p_mgpu_check.log; tail -2 gpurun_out/r2p_n2.err

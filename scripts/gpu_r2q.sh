#!/bin/bash
# round-2 run Q (N GPUs): exchange primitives, then config 4 weak+strong with the fused all-reduce kernel vs reduce-scatter + copy-engine gather
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29510 scripts/bench_peer.py --mb 64 > gpurun_out/r2q_peer_n${N}_64.json 2> gpurun_out/r2q_peer_n${N}_64.err
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --skip aw,e2e,config5,parity > gpurun_out/r2q_n${N}_fused.json 2> gpurun_out/r2q_n${N}_fused.err
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip aw,e2e,config5,parity --allgather ce > gpurun_out/r2q_n${N}_ce.json 2> gpurun_out/r2q_n${N}_ce.err
timeout 900 $TR --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --skip aw,e2e,config5,parity --nccl > gpurun_out/r2q_n${N}_nccl.json 2> gpurun_out/r2q_n${N}_nccl.err
tail -2 gpurun_out/r2q_n${N}_fused.err

#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 scripts/bench_peer.py --mb 64 > gpurun_out/r2o_peer_n$1_64.json 2> gpurun_out/r2o_peer_n$1_64.err
timeout 300 $TR --master-port 29512 scripts/bench_peer.py --mb 512 > gpurun_out/r2o_peer_n$1_512.json 2> gpurun_out/r2o_peer_n$1_512.err
tail -2 gpurun_out/r2o_peer_n$1_64.err

#!/bin/bash
# round-2 run X (8 GPUs): slab all-gather with one block per SM beside the image stage (config 4 weak + strong), peer-memory test
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "peer_memory" 2>&1 | tail -2 > gpurun_out/r2x_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --skip aw,e2e,config5 > gpurun_out/r2x_n8.json 2> gpurun_out/r2x_n8.err
echo "bench rc=$?" >> gpurun_out/r2x_pytest.log
cat gpurun_out/r2x_pytest.log

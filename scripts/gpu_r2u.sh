#!/bin/bash
# round-2 run U (2 GPUs): the whole GPU test-suite on a 2-GPU box (the torchrun test is not skipped there), then the N=2 bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2u_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip aw > gpurun_out/r2u_n2.json 2> gpurun_out/r2u_n2.err
echo "bench rc=$?" >> gpurun_out/r2u_pytest.log
cat gpurun_out/r2u_pytest.log

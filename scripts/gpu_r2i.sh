#!/bin/bash
# round-2 run I (2 GPUs): full GPU test-suite, multi-rank parity, N=2 bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2i_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_check.py > gpurun_out/r2i_mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> gpurun_out/r2i_pytest.log
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --skip aw,e2e > gpurun_out/r2i_n2.json 2> gpurun_out/r2i_n2.err
echo "bench rc=$?" >> gpurun_out/r2i_pytest.log
cat gpurun_out/r2i_pytest.log; grep "world=" gpurun_out/r2i_mgpu_check.log; tail -3 gpurun_out/r2i_n2.err

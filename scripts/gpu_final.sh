#!/bin/bash
# final measurements of the round: scripts/gpu_final.sh N  (N = 1: GPU test-suite, smoke, both bench arms; N > 1: torchrun bench)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/final_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/final_pytest.log 2>&1
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference_arm.json 2> gpurun_out/final_reference_arm.err
  python bench.py --steps 20 --warmup 5 > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err
  echo "bench rc=$?" >> gpurun_out/final_pytest.log
  python bench.py --steps 5 --warmup 3 --uniform --skip strong,config5,parity,aw,e2e,cpu > gpurun_out/final_n1_uniform.json 2> gpurun_out/final_n1_uniform.err
  cat gpurun_out/final_pytest.log
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/final_n$N.json 2> gpurun_out/final_n$N.err
  echo "bench rc=$?"
  tail -2 gpurun_out/final_n$N.err
fi

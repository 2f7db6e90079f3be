#!/bin/bash
# round-2 run B: GPU tests with the new routing / slab tests, full N=1 bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_n1.json 2> gpurun_out/r2b_n1.err
echo "bench rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log; tail -5 gpurun_out/r2b_n1.err

#!/bin/bash
# round-2 run T (1 GPU): paired record broadcasts in the dense gridder (variant 7) against the default, 20 sustained steps each; new GPU test
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "peer_memory" 2>&1 | tail -3 > gpurun_out/r2t_pytest.log
B="python bench.py --steps 20 --warmup 5 --skip strong,config5,aw,e2e,cpu"
$B --variant 7 > gpurun_out/r2t_v7.json 2> gpurun_out/r2t_v7.err
$B > gpurun_out/r2t_v0.json 2> gpurun_out/r2t_v0.err
$B --variant 7 --support 31 --nw 16 --vis 5e7 > gpurun_out/r2t_s31_v7.json 2> gpurun_out/r2t_s31_v7.err
$B --support 31 --nw 16 --vis 5e7 > gpurun_out/r2t_s31_v0.json 2> gpurun_out/r2t_s31_v0.err
cat gpurun_out/r2t_pytest.log

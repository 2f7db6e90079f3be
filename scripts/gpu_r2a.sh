#!/bin/bash
# round-2 run A: parity of the dense gridder / block-broadcast degridder, then A/B of the gridder variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/r2a_v0.json 2> gpurun_out/r2a_v0.err
$B --variant 5 > gpurun_out/r2a_v5.json 2> gpurun_out/r2a_v5.err
SKAGRID_DENSE=0 $B > gpurun_out/r2a_old.json 2> gpurun_out/r2a_old.err
$B --support 31 --nw 16 --vis 5e7 > gpurun_out/r2a_s31_v0.json 2> gpurun_out/r2a_s31_v0.err
$B --support 31 --nw 16 --vis 5e7 --variant 5 > gpurun_out/r2a_s31_v5.json 2> gpurun_out/r2a_s31_v5.err
SKAGRID_DENSE=0 $B --support 31 --nw 16 --vis 5e7 > gpurun_out/r2a_s31_old.json 2> gpurun_out/r2a_s31_old.err
tail -3 gpurun_out/r2a_pytest.log

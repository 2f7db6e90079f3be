#!/bin/bash
# final N=1 validation after the AW formation kernel change, plus its ncu capture
bash scripts/gpu_final.sh 1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_pair --launch-skip 1 -c 1 -o gpurun_out/z_conv_pair -f python scripts/bench_aw.py 100000 > gpurun_out/z_ncu.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
# round-2 run W (1 GPU): complex-to-real grid -> image against the Z2Z route: GPU test-suite, then 20 sustained steps each way
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2w_pytest.log
B="python bench.py --steps 20 --warmup 5 --skip strong,config5,aw,e2e,cpu"
$B > gpurun_out/r2w_c2r.json 2> gpurun_out/r2w_c2r.err
SKAGRID_G2I_Z2Z=1 $B > gpurun_out/r2w_z2z.json 2> gpurun_out/r2w_z2z.err
cat gpurun_out/r2w_pytest.log; tail -2 gpurun_out/r2w_c2r.err

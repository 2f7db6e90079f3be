#!/bin/bash
# launch list of the AW path at 1e6 visibilities (8 chunks): which kernels make up a chunk
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/aa_launches_aw.csv python scripts/bench_aw.py 1000000 > gpurun_out/aa_ncu.log 2>&1
echo "rc=$?"
python profiles/launch_summary.py gpurun_out/aa_launches_aw.csv | head -30

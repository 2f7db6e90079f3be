#!/bin/bash
# row-pair AW formation kernel: parity tests, A/B timing against the round-1 kernel, ncu figures
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "convolve2d or aw_gridding or smalltest" 2>&1 | tail -4 > gpurun_out/y_pytest.log
cat gpurun_out/y_pytest.log
python scripts/bench_aw.py 1000000 > gpurun_out/y_aw_pair_1e6.json 2> gpurun_out/y_aw_pair_1e6.err
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max
timeout 120 ncu --metrics $M --clock-control none -k regex:conv_ --launch-skip 1 -c 1 --csv --log-file gpurun_out/y_ncu_pair.csv python scripts/bench_aw.py 100000 > gpurun_out/y_ncu.log 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/y_aw_pair_1e6.json",):
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["device_ms"], d["device_ms_pinned"], d["wall_ms_pinned"])
PY
grep -v "^==" gpurun_out/y_ncu_pair.csv | cut -d, -f5,13- | tail -12

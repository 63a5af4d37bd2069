#!/bin/bash
# round-2 re-entry, final GPU call: full GPU suite, smoke, the driver's bench command, launch list + ncu capture for profiles/
mkdir -p gpurun_out
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/f_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_steps20.json 2> gpurun_out/f_bench_steps20.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/f_bench_steps20.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms | launches %d | cpu %.2f M/s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"], d["gpu_launches"], d["cpu_baseline"]["value"] / 1e6))
    print(d["roofline"]); print(d["clocks"])
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2b_launches_step.csv python scripts/profile_step.py 2 > gpurun_out/f_launchlist.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'bag_forward|phase1|sort_onesweep|sort_histogram' -o gpurun_out/r2b_fwd_bwd_sort python scripts/profile_step.py 1 > gpurun_out/f_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv

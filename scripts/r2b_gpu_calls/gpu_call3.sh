#!/bin/bash
# round-2 re-entry, GPU call 3: in-pipeline A/B of the grid knobs with the new defaults
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "forward or planned or window_plan or backward" > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/c3_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-ab --ab-reps 3 \
  --ab "base:;sort0:CEBAG_SORT_CTAS=0;sort74:CEBAG_SORT_CTAS=74;sort111:CEBAG_SORT_CTAS=111;sort222:CEBAG_SORT_CTAS=222;fwd16:CEBAG_FWD_CTAS_PER_SM=16;bwd16:CEBAG_BWD_CTAS_PER_SM=16;bwd64:CEBAG_BWD_CTAS_PER_SM=64;bwdu8:CEBAG_BWD_UNROLL=8;prep4:CEBAG_PREP_CTAS_PER_SM=4;prep2:CEBAG_PREP_CTAS_PER_SM=2;plan2:CEBAG_PLAN_CTAS_PER_SM=2,CEBAG_SORT_HIST_CTAS_PER_SM=1;swap28:CEBAG_SWAP_CTAS=28;swap112:CEBAG_SWAP_CTAS=112;cprio:COMPUTE_PRIORITY=-2,PRIORITY=0;items16:CEBAG_SORT_ITEMS=16" \
  > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/c3_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c3_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms | e2e ab %s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"], d["e2e"].get("ab_ms_per_step")))
    for k, v in d.get("ab", {}).items():
        print("  %-10s median %.4f  %s  %s" % (k, v["median"], v["ms_per_step"], v["settings"]))
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
    print(d["roofline"])
except Exception as e:
    print("bench parse failed", e)
PY

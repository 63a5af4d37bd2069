#!/bin/bash
# round-2 re-entry, GPU call 4: is the step bound by the window's side-stream chain?  step trace + early-done / early-submit A/B
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "lookahead or planned or window_plan or reference_loop" > gpurun_out/c4_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/c4_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --trace-steps --ab-reps 4 \
  --ab "base:;late_done:EARLY_DONE=0;before:SUBMIT_BEFORE=1;before_late:SUBMIT_BEFORE=1,EARLY_DONE=0;items16:CEBAG_SORT_ITEMS=16;prep4:CEBAG_PREP_CTAS_PER_SM=4;fwdu8:CEBAG_FWD_UNROLL=8;i16p4:CEBAG_SORT_ITEMS=16,CEBAG_PREP_CTAS_PER_SM=4;i16p4b:CEBAG_SORT_ITEMS=16,CEBAG_PREP_CTAS_PER_SM=4,SUBMIT_BEFORE=1" \
  > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench rc=$?"; grep -E "^step|first timed" gpurun_out/c4_bench.err | head -30; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c4_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]))
    for k, v in d.get("ab", {}).items():
        print("  %-12s median %.4f  %s  %s" % (k, v["median"], v["ms_per_step"], v["settings"]))
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY

#!/bin/bash
# round-2 re-entry, GPU call 5: CTA-size knobs of forward / backward phase 1 in the pipeline
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "forward_kernel_variants or deterministic" > gpurun_out/c5_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/c5_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --trace-steps --ab-reps 5 \
  --ab "base:;bwd128:CEBAG_BWD_THREADS=128;fwd128:CEBAG_FWD_THREADS=128;both128:CEBAG_BWD_THREADS=128,CEBAG_FWD_THREADS=128;bwd128u8:CEBAG_BWD_THREADS=128,CEBAG_BWD_UNROLL=8;prep4:CEBAG_PREP_CTAS_PER_SM=4;b128p4:CEBAG_BWD_THREADS=128,CEBAG_PREP_CTAS_PER_SM=4" \
  > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
echo "bench rc=$?"; grep -E "^step" gpurun_out/c5_bench.err | head -24; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c5_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]))
    for k, v in d.get("ab", {}).items():
        print("  %-12s median %.4f  %s  %s" % (k, v["median"], v["ms_per_step"], v["settings"]))
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY

#!/bin/bash
# round-2 re-entry, GPU call 6: pipeline timeline (steps vs side / copy stream) + write-back A/B
mkdir -p gpurun_out
timeout 600 python bench.py --steps 28 --warmup 5 --no-cpu-baseline --trace-steps --ab-reps 3 \
  --ab "base:;dma:DMA=1;swap28:CEBAG_SWAP_CTAS=28;b128p4:CEBAG_BWD_THREADS=128,CEBAG_PREP_CTAS_PER_SM=4" \
  > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
echo "bench rc=$?"; grep -E "^step|^window" gpurun_out/c6_bench.err | head -50; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c6_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]))
    for k, v in d.get("ab", {}).items():
        print("  %-12s median %.4f  %s  %s" % (k, v["median"], v["ms_per_step"], v["settings"]))
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY

#!/bin/bash
# round-2 re-entry, GPU call 7: long-region A/B (64 timed steps = 8 windows per sample) of the remaining grid knobs
mkdir -p gpurun_out
timeout 600 python bench.py --steps 64 --warmup 8 --no-cpu-baseline --ab-reps 4 \
  --ab "base:;prep4:CEBAG_PREP_CTAS_PER_SM=4;prep2:CEBAG_PREP_CTAS_PER_SM=2;prep1:CEBAG_PREP_CTAS_PER_SM=1;b128:CEBAG_BWD_THREADS=128;b128p4:CEBAG_BWD_THREADS=128,CEBAG_PREP_CTAS_PER_SM=4;dma:DMA=1;sort222:CEBAG_SORT_CTAS=222;sort0:CEBAG_SORT_CTAS=0" \
  > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/c7_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c7_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"]))
    for k, v in d.get("ab", {}).items():
        print("  %-12s median %.4f  %s  %s" % (k, v["median"], v["ms_per_step"], v["settings"]))
except Exception as e:
    print("bench parse failed", e)
PY

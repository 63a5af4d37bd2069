#!/bin/bash
# round-2 re-entry, GPU call 1: staged ids, probe filter A/B, ncu of the cache manager's kernels
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k "lookahead or prepare_ids or capacity or rejected or kat or reference_loop or training_under or freq_aware" > gpurun_out/c1_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c1_tests.log
timeout 200 python scripts/profile_prepare.py 4 > gpurun_out/c1_prep_filter1.log 2>&1; tail -12 gpurun_out/c1_prep_filter1.log
CEBAG_PROBE_FILTER=0 timeout 200 python scripts/profile_prepare.py 4 > gpurun_out/c1_prep_filter0.log 2>&1; tail -12 gpurun_out/c1_prep_filter0.log
timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-ab \
  --ab "base:;filter0:CEBAG_PROBE_FILTER=0;prio0:PRIORITY=0;pctas2:CEBAG_PROBE_CTAS_PER_SM=2;pctas4:CEBAG_PROBE_CTAS_PER_SM=4;swap28:CEBAG_SWAP_CTAS=28;swap112:CEBAG_SWAP_CTAS=112;base2:" \
  > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/c1_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c1_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms | other e2e order %s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"], d["e2e"].get("ab_other_order")))
    print({k: round(v["ms_per_step"], 4) for k, v in d.get("ab", {}).items()})
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:'probe_kernel|lfu_count|select_hist|bitmap_count|bitmap_emit|free_flags|count_evictable|stamp_hits|count_hits' \
  -o gpurun_out/c1_prepare python scripts/profile_prepare.py 1 > gpurun_out/c1_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
timeout 60 python scripts/hbm_rw_probe.py | tee gpurun_out/c1_hbm.log

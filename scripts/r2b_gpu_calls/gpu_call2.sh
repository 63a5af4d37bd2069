#!/bin/bash
# round-2 re-entry, GPU call 2: spread hit flags, forward L1 policy / interleaved strips, narrow side-stream grids
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/c2_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c2_tests.log
timeout 300 python scripts/profile_prepare.py 3 "m1s4096:;m0s4096:CEBAG_PROBE_MODE=0;m0s1:CEBAG_PROBE_MODE=0,CEBAG_FLAG_SPREAD=1;m1s1:CEBAG_FLAG_SPREAD=1;m1s256:CEBAG_FLAG_SPREAD=256;m1s65536:CEBAG_FLAG_SPREAD=65536;prep2:CEBAG_PREP_CTAS_PER_SM=2;prep1:CEBAG_PREP_CTAS_PER_SM=1" > gpurun_out/c2_prepare.log 2>&1
cat gpurun_out/c2_prepare.log | tail -20
timeout 300 python scripts/profile_step.py 10 65536 "ld1:;ld0:CEBAG_FWD_LD=0;ld2:CEBAG_FWD_LD=2;ilv_ld1:CEBAG_FWD_ILV=1;ilv_ld0:CEBAG_FWD_ILV=1,CEBAG_FWD_LD=0;ilv_u8:CEBAG_FWD_ILV=1,CEBAG_FWD_UNROLL=8;u8:CEBAG_FWD_UNROLL=8;ctas8:CEBAG_FWD_CTAS_PER_SM=8;ilv_ctas8:CEBAG_FWD_ILV=1,CEBAG_FWD_CTAS_PER_SM=8;ctas32:CEBAG_FWD_CTAS_PER_SM=32;sort148:CEBAG_SORT_CTAS=148;sort296:CEBAG_SORT_CTAS=296;ld1b:" > gpurun_out/c2_step.log 2>&1
grep -E "^\[|bag_forward|radix_sort" gpurun_out/c2_step.log | tail -45
timeout 500 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-ab --ab-reps 3 \
  --ab "base:;fwdld0:CEBAG_FWD_LD=0;ilv:CEBAG_FWD_ILV=1;prep2:CEBAG_PREP_CTAS_PER_SM=2;prep1:CEBAG_PREP_CTAS_PER_SM=1;sort148:CEBAG_SORT_CTAS=148;narrow:CEBAG_PREP_CTAS_PER_SM=2,CEBAG_SORT_CTAS=148;prio0:PRIORITY=0;oldprobe:CEBAG_PROBE_MODE=0,CEBAG_FLAG_SPREAD=1" \
  > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/c2_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c2_bench.json").read().strip().splitlines()[-1])
    print("value %.3f G/s %.3f ms | e2e %.3f G/s %.3f ms | e2e ab %s" % (d["value"] / 1e9, d["ms_per_step"], d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"], d["e2e"].get("ab_ms_per_step")))
    for k, v in d.get("ab", {}).items():
        print("  %-10s median %.4f  %s  %s" % (k, v["median"], v["ms_per_step"], v["settings"]))
    print({k: round(v["ms_per_step"] * 1e3, 1) for k, v in d["kernels"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'bag_forward' \
  -o gpurun_out/c2_forward python scripts/profile_step.py 1 65536 "ld1:;ld0:CEBAG_FWD_LD=0;ilv:CEBAG_FWD_ILV=1" > gpurun_out/c2_ncu_fwd.log 2>&1
echo "ncu fwd rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'probe_kernel|collect_hits|lfu_count' \
  -o gpurun_out/c2_probe python scripts/profile_prepare.py 1 > gpurun_out/c2_ncu_probe.log 2>&1
echo "ncu probe rc=$?"; ls -la gpurun_out/*.ncu-rep

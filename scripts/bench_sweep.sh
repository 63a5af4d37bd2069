#!/bin/bash
# usage: scripts/bench_sweep.sh "<ENV=.. ENV=..>" [bench args]; prints a one-line summary of a bench.py run
envs="$1"; shift
out=$(env $envs python bench.py --no-cpu-baseline "$@" 2>&1 | tail -1)
python - "$envs" "$*" <<PY
import json,sys
try:
    d=json.loads('''$out''')
    print(sys.argv[1], sys.argv[2], "| value %.3f G/s e2e %.3f G/s ms/step %.3f e2e_ms %.3f" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["e2e"]["ms_per_step"]))
    print("    ", {k: round(v["ms_per_step"]*1e3,1) for k,v in d["kernels"].items()})
except Exception as e:
    print("ERR", e, '''$out'''[-1500:])
PY

"""DRAM traffic per launch of the forward / backward kernels from an `ncu --set full` capture -> profiles/ncu_traffic.json
(the `roofline.traffic` field of bench.py reads that file; nothing is pasted into bench.py).

Recipe (on a B200, via gpurun; one batch of the Criteo-1TB launch shape, n = 1,703,936 lookups):
    ncu --set full --clock-control none --import-source on -k regex:"bag_forward|phase1" -s 6 -c 2 \
        -o gpurun_out/r2_fwd_bwd python scripts/profile_step.py 2
then here (no GPU needed):
    python scripts/ncu_traffic.py gpurun_out/r2_fwd_bwd.ncu-rep criteo1tb:n1:table
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"bag_forward": "bag_forward", "bag_backward_phase1": "bag_backward_phase1"}


def main(rep, key):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rec = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        for frag, label in NAMES.items():
            if frag in name and label not in rec:
                rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
                wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
                rec[label] = rd + wr
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = dict(rec, source=os.path.basename(rep), unit="bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)")
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(data[key]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

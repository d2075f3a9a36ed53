"""profiles/r02_traffic.json from an `ncu --set full` report of `python bench.py [...]`: DRAM bytes (read + write) per
launch of every kernel family of the chain, averaged over the captured launches.
    python scratch/ncu_traffic.py gpurun_out/r02_full.ncu-rep nsym seg start err > profiles/r02_traffic.json"""
import csv
import json
import subprocess
import sys

rep, nsym, seg, start, err = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5] == "1"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
ci = {n: i for i, n in enumerate(hdr)}
fam = {"train": "train_la_kernel", "apply": "apply_2x2_os2_kernel", "bps": "bps_fast_kernel"}
acc = {k: [] for k in fam}
dur = {k: [] for k in fam}
for r in rows[2:]:
    name = r[ci["Kernel Name"]]
    for k, pat in fam.items():
        if pat in name:
            rd, wr = float(r[ci["dram__bytes_read.sum"]]), float(r[ci["dram__bytes_write.sum"]])
            unit_r, unit_w = rows[1][ci["dram__bytes_read.sum"]], rows[1][ci["dram__bytes_write.sum"]]
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            acc[k].append(rd * mult.get(unit_r, 1.0) + wr * mult.get(unit_w, 1.0))
            dur[k].append(float(r[ci["gpu__time_duration.sum"]]))
# the chain runs twice per step: the main group of segments and the one end-aligned extra segment (a 1-CTA launch);
# the roofline is about the main launches -> keep, per family, the launches within 2x of the largest traffic
for k in acc:
    if acc[k]:
        top = max(acc[k])
        sel = [i for i, v in enumerate(acc[k]) if v * 2 > top]
        acc[k] = [acc[k][i] for i in sel]
        dur[k] = [dur[k][i] for i in sel]
print(json.dumps({"workload": {"nsym": nsym, "seg": seg, "start": start, "err": err},
                  "source": rep.split("/")[-1] + " (ncu --set full --clock-control none; per launch, mean over captured launches)",
                  "kernels": {k: (sum(v) / len(v) if v else None) for k, v in acc.items()},
                  "launches_captured": {k: len(v) for k, v in acc.items()},
                  "duration_under_ncu": {k: (sum(v) / len(v) if v else None) for k, v in dur.items()}}, indent=1))

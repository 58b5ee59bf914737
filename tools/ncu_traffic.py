#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --page raw --csv` dump: DRAM bytes per launch of every kernel (first launch of each
name), keyed by the names bench.py uses.  usage: ncu_traffic.py <raw.csv> <out.json>"""
import csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = re.sub(r"<.*", "", r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")).strip()
    name = {"k_inject_linear": "k_inject", "k_clear_masked": "k_clear", "k_transfer_masked": "k_transfer"}.get(name, name)
    if name in out:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]]) * scale.get(units[ix[m]], 1)
    out[name] = int(tot)
json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)
print(out)

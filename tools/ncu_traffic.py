#!/usr/bin/env python3
"""profiles/traffic.json from an `ncu --page raw --csv` dump: DRAM bytes per launch of every kernel (LAST launch of each
name: the capture runs three frames and the last one is a steady-state sparse frame), keyed by the names bench.py uses; for k_cone_trace also the TEX data-pipe wavefronts of the launch (bench.py's
roofline.tex_pipe; "k_cone_trace__cone_steps" = the cone steps of the captured frame, from the bench line of the same build, is
kept from the previous file).  usage: ncu_traffic.py <raw.csv> <out.json>"""
import csv, json, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))       # (an ncu --log-file starts with ==PROF== lines)
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
body = rows[2:]
starts = [i for i, r in enumerate(body) if "k_upload_frame" in r[ix["Kernel Name"]]]
if starts:                                      # only the last frame of the capture (a frame starts with its parameter upload)
    body = body[starts[-1]:]
for r in body:
    name = re.sub(r"<.*", "", r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")).strip()
    name = {"k_inject_linear": "k_inject", "k_clear_masked": "k_clear", "k_frame_begin": "k_clear", "k_transfer_masked": "k_transfer"}.get(name, name)
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]]) * scale.get(units[ix[m]], 1)
    out[name] = int(tot)
    wf = "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum"
    if name == "k_cone_trace" and wf in ix:
        out["k_cone_trace__tex_wavefronts"] = int(float(r[ix[wf]]))
try:
    out["k_cone_trace__cone_steps"] = json.load(open(sys.argv[2]))["k_cone_trace__cone_steps"]
except Exception:
    pass
json.dump(out, open(sys.argv[2], "w"), indent=1, sort_keys=True)
print(out)

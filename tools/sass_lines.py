#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.
usage: sass_lines.py <ncu --page source --csv dump> <nvdisasm -g -c dump> <kernel substring> [top]"""
import csv, re, sys, collections
src_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: sequence of (line, inlined-at chain) per instruction of the kernel
seq = []; cur = None; inside = False
for l in open(dis):
    if l.startswith('.text.') and kern in l: inside = True; continue
    if inside and l.startswith('//----'): break
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): seq.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith('0x'): break       # the dump repeats the table per view: keep the first
    data.append(r)
print("sass rows", len(data), "disasm instr", len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
for k, r in enumerate(data):
    loc = seq[k] if k < len(seq) and seq[k] else ('?', 0, '')
    key = (loc[0], loc[1])
    agg[key][0] += int(r[ix['Instructions Executed']] or 0)
    agg[key][1] += int(r[ix['# Samples']] or 0)
    agg[key][2] += 1
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
lines = {}
for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    if f not in lines:
        try: lines[f] = open('vct_b200/csrc/' + f).read().splitlines()
        except Exception: lines[f] = []
    text = lines[f][n - 1].strip()[:110] if 0 < n <= len(lines[f]) else ''
    print(f"{100*v[1]/max(ts,1):5.1f}% smp {100*v[0]/max(ti,1):5.1f}% inst {v[2]:5d} sass  {f}:{n:<4d} {text}")

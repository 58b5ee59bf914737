#!/usr/bin/env python3
"""Print registers / spills / shared memory per kernel from the ptxas logs of the last build."""
import glob, re, subprocess, sys
for f in sorted(glob.glob("vct_b200/lib/obj/*.ptxas.log")):
    name = None
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m: stack = m.groups()
        m = re.search(r"Used (\d+) registers(.*)", line)
        if m and name:
            print(f"{name:60s} regs={m.group(1):>3s} stack/spill={'/'.join(stack)} {m.group(2).strip(', ')}")
            name = None

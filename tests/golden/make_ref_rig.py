#!/usr/bin/env python3
"""Regenerates tests/golden/warp_rig_ref.json from the REFERENCE'S OWN code: oracle/_ref/warp_rig is the `#if 0` warp
example of /root/reference/src/main.cpp:21-128 compiled in place by oracle/Makefile (needs /root/reference; the
fixture travels to boxes that do not have it).  Outputs are what the reference's C++ computes, printed with %.9g
(round-trip exact for binary32)."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
RIG = os.path.join(HERE, "..", "..", "oracle", "_ref", "warp_rig")


def run(grid):
    out = subprocess.run([RIG, str(grid)], check=True, capture_output=True, text=True).stdout
    r = {"cells": [], "partials": [], "weights": [], "warp": []}
    for line in out.splitlines():
        k, *v = line.split()
        if k == "dim": r["dim"] = int(v[0])
        elif k == "cell": r["cells"].append(float(v[2]))
        elif k == "partial": r["partials"].append([int(v[2]), int(v[3])])
        elif k == "weight": r["weights"].append([float(v[1]), float(v[2])])
        elif k == "warp": r["warp"].append([float(x) for x in v])
    return r


if __name__ == "__main__":
    g = run(24)
    g["source"] = "sfreed141/vct src/main.cpp:21-128 compiled by oracle/Makefile (_ref/warp_rig), run with grid 24"
    with open(os.path.join(HERE, "warp_rig_ref.json"), "w") as f:
        json.dump(g, f)
    print(len(g["warp"]), "points")

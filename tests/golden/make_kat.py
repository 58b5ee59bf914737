#!/usr/bin/env python3
"""Regenerates tests/golden/kat.json: known answers for the oracle, from an INDEPENDENT pure-Python (numpy fp32)
restatement of the reference lines they come from.  The reference stores no expected outputs anywhere
(SURVEY.md §4); these are what its own arithmetic yields:

  rgba8_avg      shaders/voxelize.frag:111-139   imageAtomicRGBA8Avg, sequential semantics
  atomic_max     shaders/voxelize.frag:271-274   packUnorm4x8 + imageAtomicMax
  weight_table   src/Application.cpp:346-370     warp weights for k occupied cells of 32
  warp_rig       src/main.cpp:20-127             the `#if 0` 4x4 worked example (fixed l = 0.5)
  cone_fog       shaders/phong.frag:135-180      step count of a 60-degree cone from the volume centre
  diffuse_schedule                                 h and lambda per step of the diffuse cone

Run:  python tests/golden/make_kat.py      (writes kat.json next to this file; no reference checkout needed)
"""
import json
import math
import os

import numpy as np

f32 = np.float32


def rgba8_avg_py(stored, r, g, b):
    """voxelize.frag:111-139 with uint(float) = truncation and `& 0xFF` packing."""
    val = [f32(r) * f32(255), f32(g) * f32(255), f32(b) * f32(255), f32(1)]
    if stored != 0:
        rv = [f32(stored & 255), f32((stored >> 8) & 255), f32((stored >> 16) & 255), f32(stored >> 24)]
        rv[0] *= rv[3]; rv[1] *= rv[3]; rv[2] *= rv[3]
        cur = [rv[i] + val[i] for i in range(4)]
        cur[0] = cur[0] / cur[3]; cur[1] = cur[1] / cur[3]; cur[2] = cur[2] / cur[3]
        val = cur
    u = [int(v) & 255 for v in val]
    return u[3] << 24 | u[2] << 16 | u[1] << 8 | u[0]


def pack_unorm(c):
    out = 0
    for i, v in enumerate(c):
        v = min(max(f32(v), f32(0)), f32(1))
        out |= int(np.rint(v * f32(255))) << (8 * i)
    return out


def weight_table(dim, high, low):
    lo, hi = [], []
    for occ in range(dim + 1):
        if occ in (0, dim):
            lo.append(1.0); hi.append(1.0); continue
        empty = dim - occ
        h = f32(high); l = (f32(dim) - h * f32(occ)) / f32(empty)
        if l < f32(low):
            l = f32(low); h = (f32(dim) - l * f32(empty)) / f32(occ)
        lo.append(float(l)); hi.append(float(h))
    return lo, hi


def warp_rig():
    n = 4
    cells = [[0, 1, 1, 1], [0, 1, 0, 1], [0, 1, 1, 0], [0, 1, 0, 0]]
    px = [[0] * n for _ in range(n)]; py = [[0] * n for _ in range(n)]
    for row in range(n):
        s = 0
        for x in range(n):
            s += 1 if cells[row][x] > 0.5 else 0; px[row][x] = s
    for col in range(n):
        s = 0
        for y in range(n):
            s += 1 if cells[y][col] > 0.5 else 0; py[y][col] = s
    wl, wh = [0.0] * (n + 1), [0.0] * (n + 1)
    for occ in range(n + 1):
        if occ in (0, n):
            wl[occ] = wh[occ] = 1.0
        else:
            l = f32(0.5); empty = n - occ
            wl[occ] = float(l); wh[occ] = float((f32(n) - l * f32(empty)) / f32(occ))

    def warp(tc):
        lt = [f32(tc[0]) * f32(n), f32(tc[1]) * f32(n)]
        idx = [f32(np.trunc(v)) for v in lt]; pos = [lt[i] - idx[i] for i in range(2)]
        x, y = int(idx[0]), int(idx[1])
        occd = cells[y][x] > 0.5
        tot = [px[y][n - 1], py[n - 1][x]]; part = [px[y][x], py[y][x]]
        out = []
        for k in range(2):
            lo, hi = f32(wl[tot[k]]), f32(wh[tot[k]])
            res = hi if occd else lo
            prev = f32(part[k]) - f32(1) if occd else f32(part[k])
            off = lo * (idx[k] - prev) + hi * prev
            out.append(float((off + pos[k] * res) / f32(n)))
        return out

    pts = [(0.1, 0.1), (0.3, 0.1), (0.6, 0.1), (0.9, 0.1), (0.375, 0.375), (0.625, 0.375), (0.875, 0.375), (0.625, 0.625),
           (0.875, 0.875), (0.5, 0.5)]
    return {"cells": cells, "partials_x": px, "partials_y": py, "weights_low": wl, "weights_high": wh,
            "points": [[list(p), warp(p)] for p in pts]}


def main():
    seq, w = [], 0
    for c in [(1.0, 0.5, 0.25), (0.0, 0.0, 0.0), (0.2, 0.4, 0.6)]:
        w = rgba8_avg_py(w, *c); seq.append(f"0x{w:08X}")
    order = []
    for o in [(0.1, 0.1, 0.3), (0.1, 0.3, 0.1), (0.3, 0.1, 0.1)]:
        w = 0
        for v in o:
            w = rgba8_avg_py(w, v, 0, 0)
        order.append([list(o), w & 255])
    grey, w = [], 0
    for i in range(257):
        w = rgba8_avg_py(w, 0.5, 0.5, 0.5)
        if i in (0, 254, 255, 256):
            grey.append(f"0x{w:08X}")
    lo, hi = weight_table(32, 2.0, 0.5)
    h, t = f32(1.0), f32(math.tan(f32(math.radians(60.0)) / f32(2)))
    hs, lam = [], []
    for _ in range(8):
        r = h * t
        hs.append(round(float(h), 4)); lam.append(round(float(np.log2(max(f32(1), f32(2) * r)) + f32(0.5)), 4)); h = h + r
    h, n = 1.0, 0
    while n < 16 and 0.5 + (1.0 + h) / 256.0 <= 1.0:
        h += h * math.tan(math.radians(30.0)); n += 1
    gold = {
        "rgba8_avg": {"abc_words": seq, "order_dependence": order, "grey_wrap": grey},
        "atomic_max": {"words": [f"0x{pack_unorm(c + (1.0,)):08X}" for c in [(1.0, 0.5, 0.25), (0.0, 0.0, 0.0), (0.2, 0.4, 0.6)]]},
        "weight_table_32": {str(k): [round(lo[k], 6), round(hi[k], 6)] for k in (0, 1, 4, 8, 10, 11, 16, 24, 31, 32)},
        "warp_rig": warp_rig(),
        "cone_fog": {"steps_256": n},
        "diffuse_schedule": {"h": hs, "lambda": lam},
    }
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
        json.dump(gold, f, indent=1)
    return gold


if __name__ == "__main__":
    g = main()
    print(json.dumps(g["rgba8_avg"]), json.dumps(g["weight_table_32"]), json.dumps(g["warp_rig"]["points"]), sep="\n")

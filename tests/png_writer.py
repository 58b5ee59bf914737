"""Minimal PNG encoder for the ingest tests: writes seeded synthetic images of every colour type / bit depth the format
allows, with mixed row filters, optional tRNS, optional Adam7 interlacing, stored / fixed / dynamic deflate blocks and
IDAT data split over several chunks — so that vct::decode_png (vct_b200/host/vct_ingest_image.hpp) is exercised well
beyond the four PNG flavours the reference's assets use.  Test infrastructure only."""
import os
import struct
import zlib

import numpy as np


def _chunk(tag, body):
    return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if pa <= pb and pa <= pc else (b if pb <= pc else c)


def _filter_rows(rows, bpp, rng):
    """rows: list of bytes objects (packed scanlines).  Applies a random filter 0..4 to every row."""
    out, prev = bytearray(), bytes(len(rows[0])) if rows else b""
    for row in rows:
        f = int(rng.integers(0, 5))
        enc = bytearray(len(row))
        for i, v in enumerate(row):
            a = row[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[f]
            enc[i] = (v - pred) & 255
        out.append(f); out += enc
        prev = row
    return bytes(out)


def _pack_rows(samples, depth):
    """samples: (h, w, c) uint16 array of sample values -> list of packed scanlines."""
    h, w, c = samples.shape
    rows = []
    for y in range(h):
        flat = samples[y].reshape(-1)
        if depth == 16:
            rows.append(flat.astype(">u2").tobytes())
        elif depth == 8:
            rows.append(flat.astype(np.uint8).tobytes())
        else:
            per = 8 // depth
            padded = np.zeros((len(flat) + per - 1) // per * per, np.uint8); padded[:len(flat)] = flat
            acc = np.zeros(len(padded) // per, np.uint8)
            for k in range(per):
                acc |= (padded[k::per] << (8 - depth * (k + 1))).astype(np.uint8)
            rows.append(acc.tobytes())
    return rows


def encode(samples, color, depth, rng, palette=None, trns=None, interlace=False, level=6, split=3):
    h, w, c = samples.shape
    bpp = max(1, c * depth // 8)
    if interlace:
        raw = b""
        for xo, yo, xs, ys in ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)):
            sub = samples[yo::ys, xo::xs]
            if sub.shape[0] and sub.shape[1]:
                raw += _filter_rows(_pack_rows(sub, depth), bpp, rng)
    else:
        raw = _filter_rows(_pack_rows(samples, depth), bpp, rng)
    comp = zlib.compressobj(level, zlib.DEFLATED, 15, 9, zlib.Z_FIXED if level == 1 else zlib.Z_DEFAULT_STRATEGY)
    data = comp.compress(raw) + comp.flush()
    out = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, color, 0, 0, 1 if interlace else 0))
    out += _chunk(b"gAMA", struct.pack(">I", 45455))                     # ancillary chunk: must be skipped
    if palette is not None:
        out += _chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if trns is not None:
        out += _chunk(b"tRNS", trns)
    n = max(1, len(data) // split)
    for i in range(0, len(data), n):
        out += _chunk(b"IDAT", data[i:i + n])
    return out + _chunk(b"IEND", b"")


def cases():
    """name -> PNG bytes; deterministic."""
    rng = np.random.default_rng(0x504E47)
    out = {}
    sizes = {0: (13, 7), 1: (32, 5), 2: (1, 1), 3: (9, 17)}
    k = 0
    for color, channels, depths in ((0, 1, (1, 2, 4, 8, 16)), (2, 3, (8, 16)), (3, 1, (1, 2, 4, 8)), (4, 2, (8, 16)), (6, 4, (8, 16))):
        for depth in depths:
            for interlace in (False, True):
                w, h = sizes[k % 4]; k += 1
                hi = 1 << depth
                palette = trns = None
                if color == 3:
                    n = min(hi, 1 + int(rng.integers(1, 256)))
                    palette = rng.integers(0, 256, (n, 3))
                    s = rng.integers(0, n, (h, w, 1))
                    if k % 2:
                        trns = bytes(rng.integers(0, 256, int(rng.integers(1, n + 1)), dtype=np.uint8))
                else:
                    s = rng.integers(0, hi, (h, w, channels))
                    if depth == 16:                                       # make colour keys hit: few distinct values
                        s = rng.integers(0, 3, (h, w, channels)) * 21845 + (rng.integers(0, 2, (h, w, channels)) if k % 3 == 0 else 0)
                    elif depth == 8 and color in (0, 2):
                        s = rng.integers(0, 4, (h, w, channels)) * 85
                    if color in (0, 2) and k % 2:
                        key = s[h // 2, w // 2]
                        trns = b"".join(struct.pack(">H", int(v)) for v in key)
                level = (0, 1, 6, 9)[k % 4]
                name = f"c{color}_d{depth}_{'adam7' if interlace else 'plain'}_{w}x{h}{'_trns' if trns is not None else ''}_z{level}.png"
                out[name] = encode(s.astype(np.uint16), color, depth, rng, palette, trns, interlace, level, split=1 + k % 4)
    big = rng.integers(0, 256, (67, 131, 3)); big[20:50, 10:100] = (200, 30, 90)        # long matches, several deflate blocks
    out["c2_d8_plain_131x67_big_z9.png"] = encode(big.astype(np.uint16), 2, 8, rng, level=9, split=5)
    noise = rng.integers(0, 256, (300, 300, 4))
    out["c6_d8_plain_300x300_noise_z6.png"] = encode(noise.astype(np.uint16), 6, 8, rng, level=6, split=7)
    return out


def write_all(directory):
    names = []
    for name, data in cases().items():
        open(os.path.join(directory, name), "wb").write(data)
        names.append(name)
    return names

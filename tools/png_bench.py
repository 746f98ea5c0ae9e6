#!/usr/bin/env python3
"""PNG encode of a frame-sized RGB8 image, single stripe (what one libpng/zlib call
does, the reference's src/writers.cpp:4-21) against the host writer's striped encoder on all host threads.
CPU only.  Usage: python tools/png_bench.py [width height]"""
import importlib.util
import os
import sys
import time
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("cs184_raytracer_b200", ROOT / "cs184-raytracer_b200/__init__.py",
                                              submodule_search_locations=[str(ROOT / "cs184-raytracer_b200")])
pkg = importlib.util.module_from_spec(spec)
sys.modules["cs184_raytracer_b200"] = pkg
spec.loader.exec_module(pkg)

w = int(sys.argv[1]) if len(sys.argv) > 1 else 7680
h = int(sys.argv[2]) if len(sys.argv) > 2 else 4320
# a frame-like image: smooth Phong-ish shading with specular blobs and a little dither
yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
base = 0.5 + 0.4 * np.sin(xx * 0.003) * np.cos(yy * 0.002)
spec_ = np.exp(-(((xx % 700) - 350) ** 2 + ((yy % 500) - 250) ** 2) / 4000.0)
rgb = np.stack([base + spec_, base * 0.8 + spec_, base * 0.6 + spec_], axis=-1)
rgb += np.random.default_rng(1).normal(0, 0.004, rgb.shape).astype(np.float32)
img = (np.clip(rgb, 0, 1) * 255).astype(np.uint8)

def run(threads):
    t0 = time.perf_counter()
    data = pkg.encode_png(img, threads)
    return time.perf_counter() - t0, data

n = os.cpu_count() or 1
t1, d1 = run(1)
tn, dn = run(n)
print(f"{w}x{h} RGB8 ({img.nbytes / 1e6:.1f} MB raw)")
print(f"1 stripe  : {t1 * 1e3:8.1f} ms  {len(d1) / 1e6:7.2f} MB")
print(f"{n:2d} stripes: {tn * 1e3:8.1f} ms  {len(dn) / 1e6:7.2f} MB   speed-up {t1 / tn:.1f}x, size +{100 * (len(dn) / len(d1) - 1):.2f} %")
# both decode to the same pixels
def pixels(data):
    pos, idat = 8, b""
    while pos < len(data):
        ln = int.from_bytes(data[pos:pos + 4], "big")
        if data[pos + 4:pos + 8] == b"IDAT":
            idat += data[pos + 8:pos + 8 + ln]
        pos += 12 + ln
    return np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)[:, 1:]
assert np.array_equal(pixels(d1), pixels(dn)) and np.array_equal(pixels(dn).reshape(h, w, 3), img)
print("decoded pixels identical")

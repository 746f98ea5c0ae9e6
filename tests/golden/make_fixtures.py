#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference tree (run in the build container only).

  inputs/, excess_inputs/   scene DATA files copied verbatim (the .rti/.obj the reference ships;
                            no reference source code is copied)
  outputs/                  the reference's nine golden images outputs/image-0N.png
  ref/<scene>_<W>x<H>.npz   outputs of the UNMODIFIED reference hot path (oracle/_ref/libref*.so,
                            built by oracle/Makefile from /root/reference/src): raw FP64
                            framebuffer, primary-hit geometry ids, castRay call count

The GPU box has no /root/reference; the tests read only what this script wrote.
Usage: python tests/golden/make_fixtures.py [--ref /root/reference]
"""
import argparse
import ctypes as C
import hashlib
import json
import shutil
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent

SCENES = {
    # name: (relative .rti path, depth)
    **{f"input-{i:02d}": (f"inputs/input-{i:02d}.rti", 10) for i in range(1, 10)},
    "refraction3": ("excess_inputs/refraction3.rti", 10),
    "refraction": ("excess_inputs/refraction.rti", 10),
    "test": ("excess_inputs/test.rti", 10),
    "example5": ("excess_inputs/example5.rti", 10),
    "reflective_specular_test": ("excess_inputs/reflective_specular_test.rti", 10),
    "bunny4": ("excess_inputs/bunny4.rti", 10),
}
SIZES = [(96, 96), (80, 45)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    ref = Path(args.ref)
    for sub, pats in (("inputs", ("*.rti", "*.obj")), ("excess_inputs", ("*.rti", "test.obj", "bunny.obj")),
                      ("outputs", ("image-0*.png",))):
        (HERE / sub).mkdir(exist_ok=True)
        for pat in pats:
            for f in sorted((ref / sub).glob(pat)):
                if f.name in ("minicooper.rti",):
                    continue   # its mesh is not shipped (.MISSING_LARGE_BLOBS)
                shutil.copyfile(f, HERE / sub / f.name)
    lib = C.CDLL(str(ROOT / "oracle/_ref/libref_count.so"))
    lib.ref_scene_load.restype = C.c_void_p
    lib.ref_scene_load.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int]
    (HERE / "ref").mkdir(exist_ok=True)
    manifest = {}
    for name, (rel, depth) in SCENES.items():
        path = HERE / rel
        arr = (C.c_char_p * 1)(str(path).encode())
        err = C.create_string_buffer(256)
        h = lib.ref_scene_load(arr, 1, err, 256)
        assert h, (name, err.value)
        h = C.c_void_p(h)
        for (w, hh) in SIZES:
            rgb = np.zeros((hh, w, 3))
            ids = np.zeros((hh, w), dtype=np.int32)
            sec = C.c_double()
            calls = C.c_uint64()
            lib.ref_render(h, w, hh, depth, 0, 8, rgb.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p),
                           C.byref(sec), C.byref(calls))
            io = np.zeros((hh, w, 3))
            lib.ref_render(h, w, hh, depth, 1, 8, io.ctypes.data_as(C.c_void_p), None, None, None)
            out = HERE / "ref" / f"{name}_{w}x{hh}.npz"
            np.savez_compressed(out, rgb=rgb, geom=ids, castray_calls=np.uint64(calls.value), depth=depth,
                                intersection_only=io[..., 0].copy())
            manifest[out.name] = {"scene": rel, "depth": depth, "castray_calls": int(calls.value),
                                  "rgb_sha256": hashlib.sha256(rgb.tobytes()).hexdigest()[:16]}
            print(out.name, manifest[out.name])
    (HERE / "ref" / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()

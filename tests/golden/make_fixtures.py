#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference tree (run in the build container only).

  inputs/, excess_inputs/   scene DATA files copied verbatim (the .rti/.obj the reference ships;
                            no reference source code is copied)
  outputs/                  the reference's nine golden images outputs/image-0N.png
  ref/<scene>_<W>x<H>.npz   outputs of the UNMODIFIED reference hot path (oracle/_ref/libref*.so,
                            built by oracle/Makefile from /root/reference/src): raw FP64
                            framebuffer, primary-hit geometry AND face ids (ref_primary_faces:
                            the reference's own Mesh code run on one face at a time), castRay
                            call count
  ref/synthetic_<W>x<H>.npz the same for the benchmark scene (generated in memory by the host
                            library, handed to the reference through ref_scene_from_flat)
  ref/big_<scene>_<W>x<H>.npz  the BASELINE.json configurations at their full sizes (the sizes the
                            stock reference binary aborts on): ids, ray count, the 8-bit frame
                            (src/writers.cpp:7 applied to the reference's doubles) and every
                            8th pixel of the FP64 frame in both directions

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
# BASELINE.json configs[1..3] at full size (+ bunny4 at a quarter: 10 minutes of CPU at 4K)
BIG = [("input-02", 1920, 1080), ("refraction3", 3840, 2160), ("bunny4", 960, 540)]
SYNTHETIC = (708, 1000, 184)          # grid cells, spheres, seed: bench.py's "synthetic" workload
SYNTHETIC_SIZES = [(96, 54, 5)]       # width, height, depth
BIG_STRIDE = 8


def quantize(rgb):
    """(uint8)(clamp(v,0,1)*255.0), truncating: src/writers.cpp:7 in numpy."""
    v = np.where(1.0 < rgb, 1.0, rgb)
    v = np.where(v < 0.0, 0.0, v)
    v = np.where(np.isnan(v), 0.0, v)
    return (v * 255.0).astype(np.uint8)


def render_ref(lib, h, w, hh, depth, threads=8):
    rgb = np.zeros((hh, w, 3))
    ids = np.zeros((hh, w), dtype=np.int32)
    faces = np.zeros((hh, w), dtype=np.int32)
    sec = C.c_double()
    calls = C.c_uint64()
    lib.ref_render(h, w, hh, depth, 0, threads, rgb.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p),
                   C.byref(sec), C.byref(calls))
    bad = lib.ref_primary_faces(h, w, hh, ids.ctypes.data_as(C.c_void_p), faces.ctypes.data_as(C.c_void_p), threads)
    assert bad == 0, f"{bad} primary hits whose face could not be identified"
    return rgb, ids, faces, int(calls.value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--no-big", action="store_true", help="skip the full-size BASELINE frames (about a minute of CPU)")
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
            rgb, ids, faces, calls = render_ref(lib, h, w, hh, depth)
            io = np.zeros((hh, w, 3))
            lib.ref_render(h, w, hh, depth, 1, 8, io.ctypes.data_as(C.c_void_p), None, None, None)
            out = HERE / "ref" / f"{name}_{w}x{hh}.npz"
            np.savez_compressed(out, rgb=rgb, geom=ids, face=faces, castray_calls=np.uint64(calls), depth=depth,
                                intersection_only=io[..., 0].copy())
            manifest[out.name] = {"scene": rel, "depth": depth, "castray_calls": calls,
                                  "rgb_sha256": hashlib.sha256(rgb.tobytes()).hexdigest()[:16]}
            print(out.name, manifest[out.name])
        for (bname, w, hh) in BIG:
            if bname != name or args.no_big:
                continue
            rgb, ids, faces, calls = render_ref(lib, h, w, hh, depth)
            out = HERE / "ref" / f"big_{name}_{w}x{hh}.npz"
            np.savez_compressed(out, rgb8=quantize(rgb), rgb_sub=rgb[::BIG_STRIDE, ::BIG_STRIDE].copy(), stride=BIG_STRIDE,
                                geom=ids.astype(np.int16), face=faces, castray_calls=np.uint64(calls), depth=depth)
            manifest[out.name] = {"scene": rel, "depth": depth, "castray_calls": calls,
                                  "rgb_sha256": hashlib.sha256(rgb.tobytes()).hexdigest()[:16]}
            print(out.name, manifest[out.name])
    # ---- the benchmark scene: host library -> flat descriptor -> the reference's own object graph
    import importlib.util
    import sys
    pkg_dir = ROOT / "cs184-raytracer_b200"
    spec = importlib.util.spec_from_file_location("cs184_raytracer_b200", pkg_dir / "__init__.py",
                                                  submodule_search_locations=[str(pkg_dir)])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["cs184_raytracer_b200"] = pkg
    spec.loader.exec_module(pkg)
    lib.ref_scene_from_flat.restype = C.c_void_p
    lib.ref_scene_from_flat.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    host_scene = pkg.HostScene.synthetic(*SYNTHETIC)
    err = C.create_string_buffer(256)
    h = lib.ref_scene_from_flat(C.cast(host_scene.flat, C.c_void_p), err, 256)
    assert h, err.value
    h = C.c_void_p(h)
    for (w, hh, depth) in SYNTHETIC_SIZES:
        rgb, ids, faces, calls = render_ref(lib, h, w, hh, depth)
        out = HERE / "ref" / f"synthetic_{w}x{hh}.npz"
        np.savez_compressed(out, rgb=rgb, geom=ids, face=faces, castray_calls=np.uint64(calls), depth=depth,
                            synthetic=np.array(SYNTHETIC, dtype=np.int64))
        manifest[out.name] = {"scene": f"synthetic{SYNTHETIC}", "depth": depth, "castray_calls": calls,
                              "rgb_sha256": hashlib.sha256(rgb.tobytes()).hexdigest()[:16]}
        print(out.name, manifest[out.name])
    (HERE / "ref" / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()

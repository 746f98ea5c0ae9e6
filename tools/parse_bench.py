#!/usr/bin/env python3
"""ingest of a million-triangle .obj.  Writes the synthetic benchmark scene as
.rti/.obj text and times (a) our single-pass parser and (b) the reference's parser
(oracle/_ref/libref.so, per-line istringstream + per-token ostringstream) on the same files,
then checks both produce the same flat scene.  CPU only.  Usage: python tools/parse_bench.py [cells]"""
import ctypes as C
import importlib.util
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("cs184_raytracer_b200", ROOT / "cs184-raytracer_b200/__init__.py",
                                              submodule_search_locations=[str(ROOT / "cs184-raytracer_b200")])
pkg = importlib.util.module_from_spec(spec)
sys.modules["cs184_raytracer_b200"] = pkg
spec.loader.exec_module(pkg)

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 708
with tempfile.TemporaryDirectory() as tmp:
    rti, obj = os.path.join(tmp, "syn.rti"), os.path.join(tmp, "syn.obj")
    err = C.create_string_buffer(256)
    t0 = time.perf_counter()
    assert pkg.load_host().as2_write_synthetic(rti.encode(), obj.encode(), cells, 1000, 184, err, 256) == 0
    t_write = time.perf_counter() - t0
    size = os.path.getsize(obj)
    t0 = time.perf_counter()
    ours = pkg.HostScene.load(rti)
    t_ours = time.perf_counter() - t0
    a = pkg.flat_arrays(ours.flat)
    print(f"{cells}x{cells} cells, {a['num_faces']} faces, .obj {size / 1e6:.1f} MB (written in {t_write:.2f} s)")
    print(f"this host : parse + build object model + flatten  {t_ours:.2f} s  ({size / 1e6 / t_ours:.0f} MB/s)")
    ref = C.CDLL(str(ROOT / "oracle/_ref/libref.so"))
    ref.ref_scene_load.restype = C.c_void_p
    ref.ref_scene_flatten.restype = C.c_void_p
    ref.ref_scene_flatten.argtypes = [C.c_void_p]
    arr = (C.c_char_p * 1)(rti.encode())
    t0 = time.perf_counter()
    h = ref.ref_scene_load(arr, 1, err, 256)
    t_ref = time.perf_counter() - t0
    assert h, err.value
    b = pkg.flat_arrays(C.c_void_p(ref.ref_scene_flatten(C.c_void_p(h))))
    same = all((a[k] == b[k]).all() if hasattr(a[k], "all") else a[k] == b[k] for k in a)
    print(f"reference : RTIParser/OBJParser                          {t_ref:.2f} s  ({size / 1e6 / t_ref:.0f} MB/s)")
    print(f"speed-up {t_ref / t_ours:.1f}x, flat scenes byte-identical: {same}")

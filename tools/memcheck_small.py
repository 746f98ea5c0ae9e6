#!/usr/bin/env python3
"""compute-sanitizer target: a small render that exercises every kernel (LBVH build, trace, hit sort,
shade, shadow on its own stream, resolve) on the teapot and bunny scenes.
  RT_HIT_SORT_MIN_RAYS=1 compute-sanitizer --tool memcheck python tools/memcheck_small.py"""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
spec = importlib.util.spec_from_file_location("cs184_raytracer_b200", ROOT / "cs184-raytracer_b200/__init__.py",
                                              submodule_search_locations=[str(ROOT / "cs184-raytracer_b200")])
pkg = importlib.util.module_from_spec(spec)
sys.modules["cs184_raytracer_b200"] = pkg
spec.loader.exec_module(pkg)
r = pkg.Renderer(0)
for rel, depth in (("inputs/input-02.rti", 6), ("excess_inputs/bunny4.rti", 4)):
    r.upload(pkg.HostScene.load(ROOT / "tests/golden" / rel))
    rgb = r.render(160, 120, depth)
    st = r.stats()
    print(rel, rgb.shape, float(rgb.sum()), st["rays_primary"], st["rays_shadow"], st["rays_secondary"], st["kernel_launches"])
    r.primary_ids(160, 120)
r.close()

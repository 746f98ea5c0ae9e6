"""cs184-raytracer_b200 — B200-native trace loop behind CS184-Raytracer's surface.

Layout:
  csrc/   hand-written sm_100a CUDA (wavefront kernels, LBVH build) + the C ABI
  host/   C++ host mirroring the reference's parsers / object model / CLI / PNG writer
  binding.py  ctypes glue used by tests/, bench.py and __graft_entry__.py
"""
from .binding import (HostScene, Renderer, RtError, flat_arrays, load_host, load_rt, make_params, quantize_rgb8,
                      tile_counts, write_png, encode_png, rt_params, rt_scene, rt_stats, RT_FLAG_BRUTE_FORCE, RT_FLAG_COUNT_WORK, RT_FLAG_TIME_KERNELS, RT_FLAG_SERIAL, RT_FLAG_FULL_FRAME,
                      RT_SCENE_FACES_ON_DEVICE, RT_ERR_LIMIT,
                      RT_SYMBOLS, RT_TILE_PIXELS, PKG_DIR, LIB_DIR)

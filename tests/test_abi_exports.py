"""The C-ABI library loads without a GPU and exports every symbol include/rt_b200.h declares;
without a device the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import re

import pytest

from conftest import ROOT


def _declared_functions():
    text = (ROOT / "include" / "rt_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)) - {"rt_progress_fn"})


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_rt()
    declared = _declared_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rt_b200.h but not exported"
    assert sorted(pkg.RT_SYMBOLS) == declared
    assert lib.rt_abi_version() == 2


def test_host_library_exports(pkg):
    lib = pkg.load_host()
    for name in ("as2_scene_load", "as2_scene_synthetic", "as2_write_synthetic", "as2_scene_free", "as2_scene_flatten",
                 "as2_scene_render", "as2_write_png_rgb8", "as2_write_png_f64", "as2_quantize_rgb8",
                 "as2_encode_png_rgb8"):
        assert hasattr(lib, name)


def test_tile_partition_is_a_partition(pkg):
    """Pure host arithmetic of the interleaved-tile sharding (no compute call)."""
    for (w, h) in [(1, 1), (33, 31), (1920, 1080), (7680, 4320), (500, 500)]:
        for world in (1, 2, 3, 4, 8):
            counts = []
            for rank in range(world):
                own, mx, total = pkg.tile_counts(pkg.make_params(w, h, 5, tile_rank=rank, tile_world=world))
                counts.append(own)
                assert own <= mx
            assert sum(counts) == total == ((w + 31) // 32) * ((h + 31) // 32)
            assert max(counts) - min(counts) <= (h + 31) // 32    # near-even split


def test_no_gpu_means_loud_failure(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.RtError) as e:
        pkg.Renderer(0)
    assert "no usable CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    import subprocess
    pkg_dir = ROOT / "cs184-raytracer_b200"
    for path in list(pkg_dir.rglob("*.py")) + list(pkg_dir.rglob("*.cu")) + list(pkg_dir.rglob("*.cuh")) + \
            list(pkg_dir.rglob("*.cpp")) + list(pkg_dir.rglob("*.h")) + [pkg_dir / "Makefile"]:
        text = path.read_text()
        assert "liboracle" not in text and "whitted_oracle" not in text and "libref" not in text, path
    for so in (pkg_dir / "lib").glob("*.so"):
        out = subprocess.run(["ldd", str(so)], capture_output=True, text=True).stdout
        assert "oracle" not in out and "libref" not in out


def test_sub_rank_tiles_refine_a_ranks_share(pkg):
    """rt_params.n_gpus inside a rank of a multi-process job: device k of n takes the tiles
    (tx + ty) % (world * n) == rank + k * world, which must partition exactly the rank's own share."""
    for (w, h) in [(33, 31), (1920, 1080), (7680, 4320), (500, 777)]:
        for world in (1, 2, 3, 8):
            for n in (2, 3, 4):
                for rank in range(world):
                    own = pkg.tile_counts(pkg.make_params(w, h, 5, tile_rank=rank, tile_world=world))[0]
                    parts = [pkg.tile_counts(pkg.make_params(w, h, 5, tile_rank=rank + k * world, tile_world=world * n))[0]
                             for k in range(n)]
                    assert sum(parts) == own


def test_abi_struct_layout_matches_the_header(pkg):
    """ctypes mirrors of rt_params / rt_scene against the C compiler's view of include/rt_b200.h."""
    import subprocess
    import tempfile
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "rt_b200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(rt_params), offsetof(rt_params, n_gpus), sizeof(rt_scene),
           offsetof(rt_scene, flags), offsetof(rt_scene, face_normals), sizeof(rt_stats), sizeof(rt_geometry));
    return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        (ROOT / "include").exists()
        c = f"{d}/layout.c"
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), c, "-o", f"{d}/layout"], check=True)
        out = subprocess.run([f"{d}/layout"], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [C.sizeof(pkg.rt_params), pkg.rt_params.n_gpus.offset, C.sizeof(pkg.rt_scene), pkg.rt_scene.flags.offset,
            pkg.rt_scene.face_normals.offset, C.sizeof(pkg.rt_stats), C.sizeof(pkg.binding.rt_geometry)]
    assert got == want

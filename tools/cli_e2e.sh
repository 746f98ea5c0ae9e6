#!/bin/bash
# End-to-end run of the drop-in CLI on the synthetic benchmark scene written as .rti/.obj text
# (SURVEY section 8 f-1/f-2): ingest, render and PNG encode times.  Usage: tools/cli_e2e.sh [w h]
set -e
W=${1:-7680}; H=${2:-4320}
D=$(mktemp -d)
python - "$D" <<'PY'
import ctypes as C, importlib.util, sys
from pathlib import Path
ROOT = Path.cwd()
spec = importlib.util.spec_from_file_location("cs184_raytracer_b200", ROOT / "cs184-raytracer_b200/__init__.py",
                                              submodule_search_locations=[str(ROOT / "cs184-raytracer_b200")])
pkg = importlib.util.module_from_spec(spec); sys.modules["cs184_raytracer_b200"] = pkg; spec.loader.exec_module(pkg)
err = C.create_string_buffer(256)
d = sys.argv[1]
assert pkg.load_host().as2_write_synthetic(f"{d}/syn.rti".encode(), f"{d}/syn.obj".encode(), 708, 1000, 184, err, 256) == 0
PY
ls -la $D
for i in 1 2; do
  AS2_TIMING=1 cs184-raytracer_b200/bin/as2 -w $W -h $H --bdepth 5 -o $D/out.png $D/syn.rti 2>&1 | grep -v "^Progress\|^$" | tail -2
done
ls -la $D/out.png
python -c "
from PIL import Image; import sys
Image.MAX_IMAGE_PIXELS=None
im=Image.open('$D/out.png'); print('decoded', im.size, im.mode)"
rm -rf $D

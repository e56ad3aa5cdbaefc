"""A short run of tools/fuzz_parity.py on every GPU test pass: random soups / grids / degenerate meshes at random scales
(1e-6 ... 1e6) and offsets, random BVHBuildOptions, device-built trees against the oracle's, closest hits and traversal
counters bit-identical, occlusion exact with tmax at t * {0.9, 1, 1.1} and t -+ 2 ulp.  (Run long:
python tools/fuzz_parity.py 300 <seed>; 6 282 scenes / 32 M rays passed when this test was added.)"""
import importlib.util
import os

import pytest

from tests import common as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_parity_short(seed):
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(os.path.dirname(T.HERE), "tools", "fuzz_parity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    scenes, rays = mod.run(8.0, seed, verbose=False)
    assert scenes >= 20 and rays >= 20 * 4096


def test_fuzz_frames_short():
    """tools/fuzz_frames.py for a few seconds: random image sizes, sample counts, shaders, rectangles and row bands through
    mb200_render_frame against the oracle's accumulated passes, with batches small enough (MB200_FRAME_BATCH_ITEMS, read
    once per process -> subprocess) that every frame is cut by tile rows into many batches and copied back by chunks.
    (43 139 configurations passed in two long runs when this test was added.)"""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(T.HERE), "tools", "fuzz_frames.py")
    for items, seed in (("30000", "21"), ("4000", "22")):
        r = subprocess.run([sys.executable, tool, "6", seed], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, MB200_FRAME_BATCH_ITEMS=items))
        assert r.returncode == 0 and "FUZZ FRAMES OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
        assert int(r.stdout.split("FUZZ FRAMES OK: ")[1].split()[0]) >= 50

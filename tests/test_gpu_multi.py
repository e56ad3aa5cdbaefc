"""Single-process multi-GPU frames (mb200_render_frame_multi, RenderConfig::num_gpus): need >= 2 GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box.  The assembled frame must be bit-identical to the 1-GPU frame."""
import os
import subprocess

import numpy as np
import pytest

import mallie_b200 as M
from tests import common as T
from tests.test_gpu_host_api import PKG, write_quad_sphere_obj

pytestmark = pytest.mark.gpu


def need_gpus(n):
    if M.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("shader", [M.SHADER_PRIMARY_SHADOW, M.SHADER_PATHTRACE])
def test_frame_over_two_gpus_is_the_single_gpu_frame(shader):
    need_gpus(2)
    G = min(M.device_count(), 4)
    W, H = 333, 150                       # ragged: last band is short, rows not a multiple of bands * GPUs
    m = T.load_mesh("sphere40")
    scenes = [M.Scene(m["vertices"], m["faces"], device=g) for g in range(G)]
    fg = M.camera_frame((0.2, 0.1, 3.0), (0, 0, 0), width=W, height=H)
    p = scenes[0].render_params(fg, W, H, shader=shader, light=(2, 4, 3), pass_index=2, max_path_length=4)
    want, wcnt, wst = scenes[0].render_frame(p, 3)
    for n in range(2, G + 1):
        img, cnt, st = M.render_frame_multi(scenes[:n], p, 3, band_rows=8)
        assert img.tobytes() == want.tobytes() and np.array_equal(cnt, wcnt), n
        assert st == wst, (st, wst)
    # device-resident framebuffer on the first GPU
    import torch
    d_img = torch.full((H, W, 3), -1.0, dtype=torch.float32, device="cuda:0")
    d_cnt = torch.zeros((H, W), dtype=torch.int32, device="cuda:0")
    M.render_frame_multi(scenes[:2], p, 3, band_rows=12, image=d_img.data_ptr(), count=d_cnt.data_ptr(), stats=False)
    torch.cuda.synchronize()
    assert d_img.cpu().numpy().tobytes() == want.tobytes() and (d_cnt.cpu().numpy() == 3).all()
    bad = scenes[0].render_params(fg, W, H, tile=(0, 0, W, H - 1))
    with pytest.raises(M.MallieB200Error):
        M.render_frame_multi(scenes[:2], bad, 1)
    with pytest.raises(M.MallieB200Error):
        M.render_frame_multi([scenes[0], scenes[0]], p, 1)
    for s in scenes:
        s.close()


def test_cpp_render_with_num_gpus(tmp_path):
    need_gpus(2)
    exe = os.path.join(PKG, "host_api_check")
    obj = str(tmp_path / "scene.obj")
    write_quad_sphere_obj(obj)
    outs = []
    for gpus in (1, 2):
        out = str(tmp_path / f"out{gpus}.bin")
        r = subprocess.run([exe, obj, out, "96", "64", "1", str(gpus)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]


def test_cloned_replicas_render_the_same_frame():
    """Replicas made by mb200_scene_clone (device-to-device copy of a device-built scene) instead of N host uploads."""
    need_gpus(2)
    G = min(M.device_count(), 4)
    W, H = 200, 120
    m = T.load_mesh("teapot")
    first = M.Scene.build(m["vertices"], m["faces"], m["material_ids"], m["normals"], m["uvs"])
    scenes = [first] + [first.clone(g) for g in range(1, G)]
    for s in scenes[1:]:
        assert s.layout()[1].tobytes() == first.layout()[1].tobytes() and s.layout()[2].tobytes() == first.layout()[2].tobytes()
    fg = M.camera_frame((5, 40, 150), (5, 40, 0), width=W, height=H)
    p = first.render_params(fg, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(100.0, 200.0, 150.0), pass_index=1)
    want, wcnt, _ = first.render_frame(p, 2)
    img, cnt, _ = M.render_frame_multi(scenes, p, 2, band_rows=4)
    assert img.tobytes() == want.tobytes() and np.array_equal(cnt, wcnt)
    for s in scenes:
        s.close()

"""Output resolve: HDRToLDR (main_console.cc:25-43) and the SDL viewer's gamma-2.2 Display
(main_sdl.cc:156-165,420-477).

CPU: the oracle's restatement of HDRToLDR against the reference's own function (compiled into oracle/_ref through
oracle/ref_console.cc), edge values included.  GPU: mb200_resolve_ldr / mb200_render_frame_ldr against both."""
import numpy as np
import pytest

from oracle import orabind as O
from oracle import refbind as R


def edge_frame(seed=5, h=37, w=53):
    """Random radiance sums and counts plus the values where the conversions are delicate: exact quantisation
    boundaries, negatives, huge values, inf / NaN (x86's cvttsd2si gives INT_MIN -> 0), count 0 (x / 0)."""
    rng = np.random.default_rng(seed)
    cnt = rng.integers(1, 70, size=(h, w)).astype(np.int32)
    img = (rng.random((h, w, 3)) * 1.3 * cnt[..., None]).astype(np.float32)
    flat = img.reshape(-1)
    k = np.arange(256, dtype=np.float64)
    bounds = np.concatenate([(k / 255.5), np.nextafter((k / 255.5).astype(np.float32), 0), np.nextafter((k / 255.5).astype(np.float32), 9)])
    flat[:bounds.size] = bounds.astype(np.float32) * np.repeat(cnt.reshape(-1), 3)[:bounds.size]
    special = np.array([-1.0, -0.0, 0.0, 1e-30, 1.0, 1.0000001, 255.0, 1e7, 8.5e6, 3e9, 1e30, np.inf, -np.inf, np.nan], np.float32)
    flat[-special.size:] = special
    cnt.reshape(-1)[5] = 0                       # 0 / 0 = NaN, x / 0 = inf
    cnt.reshape(-1)[6] = 0
    flat[3 * 6:3 * 6 + 3] = (0.0, 2.0, -2.0)
    return img, cnt


@pytest.mark.skipif(not R.available(), reason="reference not compiled here")
def test_oracle_hdr_to_ldr_is_the_reference_function():
    img, cnt = edge_frame()
    with np.errstate(all="ignore"):
        assert O.hdr_to_ldr(img, cnt).tobytes() == R.hdr_to_ldr(img, cnt).tobytes()
    out = O.hdr_to_ldr(img, cnt)
    assert out.min() == 0 and out.max() == 255


def test_display_restatement_layout_and_gamma():
    img = np.zeros((2, 2, 3), np.float32)
    cnt = np.full((2, 2), 4, np.int32)
    img[0, 0] = (4.0, 2.0, 0.0)                  # R = 1.0, G = 0.5, B = 0
    out = O.display_bgra(img, cnt)
    assert out.shape == (2, 2, 4) and (out[..., 3] == 255).all()
    assert out[0, 0, 2] == 255 and out[0, 0, 0] == 0                    # BGRA: R at byte 2, B at byte 0
    assert out[0, 0, 1] == int(np.float32(0.5) ** np.float32(1 / 2.2) * 255.5)


@pytest.mark.gpu
def test_device_ldr_resolve_matches_reference_and_oracle():
    import mallie_b200 as M
    from tests import common as T
    m = T.load_mesh("sphere40")
    sc = M.Scene(m["vertices"], m["faces"])
    img, cnt = edge_frame(h=61, w=47)
    with np.errstate(all="ignore"):
        want0 = R.hdr_to_ldr(img, cnt) if R.available() else O.hdr_to_ldr(img, cnt)
        want1 = O.display_bgra(img, cnt)
    got0 = sc.resolve_ldr(img, cnt, 47, 61, M.capi.LDR_RGB8_LINEAR)
    assert got0.tobytes() == want0.tobytes(), "HDRToLDR on the device differs from the reference's"
    got1 = sc.resolve_ldr(img, cnt, 47, 61, M.capi.LDR_BGRA8_GAMMA22)
    diff = np.abs(got1.astype(int) - want1.astype(int))
    # CUDA's pow vs glibc's powf: a last-ulp difference can move a value across a quantisation boundary
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-4, (diff.max(), (diff != 0).mean())
    # a rendered frame: device-resident float frame -> 8-bit, and the one-call form
    W, H = 320, 200
    fg = M.camera_frame((0.3, 0.2, 3.0), (0, 0, 0), width=W, height=H)
    p = sc.render_params(fg, W, H, shader=M.SHADER_PRIMARY_SHADOW, light=(2.0, 4.0, 3.0), pass_index=1)
    frame, fcnt, st = sc.render_frame(p, 5)
    want = R.hdr_to_ldr(frame, fcnt) if R.available() else O.hdr_to_ldr(frame, fcnt)
    ldr, st2 = sc.render_frame_ldr(p, 5, M.capi.LDR_RGB8_LINEAR)
    assert ldr.tobytes() == want.tobytes() and st2["primary_rays"] == st["primary_rays"] == 5 * W * H
    assert ldr.max() > 100 and ldr.min() == 0
    bgra, _ = sc.render_frame_ldr(p, 5, M.capi.LDR_BGRA8_GAMMA22)
    wantd = O.display_bgra(frame, fcnt)
    d = np.abs(bgra.astype(int) - wantd.astype(int))
    assert d.max() <= 1 and (d != 0).mean() <= 1e-4
    # device buffers in, device buffer out (enqueue-only)
    torch = pytest.importorskip("torch")
    d_img = torch.from_numpy(frame).cuda()
    d_cnt = torch.from_numpy(fcnt).cuda()
    d_out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    sc.resolve_ldr(d_img.data_ptr(), d_cnt.data_ptr(), W, H, 0, out=d_out.data_ptr())
    sc.synchronize()
    assert d_out.cpu().numpy().tobytes() == want.tobytes()
    with pytest.raises(M.MallieB200Error):
        sc.resolve_ldr(img, cnt, 47, 61, 7)
    sc.close()

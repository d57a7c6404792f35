"""Image export (RendererCore::saveImage, src/RendererCore.cpp:608-646): the product's own encoders behind
yune_write_image, decoded again by independent readers (PIL for .png / .jpg / .ppm, a few lines of numpy for .hdr / .pfm).
The reference's conventions under test: first file row = top of the picture (stbi_flip_vertically_on_write), 8-bit values =
clamp(v, 0, 1) * 255 rounded (the GL_UNSIGNED_BYTE read-back), .hdr keeps the float image, alpha is dropped."""
import os

import numpy as np
import pytest

import yune_b200 as yb

Image = pytest.importorskip("PIL.Image")


def _picture(H, W, seed=0):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W, 4), np.float32)
    img[..., 0] = x / max(W - 1, 1); img[..., 1] = y / max(H - 1, 1)
    img[..., 2] = 0.5 + 0.5 * np.sin(x * 0.7) * np.cos(y * 0.4)
    img[..., :3] += rng.normal(0, 0.02, (H, W, 3)).astype(np.float32)
    img[..., 3] = 64                                      # sample count in alpha: must not reach the file
    if H > 8 and W > 8:
        img[2:5, 2:5, :3] = [1.7, -0.2, np.nan]           # clamped / NaN -> 0
    return img


def _as_u8_top_down(img):
    v = np.clip(np.nan_to_num(img[::-1, :, :3], nan=0.0), 0, 1)
    return (v * np.float32(255) + np.float32(0.5)).astype(np.uint8)


@pytest.mark.parametrize("H,W", [(37, 53), (1, 1), (8, 8), (200, 301)])       # 200 x 301: several 64 KB deflate blocks
def test_png_and_ppm_are_exact(tmp_path, H, W):
    img = _picture(H, W)
    for ext in (".png", ".ppm", ".PNG"):
        p = str(tmp_path / ("a" + ext))
        assert yb.write_image(p, img)
        got = np.asarray(Image.open(p).convert("RGB"))
        assert got.shape == (H, W, 3) and (got == _as_u8_top_down(img)).all(), ext


@pytest.mark.parametrize("H,W", [(37, 53), (1, 1), (8, 8), (64, 80)])
def test_jpg_decodes_to_the_picture(tmp_path, H, W):
    """Baseline JPEG with quantisation tables of ones (quality 100 in stb_image_write): only the YCbCr / DCT rounding is lost."""
    img = _picture(H, W, seed=1)
    if (H, W) == (64, 80):
        img[..., :3] = np.random.default_rng(2).random((H, W, 3))          # noise: long codes, ZRL runs, 0xFF byte stuffing
    p = str(tmp_path / "a.jpg")
    assert yb.write_image(p, img)
    raw = open(p, "rb").read()
    assert raw[:4] == b"\xff\xd8\xff\xe0" and raw[6:11] == b"JFIF\0" and raw[-2:] == b"\xff\xd9"
    im = Image.open(p)
    assert im.format == "JPEG" and im.size == (W, H)
    d = np.abs(np.asarray(im.convert("RGB")).astype(int) - _as_u8_top_down(img).astype(int))
    assert d.max() <= 4 and d.mean() < 1.0


def test_hdr_and_pfm_keep_the_float_image(tmp_path):
    H, W = 23, 31
    img = _picture(H, W, seed=3)
    img[..., :3] = np.abs(np.nan_to_num(img[..., :3])) * np.float32(40.0)       # radiance well above 1
    img[0, 0, :3] = 0
    p = str(tmp_path / "a.pfm")
    assert yb.write_image(p, img)
    raw = open(p, "rb").read()
    head = ("PF\n%d %d\n-1.0\n" % (W, H)).encode()
    assert raw.startswith(head)
    got = np.frombuffer(raw[len(head):], "<f4").reshape(H, W, 3)                  # PFM rows are bottom-up, like ours
    assert (got == img[..., :3]).all()
    p = str(tmp_path / "a.hdr")
    assert yb.write_image(p, img)
    raw = open(p, "rb").read()
    head = ("#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (H, W)).encode()
    assert raw.startswith(head)
    px = np.frombuffer(raw[len(head):], np.uint8).reshape(H, W, 4).astype(np.float64)
    val = px[..., :3] * np.exp2(px[..., 3:4] - 136.0) * (px[..., 3:4] > 0)       # mantissa / 256 * 2^(e - 128)
    want = img[::-1, :, :3].astype(np.float64)
    m = want.max(-1, keepdims=True)
    assert (np.abs(val - want) <= m / 128.0 + 1e-30).all()                        # 8-bit mantissa shared by the pixel
    assert (val[-1, 0] == 0).all()
    cv2 = pytest.importorskip("cv2")
    dec = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    if dec is not None:                                                           # OpenCV built with the Radiance codec
        assert dec.shape == (H, W, 3) and (np.abs(dec[..., ::-1] - want) <= m / 64.0 + 1e-30).all()


def test_error_codes(tmp_path):
    from yune_b200 import _native
    lib = _native.load()
    img = _picture(4, 4)
    ok = img.ctypes.data
    assert lib.yune_write_image(None, None, ok, 4, 4) == -1
    assert lib.yune_write_image(str(tmp_path / "a.png").encode(), None, None, 4, 4) == -1
    assert lib.yune_write_image(str(tmp_path / "a.png").encode(), None, ok, 0, 4) == -1
    assert lib.yune_write_image(str(tmp_path / "a.bmp").encode(), None, ok, 4, 4) == -2
    assert lib.yune_write_image(str(tmp_path / "noext").encode(), None, ok, 4, 4) == -2
    assert lib.yune_write_image(str(tmp_path / "no" / "such" / "dir.png").encode(), None, ok, 4, 4) == -3
    assert not os.path.exists(str(tmp_path / "a.bmp"))
    # explicit format, file name as given (RendererCore::saveImage(save_fn, save_ext), "Save At Samples")
    assert lib.yune_write_image(str(tmp_path / "shot").encode(), b".png", ok, 4, 4) == 0
    assert open(str(tmp_path / "shot"), "rb").read(8) == b"\x89PNG\r\n\x1a\n"
    assert yb.write_image(str(tmp_path / "shot2"), img, ".jpg") and open(str(tmp_path / "shot2"), "rb").read(2) == b"\xff\xd8"
    with pytest.raises(ValueError):
        yb.write_image(str(tmp_path / "a.png"), np.zeros((4, 4, 3), np.float32))

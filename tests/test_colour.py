"""Colour path: ImageIO::rgb2yuv / yuv2rgb (image/ImageIO.cpp:40-84) and the three planes of main.cpp:193-200."""
import os

import numpy as np
import pytest

from tests.cases import assert_items_equal


def _rgb(seed, h, w):
    """Smooth colour image + noise (random pixels alone never come close to a block match)."""
    rs = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 90 * np.sin(xx / 23.0 + seed) * np.cos(yy / 31.0), 128 + 100 * np.sin((xx + yy) / 40.0), 128 + 80 * np.cos(yy / 17.0 - xx / 29.0)], -1)
    return np.clip(base + rs.normal(0, 6, (h, w, 3)), 0, 255).astype(np.uint8)


def test_colour_conversion_restatement_equals_reference(fo):
    """Our C restatement against the REAL reference compiled from /root/reference, both FMA modes."""
    from oracle import pyoracle as po
    rs = np.random.default_rng(0)
    for shape in ((64, 96, 3), (130, 258, 3)):
        rgb = rs.integers(0, 256, shape, dtype=np.uint8)
        for fma in (False, True):
            ref = po.reference(fma=fma)
            if ref is None:
                pytest.skip("compiled reference absent")
            a, b = fo.rgb2yuv(rgb, fma), ref.rgb2yuv(rgb, fma)
            assert all((x == y).all() for x, y in zip(a, b))
            assert (fo.yuv2rgb(*b, fma) == ref.yuv2rgb(*b, fma)).all()
    # chroma subsampling: the last pixel of a 2x2 cell wins
    rgb = np.zeros((2, 2, 3), np.uint8)
    rgb[1, 1] = (255, 0, 0)
    _, u, v = fo.rgb2yuv(rgb)
    assert u[0, 0] == int(-0.169 * 255 + 128) and v[0, 0] == 255


@pytest.mark.gpu
def test_colour_conversion_gpu(ctx, fo):
    import fractencode_b200 as fb
    rs = np.random.default_rng(1)
    for shape in ((64, 96, 3), (256, 512, 3)):
        rgb = rs.integers(0, 256, shape, dtype=np.uint8)
        for fma in (False, True):
            got, want = ctx.rgb_to_yuv420(rgb, fma), fo.rgb2yuv(rgb, fma)
            assert all((x == y).all() for x, y in zip(got, want)), (shape, fma)
            assert (ctx.yuv420_to_rgb(*want, fma) == fo.yuv2rgb(*want, fma)).all()
    with pytest.raises(fb.FractencodeError):
        ctx.rgb_to_yuv420(np.zeros((63, 96, 3), np.uint8))


@pytest.mark.gpu
def test_three_plane_colour_encode_decode_round_trip(ctx, fo):
    """RGB -> YUV 4:2:0 -> the three planes quadtree-encoded as one pipelined call -> decoded -> RGB, every step compared with
    the oracle (the reference runs encode_image2 once per plane)."""
    import fractencode_b200 as fb
    rgb = _rgb(3, 256, 256)
    planes = ctx.rgb_to_yuv420(rgb)
    assert all((a == b).all() for a, b in zip(planes, fo.rgb2yuv(rgb)))
    p = fb.Params(6.0, -1.0, True)
    lists = ctx.encode_planes(planes, 16, 4, p)
    decoded = []
    for pl, got in zip(planes, lists):
        want, _ = fo.encode_quadtree(np.ascontiguousarray(pl), 16, 4, fo.params(6.0, -1.0, True))
        assert_items_equal(got, want)
        d, it, rms = ctx.decode(got, pl.shape[1], pl.shape[0])
        od, oit, orms = fo.decode(want, pl.shape[1], pl.shape[0])
        assert (d == od).all() and it == oit and rms == orms
        decoded.append(d)
    out = ctx.yuv420_to_rgb(*decoded)
    assert (out == fo.yuv2rgb(*decoded)).all()
    err = np.abs(out.astype(np.int32) - rgb.astype(np.int32)).mean()
    assert err < 12.0, err

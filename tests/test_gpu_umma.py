"""tcgen05/TMEM search path: integer-exactness known-answer tests at maximum magnitude, and A/B parity
with the exact integer (dp4a) path and the oracle.  All through the C ABI (fe_params.search_impl)."""
import numpy as np
import pytest

from tests.cases import assert_items_equal

pytestmark = pytest.mark.gpu

EXACT, UMMA = 1, 2


def _encode(ctx, img, S, T, impl, thr=0.0, cls=False, smax=-1.0, fma=False):
    import fractencode_b200 as fb
    H, W = img.shape
    ctx.set_image(img)
    dom, rng = fb.uniform_grid(W, H, S, S // 2), fb.uniform_grid(W, H, T, T)
    ctx.stats_reset()
    out = ctx.encode_level(dom, rng, fb.Params(thr, smax, cls, fma, impl))
    return out, ctx.stats()


def extreme_images():
    """Saturated content: the accumulator magnitudes the exactness proof is about."""
    rs = np.random.default_rng(3)
    W = H = 128
    imgs = {}
    imgs["black"] = np.zeros((H, W), np.uint8)
    imgs["white"] = np.full((H, W), 255, np.uint8)
    half = np.zeros((H, W), np.uint8)
    half[:, W // 2:] = 255
    imgs["half"] = half
    imgs["checker1"] = ((np.add.outer(np.arange(H), np.arange(W)) & 1) * 255).astype(np.uint8)
    imgs["checker8"] = (((np.add.outer(np.arange(H) // 8, np.arange(W) // 8)) & 1) * 255).astype(np.uint8)
    imgs["binary_noise"] = (rs.integers(0, 2, (H, W)) * 255).astype(np.uint8)
    blocks = np.kron(rs.integers(0, 2, (H // 16, W // 16)) * 255, np.ones((16, 16))).astype(np.uint8)
    imgs["binary_blocks"] = blocks
    near = blocks.copy()
    near[::3, ::5] ^= 1  # almost saturated, odd parities of sum(b^2)
    imgs["near_saturated"] = near
    return imgs


@pytest.mark.parametrize("name", list(extreme_images().keys()))
@pytest.mark.parametrize("T", [4, 8, 16, 32])
def test_umma_max_magnitude_kat(ctx, fo, name, T):
    """Known-answer test of the fp32 accumulator's integer exactness (SURVEY hard part 1): saturated
    blocks drive |sum a*b| to N*510^2.  AUTO must give the oracle's list bit for bit (falling back to the
    exact kernel only when the inexact band is flagged), and where UMMA runs it must equal the dp4a path."""
    img = extreme_images()[name]
    want = fo.encode_level(img, img, fo.uniform_grid(128, 128, 2 * T, T), fo.uniform_grid(128, 128, T, T), fo.params(0.0))
    got_auto, st = _encode(ctx, img, 2 * T, T, 0)
    assert_items_equal(got_auto, want, "%s T=%d auto" % (name, T))
    got_exact, _ = _encode(ctx, img, 2 * T, T, EXACT)
    assert_items_equal(got_exact, want, "%s T=%d exact" % (name, T))


@pytest.mark.parametrize("T", [4, 8, 16, 32])
@pytest.mark.parametrize("cls", [False, True])
@pytest.mark.parametrize("thr", [-1.0, 0.0, 30.0])
def test_umma_vs_exact_and_oracle(ctx, fo, T, cls, thr):
    rs = np.random.default_rng(100 + T)
    img = fo.synth_image(256, 192, 77, 0).copy()
    img[40:72, 60:100] = np.kron(rs.integers(0, 256, (4, 5), dtype=np.uint8), np.ones((8, 8), np.uint8))  # flat patches -> ties
    got_u, st = _encode(ctx, img, 2 * T, T, UMMA, thr=thr, cls=cls)
    assert st.umma_levels == 1 and st.exact_levels == 0
    got_e, st2 = _encode(ctx, img, 2 * T, T, EXACT, thr=thr, cls=cls)
    assert st2.exact_levels == 1
    assert_items_equal(got_u, got_e, "umma vs exact")
    dom, rng = fo.uniform_grid(256, 192, 2 * T, T), fo.uniform_grid(256, 192, T, T)
    if cls:
        dom, rng = fo.preclassify(img, dom), fo.preclassify(img, rng)
    want = fo.encode_level(img, img, dom, rng, fo.params(thr, -1.0, cls))
    assert_items_equal(got_u, want, "umma vs oracle")


def test_umma_pattern_ties(ctx, fo):
    """Pattern image: massive exact ties and zero distances -> exercises the (V, parity, column) tie rule."""
    img = fo.synth_image(128, 128, 1, 2)
    for T in (4, 8, 16):
        for thr in (-1.0, 0.0, 3.0):
            got, st = _encode(ctx, img, 2 * T, T, UMMA, thr=thr)
            want = fo.encode_level(img, img, fo.uniform_grid(128, 128, 2 * T, T), fo.uniform_grid(128, 128, T, T), fo.params(thr))
            assert_items_equal(got, want, "pattern T=%d thr=%g" % (T, thr))


def test_umma_few_ranges_many_domains(ctx, fo):
    """Few range blocks against a large pool: the column-chunked work split (atomicMin merge)."""
    import fractencode_b200 as fb
    img = fo.synth_image(512, 512, 9, 0)
    ctx.set_image(img)
    dom = fb.uniform_grid(512, 512, 8, 4)
    rng = fb.uniform_grid(512, 512, 4, 4)[5:3000:411]
    got = ctx.encode_level(dom, rng, fb.Params(0.0, -1.0, False, False, UMMA))
    want = fo.encode_level(img, img, dom, rng, fo.params(0.0))
    assert_items_equal(got, want)


def test_auto_uses_tensor_path(ctx, lenna):
    import fractencode_b200 as fb
    ctx.set_image(lenna)
    ctx.stats_reset()
    ctx.encode_quadtree(16, 4, fb.Params(5.0, -1.0, True))
    st = ctx.stats()
    assert st.umma_levels == 3 and st.exact_levels == 0  # T=16 on kind::i8, T=8 and T=4 on kind::f16


def test_i8_kind_small_blocks_and_odd_sizes(ctx, fo):
    """kind::i8 is exact for any K: non power-of-two blocks (T = 6, 12, 24: K padded to 64 / 160 / 768 bytes)."""
    import fractencode_b200 as fb
    img = fo.synth_image(192, 144, 21, 0)
    for T in (6, 12, 24):
        got, st = _encode(ctx, img, 2 * T, T, UMMA, thr=10.0)
        assert st.umma_levels == 1
        want = fo.encode_level(img, img, fo.uniform_grid(192, 144, 2 * T, T), fo.uniform_grid(192, 144, T, T), fo.params(10.0))
        assert_items_equal(got, want, "T=%d" % T)


@pytest.mark.parametrize("T", [4, 8, 16, 32])
def test_tensor_paths_at_scale_vs_exact(ctx, fo, T):
    """1024^2 image: many row tiles per CTA, many column tiles per work item, multi-stage K (T=32) -- the pipeline
    paths small images never reach.  The tensor result must equal the exact integer kernel's bit for bit, and a
    sample of ranges must equal the oracle."""
    import fractencode_b200 as fb
    W = H = 1024
    ctx.set_synthetic_image(W, H, 4321, 0)
    img = ctx.get_image()
    dom, rng = fb.uniform_grid(W, H, 2 * T, T), fb.uniform_grid(W, H, T, T)
    if T == 4:
        rng = rng[::7]  # keep the dp4a comparison run short
    outs = {}
    for impl in (UMMA, EXACT):
        ctx.stats_reset()
        outs[impl] = ctx.encode_level(dom, rng, fb.Params(20.0, -1.0, False, False, impl))
        st = ctx.stats()
        assert (st.umma_levels, st.exact_levels) == ((1, 0) if impl == UMMA else (0, 1))
    assert_items_equal(outs[UMMA], outs[EXACT], "T=%d tensor vs exact" % T)
    sub = rng[:: max(1, len(rng) // 24)][:24]
    want = fo.encode_level(img, img, dom, sub, fo.params(20.0))
    got = ctx.encode_level(dom, sub, fb.Params(20.0, -1.0, False, False, UMMA))
    assert_items_equal(got, want, "T=%d tensor vs oracle sample" % T)

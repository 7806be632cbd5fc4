"""Host-side planning of a threshold search (fe_plan_threshold: no GPU needed): the integer threshold and the brightness
bins must be SAFE -- every candidate the reference would accept lies inside the bins the GPU search looks at."""
import numpy as np
import pytest

import fractencode_b200 as fb
from fractencode_b200.capi import plan_threshold


def ref_distance(n16, S):
    """image/metrics.h:38-49 for an exact SSE = n16 / 16 below 2^20: float running sum, then double / (S*S)."""
    return float(np.float32(n16 / 16.0)) / float(S * S)


@pytest.mark.parametrize("T", [4, 8, 16, 32])
@pytest.mark.parametrize("thr", [0.0, 0.37, 5.0, 25.0, 100.0])
def test_integer_threshold_is_the_reference_threshold(T, thr):
    S = 2 * T
    pl = plan_threshold(thr, S, T)
    assert pl.use_threshold == 1
    assert ref_distance(pl.thr16, S) <= thr
    assert ref_distance(pl.thr16 + 1, S) > thr
    assert pl.radius ** 2 <= T * T * pl.thr16 < (pl.radius + 1) ** 2


def test_negative_threshold_never_hits_and_huge_threshold_is_refused():
    assert plan_threshold(-1.0, 16, 8).use_threshold == 0
    with pytest.raises(fb.FractencodeError):
        plan_threshold(2.0 ** 20 / 256.0, 16, 8)      # SSE >= 2^20: the reference's float sum starts to round


@pytest.mark.parametrize("T,thr", [(4, 25.0), (8, 25.0), (8, 3.0), (16, 25.0), (16, 80.0), (32, 25.0)])
def test_bins_contain_every_candidate_under_the_threshold(T, thr, fo):
    """Random range blocks, domains built to be close to them (down to exact copies) at every brightness: whenever the
    pair is under the threshold its two bins are at most bin_span apart."""
    S = 2 * T
    pl = plan_threshold(thr, S, T)
    assert pl.n_bins >= 2 * (2 * pl.bin_span + 1), "bins must be in use for these cases"
    assert pl.n_bins <= 64 and (1020 * T * T) // pl.bin_width + 1 == pl.n_bins
    rs = np.random.default_rng(T * 1000 + int(thr))
    hits = 0
    worst = 0
    for trial in range(4000):
        base = int(rs.integers(0, 256))
        r = np.clip(base + rs.integers(-40, 41, (T, T)), 0, 255).astype(np.int64)
        # domain box sums D = 4 r + noise, noise scaled so that a good share of the pairs straddles the threshold
        amp = int(rs.integers(0, 4 * int(np.sqrt(thr * 4) + 2)))
        D = np.clip(4 * r + rs.integers(-amp, amp + 1, (T, T)) + int(rs.integers(-amp, amp + 1)), 0, 1020)
        n16 = int(((4 * r - D) ** 2).sum())
        if n16 > pl.thr16:
            continue
        hits += 1
        sa, sb = int(4 * r.sum()), int(D.sum())
        assert abs(sa - sb) <= pl.radius
        ba, bb = min(sa // pl.bin_width, 63), min(sb // pl.bin_width, 63)
        assert abs(ba - bb) <= pl.bin_span
        worst = max(worst, abs(ba - bb))
    assert hits > 200, "the generator must produce candidates under the threshold (%d)" % hits
    assert worst >= 1, "and pairs that sit in different bins"


def test_threshold_matches_the_oracle_distance(fo):
    """One real pair through the oracle's estimate(): its distance sits on the same side of the threshold as n16 vs thr16."""
    T, S = 8, 16
    rs = np.random.default_rng(3)
    img = rs.integers(0, 256, (64, 64), dtype=np.uint8)
    dom, rng = fo.uniform_grid(64, 64, S, T), fo.uniform_grid(64, 64, T, T)
    out = fo.encode_level(img, img, dom, rng[:8], fo.params(-1.0))
    for it in out:
        for thr in (float(it["distance"]) * 0.999, float(it["distance"]), float(it["distance"]) * 1.001):
            pl = plan_threshold(thr, S, T)
            n16 = round(float(it["distance"]) * S * S * 16)   # exact regime: distance = (n16 / 16) / S^2
            assert (n16 <= pl.thr16) == (float(it["distance"]) <= thr)

#!/usr/bin/env python
"""Generate tests/golden/* from the REAL reference (oracle/_ref, built from /root/reference).

Run in the build container (needs /root/reference for the lenna PNG and for
oracle/_ref):   python tests/golden/make_golden.py
Outputs (committed):
  lenna512_luma.npz     raw luma of tests/input/lenna512x512.png as the reference's ImageIO
                        produces it in a non-FMA build (never convert at test time)
  goldens.json          per case: item count, md5 of the SURVEY-8c dump for the non-FMA and the
                        FMA build of the reference, quadtree level counts, decode iterations /
                        rms / FNV-1a64 of the decoded image
  items_<case>.npz      full 64-byte encode_item_t lists (non-FMA build) for the small cases
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FRAC_REFERENCE", "/root/reference")


def fnv1a64(b: bytes) -> str:
    h = 0xCBF29CE484222325
    for byte in b:
        h = ((h ^ byte) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def md5(items) -> str:
    return hashlib.md5(po.dump_lines(items).encode()).hexdigest()


# name, image, mode, params
CASES = [
    # BASELINE config 1 and its classifier twin; the reference's default 16->4 geometry
    dict(name="lenna_16_8", image="lenna", S=16, T=8),
    dict(name="lenna_16_8_cls", image="lenna", S=16, T=8, cls=True),
    dict(name="lenna_16_4", image="lenna", S=16, T=4),
    dict(name="lenna_8_4_cls", image="lenna", S=8, T=4, cls=True),
    dict(name="lenna_32_16", image="lenna", S=32, T=16),
    dict(name="lenna_64_32", image="lenna", S=64, T=32),
    dict(name="lenna_64_32_cls", image="lenna", S=64, T=32, cls=True),
    # BASELINE config 2: quadtree 16->4 + Classifier2 (threshold picked so all levels are populated)
    dict(name="lenna_qt_16_4_cls_thr5", image="lenna", qt=(16, 4), cls=True, thr=5.0),
    dict(name="lenna_qt_16_4_cls_thr2", image="lenna", qt=(16, 4), cls=True, thr=2.0),
    dict(name="lenna_qt_32_8_thr10", image="lenna", qt=(32, 8), thr=10.0),
    # synthetic 256x256 (SURVEY 8d generators, seed 1234)
    dict(name="nat256_16_8", image="natural256", S=16, T=8),
    dict(name="nat256_8_4_cls", image="natural256", S=8, T=4, cls=True),
    dict(name="nat256_32_16", image="natural256", S=32, T=16),
    dict(name="nat256_64_32", image="natural256", S=64, T=32),
    dict(name="nat256_qt_32_4_thr8", image="natural256", qt=(32, 4), thr=8.0),
    dict(name="nat256_qt_32_4_cls_thr20_smax1", image="natural256", qt=(32, 4), cls=True, thr=20.0, smax=1.0),
    dict(name="noise256_16_8", image="noise256", S=16, T=8),
    dict(name="noise256_32_16", image="noise256", S=32, T=16),  # best SSE >= 2^20: fp32 sequential-sum regime
    dict(name="noise256_64_32_cls", image="noise256", S=64, T=32, cls=True),
    dict(name="pattern256_16_8", image="pattern256", S=16, T=8),  # massive ties and zero distances
    dict(name="pattern256_8_4_cls", image="pattern256", S=8, T=4, cls=True),
    dict(name="pattern256_16_8_thr3", image="pattern256", S=16, T=8, thr=3.0),
    dict(name="pattern256_qt_16_4_thr1", image="pattern256", qt=(16, 4), thr=1.0),
    # non power-of-two block size (S*S is not a power of two -> the final divide rounds)
    dict(name="nat240_12_6", image="natural240", S=12, T=6),
    dict(name="nat240_24_12_cls_thr30", image="natural240", S=24, T=12, cls=True, thr=30.0),
    # rho = 4 geometries (reference default family)
    dict(name="nat256_16_4", image="natural256", S=16, T=4),
    dict(name="nat256_32_8_cls", image="natural256", S=32, T=8, cls=True),
]
FULL_ITEMS = {"lenna_16_8", "lenna_16_8_cls", "lenna_qt_16_4_cls_thr5", "nat256_32_16", "noise256_32_16",
              "pattern256_16_8", "nat240_12_6", "nat256_qt_32_4_thr8", "lenna_64_32", "nat256_16_4"}


def images(fo, fr):
    luma = fr.load_luma(os.path.join(REF, "tests/input/lenna512x512.png"))
    assert hashlib.md5(luma.tobytes()).hexdigest() == "4d1651d0106bf6167b6c7548c0865f34"
    return {
        "lenna": luma,
        "natural256": fo.synth_image(256, 256, 1234, 0),
        "noise256": fo.synth_image(256, 256, 1234, 1),
        "pattern256": fo.synth_image(256, 256, 1234, 2),
        "natural240": fo.synth_image(240, 240, 1234, 0),
    }


def run_case(lib, img, c, fma):
    H, W = img.shape
    p = lib.params(c.get("thr", 0.0), c.get("smax", -1.0), c.get("cls", False), fma)
    if "qt" in c:
        items, counts = lib.encode_quadtree(img, c["qt"][0], c["qt"][1], p)
        return items, counts
    dom = lib.uniform_grid(W, H, c["S"], c["S"] // 2)
    rng = lib.uniform_grid(W, H, c["T"], c["T"])
    if c.get("cls", False):
        dom, rng = lib.preclassify(img, dom), lib.preclassify(img, rng)
    return lib.encode_level(img, img, dom, rng, p), None


def main():
    po.build(ref=True)
    fo, fr, frf = po.restatement(), po.reference(False), po.reference(True)
    assert fr is not None and frf is not None, "oracle/_ref missing: needs /root/reference"
    imgs = images(fo, fr)
    np.savez_compressed(os.path.join(OUT, "lenna512_luma.npz"), luma=imgs["lenna"])
    gold = {"_generator": "tests/golden/make_golden.py", "_reference": fr.version(), "cases": {}}
    for c in CASES:
        img = imgs[c["image"]]
        items, counts = run_case(fr, img, c, False)
        items_f, _ = run_case(frf, img, c, True)
        dec, it, rms = fr.decode(items, img.shape[1], img.shape[0])
        entry = dict(c)
        entry.update(
            n_items=int(len(items)), md5_nofma=md5(items), md5_fma=md5(items_f), level_counts=counts,
            image_md5=hashlib.md5(img.tobytes()).hexdigest(),
            decode_iterations=int(it), decode_rms=float(rms), decode_fnv1a64=fnv1a64(dec.tobytes()),
            n_brightness_differs_fma=int((items["brightness"] != items_f["brightness"]).sum()),
        )
        if "qt" in entry:
            entry["qt"] = list(entry["qt"])
        gold["cases"][c["name"]] = entry
        if c["name"] in FULL_ITEMS:
            np.savez_compressed(os.path.join(OUT, "items_%s.npz" % c["name"]), items=po.sort_items(items), items_fma=po.sort_items(items_f))
        print(c["name"], entry["n_items"], entry["md5_nofma"], counts, it, flush=True)
    with open(os.path.join(OUT, "goldens.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

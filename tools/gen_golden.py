#!/usr/bin/env python3
"""Generates tests/golden/*.npz (run in the build container, where /root/reference is mounted).

Extractor vectors come from oracle/_ref/orb_ref = the reference's OWN src/ORBextractor.cc compiled in place against the
OpenCV stand-in (oracle/ref_shim), i.e. they are outputs of the reference itself, not of our restatement.
Matcher vectors: matcher_ref_vectors.npz holds outputs of oracle/_ref/libmatch_ref.so = the reference's OWN
src/ORBmatcher.cc compiled in place against the stand-ins of oracle/ref_shim/matcher (seeded cases of
tests/matcher_cases.py); matcher_vectors.npz holds the grid CSR and area queries of the reference's own Frame.cc / KeyFrame.cc text
and oracle-generated brute-force vectors (its pinned_by field says which is which).
Stereo vectors: stereo_ref_vectors.npz holds outputs of oracle/_ref/libstereo_ref.so = the reference's OWN text of
Frame::ComputeStereoMatches compiled in place (seeded cases of tests/stereo_cases.py; --stereo-only regenerates just these).
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from orbb200.synth import synth_frame, shifted_pair  # noqa: E402
from oracle_py import Oracle  # noqa: E402
import ref_runner  # noqa: E402
from datagen import planted_descriptors  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.makedirs(OUT, exist_ok=True)
    assert ref_runner.ref_binary("orb_ref"), "build oracle/_ref first (make -C oracle/ref_shim)"
    # ---- extractor: reference outputs
    cases = [("euroc_s0", 0, 752, 480, 1000, False), ("small_s5", 5, 400, 300, 1500, False),
             ("qvga_noise", 9, 320, 240, 500, True), ("kitti_s2", 2, 1241, 376, 2000, False)]
    for name, seed, w, h, nf, noise in cases:
        img = synth_frame(seed, w, h, noise_only=noise)
        r = ref_runner.ref_extract([img], nfeatures=nf, dump_levels=True)[0]
        small = w * h <= 400 * 300
        np.savez_compressed(os.path.join(OUT, "extract_%s.npz" % name),
                            seed=seed, w=w, h=h, nfeatures=nf, noise=noise, image_sha256=sha(img),
                            image=img if small else np.zeros(0, np.uint8),
                            kps=r["kps"], desc=r["desc"],
                            level_sha256=np.array([sha(l) for l in r["levels"]]),
                            level_shapes=np.array([l.shape for l in r["levels"]]), pinned_by="reference (oracle/_ref)")
        print(name, len(r["kps"]))
    # ---- quadtree KATs: reference DistributeOctTree on hand-made key sets
    o = Oracle()
    rng = np.random.default_rng(4)
    kats = {}

    def keys(xy, resp):
        k = np.zeros(len(xy), ref_runner.KP_DTYPE)
        k["x"], k["y"], k["response"] = xy[:, 0], xy[:, 1], resp
        k["size"], k["angle"], k["class_id"] = 7, -1, -1
        return k

    W, H = 720, 448
    kats["ties"] = (keys(np.stack([np.repeat(np.arange(8) * 90 + 3, 8) + np.tile(np.arange(8), 8) * 2.0,
                                   np.tile(np.arange(8) * 50 + 3.0, 8)], 1), np.full(64, 30.0)), 40)
    kats["equal_response"] = (keys(rng.integers(3, [W - 3, H - 3], (3000, 2)).astype(np.float64), np.full(3000, 25.0)), 217)
    kats["stop_mid_round"] = (keys(rng.integers(3, [W - 3, H - 3], (5000, 2)).astype(np.float64), rng.integers(7, 120, 5000)), 217)
    kats["single_key_roots"] = (keys(np.array([[10.0, 10.0], [600.0, 300.0]]), [20, 40]), 100)
    kats["one_root_empty"] = (keys(np.stack([rng.integers(3, 300, 500), rng.integers(3, H - 3, 500)], 1).astype(np.float64),
                                   rng.integers(7, 200, 500)), 60)
    kats["fewer_keys_than_n"] = (keys(rng.integers(3, [W - 3, H - 3], (90, 2)).astype(np.float64), rng.integers(7, 90, 90)), 217)
    kats["clustered"] = (keys(np.clip(rng.normal([360, 224], [25, 18], (4000, 2)), 3, [W - 4, H - 4]).round(),
                              rng.integers(7, 250, 4000)), 151)
    kats["n_zero"] = (keys(rng.integers(3, [W - 3, H - 3], (300, 2)).astype(np.float64), rng.integers(7, 90, 300)), 0)
    save = {}
    for name, (k, n) in kats.items():
        out = ref_runner.ref_distribute(k, 16, 16 + W, 16, 16 + H, n)
        assert out.tobytes() == o.distribute(k, 16, 16 + W, 16, 16 + H, n).tobytes(), name
        save[name + "_in"], save[name + "_out"], save[name + "_n"] = k, out, n
        print("octree", name, len(k), "->", len(out))
    np.savez_compressed(os.path.join(OUT, "octree_kats.npz"), win=np.array([16, 16 + W, 16, 16 + H]),
                        pinned_by="reference (oracle/_ref)", **save)
    # ---- matcher vectors (oracle-generated)
    a, b = shifted_pair(3, 400, 300)
    oe = o.extractor(800)
    ka, da = oe.extract(a)
    kb, db = oe.extract(b)
    bounds = (0.0, 0.0, 400.0, 300.0)
    f1, f2 = o.frame(ka, da, bounds), o.frame(kb, db, bounds)
    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
    n, m12, p = f1.search_init(f2, prev, 100, 0.9, True)
    q, qa, t, ta = planted_descriptors(np.random.default_rng(1), 300, 280)
    bn, best, second, idx, bm12 = o.bruteforce(q, qa, t, ta, 0.9, True)
    # ---- the grid and its area queries come from the reference's own Frame.cc / KeyFrame.cc text (oracle/ref_shim/grid)
    import ref_matcher
    from matcher_cases import CASES, run_case
    if not ref_matcher.available():
        ref_matcher.build()
    rm = ref_matcher.RefMatcher()
    gs, gi = rm.grid_csr(kb, bounds)
    rng = np.random.default_rng(5)
    aq = np.zeros(64, [("x", "f4"), ("y", "f4"), ("r", "f4"), ("min_level", "i4"), ("max_level", "i4")])
    aq["x"], aq["y"] = rng.uniform(-30, 430, 64), rng.uniform(-30, 330, 64)
    aq["r"] = rng.choice([3, 15, 40, 100], 64)
    aq["min_level"], aq["max_level"] = rng.choice([-1, 0, 2, 4], 64), rng.choice([-1, 1, 3, 7], 64)
    area = [rm.features_in_area(kb, bounds, float(a["x"]), float(a["y"]), float(a["r"]), int(a["min_level"]), int(a["max_level"]))
            for a in aq]
    area_kf = [rm.features_in_area(kb, bounds, float(a["x"]), float(a["y"]), float(a["r"]), keyframe=True) for a in aq]
    np.savez_compressed(os.path.join(OUT, "matcher_vectors.npz"), ka=ka, da=da, kb=kb, db=db, bounds=np.array(bounds),
                        init_n=n, init_m12=m12, init_prev=p, grid_start=gs, grid_idx=gi,
                        area_queries=aq, area_start=np.cumsum([0] + [len(a) for a in area]), area_idx=np.concatenate(area),
                        area_kf_start=np.cumsum([0] + [len(a) for a in area_kf]), area_kf_idx=np.concatenate(area_kf),
                        bf_q=q, bf_qa=qa, bf_t=t, bf_ta=ta, bf_n=bn, bf_best=best, bf_second=second, bf_idx=idx, bf_m12=bm12,
                        pinned_by="grid CSR and area queries: reference (oracle/_ref/libmatch_ref.so = the text of Frame.cc:574-589, "
                                  "671-736 and KeyFrame.cc:1138-1177 compiled in place); init_*: equal to the reference's "
                                  "SearchForInitialization case in matcher_ref_vectors.npz; bf_*: oracle (brute force has no reference "
                                  "function of its own: DescriptorDistance + the accept rules of ORBmatcher.cc:432-512)")
    print("matcher vectors: init", n, "bruteforce", bn, "grid members", len(gi), "area hits", sum(len(a) for a in area))
    # ---- matcher vectors from the reference's own ORBmatcher.cc
    r1, r2 = rm.frame(ka, da, bounds), rm.frame(kb, db, bounds)
    save = {}
    for c in CASES:
        out = run_case(c, r1, r2, ka, da, kb, db)
        for j, a in enumerate(out):
            save["%s__%d" % (c, j)] = np.asarray(a)
        print("reference", c, int(out[0]))
    np.savez_compressed(os.path.join(OUT, "matcher_ref_vectors.npz"), cases=np.array(CASES),
                        pinned_by="reference (oracle/_ref/libmatch_ref.so = /root/reference/src/ORBmatcher.cc compiled in place)",
                        **save)
    stereo_golden()


def stereo_golden():
    """stereo_ref_vectors.npz: mvuRight / mvDepth of the reference's own Frame::ComputeStereoMatches text
    (oracle/_ref/libstereo_ref.so) on the seeded cases of tests/stereo_cases.py"""
    import ref_stereo
    from stereo_cases import STEREO_CASES, images, oracle_inputs
    if not ref_stereo.available():
        ref_stereo.build()
    o = Oracle()
    save = {}
    for name in sorted(STEREO_CASES):
        left, right = images(name)
        ur, depth, kept = ref_stereo.stereo(*oracle_inputs(o, name))
        save[name + "/u_right"] = ur
        save[name + "/depth"] = depth
        save[name + "/kept"] = np.int32(kept)
        save[name + "/left_sha256"] = sha(left)
        save[name + "/right_sha256"] = sha(right)
        print("reference stereo", name, kept, "of", len(ur))
    np.savez_compressed(os.path.join(OUT, "stereo_ref_vectors.npz"),
                        pinned_by="reference (oracle/_ref/libstereo_ref.so = Frame::ComputeStereoMatches of "
                                  "/root/reference/src/Frame.cc compiled in place)", **save)


if __name__ == "__main__":
    if "--stereo-only" in sys.argv:
        stereo_golden()
    else:
        main()

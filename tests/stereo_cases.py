"""Seeded stereo cases shared by the reference pin (test_stereo_ref.py), the golden generator (tools/gen_golden.py) and
the GPU parity tests: name -> (width, height, nfeatures, mb, mbf, seed, planted disparities, identical_rows)."""
from datagen import stereo_pair

# EuRoC rig of Examples/Stereo/EuRoC.yaml (bf 47.906, fx 435.2), KITTI 00-02 (bf 386.1448, fx 718.856)
STEREO_CASES = {
    "euroc_s1": (752, 480, 1200, 47.90639384423901 / 435.2046959714599, 47.90639384423901, 1, (6, 14, 27, 41), 0),
    "euroc_s2": (752, 480, 1200, 47.90639384423901 / 435.2046959714599, 47.90639384423901, 2, (3, 9, 33, 60), 0),
    "kitti_s3": (1241, 376, 2000, 386.1448 / 718.856, 386.1448, 3, (6, 14, 27, 41), 0),
    # zero disparity with the top rows of both views identical: reaches the reference's disparity <= 0 branch
    "euroc_zero": (752, 480, 1200, 47.90639384423901 / 435.2046959714599, 47.90639384423901, 4, (0, 0, 0, 0), 150),
    # far larger disparity than mbf/mb allows in the lowest band: candidates rejected by the uR window
    "small_wide": (400, 300, 600, 0.5, 40.0, 5, (2, 20, 70, 95), 0),
}


def images(name):
    w, h, nfeat, mb, mbf, seed, disp, same_rows = STEREO_CASES[name]
    left, right = stereo_pair(seed, w, h, disparities=disp)
    if same_rows:
        right[:same_rows] = left[:same_rows]
    return left, right


def oracle_inputs(oracle, name):
    """(keysL, descL, keysR, descR, levelsL, levelsR, scale, inv_scale, mb, mbf) from the oracle extractor"""
    w, h, nfeat, mb, mbf = STEREO_CASES[name][:5]
    left, right = images(name)
    el, er = oracle.extractor(nfeat), oracle.extractor(nfeat)
    kl, dl = el.extract(left)
    kr, dr = er.extract(right)
    t = el.tables()
    LL = [el.level_padded(i) for i in range(8)]
    RR = [er.level_padded(i) for i in range(8)]
    return kl, dl, kr, dr, LL, RR, t["scale"], t["inv_scale"], mb, mbf

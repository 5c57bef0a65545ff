#!/usr/bin/env python3
"""Benchmark of the ORB front-end hot path on B200 (BASELINE.json metric: "ORB extract frames/s @752x480 1k kp;
Hamming compares/s; % roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl ours|reference]

One step = one pass of ORBextractor::operator() over a batch of B synthetic EuRoC-shaped frames per GPU
(BASELINE.json configs[2] shape; ORBextractor(1000, 1.2, 8, 20, 7)).  Prints ONE JSON line (rank 0):
  value        whole-job frames/s with the batch resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e          the same through the C-ABI host entry point orbx_extract_batch: pinned host input, H2D, kernels, D2H of
               keypoints/descriptors/counts inside the timed region
  roofline     dominant kernel of the extraction: algorithmic bytes per launch / its CUDA-event time, vs measured HBM peak
  hamming      config-4 brute-force matcher (1024 pairs of 2000x2000): compares/s and fraction of the measured POPC peak
  cpu_baseline the reference's CPU extractor (oracle/_ref, else the oracle port) on a bounded sample, single thread
`--impl reference` times the CPU reference on all host cores instead (rank 0 only).
Multi-GPU: frames are independent, so ranks shard the batch with no data-path collective (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "vi-orb-slam-icra2018_b200"))

W, H, NFEAT, NLEVELS, SCALE, INI_TH, MIN_TH = 752, 480, 1000, 8, 1.2, 20, 7
METRIC = "ORB extract frames/s @752x480 1k kp"


def level_sizes(w, h):
    import numpy as np
    s = np.float32(1.0)
    out = []
    for l in range(NLEVELS):
        inv = np.float32(1.0) / s
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
        s = np.float32(np.float64(s) * np.float64(np.float32(SCALE)))
    return out


def algorithmic_bytes(w, h, nkp):
    """SURVEY.md section 8d staged model, bytes per frame and per kernel group."""
    lv = level_sizes(w, h)
    px = sum(a * b for a, b in lv)
    padded = sum((a + 38) * (b + 38) for a, b in lv)
    pyr = w * h + sum(a * b for a, b in lv[:-1]) + padded
    return {"pyramid": pyr, "fast": px, "quadtree": 0, "blur": 2 * px, "brief": nkp * 60,
            "total": pyr + px + 2 * px + nkp * 60}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.p:
            self.p.terminate()
        # samples taken while the timed region ran (nvidia-smi reports ~100 ms late: widen the window a little)
        inside = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.15)]
        rows = inside or [r for _, r in self.rows]
        sm = sorted(int(float(r[0])) for r in rows if r and r[0].replace(".", "").isdigit())
        mx = max([int(float(r[1])) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4) if r[3 + i].startswith("Active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm), "samples_in_timed_region": len(inside)}


def cpu_reference(frames, procs, seconds_budget, native=True):
    """frames/s of the reference CPU extractor on `procs` cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    try:
        import ref_runner
        kind = "orb_ref_o3" if native and ref_runner.ref_binary("orb_ref_o3") else "orb_ref"
        if ref_runner.ref_binary(kind):
            probe = ref_runner.ref_bench(frames[:2], 1, nfeatures=NFEAT, kind=kind, procs=1)
            iters = max(1, int(seconds_budget / max(probe["ms_per_frame"] * 1e-3 * len(frames), 1e-3)))
            r = ref_runner.ref_bench(frames, iters, nfeatures=NFEAT, kind=kind, procs=procs)
            return {"value": r["frames_per_s"], "unit": "frames/s", "cores": procs, "kind": "reference",
                    "sample": "%d synthetic 752x480 frames x %d passes per core through oracle/_ref/%s (the reference's own "
                              "ORBextractor.cc on the cv2-pinned primitive models)" % (len(frames), iters, kind),
                    "ms_per_frame_one_core": r["ms_per_frame"],
                    "stage_ms": [r["ms_pyramid"], r["ms_keypoints"], r["ms_descriptors"]]}
    except Exception as ex:  # the binary may not run on this host; fall back to the port
        sys.stderr.write("oracle/_ref unavailable (%s), timing the oracle port\n" % ex)
    from oracle_py import Oracle
    ex = Oracle().extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < seconds_budget:
        ex.extract(frames[n % len(frames)])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": "%d extractions of synthetic 752x480 frames through oracle/liborb_oracle.so" % n}


def cpu_matcher_reference():
    """CPU side of the matcher numbers (part of the cpu_baseline leg): the reference's own ORBmatcher.cc
    (oracle/_ref/libmatch_ref.so, built with the parity flags -O2) on the inputs of the latency leg, one core, search call
    only; plus the oracle's brute-force loop (ORBmatcher.cc:432-461 restated) on one 2000x2000 pair."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from orbb200.synth import shifted_pair
    from oracle_py import Oracle
    out = {"cores": 1}
    o = Oracle()
    rng = np.random.default_rng(7)
    q = rng.integers(0, 256, (2000, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (2000, 32), dtype=np.uint8)
    qa = (rng.random(2000) * 360).astype(np.float32)
    ta = (rng.random(2000) * 360).astype(np.float32)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 1.5:
        o.bruteforce(q, qa, t, ta, 0.9, True)
        reps += 1
    out["bruteforce_compares_per_s"] = reps * 2000 * 2000 / (time.perf_counter() - t0)
    out["bruteforce_kind"] = "port (oracle/match_oracle.cpp, SWAR popcount as ORBmatcher.cc:1675-1691), parity build -O2"
    try:   # the reference's own flags (-O3 -march=native, CMakeLists.txt:10-11), compiled on THIS host: GCC turns the SWAR
        # count into vector popcounts where the CPU has them, which is worth up to 10x
        native = os.path.join(ROOT, "oracle", "liborb_oracle_native.so")
        if os.path.exists(native):
            os.remove(native)                    # never trust a binary built for another CPU
        on = Oracle(native=True)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 1.5:
            on.bruteforce(q, qa, t, ta, 0.9, True)
            reps += 1
        out["bruteforce_native_compares_per_s"] = reps * 2000 * 2000 / (time.perf_counter() - t0)
        out["bruteforce_native_kind"] = "same port, -O3 -march=native compiled on this host"
    except Exception as ex:
        out["bruteforce_native_kind"] = "unavailable: %s" % ex
    try:
        import ref_matcher
        if ref_matcher.available():
            rm = ref_matcher.RefMatcher()
            rm.lib.refm_last_search_ms.restype = __import__("ctypes").c_double
            sf = np.array([1.2 ** i for i in range(8)], np.float32)
            for name, (w, h) in (("euroc_752x480", (752, 480)), ("kitti_1241x376", (1241, 376))):
                a, b = shifted_pair(3, w, h)
                oe = o.extractor(2000, SCALE, NLEVELS, INI_TH, MIN_TH)
                ka, da = oe.extract(a)
                kb, db = oe.extract(b)
                bounds = (0.0, 0.0, float(w), float(h))
                f1, f2 = rm.frame(ka, da, bounds), rm.frame(kb, db, bounds)
                ts = []
                if name.startswith("euroc"):
                    prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
                    for _ in range(7):
                        f1.search_init(f2, prev, 100, 0.9, True)
                        ts.append(rm.lib.refm_last_search_ms())
                    out["search_for_initialization_ms"] = float(np.median(ts))
                else:
                    from oracle_py import PROJ_QUERY_DTYPE
                    pq = np.zeros(len(ka), PROJ_QUERY_DTYPE)
                    pq["u"], pq["v"], pq["invz"], pq["octave"], pq["valid"], pq["obsPositive"], pq["angle"] = \
                        ka["x"] + 7, ka["y"] + 3, 1.0, ka["octave"], 1, 1, ka["angle"]
                    for _ in range(7):
                        f2.search_projection(sf, pq, da, 15.0, 0, None, None, 0.0, True)
                        ts.append(rm.lib.refm_last_search_ms())
                    out["search_by_projection_th15_ms"] = float(np.median(ts))
            out["search_kind"] = "reference (ORBmatcher.cc compiled in place, -O2; time of the search call alone)"
    except Exception as ex:  # the .so may be missing on a box that never saw /root/reference
        out["search_kind"] = "unavailable: %s" % ex
    try:   # CPU side of the section-8f rows: the reference's own ComputeStereoMatches text, the oracle's distinctive loop
        import ref_stereo
        from orbb200.synth import stereo_pair
        left, right = stereo_pair(1, 752, 480)
        el, er = o.extractor(1200, SCALE, NLEVELS, INI_TH, MIN_TH), o.extractor(1200, SCALE, NLEVELS, INI_TH, MIN_TH)
        kl, dl = el.extract(left)
        kr, dr = er.extract(right)
        tb = el.tables()
        LL = [el.level_padded(i) for i in range(8)]
        RR = [er.level_padded(i) for i in range(8)]
        mb, mbf = 47.90639384423901 / 435.2046959714599, 47.90639384423901
        fn = ref_stereo.stereo if ref_stereo.available() else (lambda *a: o.stereo(*a))
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            fn(kl, dl, kr, dr, LL, RR, tb["scale"], tb["inv_scale"], mb, mbf)
            ts.append((time.perf_counter() - t0) * 1e3)
        out["stereo_matches_euroc_752x480_ms"] = float(np.median(ts))
        out["stereo_kind"] = ("reference (Frame::ComputeStereoMatches compiled in place behind cv::Mat stand-ins, -O2)"
                              if ref_stereo.available() else "port (oracle/match_oracle.cpp)")
        rng = np.random.default_rng(3)
        sizes = rng.integers(2, 41, 2000)
        start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
        t0 = time.perf_counter()
        o.distinctive(desc, start)
        out["distinctive_descriptors_20000_points_ms"] = (time.perf_counter() - t0) * 1e3 * 10
        out["distinctive_kind"] = "port (oracle/match_oracle.cpp), 2000 points timed and scaled x10"
        from orbb200.synth import make_vocab
        import ref_bow
        voc = make_vocab(seed=1, k=10, L=4)
        kw = dict(lib=ref_bow.lib(), fn="ref_bow_transform") if ref_bow.available() else {}
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            o.bow_transform(voc, dl, 4, **kw)
            ts.append((time.perf_counter() - t0) * 1e3)
        out["bow_transform_%d_features_k10_L4_ms" % len(dl)] = float(np.median(ts))
        out["bow_kind"] = ("reference (DBoW2 transform / FORB::distance / BowVector.cpp / FeatureVector.cpp compiled in place, "
                           "-O2, one thread)" if ref_bow.available() else "port (oracle/bow_oracle.cpp)")
    except Exception as ex:
        out["frame_side_kind"] = "unavailable: %s" % ex
    return out


def single_frame_latency(device, iters=30):
    """configs[0]: one 752x480 frame, ORBextractor(2000, ...) as the initialiser uses (Tracking.cc:822) + SearchForInitialization
    against the same scene shifted by (+7,+3); configs[1]: one 1241x376 frame, 2000 features + SearchByProjection(th=15).
    Host buffers in, host buffers out, median of `iters` calls in milliseconds."""
    import numpy as np
    import orbb200
    from orbb200.synth import shifted_pair

    def med(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            t0 = time.perf_counter()
            fn()
            ts.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(ts))

    out = {}
    m = orbb200.Matcher(device)
    sf = np.array([1.2 ** i for i in range(8)], np.float32)
    for name, (w, h) in (("euroc_752x480", (752, 480)), ("kitti_1241x376", (1241, 376))):
        a, b = shifted_pair(3, w, h)
        ex = orbb200.Extractor(2000, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=1, device=device)
        out["extract_%s_2000kp" % name] = med(lambda: ex(a))
        ka, da = ex(a)
        kb, db = ex(b)
        st = ex.stage_times()
        out["extract_%s_stage_ms" % name] = [float(x) for x in st]
        bounds = (0.0, 0.0, float(w), float(h))
        out["frame_grid_%s" % name] = med(lambda: m.frame(kb, db, bounds).close())
        f1, f2 = m.frame(ka, da, bounds), m.frame(kb, db, bounds)
        if name.startswith("euroc"):
            prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32)
            out["search_for_initialization"] = med(lambda: m.search_for_initialization(f1, f2, prev, 100, 0.9, True))
            out["search_for_initialization_matches"] = int(m.search_for_initialization(f1, f2, prev, 100, 0.9, True)[0])
        else:
            q = np.zeros(len(ka), orbb200.PROJ_QUERY_DTYPE)
            q["u"], q["v"], q["invz"], q["octave"], q["valid"], q["obs_positive"], q["angle"] = \
                ka["x"] + 7, ka["y"] + 3, 0.1, ka["octave"], 1, 1, ka["angle"]
            out["search_by_projection_th15"] = med(lambda: m.search_by_projection(f2, sf, q, da, 15.0, 0, None, None, 0.0, True))
            out["search_by_projection_matches"] = int(m.search_by_projection(f2, sf, q, da, 15.0, 0, None, None, 0.0, True)[0])
        f1.close(); f2.close(); ex.close()
    try:
        out.update(frame_side_latency(device, m, med))
    except Exception as exn:   # the widened rows (section 8f) must never cost the headline line
        out["frame_side_error"] = repr(exn)
    m.close()
    return out


def frame_side_latency(device, m, med):
    """SURVEY section 8f ranks 2-4, host to host unless stated: Frame post-extraction kept on the device
    (orbm_frame_create_device: undistortion + grid, EuRoC camera of Examples/Monocular/EuRoC.yaml), ComputeStereoMatches on
    an EuRoC-shaped rectified pair, ComputeDistinctiveDescriptors for a LocalMapping-sized batch of map points."""
    import numpy as np
    import torch
    import orbb200
    from orbb200.synth import stereo_pair
    out = {}
    w, h = 752, 480
    left, right = stereo_pair(1, w, h)
    mb, mbf = 47.90639384423901 / 435.2046959714599, 47.90639384423901
    exl = orbb200.Extractor(1200, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=1, device=device)
    exr = orbb200.Extractor(1200, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=1, device=device)
    kl, dl = exl(left)
    kr, dr = exr(right)
    out["stereo_matches_euroc_752x480"] = med(lambda: exl.stereo_matches(exr, kl, dl, kr, dr, mb, mbf))
    out["stereo_matches_kept"] = int(exl.stereo_matches(exr, kl, dl, kr, dr, mb, mbf)[2])
    out["stereo_frame_euroc_752x480"] = med(lambda: (exl(left), exr(right), exl.stereo_matches(exr, kl, dl, kr, dr, mb, mbf)))
    ex2 = orbb200.Extractor(1200, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=2, device=device)
    both = np.stack([left, right])

    def one_handle():
        (a, da), (b, db) = ex2.extract_batch(both)
        ex2.stereo_matches(ex2, a, da, b, db, mb, mbf, frame_l=0, frame_r=1)
    out["stereo_frame_one_batch_of_two"] = med(one_handle)
    ex2.close()
    # device-resident hand-over: keypoints stay where the extractor wrote them
    cam = orbb200.camera(458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)
    cap = exl.capacity
    with torch.cuda.device(device):
        d_img = torch.from_numpy(left[None]).cuda()
        d_k = torch.zeros((1, cap, 7), dtype=torch.int32, device="cuda")
        d_d = torch.zeros((1, cap, 32), dtype=torch.uint8, device="cuda")
        d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    exl.extract_batch_device(d_img, d_k, d_d, d_n)
    exl.synchronize()
    bounds = m.image_bounds(cam, w, h)
    out["frame_create_device_undistort_grid"] = med(
        lambda: orbb200.Frame.from_device(m, d_k[0], d_d[0], d_n, cap, bounds, cam).close())
    xy = np.stack([kl["x"], kl["y"]], 1)

    def host_chain():
        ku = kl.copy()
        un = m.undistort_points(cam, xy)
        ku["x"], ku["y"] = un[:, 0], un[:, 1]
        m.frame(ku, dl, bounds).close()
    out["frame_create_host_undistort_grid"] = med(host_chain)
    # distinctive descriptors: 20 000 map points with 2..40 observations
    rng = np.random.default_rng(3)
    sizes = rng.integers(2, 41, 20000)
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    desc = rng.integers(0, 256, (int(start[-1]), 32), dtype=np.uint8)
    out["distinctive_descriptors_20000_points"] = med(lambda: m.distinctive_descriptors(desc, start))
    out["distinctive_descriptors_observations"] = int(start[-1])
    # Frame::ComputeBoW: descent + BowVector / FeatureVector assembly, synthetic k=10 L=4 tree (ORBvoc itself is k=10 L=6)
    from orbb200.synth import make_vocab
    voc = make_vocab(seed=1, k=10, L=4)
    v = m.vocabulary(voc)
    out["bow_transform_%d_features_k10_L4" % len(dl)] = med(lambda: m.bow_transform(v, dl, 4))
    m.vocabulary_destroy(v)
    exl.close()
    exr.close()
    return out


def bind_to_gpu_numa(local):
    """Pin this process to the CPU cores next to its GPU before any pinned buffer is allocated: pinned pages are placed
    by first touch, and eight ranks copying from one NUMA node's memory is what halves the per-GPU H2D rate."""
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(path + "numa_node").read())
        cpus = set()
        for part in open(path + "local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        info.update({"numa_node": node, "local_cpus": len(cpus)})
        if cpus and node >= 0:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as ex:     # not fatal: the benchmark runs unbound
        info["error"] = str(ex)[:80]
    return info


def allpairs_leg(orbb200, dist, world, rank, local, dev, stream, barrier, max_over_ranks, steps):
    """configs[4]: all-pairs keyframe matching, 8192 keyframes x 1000 descriptors sharded by query block over the ranks, the
    descriptor table all-gathered in chunks over NVLink (NCCL) under the matching of the chunks that have landed.  The full
    matrix is 6.7e13 compares (15 s on 8 GPUs), so every rank matches a SAMPLE of 64 of its query keyframes against ALL
    8192 db keyframes; the collective is the full one (every rank receives the other ranks' whole blocks)."""
    import torch
    from orbb200 import shard
    n_kf, n_desc, q_per_rank, chunk_kf = 8192, 1000, 64, 512
    per = [e - b for b, e in (shard.block_range(n_kf, r, world) for r in range(world))]
    n_loc = per[rank]
    m = orbb200.Matcher(local)
    if dist:
        comm = shard.make_comm(dist, local)
    else:
        comm = orbb200.Comm(orbb200.Comm.unique_id(), 0, 1, local)
    g = torch.Generator(device=dev)
    g.manual_seed(4242)
    # a shared "map" of 1000 descriptors that every keyframe sees through 25 % bit noise in its first 400 slots, so that
    # keyframe pairs do have matches and the sequential one-to-one rule runs
    world_desc = torch.randint(0, 256, (n_desc, 32), dtype=torch.uint8, device=dev, generator=g)
    g.manual_seed(4243 + rank)
    desc = torch.randint(0, 256, (n_loc, n_desc, 32), dtype=torch.uint8, device=dev, generator=g)
    noise = (torch.randint(0, 256, (n_loc, 400, 32), dtype=torch.uint8, device=dev, generator=g)
             & torch.randint(0, 256, (n_loc, 400, 32), dtype=torch.uint8, device=dev, generator=g)
             & torch.randint(0, 256, (n_loc, 400, 32), dtype=torch.uint8, device=dev, generator=g))
    desc[:, :400] = world_desc[None, :400] ^ noise
    ang = torch.rand((n_loc, n_desc), device=dev, generator=g) * 360
    counts = torch.empty((q_per_rank, n_kf), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    def step():
        m.allpairs_sharded(comm, desc, ang, per, 0.75, True, counts, chunk_kf=chunk_kf, q_count=q_per_rank, stream=stream.cuda_stream)

    step()
    stream.synchronize()
    popc = m.popc_peak()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    stream.synchronize()
    gather = comm.last_gather()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    gather_ms = max_over_ranks(gather["ms"])
    compares = world * q_per_rank * n_kf * n_desc * n_desc
    out = {"workload": "configs[4]: %d keyframes x %d descriptors, query blocks of %d keyframes per rank, %d sampled query "
                       "keyframes per rank against all db keyframes; table all-gathered in %d chunks of %d keyframes per rank"
                       % (n_kf, n_desc, n_loc, q_per_rank, gather["chunks"], chunk_kf),
           "value": compares / (ms * 1e-3), "unit": "compares/s", "ms_per_step": ms, "scaling": "weak (queries per rank fixed)",
           "roofline": {"bound": "popc", "achieved": compares / world / (ms * 1e-3) * 8 / 1e9, "peak": popc / 1e9, "unit": "GPOPC32/s",
                        "frac": compares / world / (ms * 1e-3) * 8 / popc, "peak_source": "orbm_popc_peak microbenchmark in this run"},
           "collective": {"op": "ncclAllGather (C++ host, orbm_allpairs_sharded)", "nccl_version": comm.nccl_version(),
                          "bytes_received_per_rank": gather["bytes_received"], "gather_ms": gather_ms,
                          "nvlink_gbs_per_rank": (gather["bytes_received"] / (gather_ms * 1e-3) / 1e9) if gather_ms > 0 else None,
                          "gather_share_of_step": gather_ms / ms if ms > 0 else None},
           "gpu_launches": m.launch_count() * steps, "matches_mean": float(counts.float().mean().item())}
    # correctness on a small table: the sharded tile equals the single-GPU kernel on the gathered table
    if dist:
        small_kf, small_desc = 32 * world, 300
        sper = [e - b for b, e in (shard.block_range(small_kf, r, world) for r in range(world))]
        sd = desc[: sper[rank], :small_desc].contiguous()
        sa = ang[: sper[rank], :small_desc].contiguous()
        tile = torch.full((sper[rank], small_kf), -1, dtype=torch.int32, device=dev)
        ref = torch.full((sper[rank], small_kf), -1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()     # the fills above run on torch's current stream, the matcher on `stream`
        m.allpairs_sharded(comm, sd, sa, sper, 0.75, True, tile, chunk_kf=8, stream=stream.cuda_stream)
        stream.synchronize()
        gd = torch.empty((small_kf, small_desc, 32), dtype=torch.uint8, device=dev)
        ga = torch.empty((small_kf, small_desc), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gd, sd)
        dist.all_gather_into_tensor(ga, sa)
        torch.cuda.synchronize()
        q0 = sum(sper[:rank])
        m.allpairs_device(gd, ga, q0, q0 + sper[rank], 0, small_kf, 0.75, True, ref, stream=stream.cuda_stream)
        stream.synchronize()
        same = torch.tensor([int(torch.equal(tile, ref))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out["check"] = {"sharded_equals_single_gpu": bool(same.item()), "keyframes": small_kf, "descriptors": small_desc}
    comm.close()
    m.close()
    if dist and rank == 0:
        try:
            import glob
            import re
            lines = []
            for f in sorted(glob.glob("/tmp/orbb_nccl_rank0_*.log"), key=os.path.getmtime)[-1:]:
                for ln in open(f, errors="replace"):
                    if re.search(r"NVLS|Channel 0[01]/|Connected all|nRanks|via P2P|Trees|Rings|comm 0x", ln):
                        lines.append(ln.strip()[-160:])
            out["collective"]["nccl_log"] = lines[:6] + lines[-6:] if len(lines) > 12 else lines
        except Exception as ex:
            out["collective"]["nccl_log"] = ["unavailable: %s" % str(ex)[:80]]
    return out


def kitti_leg(orbb200, rank, local, dev, stream, barrier, max_over_ranks, steps, hbm):
    """configs[1]: KITTI-shaped 1241x376 frames, ORBextractor(2000, 1.2, 8, 20, 7): extraction throughput of a device-resident
    batch with its own HBM roofline, and windowed SearchByProjection (th = 15, mono) as a batch of frame pairs through
    orbm_search_by_projection_batch (host query arrays in, host matches out: the call a tracker would make for a batch)."""
    import numpy as np
    import torch
    from orbb200.synth import synth_frames_torch, shifted_pair
    w, h, nfeat, nb = 1241, 376, 2000, 1024
    out = {}
    d_img = torch.cat([synth_frames_torch(min(128, nb - i), w, h, seed=5000 + 1000 * rank + i, device=dev) for i in range(0, nb, 128)])
    ex = orbb200.Extractor(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=nb, device=local)
    cap = ex.capacity
    d_kps = torch.empty((nb, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.empty((nb, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.empty(nb, dtype=torch.int32, device=dev)
    for _ in range(3):
        ex.extract_batch_device(d_img, d_kps, d_desc, d_n, stream=stream.cuda_stream)
    stream.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ex.extract_batch_device(d_img, d_kps, d_desc, d_n, stream=stream.cuda_stream)
    e1.record(stream)
    stream.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    nkp = float(d_n.float().mean().item())
    alg = algorithmic_bytes(w, h, nkp)
    fps = nb / (ms * 1e-3)
    out["extraction"] = {"workload": "batch of %d synthetic KITTI 1241x376 frames per GPU, ORBextractor(2000,1.2,8,20,7), device resident" % nb,
                         "value_per_gpu": fps, "unit": "frames/s", "ms_per_step": ms, "keypoints_per_frame": nkp,
                         "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": alg["total"], "achieved_gbs": alg["total"] * fps / 1e9,
                                      "peak": hbm, "frac": alg["total"] * fps / 1e9 / hbm}}
    ex.close()
    del d_img, d_kps, d_desc
    # ---- SearchByProjection on 64 distinct frame pairs (shifted scenes), the batch repeated 4 x = 256 jobs per call
    m = orbb200.Matcher(local)
    exs = orbb200.Extractor(nfeat, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=w, max_height=h, max_batch=2, device=local)
    sf = np.array([1.2 ** i for i in range(8)], np.float32)
    bounds = (0.0, 0.0, float(w), float(h))
    jobs, keep = [], []
    for k in range(64):
        a, b = shifted_pair(100 + 64 * rank + k, w, h)
        (ka, da), (kb, db) = exs.extract_batch(np.stack([a, b]))
        f2 = m.frame(kb, db, bounds)
        q = np.zeros(len(ka), orbb200.PROJ_QUERY_DTYPE)
        q["u"], q["v"], q["invz"], q["octave"], q["valid"], q["obs_positive"], q["angle"] = \
            ka["x"] + 7, ka["y"] + 3, 0.1, ka["octave"], 1, 1, ka["angle"]
        keep.append(f2)
        jobs.append((f2, q, da, None, None))
    jobs = jobs * 4
    res, cand = m.search_by_projection_batch(jobs, sf, 15.0, mode=0, mbf=0.0, check_ori=True)
    n1, m1 = m.search_by_projection(jobs[5][0], sf, jobs[5][1], jobs[5][2], 15.0, 0, None, None, 0.0, True)
    same = bool(n1 == res[5][0] and np.array_equal(m1, res[5][1]))
    barrier()
    t0 = time.perf_counter()
    reps = max(2, steps // 2)
    for _ in range(reps):
        res, cand = m.search_by_projection_batch(jobs, sf, 15.0, mode=0, mbf=0.0, check_ori=True)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0) / reps
    nq = sum(len(j[1]) for j in jobs)
    out["search_by_projection"] = {"workload": "%d frame pairs per call (64 distinct, 2000 features each), th = 15, mono, rotation check; host query "
                                               "arrays in, host matches out" % len(jobs),
                                   "pairs_per_s_per_gpu": len(jobs) / dt, "queries_per_s_per_gpu": nq / dt,
                                   "candidate_compares_per_s_per_gpu": cand / dt, "candidates_per_query": cand / max(nq, 1),
                                   "ms_per_call": dt * 1e3, "matches_per_pair": float(np.mean([r[0] for r in res])),
                                   "gpu_launches_per_call": m.launch_count(), "batch_equals_single_call": same,
                                   "single_call_ms": None}
    t0 = time.perf_counter()
    for k in range(16):
        m.search_by_projection(jobs[k][0], sf, jobs[k][1], jobs[k][2], 15.0, 0, None, None, 0.0, True)
    out["search_by_projection"]["single_call_ms"] = (time.perf_counter() - t0) / 16 * 1e3
    for f in keep:
        f.close()
    exs.close()
    m.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from orbb200.synth import synth_frame
    frames = [synth_frame(s, W, H) for s in range(4)]
    procs = os.cpu_count() or 1
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.warmup + args.steps):
        vals.append(cpu_reference(frames, procs, seconds_budget=max(2.0, 40.0 / (args.warmup + args.steps))))
    base = vals[-1]
    timed = vals[args.warmup:]
    value = sum(v["value"] for v in timed) / len(timed)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t0) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[2] shape: synthetic EuRoC 752x480 frames, ORBextractor(1000,1.2,8,20,7), "
                                   "CPU reference on all host cores, bounded sample per step"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU per step (BASELINE.json configs[2]: batch of 4096)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-frame latency leg")
    ap.add_argument("--no-allpairs", action="store_true", help="skip the configs[4] all-pairs leg")
    ap.add_argument("--no-kitti", action="store_true", help="skip the configs[1] KITTI extraction + batched search leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import orbb200
    from orbb200.synth import synth_frames_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local)
    dist = None
    if world > 1:
        # NCCL's own communicator lines (rings / NVLS, channel counts) go to a file per rank; rank 0 quotes them in the
        # JSON line (allpairs.collective.nccl_log) so that the transport the gather used is on record
        os.environ["NCCL_DEBUG"] = "INFO"
        os.environ["NCCL_DEBUG_SUBSYS"] = "INIT,GRAPH"
        os.environ["NCCL_DEBUG_FILE"] = "/tmp/orbb_nccl_rank%d_%%p.log" % rank
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B, K, Wm = args.frames, args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local)
    sampler.start()      # nvidia-smi takes ~1 s to produce its first line: start it before the inputs are generated

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs: B distinct frames per rank, resident in HBM (B*361 KB > 126 MB L2 for B >= 350) and pinned on the host
    chunks = [synth_frames_torch(min(128, B - i), W, H, seed=1000 * rank + i, device=dev) for i in range(0, B, 128)]
    d_images = torch.cat(chunks)
    del chunks
    h_images = torch.empty((B, H, W), dtype=torch.uint8, pin_memory=True)
    h_images.copy_(d_images)
    ex = orbb200.Extractor(NFEAT, SCALE, NLEVELS, INI_TH, MIN_TH, max_width=W, max_height=H, max_batch=B, device=local)
    cap = ex.capacity
    d_kps = torch.empty((B, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)

    def step():
        ex.extract_batch_device(d_images, d_kps, d_desc, d_n, stream=stream.cuda_stream)

    for _ in range(Wm):
        step()
    stream.synchronize()
    launches_per_step = ex.launch_count()
    nkp = float(d_n.float().mean().item())

    # ---- timed region: K steps, kernels bracketed by events on the launching stream
    barrier()
    sampler.mark_begin()
    ex.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    stream.synchronize()
    sampler.mark_end()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    kernel_ms, ncalls = ex.kernel_times()
    ex.set_profiling(False)
    clocks = sampler.stop()
    value = world * B * K / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C ABI (pinned input, H2D + kernels + D2H per step)
    h_kps = torch.empty((B, cap, 7), dtype=torch.int32, pin_memory=True)
    h_desc = torch.empty((B, cap, 32), dtype=torch.uint8, pin_memory=True)
    h_n = torch.empty(B, dtype=torch.int32, pin_memory=True)
    import ctypes as C

    def e2e_step():
        orbb200._check(ex.L.orbx_extract_batch(ex.h, C.c_void_p(h_images.data_ptr()), B, W, H, W, C.c_size_t(W * H),
                                               C.c_void_p(h_kps.data_ptr()), C.c_void_p(h_desc.data_ptr()), cap,
                                               C.c_void_p(h_n.data_ptr())))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, K // 2)
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * e2e_steps / e2e_s
    # what the link alone gives: the same pinned input copied with nothing else running (context for the e2e number)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d_images.copy_(h_images, non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(3):
        d_images.copy_(h_images, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * B * W * H / (c0.elapsed_time(c1) * 1e-3) / 1e9
    # ... and with every rank copying at the same time: the ceiling the e2e number of this N can reach
    barrier()
    c0.record()
    for _ in range(3):
        d_images.copy_(h_images, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    conc_ms = max_over_ranks(c0.elapsed_time(c1))
    h2d_conc_gbs = 3 * B * W * H / (conc_ms * 1e-3) / 1e9      # per GPU, slowest rank
    e2e_launches = ex.launch_count() * e2e_steps

    # ---- strong scaling of configs[2]: the SAME 4096 frames in total, cut over the ranks (device resident)
    Bs = max(1, B // world)
    for _ in range(2):
        ex.extract_batch_device(d_images[:Bs], d_kps, d_desc, d_n, stream=stream.cuda_stream)
    stream.synchronize()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for _ in range(K):
        ex.extract_batch_device(d_images[:Bs], d_kps, d_desc, d_n, stream=stream.cuda_stream)
    s1.record(stream)
    stream.synchronize()
    barrier()
    strong_ms = max_over_ranks(s0.elapsed_time(s1))

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        alg = algorithmic_bytes(W, H, nkp)
        names = ["pyramid", "fast", "quadtree", "blur", "brief"]
        per_kernel = {n: kernel_ms[i] / max(ncalls, 1) for i, n in enumerate(names)}
        dom = max(names, key=lambda n: per_kernel[n])
        # the quadtree kernel moves almost no bytes (latency-bound list surgery): its roofline entry is the time share
        dom_bw = "fast" if dom == "quadtree" else dom
        ach = alg[dom_bw] * B / (per_kernel[dom_bw] * 1e-3) / 1e9
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per frame (profiles/)
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tr[dom_bw]["dram_bytes_per_frame"] * B
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[2] shape: batch of %d synthetic EuRoC 752x480 frames per GPU, "
                                   "ORBextractor(1000,1.2,8,20,7), extraction sharded by frame" % B,
                       "frames_per_gpu_per_step": B, "l2": "inputs larger than L2 (%d MB per step per GPU)" % (B * W * H >> 20),
                       "keypoints_per_frame": nkp},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * W * H,
                    "d2h_bytes_per_step": B * (cap * 60 + 4), "steps": e2e_steps,
                    "h2d_gbs_in_e2e": e2e_value / world * W * H / 1e9, "h2d_gbs_link_alone": h2d_gbs,
                    "h2d_gbs_all_ranks_copying": h2d_conc_gbs,
                    "frames_per_s_ceiling_of_the_link": world * h2d_conc_gbs * 1e9 / (W * H),
                    "frac_of_link_ceiling": (e2e_value / world * W * H / 1e9) / h2d_conc_gbs, "numa": numa},
            "strong_scaling": {"workload": "configs[2]: %d frames in total, %d per GPU" % (Bs * world, Bs),
                               "value": Bs * world * K / (strong_ms * 1e-3), "unit": "frames/s", "ms_per_step": strong_ms / K},
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": "hbm", "kernel": dom_bw, "achieved": ach, "peak": hbm, "unit": "GB/s",
                         "frac": ach / hbm, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "algorithmic_bytes_per_launch": alg[dom_bw] * B, "launch_ms": per_kernel[dom_bw],
                         "dominant_by_time": dom},
            "kernels_ms_per_step": per_kernel,
            # every stage against the HBM roofline (north star: pyramid, blur and extraction as fractions of HBM GB/s)
            "stage_rooflines": {n: {"algorithmic_bytes_per_frame": alg[n], "ms_per_step": per_kernel[n],
                                    "achieved_gbs": alg[n] * B / (per_kernel[n] * 1e-3) / 1e9,
                                    "frac_of_hbm": alg[n] * B / (per_kernel[n] * 1e-3) / 1e9 / hbm}
                                for n in names if alg[n] > 0},
            "extract_roofline": {"algorithmic_bytes_per_frame": alg["total"],
                                 "achieved_gbs": alg["total"] * value / world / 1e9,
                                 "frac_of_hbm": alg["total"] * value / world / 1e9 / hbm},
        }

    # ---- Hamming matcher, config 4: 1024 pairs of 2000 x 2000 descriptors, ratio test + rotation histogram
    if not args.no_hamming:
        m = orbb200.Matcher(local)
        P, n = 1024, 2000
        g = torch.Generator(device=dev)
        g.manual_seed(7 + rank)
        q = torch.randint(0, 256, (P, n, 32), dtype=torch.uint8, device=dev, generator=g)
        t = torch.randint(0, 256, (P, n, 32), dtype=torch.uint8, device=dev, generator=g)
        t[:, : n // 2] = q[:, : n // 2] ^ (torch.randint(0, 256, (P, n // 2, 32), dtype=torch.uint8, device=dev, generator=g)
                                          & torch.randint(0, 256, (P, n // 2, 32), dtype=torch.uint8, device=dev, generator=g)
                                          & torch.randint(0, 256, (P, n // 2, 32), dtype=torch.uint8, device=dev, generator=g))
        qa = torch.rand((P, n), device=dev, generator=g) * 360
        ta = torch.rand((P, n), device=dev, generator=g) * 360
        outs = [torch.empty((P, n), dtype=torch.int32, device=dev) for _ in range(4)]
        nm = torch.empty(P, dtype=torch.int32, device=dev)

        def hstep():
            m.bruteforce_device(q, qa, t, ta, 0.9, True, outs[0], outs[1], outs[2], outs[3], nm, stream=stream.cuda_stream)

        for _ in range(3):
            hstep()
        stream.synchronize()
        popc = m.popc_peak()
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        for _ in range(K):
            hstep()
        h1.record(stream)
        stream.synchronize()
        barrier()
        hms = max_over_ranks(h0.elapsed_time(h1))
        cps = world * P * n * n * K / (hms * 1e-3)
        if rank == 0:
            line["hamming"] = {"workload": "configs[3]: %d pairs/GPU of 2000x2000 descriptors, ratio 0.9 + rotation histogram" % P,
                               "value": cps, "unit": "compares/s", "ms_per_step": hms / K,
                               "roofline": {"bound": "popc", "achieved": cps / world * 8 / 1e9, "peak": popc / 1e9,
                                            "unit": "GPOPC32/s", "frac": cps / world * 8 / popc,
                                            "peak_source": "orbm_popc_peak microbenchmark in this run"},
                               "gpu_launches": 2 * K, "matches_per_pair": float(nm.float().mean().item())}
        m.close()

    # ---- configs[4]: all-pairs keyframe matching with the NCCL all-gather of the descriptor table
    if not args.no_allpairs:
        ap_line = allpairs_leg(orbb200, dist, world, rank, local, dev, stream, barrier, max_over_ranks, max(2, K // 5))
        if rank == 0:
            line["allpairs"] = ap_line

    # ---- configs[1]: KITTI-shaped extraction + batched windowed SearchByProjection
    if not args.no_kitti:
        kl = kitti_leg(orbb200, rank, local, dev, stream, barrier, max_over_ranks, K,
                       float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
                       if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0)
        if rank == 0:
            line["kitti"] = kl

    # ---- single-frame latency of configs[0] / configs[1] through the host C ABI (what a live SLAM loop sees)
    if rank == 0 and world == 1 and not args.no_latency:
        line["latency_ms"] = single_frame_latency(local)

    # ---- CPU baseline beside it (rank 0, N=1 only): bounded sample, single thread
    if rank == 0 and world == 1 and not args.no_cpu:
        frames = [h_images[i].numpy().copy() for i in range(4)]
        line["cpu_baseline"] = cpu_reference(frames, 1, seconds_budget=12.0)
        line["cpu_baseline"]["matcher"] = cpu_matcher_reference()
    if rank == 0:
        print(json.dumps(line))
    ex.close()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

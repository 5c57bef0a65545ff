"""TEST INFRASTRUCTURE ONLY -- runs oracle/_ref/orb_ref (the reference's own ORBextractor.cc built against the shim)."""
import json
import os
import struct
import subprocess
import tempfile

import numpy as np

from oracle_py import KP_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_binary(kind="orb_ref"):
    p = os.path.join(HERE, "_ref", kind)
    return p if os.path.exists(p) else None


def _request(mode, w, h, nframes, nfeatures, nlevels, ini, mn, dump, scale):
    return struct.pack("<10if", 0x0B1, mode, w, h, nframes, nfeatures, nlevels, ini, mn, dump, scale)


def ref_extract(frames, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, dump_levels=False, kind="orb_ref"):
    """frames: list of equally sized uint8 arrays -> list of dicts(kps, desc[, levels])."""
    exe = ref_binary(kind)
    assert exe, "oracle/_ref not built"
    h, w = frames[0].shape
    with tempfile.TemporaryDirectory() as d:
        rq, rp = os.path.join(d, "rq.bin"), os.path.join(d, "rp.bin")
        with open(rq, "wb") as f:
            f.write(_request(0, w, h, len(frames), nfeatures, nlevels, ini, mn, int(dump_levels), scale))
            for fr in frames:
                f.write(np.ascontiguousarray(fr, np.uint8).tobytes())
        subprocess.check_call([exe, rq, rp])
        buf = open(rp, "rb").read()
    out, off = [], 0
    for _ in frames:
        n = struct.unpack_from("<i", buf, off)[0]; off += 4
        kps = np.frombuffer(buf, KP_DTYPE, n, off).copy(); off += 28 * n
        desc = np.frombuffer(buf, np.uint8, 32 * n, off).reshape(n, 32).copy(); off += 32 * n
        r = dict(kps=kps, desc=desc)
        if dump_levels:
            r["levels"] = []
            for _l in range(nlevels):
                lw, lh = struct.unpack_from("<2i", buf, off); off += 8
                sz = (lw + 38) * (lh + 38)
                r["levels"].append(np.frombuffer(buf, np.uint8, sz, off).reshape(lh + 38, lw + 38).copy()); off += sz
        out.append(r)
    return out


def ref_distribute(keys, minX, maxX, minY, maxY, N, kind="orb_ref"):
    exe = ref_binary(kind)
    assert exe, "oracle/_ref not built"
    keys = np.ascontiguousarray(keys, KP_DTYPE)
    with tempfile.TemporaryDirectory() as d:
        rq, rp = os.path.join(d, "rq.bin"), os.path.join(d, "rp.bin")
        with open(rq, "wb") as f:
            f.write(_request(1, 0, 0, 0, 1000, 8, 20, 7, 0, 1.2))
            f.write(struct.pack("<6i", minX, maxX, minY, maxY, N, len(keys)))
            f.write(keys.tobytes())
        subprocess.check_call([exe, rq, rp])
        buf = open(rp, "rb").read()
    n = struct.unpack_from("<i", buf, 0)[0]
    return np.frombuffer(buf, KP_DTYPE, n, 4).copy()


def ref_bench(frames, iters, nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7, kind="orb_ref", procs=1):
    """Times the reference extractor; `procs` copies run concurrently (one per host core). Returns aggregated dict."""
    exe = ref_binary(kind)
    assert exe, "oracle/_ref not built"
    h, w = frames[0].shape
    with tempfile.TemporaryDirectory() as d:
        rq = os.path.join(d, "rq.bin")
        with open(rq, "wb") as f:
            f.write(_request(0, w, h, len(frames), nfeatures, nlevels, ini, mn, 0, scale))
            for fr in frames:
                f.write(np.ascontiguousarray(fr, np.uint8).tobytes())
        import time
        t0 = time.perf_counter()
        ps = [subprocess.Popen([exe, rq, os.path.join(d, "x"), str(iters)], stdout=subprocess.PIPE) for _ in range(procs)]
        outs = [json.loads(p.communicate()[0].decode()) for p in ps]
        wall = time.perf_counter() - t0
    frames_done = sum(o["frames"] for o in outs)
    return dict(procs=procs, frames=frames_done, wall_s=wall, frames_per_s=frames_done / wall,
                ms_per_frame=float(np.mean([o["ms_per_frame"] for o in outs])),
                ms_pyramid=float(np.mean([o["ms_pyramid"] for o in outs])),
                ms_keypoints=float(np.mean([o["ms_keypoints"] for o in outs])),
                ms_descriptors=float(np.mean([o["ms_descriptors"] for o in outs])),
                kp_per_frame=float(np.mean([o["kp_per_frame"] for o in outs])))

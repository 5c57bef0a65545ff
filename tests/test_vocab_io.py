"""adapter/ORBVocabulary.h reads and writes the reference's two vocabulary file formats (System.cc:334-339 ->
DBoW2 loadFromTextFile / loadFromBinaryFile).  CPU only.  The tree it builds must equal, array for array, the tree DBoW2's
OWN readers build from the same file (oracle/_ref/libvocio_ref.so = those functions compiled in place), including the two
artefacts of their eof loops (the node made from the empty last line of a text file, the last binary record stored twice),
and the files it writes must equal DBoW2's own writers' byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_vocio  # noqa: E402
from bow_cases import make_vocab  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_vocio.available() and not os.path.isdir("/root/reference"),
                               reason="oracle/_ref/libvocio_ref.so is built only where /root/reference is mounted")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("vocio") / "vocab_io")
    pkg = os.path.join(ROOT, "vi-orb-slam-icra2018_b200")
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-I", os.path.join(pkg, "adapter"), "-o", out,
                           os.path.join(ROOT, "tests", "vocab_io.cpp"), "-L", pkg, "-lorbb200", "-Wl,-rpath," + pkg])
    return out


def write_text(voc, path, trailing_newline=True, weight_fmt="%.6g"):
    """the layout saveToTextFile produces (TemplatedVocabulary.h:1654-1676)"""
    parent = np.zeros(voc["n_nodes"], np.int64)
    for p in range(voc["n_nodes"]):
        parent[voc["children"][voc["child_start"][p]:voc["child_start"][p + 1]]] = p
    lines = ["%d %d  %d %d" % (voc["k"], voc["L"], 0, 0)]
    for i in range(1, voc["n_nodes"]):
        leaf = voc["child_start"][i + 1] == voc["child_start"][i]
        lines.append("%d %d %s  %s" % (parent[i], int(leaf), " ".join(str(int(b)) for b in voc["desc"][i]),
                                       weight_fmt % voc["weight"][i]))
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if trailing_newline else ""))


def read_dump(path):
    raw = open(path, "rb").read()
    n, k, L, sc, wt, ne = np.frombuffer(raw, np.int32, 6)
    o = 24
    def take(dtype, count):
        nonlocal o
        a = np.frombuffer(raw, dtype, count, o)
        o += a.nbytes
        return a
    d = dict(n_nodes=int(n), k=int(k), L=int(L), scoring=int(sc), weighting=int(wt))
    d["desc"] = take(np.uint8, n * 32).reshape(n, 32)
    d["parent"] = take(np.int32, n)
    d["child_start"] = take(np.int32, n + 1)
    d["children"] = take(np.int32, ne)
    d["word_id"] = take(np.int32, n)
    d["is_word"] = take(np.int32, n)
    d["weight"] = take(np.float64, n)
    return d


def same_tree(a, b):
    for key in ("n_nodes", "k", "L", "scoring", "weighting"):
        assert a[key] == b[key], key
    for key in ("desc", "parent", "child_start", "children", "word_id"):
        assert np.array_equal(a[key], b[key]), key
    assert int(a["is_word"].sum()) == b["n_words"]
    assert np.array_equal(a["weight"].view(np.uint64), b["weight"].view(np.uint64))


def run(exe, src, kind, tmp_path, tag):
    dump, otxt, obin = (str(tmp_path / (tag + s)) for s in (".dump", ".txt", ".bin"))
    r = subprocess.run([exe, src, kind, dump, otxt, obin], capture_output=True, text=True)
    return r.returncode, dump, otxt, obin


@needs_ref
@pytest.mark.parametrize("spec", [dict(seed=1, k=10, L=3), dict(seed=2, k=6, L=4), dict(seed=3, k=9, L=4, ragged=True)])
def test_round_trip_equals_dbow2(exe, tmp_path, spec):
    if not ref_vocio.available():
        ref_vocio.build()
    voc = make_vocab(**spec)
    src = str(tmp_path / "voc.txt")
    write_text(voc, src)
    # text -> arrays
    rc, dump, otxt, obin = run(exe, src, "text", tmp_path, "a")
    assert rc == 0
    mine, ref = read_dump(dump), ref_vocio.load(src, False)
    same_tree(mine, ref)
    assert mine["n_nodes"] == voc["n_nodes"] + 1                 # + the node made from the empty last line
    # ... which inherits the last line's parent and leaf flag, with weight 0 and no descriptor bytes
    assert mine["parent"][-1] == mine["parent"][-2] and mine["is_word"][-1] == 1 and not mine["desc"][-1].any()
    assert mine["weight"][-1] == 0
    assert np.array_equal(mine["desc"][:-1], voc["desc"]) and np.array_equal(mine["children"][mine["children"] < voc["n_nodes"]],
                                                                              voc["children"])
    # the writers: byte for byte what DBoW2's writers produce from the same tree
    rtxt, rbin = str(tmp_path / "r.txt"), str(tmp_path / "r.bin")
    assert ref_vocio.resave(src, False, rtxt, rbin) == mine["n_nodes"]
    assert open(otxt, "rb").read() == open(rtxt, "rb").read()
    assert open(obin, "rb").read() == open(rbin, "rb").read()
    # binary -> arrays (the last record is stored twice), and text written by us -> arrays again
    rc, dump2, otxt2, obin2 = run(exe, obin, "binary", tmp_path, "b")
    assert rc == 0
    mine2, ref2 = read_dump(dump2), ref_vocio.load(obin, True)
    same_tree(mine2, ref2)
    assert mine2["n_nodes"] == mine["n_nodes"] + 1 and np.array_equal(mine2["desc"][-1], mine2["desc"][-2])
    rc, dump3, _, _ = run(exe, otxt, "text", tmp_path, "c")
    assert rc == 0
    same_tree(read_dump(dump3), ref_vocio.load(otxt, False))


@needs_ref
def test_text_without_final_newline_and_odd_tokens(exe, tmp_path):
    """no trailing newline -> no extra node; a malformed byte token stops the descriptor parse where FORB::fromString stops"""
    voc = make_vocab(seed=5, k=4, L=2)
    src = str(tmp_path / "voc.txt")
    write_text(voc, src, trailing_newline=False, weight_fmt="%.17g")
    lines = open(src).read().split("\n")
    toks = lines[3].split(" ")
    toks[5] = "12abc"                                            # parsed as 12, then the stream fails: the rest stays 0
    lines[3] = " ".join(toks)
    open(src, "w").write("\n".join(lines))
    rc, dump, _, _ = run(exe, src, "text", tmp_path, "a")
    assert rc == 0
    mine = read_dump(dump)
    same_tree(mine, ref_vocio.load(src, False))
    assert mine["n_nodes"] == voc["n_nodes"]
    assert mine["desc"][3][3] == 12 and not mine["desc"][3][4:].any()


def test_refused_files(exe, tmp_path):
    bad = str(tmp_path / "bad.txt")
    open(bad, "w").write("25 6  0 0\n0 0 " + "1 " * 32 + " 0.5\n")          # k > 20: "not a correct text file"
    assert run(exe, bad, "text", tmp_path, "x")[0] == 3
    assert run(exe, str(tmp_path / "missing.txt"), "text", tmp_path, "y")[0] == 3
    orphan = str(tmp_path / "orphan.txt")
    open(orphan, "w").write("10 2  0 0\n7 1 " + "1 " * 32 + " 0.5")          # parent 7 does not exist yet
    assert run(exe, orphan, "text", tmp_path, "z")[0] == 3
    assert run(exe, str(tmp_path / "missing.bin"), "binary", tmp_path, "w")[0] == 3
    if ref_vocio.available():
        assert ref_vocio.load(bad, False) is None

"""Pins the oracle's BoW transform (oracle/bow_oracle.cpp) against DBoW2's OWN text as vendored in the reference:
TemplatedVocabulary::transform(feature, ...) and FORB::distance streamed into the compiler, BowVector.cpp and
FeatureVector.cpp compiled in place (oracle/_ref/libbow_ref.so).  Word ids, node ids, idf weights per feature, the
L1-normalised BowVector (doubles, bit for bit) and the FeatureVector must be identical."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_bow  # noqa: E402
from bow_cases import VOCABS, features_for, make_vocab  # noqa: E402


def same_bow(a, b):
    for k in ("word", "node", "bow_word", "fv_node", "fv_start", "fv_idx"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["weight"].view(np.uint64), b["weight"].view(np.uint64))
    assert np.array_equal(a["bow_value"].view(np.uint64), b["bow_value"].view(np.uint64))


@pytest.mark.skipif(not ref_bow.available() and not os.path.isdir("/root/reference"),
                    reason="oracle/_ref/libbow_ref.so is built only where /root/reference is mounted")
@pytest.mark.parametrize("name", sorted(VOCABS))
def test_oracle_equals_dbow2(oracle, name):
    if not ref_bow.available():
        ref_bow.build()
    voc = make_vocab(**VOCABS[name])
    desc = features_for(voc, 11, 1500)
    for levelsup in (4, 2, 0):
        got = oracle.bow_transform(voc, desc, levelsup)
        want = oracle.bow_transform(voc, desc, levelsup, lib=ref_bow.lib(), fn="ref_bow_transform")
        same_bow(got, want)
    assert len(got["bow_word"]) > 20 and abs(got["bow_value"].sum() - 1.0) < 1e-12


def test_known_answers(oracle):
    """two-level tree by hand: ties go to the FIRST child, stopped words (weight 0) are left out, levelsup >= L gives the
    root as node id, an empty feature list gives empty vectors"""
    desc = np.zeros((7, 32), np.uint8)
    desc[1, 0] = 0b0001          # children of the root: nodes 1, 2
    desc[2, 0] = 0b0010
    desc[3, 0] = 0b0001          # children of 1: leaves 3, 4 (3 and 4 equidistant from 0b0011 -> 3 wins)
    desc[4, 0] = 0b0010
    desc[5, 0] = 0b0110          # children of 2: leaves 5, 6
    desc[6, 0] = 0b1010
    voc = dict(n_nodes=7, L=2, desc=desc, child_start=np.array([0, 2, 4, 6, 6, 6, 6, 6], np.int32),
               children=np.array([1, 2, 3, 4, 5, 6], np.int32), word_id=np.array([0, 0, 0, 0, 1, 2, 3], np.int32),
               weight=np.array([0, 0, 0, 2.0, 3.0, 0.0, 5.0]))
    f = np.zeros((4, 32), np.uint8)
    f[0, 0] = 0b0011             # root: d(1)=1, d(2)=1 -> node 1; then d(3)=1, d(4)=1 -> leaf 3 (word 0, weight 2)
    f[1, 0] = 0b0010             # root: d(1)=2, d(2)=0 -> node 2; d(5)=1, d(6)=1 -> leaf 5 (word 2, stopped)
    f[2, 0] = 0b1010             # node 2, leaf 6 (word 3, weight 5)
    f[3, 0] = 0b0011             # again word 0
    r = oracle.bow_transform(voc, f, 1)
    assert r["word"].tolist() == [0, 2, 3, 0] and r["node"].tolist() == [1, 2, 2, 1]
    assert r["bow_word"].tolist() == [0, 3] and np.allclose(r["bow_value"], [4.0 / 9.0, 5.0 / 9.0])
    assert r["fv_node"].tolist() == [1, 2] and r["fv_idx"].tolist() == [0, 3, 2] and r["fv_start"].tolist() == [0, 2, 3]
    assert oracle.bow_transform(voc, f, 4)["node"].tolist() == [0, 0, 0, 0]
    e = oracle.bow_transform(voc, f[:0], 1)
    assert len(e["bow_word"]) == 0 and len(e["fv_node"]) == 0

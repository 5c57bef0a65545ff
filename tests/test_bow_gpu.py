"""GPU parity of the BoW transform (Frame::ComputeBoW -> DBoW2 TemplatedVocabulary::transform, SURVEY.md section 8f
rank 4) against the oracle, itself pinned to DBoW2's own text by tests/test_bow_ref.py.  Word / node ids exact, idf weights
and the L1-normalised BowVector bit for bit (doubles), FeatureVector CSR identical; and the chain extraction -> transform ->
SearchByBoW on the resulting feature vectors."""
import numpy as np
import pytest

import orbb200
from bow_cases import VOCABS, features_for, make_vocab
from orbb200.synth import shifted_pair
from test_bow_ref import same_bow

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(VOCABS))
def test_transform_equals_oracle(matcher, oracle, name):
    voc = make_vocab(**VOCABS[name])
    v = matcher.vocabulary(voc)
    desc = features_for(voc, 11, 1500)
    for levelsup in (4, 2, 0):
        got = matcher.bow_transform(v, desc, levelsup)
        assert matcher.launch_count() == 1
        same_bow(got, oracle.bow_transform(voc, desc, levelsup))
    e = matcher.bow_transform(v, desc[:0], 4)
    assert len(e["bow_word"]) == 0 and len(e["fv_node"]) == 0 and e["fv_start"].tolist() == [0]
    matcher.vocabulary_destroy(v)


def test_extract_transform_search_by_bow(matcher, oracle):
    """Frame::ComputeBoW on extracted descriptors, then SearchByBoW(KF, F) gated by the feature vectors it produced"""
    torch = pytest.importorskip("torch")
    a, b = shifted_pair(8, 752, 480)
    ex = orbb200.Extractor(1500, max_width=752, max_height=480, max_batch=2)
    (ka, da), (kb, db) = ex.extract_batch(np.stack([a, b]))
    voc = make_vocab(seed=21, k=10, L=4)
    # make the tree about THESE descriptors: first-level centres are real descriptors, so the nodes gate meaningfully
    voc["desc"][1:11] = da[:: len(da) // 10][:10]
    v = matcher.vocabulary(voc)
    bounds = (0.0, 0.0, 752.0, 480.0)
    out = []
    for k, d in ((ka, da), (kb, db)):
        g = matcher.bow_transform(v, d, 2)
        same_bow(g, oracle.bow_transform(voc, d, 2))
        out.append(g)
        # the descent alone on device-resident descriptors
        dd = torch.from_numpy(d).cuda()
        w = torch.zeros(len(d), dtype=torch.int32, device="cuda")
        wt = torch.zeros(len(d), dtype=torch.float64, device="cuda")
        nd = torch.zeros(len(d), dtype=torch.int32, device="cuda")
        matcher.bow_transform_device(v, dd, len(d), 2, w, wt, nd)
        matcher.synchronize()
        assert np.array_equal(w.cpu().numpy(), g["word"]) and np.array_equal(nd.cpu().numpy(), g["node"])
        assert np.array_equal(wt.cpu().numpy().view(np.uint64), g["weight"].view(np.uint64))
    fv1 = (out[0]["fv_node"], out[0]["fv_start"], out[0]["fv_idx"])
    fv2 = (out[1]["fv_node"], out[1]["fv_start"], out[1]["fv_idx"])
    g1, g2 = matcher.frame(ka, da, bounds), matcher.frame(kb, db, bounds)
    o1, o2 = oracle.frame(ka, da, bounds), oracle.frame(kb, db, bounds)
    valid1 = np.ones(len(ka), np.uint8)
    n, m12, m21 = matcher.search_by_bow(g1, g2, fv1, fv2, valid1, None, 0.7, True, False)
    rn, rm12, rm21 = o1.search_bow(o2, fv1, fv2, valid1, None, 0.7, True, False)
    assert n == rn and n > 30 and np.array_equal(m12, rm12) and np.array_equal(m21, rm21)
    matcher.vocabulary_destroy(v)
    g1.close(); g2.close(); ex.close()


def test_vocabulary_validation(matcher):
    voc = make_vocab(seed=1, k=4, L=2)
    bad = dict(voc)
    bad["children"] = voc["children"].copy()
    bad["children"][3] = bad["children"][2]            # a node listed twice: not a tree
    with pytest.raises(orbb200.OrbError):
        matcher.vocabulary(bad)
    cyc = dict(voc)
    cyc["children"] = voc["children"].copy()
    cyc["children"][-1] = 0                            # back edge to the root
    with pytest.raises(orbb200.OrbError):
        matcher.vocabulary(cyc)
    v = matcher.vocabulary(voc)
    other = orbb200.Matcher(0)
    with pytest.raises(orbb200.OrbError):              # a vocabulary belongs to the matcher that created it
        other.bow_transform(v, np.zeros((1, 32), np.uint8), 1)
    other.close()
    matcher.vocabulary_destroy(v)

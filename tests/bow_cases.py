"""Synthetic vocabulary trees for the BoW-transform tests.  The reference's Vocabulary/ORBvoc.bin (k = 10, L = 6, about
a million words) is not in the checkout, and parity does not depend on how a tree was trained, only on its arrays: node
descriptors, child lists, word ids, idf weights."""
import numpy as np


def make_vocab(seed, k=10, L=4, ragged=False, stop_fraction=0.02):
    """Flat arrays of a DBoW2 tree built like HKmeansStep numbers it (children of a node get consecutive ids in creation
    order, depth first per level); child descriptors are noisy copies of the parent's so that descents are informative.
    ragged: some inner nodes keep fewer than k children and some branches end early (leaves above level L).
    stop_fraction of the words get weight 0 (DBoW2 'stopped' words: transform skips them)."""
    rng = np.random.default_rng(seed)
    desc = [np.zeros(32, np.uint8)]
    children = [[]]
    level = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for parent in frontier:
            if ragged and lv > 1 and rng.random() < 0.08:
                continue                                # the branch ends here: `parent` stays a leaf
            nk = k if not ragged else int(rng.integers(2, k + 1))
            base = np.unpackbits(desc[parent]) if parent else rng.integers(0, 2, 256, dtype=np.uint8)
            for _ in range(nk):
                bits = base.copy() if parent else rng.integers(0, 2, 256, dtype=np.uint8)
                flip = rng.permutation(256)[:max(4, 96 >> lv)]
                bits[flip] ^= 1
                nid = len(desc)
                desc.append(np.packbits(bits))
                children.append([])
                level.append(lv)
                children[parent].append(nid)
                nxt.append(nid)
        frontier = nxt
    n = len(desc)
    word_id = np.zeros(n, np.int32)
    weight = np.zeros(n, np.float64)
    w = 0
    for i in range(n):
        if i and not children[i]:
            word_id[i] = w
            w += 1
            weight[i] = 0.0 if rng.random() < stop_fraction else float(np.log(rng.uniform(1.5, 400.0)))
    start = np.zeros(n + 1, np.int32)
    start[1:] = np.cumsum([len(c) for c in children])
    flat = np.array([c for cs in children for c in cs], np.int32)
    return dict(n_nodes=n, L=L, k=k, desc=np.ascontiguousarray(np.stack(desc)), child_start=start, children=flat,
                word_id=word_id, weight=weight, n_words=w)


def features_for(voc, seed, n):
    """descriptors near random tree nodes (so different branches are visited) plus pure noise"""
    rng = np.random.default_rng(seed)
    pick = rng.integers(1, voc["n_nodes"], n)
    bits = np.unpackbits(voc["desc"][pick], axis=1)
    for i in range(n):
        k = int(rng.integers(0, 50))
        bits[i, rng.permutation(256)[:k]] ^= 1
    out = np.packbits(bits, axis=1)
    out[::17] = rng.integers(0, 256, (len(out[::17]), 32), dtype=np.uint8)
    return np.ascontiguousarray(out)


VOCABS = {"k10_L4": dict(seed=1, k=10, L=4), "k6_L5": dict(seed=2, k=6, L=5), "ragged_k9_L5": dict(seed=3, k=9, L=5, ragged=True),
          "k10_L3_levelsup4": dict(seed=4, k=10, L=3), "k32_L2": dict(seed=5, k=32, L=2), "k40_L2": dict(seed=6, k=40, L=2)}

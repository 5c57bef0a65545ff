"""Synthetic vocabulary trees for the BoW-transform tests.  The reference's Vocabulary/ORBvoc.bin (k = 10, L = 6, about
a million words) is not in the checkout, and parity does not depend on how a tree was trained, only on its arrays: node
descriptors, child lists, word ids, idf weights."""
import numpy as np

from orbb200.synth import features_for, make_vocab  # noqa: E402,F401

VOCABS = {"k10_L4": dict(seed=1, k=10, L=4), "k6_L5": dict(seed=2, k=6, L=5), "ragged_k9_L5": dict(seed=3, k=9, L=5, ragged=True),
          "k10_L3_levelsup4": dict(seed=4, k=10, L=3), "k32_L2": dict(seed=5, k=32, L=2), "k40_L2": dict(seed=6, k=40, L=2)}

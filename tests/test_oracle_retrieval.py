"""Retrieval-metric restatements of the oracle against the reference's own numpy code (utils/mAP.py, utils/nDCG.py) and
its one known-answer test (utils/nDCG.py:154-181)."""
import importlib
import sys

import numpy as np
import pytest

from oracle import hh_oracle as O
from oracle import ref_import

from oracle.golden_cases import KNOWN_K, KNOWN_NDCG, KNOWN_REL, KNOWN_SIM, synth_retrieval  # noqa: E402


def test_known_answer():
    assert (O.calculate_k_counts(KNOWN_REL) == KNOWN_K).all()
    assert O.calculate_nDCG(KNOWN_SIM, KNOWN_REL, KNOWN_K)[0] == KNOWN_NDCG


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree only exists in the build container")
def test_against_live_reference():
    if ref_import.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_import.REFERENCE_ROOT)
    ref_map = importlib.import_module("utils.mAP")
    ref_ndcg = importlib.import_module("utils.nDCG")
    for N, M, seed in [(5, 7, 0), (40, 300, 1), (16, 2000, 2)]:
        sim, rel = synth_retrieval(N, M, seed)
        assert O.calculate_mAP(sim, rel)[0] == ref_map.calculate_mAP(sim, rel)
        assert O.calculate_nDCG(sim, rel)[0] == ref_ndcg.calculate_nDCG(sim, rel)
        assert O.calculate_mAP(sim.T.copy(), rel.T.copy())[0] == ref_map.calculate_mAP(sim.T, rel.T)

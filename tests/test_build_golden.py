"""Builder edge cases against fingerprints of the REFERENCE's own trees (tests/golden/build_golden.json, made by
tests/golden/make_build_golden.py from oracle/_ref): the host builder, the oracle's builder and the numpy model of the
level-synchronous device algorithm.  The CUDA builder is held against the same fingerprints in tests/test_gpu_build.py."""
import pytest

import mallie_b200 as M
from oracle import orabind as O
from tests import common as T
from tests import lsbuild_model as L

CASES = sorted(T.build_cases())


@pytest.mark.parametrize("name", CASES)
def test_host_and_oracle_builders_match_the_reference_tree(name):
    v, f = T.build_cases()[name]
    g = T.build_golden()[name]
    want = {k: g[k] for k in ("num_nodes", "nodes_fnv", "indices_fnv")}
    hb = M.HostBVH.build(v, f)
    assert T.tree_fingerprint(*hb.arrays()) == want and hb.stats() == g["stats"]
    ob = O.BVH.build(O.Mesh(v, f))
    assert T.tree_fingerprint(*ob.arrays()) == want and ob.stats() == g["stats"]


@pytest.mark.parametrize("name", CASES)
def test_level_synchronous_model_matches_the_reference_tree(name):
    v, f = T.build_cases()[name]
    g = T.build_golden()[name]
    assert T.tree_fingerprint(*L.build(v, f)) == {k: g[k] for k in ("num_nodes", "nodes_fnv", "indices_fnv")}

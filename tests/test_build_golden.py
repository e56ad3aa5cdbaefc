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


OPTION_CASES = T.build_option_cases()


@pytest.mark.parametrize("name,opt", OPTION_CASES, ids=[T.build_option_key(n, o) for n, o in OPTION_CASES])
def test_host_and_oracle_builders_match_the_reference_tree_under_options(name, opt):
    """Non-default BVHBuildOptions (bvh_accel.h:32-42), incl. minLeafPrimitives = 1: the reference then splits every
    single-triangle range by the object-median fallback into an EMPTY left leaf -- whose box ComputeBoundingBox seeds
    from the first vertex of the triangle at indices[leftIndex] (bvh_accel.cc:291-298) -- and the triangle again, down to
    maxTreeDepth."""
    v, f = T.build_cases()[name]
    g = T.build_golden()[T.build_option_key(name, opt)]
    want = {k: g[k] for k in ("num_nodes", "nodes_fnv", "indices_fnv")}
    hb = M.HostBVH.build(v, f, **opt)
    assert T.tree_fingerprint(*hb.arrays()) == want and hb.stats() == g["stats"]
    ob = O.BVH.build(O.Mesh(v, f), **opt)
    assert T.tree_fingerprint(*ob.arrays()) == want and ob.stats() == g["stats"]


def test_host_builder_rejects_options_the_reference_cannot_finish():
    v, f = T.build_cases()["soup_17"]
    for opt in (dict(min_leaf=0), dict(min_leaf=-3), dict(max_depth=-1), dict(bin_size=1)):
        with pytest.raises(M.MallieB200Error):
            M.HostBVH.build(v, f, **opt)

"""Peel decomposition (host/src/decomposition/peel_decomposition.cpp): the reference's S-tree
split rules applied to degree-2 separation pairs.  CPU only - no numerics here."""
import numpy as np
import pytest

import host_lib as H
import sketch_gen as S


@pytest.fixture(scope="module")
def host(built):
    built.build_host()
    return H.load()


def test_linkage_decomposes_into_n_minus_2_leaves_with_every_edge_once(host):
    el, edges = S.make_linkage(2000, seed=3)
    assert len(edges) == 2 * len(el) - 3
    n, leaves, nvirt, nreal = H.decompose(el, edges)
    assert n == len(el) - 2
    # the base triangle carries three real edges and no virtual one; every other leaf two + one
    assert (nvirt[0], nreal[0]) == (0, 3)
    assert set(zip(nvirt[1:], nreal[1:])) == {(1, 2)}
    # every real edge of the sketch lands in exactly one leaf
    assert sum(nreal) == len(edges)
    # each leaf after the base introduces exactly one new element (solve order is valid)
    seen = set(leaves[0])
    for lf in leaves[1:]:
        new = [i for i in lf if i not in seen]
        assert len(new) == 1
        seen.add(new[0])
    assert seen == set(range(len(el)))


def test_decomposed_leaves_plan_like_the_generated_leaf_list(host):
    """Mixed point/line sketches: decomposing the whole-sketch graph and planning the leaves must
    find a solver for every leaf, as the generator's own leaf list does."""
    for seed, first in ((21, 1), (22, 2), (23, 3)):
        el, lv = S.make_sketch(400, seed=seed, first_shape=first)
        edges = S.sketch_graph(el, lv)
        n, leaves, nvirt, nreal = H.decompose(el, edges)
        assert n == len(el) - 2
        ref = H.leaves_solve(el, lv, mode=2)
        assert ref["solved"] == len(lv)
        # rebuild a leaf list from the decomposition and plan it
        by_pair = {}
        for e in edges:
            by_pair[(min(e["a"], e["b"]), max(e["a"], e["b"]))] = e
        placed = set(leaves[0])
        mine = []
        for k, lf in enumerate(leaves):
            es = []
            pairs = [(lf[0], lf[1]), (lf[0], lf[2]), (lf[1], lf[2])]
            new = None if k == 0 else [i for i in lf if i not in placed][0]
            for a, b in pairs:
                e = by_pair.get((min(a, b), max(a, b)))
                if k == 0:
                    if e:
                        es.append(e)
                elif new in (a, b):
                    es.append(e)
                else:
                    es.append({"a": a, "b": b, "type": 2})
            if new is not None:
                placed.add(new)
            mine.append({"elems": list(lf), "edges": es})
        got = H.leaves_solve(el, mine, mode=2)
        assert got["solved"] == len(mine), (seed, got["solved"], len(mine))


def test_graphs_without_degree_two_elements_are_refused(host):
    # K4 on four points (6 distances): over-constrained AND without a degree-2 node
    P = lambda x, y: dict(type=0, canvas=[x, y])
    D = lambda a, b, v: dict(a=a, b=b, type=0, value=v)
    el = [P(0, 0), P(1, 0), P(0, 1), P(1, 1)]
    edges = [D(0, 1, 1), D(0, 2, 1), D(0, 3, 1.4), D(1, 2, 1.4), D(1, 3, 1), D(2, 3, 1)]
    n, *_ = H.decompose(el, edges)
    assert n == -1 and "degree-2" in H.last_error()

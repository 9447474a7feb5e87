"""Host mirror, CPU side: classification, role assignment, anchoring and canvas-side signs of the
packer (host/src/b200/leaf_batch.cpp) against the golden components produced by the reference's
own classifyAndSolve (tests/golden/components.json).  The packed row is finished by the CPU
oracle here (this is a test); the GPU twin of this file, test_gpu_host.py, sends the same
components through the CUDA path."""
import json
import os

import numpy as np
import pytest

import host_lib as H
import oracle_lib as O
from util import bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHAPE_TO_KIND = {1: 1, 2: 2, 3: 5, 4: 1, 5: 2, 6: 3, 7: 4, 8: 5}


@pytest.fixture(scope="module")
def items():
    return json.load(open(os.path.join(GOLD, "components.json")))["items"]


@pytest.fixture(scope="module")
def host(built):
    built.build_host()
    return H.load()


def finish_with_oracle(capi, kind, row, code):
    hb = capi.HostBatch(kind, 2, [np.array([row[c]]) for c in range(capi.IN_COLS[kind])], np.array([code], np.uint8))
    O.solve(hb.alloc_outputs())
    return [float(o[0]) for o in hb.out]


def same_pos(a, b):
    return all(np.float64(x).view(np.uint64) == np.float64(y).view(np.uint64) or (x != x and y != y) for x, y in zip(a, b))


def test_packer_plus_oracle_reproduces_the_reference_components(gcs, host, items):
    seen = {}
    for it in items:
        sid, kind, row, code, target, els = H.component_pack(it["elements"], it["edges"])
        if it["status"] == 1:  # reference: Unsupported
            assert sid == 0
            continue
        if it["status"] == -1:  # the reference threw (malformed leaf): so must the packer
            assert sid == -1, (it["shape"], sid)
            continue
        assert sid == it["shape"], (sid, it["shape"])
        assert kind == SHAPE_TO_KIND[sid]
        out = finish_with_oracle(gcs.capi, kind, row, code)
        els[target]["pos"] = out
        els[target]["is_set"] = True
        for got, exp in zip(els, it["expected"]):
            assert got["is_set"] == exp["is_set"]
            if exp["is_set"]:
                assert same_pos(got["pos"], exp["pos"]), (it["shape"], got, exp)
        seen[sid] = seen.get(sid, 0) + 1
    assert all(seen.get(s, 0) >= 50 for s in range(1, 9)), seen


def test_dispatch_order_and_predicates(host):
    P = lambda x, y, **k: dict(type=0, canvas=[x, y], **k)
    L = lambda a, b, c, d, **k: dict(type=1, canvas=[a, b, c, d], **k)
    D = lambda a, b, v: dict(a=a, b=b, type=0, value=v)
    A = lambda a, b, v, flip=False: dict(a=a, b=b, type=1, value=v, flip=flip)
    V = lambda a, b: dict(a=a, b=b, type=2)
    tri = [P(0, 0), P(4, 0), P(0, 3)]
    assert H.component_pack(tri, [D(0, 1, 4), D(0, 2, 3), D(1, 2, 5)])[0] == 1
    # 3 nodes / 3 edges one of which is virtual: matches() still says ZeroFixedPoints
    # (edgeCount counts it, getConstraints does not) and solve() throws on the missing value
    assert H.component_pack(tri, [D(0, 1, 4), D(0, 2, 3), V(1, 2)])[0] == -1
    # two edges only: not a zero-fixed triangle, nothing solved -> unsupported
    assert H.component_pack(tri, [D(0, 1, 4), D(0, 2, 3)])[0] == 0
    # an angle between points is not a distance -> unsupported
    assert H.component_pack(tri, [D(0, 1, 4), D(0, 2, 3), A(1, 2, 0.3)])[0] == 0
    solved = [P(0, 0, is_set=True, pos=[0, 0]), P(4, 0, is_set=True, pos=[4, 0]), P(0, 3)]
    assert H.component_pack(solved, [D(0, 2, 3), D(1, 2, 5), V(0, 1)])[0] == 4
    assert H.component_pack(solved, [D(0, 2, 3), D(1, 2, 5)])[0] == 4          # no edge count condition
    # all three solved: the reference dereferences a null free point; here a reported error
    allset = [dict(e, is_set=True, pos=e["canvas"]) for e in tri]
    assert H.component_pack(allset, [D(0, 1, 4), D(0, 2, 3), D(1, 2, 5)])[0] == -1
    assert "null free point" in H.last_error()
    ppl = [P(0, 0), L(0, 5, 9, 5), P(4, 0)]
    assert H.component_pack(ppl, [D(0, 2, 4), D(0, 1, 5), D(2, 1, 5)])[0] == 2
    llp = [L(0, 0, 10, 0), P(3, 3), L(0, 0, 7, 7)]
    assert H.component_pack(llp, [A(0, 2, 0.7), D(1, 0, 3), D(1, 2, 2)])[0] == 3
    assert H.component_pack(llp, [D(0, 2, 0.7), D(1, 0, 3), D(1, 2, 2)])[0] == 0   # no angle
    three_lines = [L(0, 0, 1, 0), L(0, 0, 0, 1), L(1, 0, 0, 1)]
    assert H.component_pack(three_lines, [A(0, 1, 1), A(1, 2, 1), A(0, 2, 1)])[0] == 0


def test_role_assignment_follows_node_order(gcs, host):
    """fixed1 / fixed2 are the solved points in ascending node order (point_point_solvers.cpp:110-123):
    swapping the two fixed elements swaps (a, ra) and (b, rb) in the packed row."""
    P = lambda x, y, **k: dict(type=0, canvas=[x, y], **k)
    D = lambda a, b, v: dict(a=a, b=b, type=0, value=v)
    els = [P(10, 10, is_set=True, pos=[1, 2]), P(30, 10), P(20, 30, is_set=True, pos=[7, 8])]
    sid, kind, row, code, target, _ = H.component_pack(els, [D(0, 1, 3.5), D(2, 1, 4.5)])
    assert (sid, kind, target) == (4, 1, 1)
    assert row[:6].tolist() == [1, 2, 3.5, 7, 8, 4.5]
    # canvas triangle (10,10),(20,30),(30,10): clockwise -> sign -1 -> code 0
    assert (code & 3) - 1 == -1
    els2 = [els[2], els[1], els[0]]
    sid, kind, row, code, target, _ = H.component_pack(els2, [D(2, 1, 3.5), D(0, 1, 4.5)])
    assert row[:6].tolist() == [7, 8, 4.5, 1, 2, 3.5] and (code & 3) - 1 == 1


def test_plan_levels_follow_data_dependences(host):
    """A fan: triangle (0,1,2) anchors, then points 3..6 each hang on two solved points; point 7
    hangs on 5 and 6.  Waves: 0 | 3,4,5,6 at 1 | 7 at 2.  No device needed for the plan."""
    P = lambda x, y: dict(type=0, canvas=[x, y])
    D = lambda a, b, v: dict(a=a, b=b, type=0, value=v)
    V = lambda a, b: dict(a=a, b=b, type=2)
    els = [P(0, 0), P(10, 0), P(5, 8), P(15, 8), P(-5, 8), P(5, -8), P(20, 0), P(12, -9)]
    leaves = [
        {"elems": [0, 1, 2], "edges": [D(0, 1, 10), D(0, 2, 9), D(1, 2, 9)]},
        {"elems": [1, 2, 3], "edges": [D(1, 3, 9), D(2, 3, 10), V(1, 2)]},
        {"elems": [0, 2, 4], "edges": [D(0, 4, 9), D(2, 4, 10), V(0, 2)]},
        {"elems": [0, 1, 5], "edges": [D(0, 5, 9), D(1, 5, 9), V(0, 1)]},
        {"elems": [1, 3, 6], "edges": [D(1, 6, 10), D(3, 6, 9)]},
        {"elems": [5, 6, 7], "edges": [D(5, 7, 7), D(6, 7, 12), V(5, 6)]},
    ]
    r = H.leaves_solve(els, leaves, mode=2)
    assert r["rc"] == 0
    assert r["solver"] == [1, 4, 4, 4, 4, 4]
    assert r["level"] == [0, 1, 1, 1, 2, 3]
    assert r["waves"] == 4 and r["solved"] == 6
    # a leaf whose two "fixed" points are not solved yet when its turn comes is unsupported,
    # exactly as in the sequential loop - and so is everything that depended on it
    early = {"elems": [5, 6, 7], "edges": leaves[5]["edges"][:2]}
    r = H.leaves_solve(els, [early] + leaves[:5], mode=2)
    assert r["solver"][0] == 0 and r["status"][0] == 1 and r["level"][0] == -1
    assert r["solver"][1:] == [1, 4, 4, 4, 4]
    # with its virtual edge the early leaf LOOKS like a zero-fixed triangle (3 nodes, 3 edges,
    # nothing solved, every stored constraint a distance) and the reference's solve() then throws
    # on the edge without a value: the loop stops there, nothing after it is solved
    r = H.leaves_solve(els, [leaves[5]] + leaves[:5], mode=2)
    assert r["solved"] == 0 and r["level"] == [-1] * 6


def test_the_digest_the_plan_reads_follows_every_change_to_a_leaf(host):
    """ConstraintGraph::triangleDigest - the summary the scheduler plans from instead of walking a
    leaf's containers - tracks edges added behind the class's back (getGraph()), constraints, virtual
    edges, removals and copies; Element serial numbers are per object."""
    assert H.load().gcs_host_selftest_digest() == 0, H.last_error()


def test_plan_does_not_depend_on_how_the_elements_are_numbered(host, monkeypatch):
    """The sweep indexes its per-element tables by Element::serial(): a subtraction when the numbers
    are close together, a hash table when they are not.  Same sketches, same plans."""
    import sketch_gen as S
    data = json.load(open(os.path.join(GOLD, "sketch_leaves.json")))["items"]
    cases = [(sk["elements"], sk["leaves"]) for sk in data]
    el, lv = S.make_sketch(3000, seed=5, first_shape=2, p_line=0.3)
    cases.append((el, lv))
    for els, leaves in cases:
        monkeypatch.delenv("GCS_HOST_SERIAL_STRIDE", raising=False)
        dense = H.leaves_solve([dict(e) for e in els], leaves, mode=2)
        monkeypatch.setenv("GCS_HOST_SERIAL_STRIDE", "1000003")
        sparse = H.leaves_solve([dict(e) for e in els], leaves, mode=2)
        assert dense["rc"] == sparse["rc"] == 0
        for key in ("solver", "level", "status", "waves", "solved"):
            assert dense[key] == sparse[key], key


def test_leaves_the_digest_cannot_describe_take_the_container_walk(host):
    """Two edges on one node pair: not a `simple` triangle, so the general classification (the
    reference's own counts over the containers) decides - next to ordinary leaves in one plan.
    The two-fixed predicates do not look at the edge count (point_point_solvers.cpp:87-95), so such a
    leaf is solved like any other; a leaf whose three points are all solved by the time its turn
    comes makes the reference dereference a null free point (:110-127): the loop stops there.
    (Expected values: the plan of the round-1 build, which walked the containers of every leaf.)"""
    P = lambda x, y: dict(type=0, canvas=[x, y])
    D = lambda a, b, v: dict(a=a, b=b, type=0, value=v)
    V = lambda a, b: dict(a=a, b=b, type=2)
    els = [P(0, 0), P(10, 0), P(5, 8), P(15, 8), P(-5, 8), P(3, 3)]
    leaves = [
        {"elems": [0, 1, 2], "edges": [D(0, 1, 10), D(0, 2, 9), D(1, 2, 9)]},
        {"elems": [1, 2, 3], "edges": [D(1, 3, 9), D(2, 3, 10), D(2, 3, 10), V(1, 2)]},  # 4 edges
        {"elems": [0, 2, 4], "edges": [D(0, 4, 9), D(2, 4, 10), V(0, 2)]},
        {"elems": [3, 4, 5], "edges": [D(3, 5, 9), D(3, 5, 9), D(4, 5, 10)]},            # a double edge, reads waves 1
        {"elems": [1, 2, 3], "edges": [D(1, 3, 9), D(1, 3, 9), D(2, 3, 10)]},            # all three solved by now
        {"elems": [0, 1, 5], "edges": [D(0, 5, 9), D(1, 5, 10)]},
    ]
    plan = H.leaves_solve(els, leaves, mode=2)
    assert plan["rc"] == 0
    assert plan["solver"] == [1, 4, 4, 4, 0, 0]
    assert plan["level"] == [0, 1, 1, 2, -1, -1]
    assert plan["status"] == [0, 0, 0, 0, 1, 1]
    assert plan["waves"] == 3 and plan["solved"] == 4


def test_wave_plan_reproduces_the_reference_loop_on_golden_sketches(gcs, host):
    """tests/golden/sketch_leaves.json: 4 sketches x (120..300) leaves solved by the reference's
    sequential loop.  Here: plan on the host (solver + wave per leaf), then replay the waves -
    every leaf of a wave packed from the element state left by the earlier waves, finished by the
    oracle - and the final element state must carry the reference's bits."""
    data = json.load(open(os.path.join(GOLD, "sketch_leaves.json")))["items"]
    for sk in data:
        els = [dict(e) for e in sk["elements"]]
        plan = H.leaves_solve(els, sk["leaves"], mode=2)
        assert plan["rc"] == 0 and plan["solved"] == len(sk["leaves"])
        assert all(s != 0 for s in plan["solver"])
        assert plan["waves"] < len(sk["leaves"])  # there is parallelism to exploit
        for w in range(plan["waves"]):
            updates = []
            for li, lf in enumerate(sk["leaves"]):
                if plan["level"][li] != w:
                    continue
                local = [els[i] for i in lf["elems"]]
                remap = {g: k for k, g in enumerate(lf["elems"])}
                edges = [dict(e, a=remap[e["a"]], b=remap[e["b"]]) for e in lf["edges"]]
                sid, kind, row, code, target, out = H.component_pack(local, edges)
                assert sid == plan["solver"][li]
                out[target]["pos"] = finish_with_oracle(gcs.capi, kind, row, code)
                out[target]["is_set"] = True
                updates.append((lf["elems"], out))
            for ids, out in updates:  # a wave's writes land after all of its reads
                for g, o in zip(ids, out):
                    if o["is_set"] and not els[g].get("is_set"):
                        els[g] = dict(els[g], is_set=True, pos=o["pos"])
        for got, exp in zip(els, sk["expected"]):
            assert got.get("is_set") == exp["is_set"]
            assert same_pos(got["pos"], exp["pos"])


def test_kernel_variant_setting_of_the_host_mirror(host):
    """Gcs::B200::setKernelVariant: the default is the bit-identical class; the setter returns the
    previous value (what KindBatch::descriptor() puts into gcs_b200_batch.variant)."""
    assert host.gcs_host_set_variant(5) == 0
    assert host.gcs_host_set_variant(0) == 5
    assert host.gcs_host_set_variant(0) == 0

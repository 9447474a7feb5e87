"""the data format on one side of the path and the step on the other.

* JSON v1 sketch files (reference: gui/src/model_serializer.cpp:19-217) read / written by
  sketch_io, with the model's acceptance rules and degree -> radian conversion
  (gui/src/constraint_model.cpp:76-84, :116-121, :133) checked against the reference's own
  ConstraintModel (tests/golden/model_transform.npz, oracle/make_golden_model.py).
* The solver -> canvas rigid motion (reference: constraint_model.cpp:394-501) of the host mirror
  against the reference's applySolverToCanvasTransform, bit for bit.
* GPU: a JSON sketch through the whole pipeline - load, decompose, batched solve, transform, save."""
import importlib
import json
import math
import os

import numpy as np
import pytest

import host_lib as H
from util import bits

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
sio = gcs.sketch_io


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "model_transform.npz"))


@pytest.fixture(scope="module")
def host(built):
    built.build_host()
    return H.load()


def test_canvas_transform_matches_the_reference_model(host, gold):
    n_case = gold["types"].shape[0]
    for k in range(n_case):
        t, c, p, s, exp = gold["types"][k], gold["canvas"][k], gold["pos"][k], gold["solved"][k], gold["out"][k]
        els = []
        for i in range(len(t)):
            w = 2 if t[i] == 0 else 4
            els.append({"type": int(t[i]), "canvas": list(c[i, :w]), "pos": list(p[i, :w]), "is_set": bool(s[i])})
        rc, got = H.canvas_transform(els)
        assert rc == 0, H.last_error()
        for i, g in enumerate(got):
            assert np.array_equal(bits(np.array(g)), bits(exp[i, :len(g)])), (k, i, g, exp[i])


def test_canvas_transform_restores_a_rigidly_moved_sketch(host):
    rng = np.random.default_rng(3)
    canvas = rng.uniform(0, 500, (6, 2))
    th = 1.1
    rot = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
    solver = (canvas - [100.0, 50.0]) @ rot.T
    els = [{"type": 0, "canvas": list(canvas[i]), "pos": list(solver[i]), "is_set": True} for i in range(6)]
    rc, got = H.canvas_transform(els)
    assert rc == 0 and np.allclose(np.array(got), canvas, atol=1e-9)


def test_constraint_acceptance_and_degree_conversion_match_the_reference_model(gold):
    types, canvas = gold["con_types"], gold["con_canvas"]
    sketch = {"elements": [], "constraints": [], "view": {"panX": 0.0, "panY": 0.0, "zoom": 1.0}}
    for t, c in zip(types, canvas):
        sketch["elements"].append({"type": "point", "x": c[0], "y": c[1]} if t == 0 else
                                  {"type": "line", "x1": c[0], "y1": c[1], "x2": c[2], "y2": c[3]})
    for (a, b, ty, flip), v in zip(gold["con"], gold["con_value"]):
        rec = {"type": "distance" if ty == 0 else "angle", "elementA": int(a), "elementB": int(b), "value": float(v)}
        if ty == 1:
            rec["flipped"] = bool(flip)
        sketch["constraints"].append(rec)
    sketch = sio.loads(sio.dumps(sketch))          # through the file format and back
    _, edges, rejected = sio.to_solver_input(sketch)
    stored = gold["con_stored"]
    assert len(edges) == int(gold["con_accepted"])
    assert rejected == [k for k in range(len(stored)) if np.isnan(stored[k])]
    kept = [k for k in range(len(stored)) if not np.isnan(stored[k])]
    for e, k in zip(edges, kept):
        assert np.float64(e["value"]).view(np.uint64) == np.float64(stored[k]).view(np.uint64), (k, e["value"], stored[k])


def test_json_v1_round_trip_and_errors(tmp_path):
    text = json.dumps({
        "version": 1,
        "elements": [{"type": "point", "x": 100, "y": 100}, {"type": "point", "x": 200.5, "y": 100},
                     {"type": "line", "x1": 0, "y1": 0, "x2": 50, "y2": 75}],
        "constraints": [{"type": "distance", "elementA": 0, "elementB": 1, "value": 3.0},
                        {"type": "angle", "elementA": 2, "elementB": 2, "value": 45.0}],   # "flipped" defaults to false
    })
    s = sio.loads(text)
    assert s["view"] == {"panX": 0.0, "panY": 0.0, "zoom": 1.0}
    assert s["constraints"][1]["flipped"] is False
    path = tmp_path / "sketch.json"
    sio.save(str(path), s)
    assert sio.load(str(path)) == s
    again = json.loads(path.read_text())
    assert again["version"] == 1 and list(again) == ["version", "elements", "constraints", "view"]
    for bad, msg in [
        ('{"elements": []}', "Missing 'version' field"),
        ('{"version": 2, "elements": []}', "Unsupported file version 2 (expected 1)"),
        ('{"version": 1}', "Missing or invalid 'elements' array"),
        ('{"version": 1, "elements": [{"type": "circle"}]}', "Unknown element type: 'circle'"),
        ('{"version": 1, "elements": [{"type": "point", "x": 1, "y": 2}], "constraints": [{"type": "distance", "elementA": 0, "elementB": 5, "value": 1}]}',
         "Constraint references invalid element index (0 or 5; 1 elements exist)"),
        ('{"version": 1, "elements": [], "constraints": [{"type": "tangent", "elementA": 0, "elementB": 0, "value": 1}]}',
         "Unknown constraint type: 'tangent'"),
        ("{not json", "JSON parse error"),
    ]:
        with pytest.raises(sio.SketchFormatError) as ei:
            sio.loads(bad)
        assert msg in str(ei.value)


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_json_sketch_through_the_whole_pipeline(gpu, host, tmp_path):
    """BASELINE config 1 as a file: 3 points, distances 3-4-5 (and a 5-point fan), loaded from JSON
    v1, decomposed, solved on the GPU, moved back over the drawing, saved."""
    sketch = {
        "elements": [{"type": "point", "x": 100.0, "y": 100.0}, {"type": "point", "x": 200.0, "y": 100.0},
                     {"type": "point", "x": 150.0, "y": 200.0}, {"type": "point", "x": 260.0, "y": 210.0},
                     {"type": "point", "x": 300.0, "y": 90.0}],
        "constraints": [{"type": "distance", "elementA": 0, "elementB": 1, "value": 4.0},
                        {"type": "distance", "elementA": 0, "elementB": 2, "value": 3.0},
                        {"type": "distance", "elementA": 1, "elementB": 2, "value": 5.0},
                        {"type": "distance", "elementA": 1, "elementB": 3, "value": 6.0},
                        {"type": "distance", "elementA": 2, "elementB": 3, "value": 4.5},
                        {"type": "distance", "elementA": 1, "elementB": 4, "value": 5.5},
                        {"type": "distance", "elementA": 3, "elementB": 4, "value": 7.0}],
        "view": {"panX": 0.0, "panY": 0.0, "zoom": 1.0},
    }
    path = tmp_path / "in.json"
    sio.save(str(path), sketch)
    loaded = sio.load(str(path))
    elements, edges, rejected = sio.to_solver_input(loaded)
    assert rejected == []
    rc, solved, stats = H.system_solve_ex(elements, edges)
    assert rc == 0, H.last_error()
    assert stats["leaves"] == 3 and stats["solved"] == 3
    pos = np.array([e["pos"] for e in solved])
    for c in loaded["constraints"]:
        d = np.linalg.norm(pos[c["elementA"]] - pos[c["elementB"]])
        assert abs(d - c["value"]) < 1e-9 * max(1.0, c["value"])          # the north star's coordinate tolerance
    # orientation of every triangle as drawn (the root-selection heuristic)
    cv = np.array([e["canvas"] for e in elements])
    def ori(q, a, b, c):
        u, v = q[b] - q[a], q[c] - q[a]
        return np.sign(u[0] * v[1] - u[1] * v[0])
    for tri in ((0, 1, 2), (1, 2, 3), (1, 3, 4)):
        assert ori(pos, *tri) == ori(cv, *tri)
    for e, s in zip(elements, solved):
        e.update(pos=s["pos"], is_set=s["is_set"])
    rc, canvas = H.canvas_transform(elements)
    assert rc == 0
    out = sio.with_canvas(loaded, canvas)
    sio.save(str(tmp_path / "out.json"), out)
    back = sio.load(str(tmp_path / "out.json"))
    q = np.array([[e["x"], e["y"]] for e in back["elements"]])
    for c in back["constraints"]:
        assert abs(np.linalg.norm(q[c["elementA"]] - q[c["elementB"]]) - c["value"]) < 1e-9 * max(1.0, c["value"])
    # a rigid motion: the drawn sketch (size ~100) cannot be matched by the solved one (size ~5), but
    # the centroids coincide and the orientation is kept
    assert np.allclose(q.mean(axis=0), cv.mean(axis=0), atol=1e-9)
    assert ori(q, 0, 1, 2) == ori(cv, 0, 1, 2)

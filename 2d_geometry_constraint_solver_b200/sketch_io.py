"""The reference GUI's on-disk sketch format, JSON v1 (reference: gui/src/model_serializer.cpp:19-217),
so that sketches saved by the reference can be fed to this path and results handed back.

    {"version": 1,
     "elements":    [{"type": "point", "x": .., "y": ..} | {"type": "line", "x1": .., "y1": .., "x2": .., "y2": ..}],
     "constraints": [{"type": "distance" | "angle", "elementA": i, "elementB": j, "value": v, "flipped": bool (angle only)}],
     "view": {"panX": .., "panY": .., "zoom": ..}}

Element indices are positions in the `elements` array.  Angles are stored in DEGREES in the file
and converted to radians when the constraint enters the model (gui/src/constraint_model.cpp:133:
angleDegrees * pi / 180.0); `to_solver_input` performs that conversion with the same expression.
Validation mirrors the reference's deserialiser (same conditions, same messages) and the model's
acceptance rules (constraint_model.cpp:76-84: no distance between two lines; :116-121: angles only
between two lines)."""
from __future__ import annotations

import json
import math

FORMAT_VERSION = 1


class SketchFormatError(ValueError):
    pass


def loads(text: str) -> dict:
    """Parse JSON v1 into {"elements": [...], "constraints": [...], "view": {...}} (file units)."""
    try:
        root = json.loads(text)
    except json.JSONDecodeError as ex:
        raise SketchFormatError(f"JSON parse error: {ex}") from ex
    if not isinstance(root, dict) or "version" not in root:
        raise SketchFormatError("Missing 'version' field")
    if root["version"] != FORMAT_VERSION:
        raise SketchFormatError(f"Unsupported file version {root['version']} (expected {FORMAT_VERSION})")
    if not isinstance(root.get("elements"), list):
        raise SketchFormatError("Missing or invalid 'elements' array")
    elements = []
    for e in root["elements"]:
        try:
            t = e["type"]
            if t == "point":
                elements.append({"type": "point", "x": float(e["x"]), "y": float(e["y"])})
            elif t == "line":
                elements.append({"type": "line", "x1": float(e["x1"]), "y1": float(e["y1"]),
                                 "x2": float(e["x2"]), "y2": float(e["y2"])})
            else:
                raise SketchFormatError(f"Unknown element type: '{t}'")
        except (KeyError, TypeError) as ex:
            raise SketchFormatError(f"JSON parse error: {ex}") from ex
    constraints = []
    if isinstance(root.get("constraints"), list):
        for c in root["constraints"]:
            try:
                t = c["type"]
                if t not in ("distance", "angle"):
                    raise SketchFormatError(f"Unknown constraint type: '{t}'")
                a, b = int(c["elementA"]), int(c["elementB"])
                rec = {"type": t, "elementA": a, "elementB": b, "value": float(c["value"])}
                if t == "angle":
                    rec["flipped"] = bool(c.get("flipped", False))
            except (KeyError, TypeError) as ex:
                raise SketchFormatError(f"JSON parse error: {ex}") from ex
            if not (0 <= a < len(elements) and 0 <= b < len(elements)):
                raise SketchFormatError(
                    f"Constraint references invalid element index ({a} or {b}; {len(elements)} elements exist)")
            constraints.append(rec)
    view = {"panX": 0.0, "panY": 0.0, "zoom": 1.0}
    if isinstance(root.get("view"), dict):
        for k in view:
            view[k] = float(root["view"].get(k, view[k]))
    return {"elements": elements, "constraints": constraints, "view": view}


def dumps(sketch: dict) -> str:
    """Serialise back to JSON v1 (2-space indent, like the reference's dump(2))."""
    els = []
    for e in sketch["elements"]:
        els.append({"type": "point", "x": e["x"], "y": e["y"]} if e["type"] == "point" else
                   {"type": "line", "x1": e["x1"], "y1": e["y1"], "x2": e["x2"], "y2": e["y2"]})
    cons = []
    for c in sketch["constraints"]:
        rec = {"type": c["type"]}
        if c["type"] == "angle":
            rec["flipped"] = bool(c.get("flipped", False))
        rec.update(elementA=c["elementA"], elementB=c["elementB"], value=c["value"])
        cons.append(rec)
    view = sketch.get("view", {"panX": 0.0, "panY": 0.0, "zoom": 1.0})
    return json.dumps({"version": FORMAT_VERSION, "elements": els, "constraints": cons, "view": view}, indent=2)


def load(path: str) -> dict:
    with open(path, "r", encoding="utf-8") as f:
        return loads(f.read())


def save(path: str, sketch: dict) -> None:
    with open(path, "w", encoding="utf-8") as f:
        f.write(dumps(sketch))


def to_solver_input(sketch: dict):
    """(elements, edges) in the record layout of the host mirror's C entry points (tests/host_lib.py,
    host/src/capi_host.cpp): element {type 0 point / 1 line, canvas}, edge {a, b, type 0 distance /
    1 angle, value (radians for angles), flip}.  Constraints the reference model rejects are dropped,
    as ConstraintModel::add*Constraint would drop them; their indices are returned third."""
    elements = []
    for e in sketch["elements"]:
        if e["type"] == "point":
            elements.append({"type": 0, "canvas": [e["x"], e["y"]]})
        else:
            elements.append({"type": 1, "canvas": [e["x1"], e["y1"], e["x2"], e["y2"]]})
    edges, rejected = [], []
    for k, c in enumerate(sketch["constraints"]):
        a, b = c["elementA"], c["elementB"]
        both_lines = elements[a]["type"] == 1 and elements[b]["type"] == 1
        if c["type"] == "distance":
            if both_lines:
                rejected.append(k)
                continue
            edges.append({"a": a, "b": b, "type": 0, "value": c["value"]})
        else:
            if not both_lines:
                rejected.append(k)
                continue
            edges.append({"a": a, "b": b, "type": 1, "value": c["value"] * math.pi / 180.0, "flip": bool(c.get("flipped", False))})
    return elements, edges, rejected


def with_canvas(sketch: dict, canvas) -> dict:
    """A copy of `sketch` whose element coordinates are replaced by `canvas` (per element x,y or
    x1,y1,x2,y2) - what the GUI shows after solveConstraintSystem()."""
    out = {"elements": [], "constraints": [dict(c) for c in sketch["constraints"]], "view": dict(sketch.get("view", {}))}
    for e, c in zip(sketch["elements"], canvas):
        if e["type"] == "point":
            out["elements"].append({"type": "point", "x": float(c[0]), "y": float(c[1])})
        else:
            out["elements"].append({"type": "line", "x1": float(c[0]), "y1": float(c[1]), "x2": float(c[2]), "y2": float(c[3])})
    return out

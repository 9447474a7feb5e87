#!/usr/bin/env python
"""Golden vectors for the step after the solve: the reference GUI
model's applySolverToCanvasTransform (gui/src/constraint_model.cpp:394-501) and the degree->radian
conversion of addAngleConstraint (:133), generated from the reference's own constraint_model.cpp
(oracle/_ref/libgcs_ref.so; the solve itself is replaced by given solver positions, see
oracle/ref_model_driver.cpp).  Run in the build container:

    python oracle/make_golden_model.py        ->  tests/golden/model_transform.npz
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_lib as R  # noqa: E402

NCASE, NEL = 64, 12


def main():
    rng = np.random.default_rng(0xCA27A5)
    types = np.zeros((NCASE, NEL), dtype=np.int32)
    canvas = np.zeros((NCASE, NEL, 4))
    pos = np.zeros((NCASE, NEL, 4))
    solved = np.zeros((NCASE, NEL), dtype=np.uint8)
    out = np.zeros((NCASE, NEL, 4))
    for k in range(NCASE):
        t = (rng.uniform(size=NEL) < 0.35).astype(np.int32)
        c = rng.uniform(0, 1000, (NEL, 4))
        # solver frame = canvas moved rigidly (sometimes mirrored: the determinant fix) + noise
        th = rng.uniform(0, 2 * math.pi)
        rot = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        if k % 5 == 2:
            rot = rot @ np.diag([1.0, -1.0])
        sh = rng.uniform(-500, 500, 2)
        p = np.concatenate([(c[:, :2] - sh) @ rot, (c[:, 2:] - sh) @ rot], axis=1) + rng.normal(0, 2.0, (NEL, 4)) * (k % 2)
        s = (rng.uniform(size=NEL) < 0.8).astype(np.uint8)
        if k == 3:
            s[:] = 0                       # nothing solved: canvas untouched
        if k == 4:
            s[:] = 0
            s[np.nonzero(t == 0)[0][0]] = 1  # one solved point: translation only
            s[t == 1] = 1                  # ... applied to the solved lines as well
        if k == 6:
            s[t == 0] = 0                  # solved lines only: no point pairs -> nothing happens
        if k == 7:
            p[t == 0, :2] = p[np.nonzero(t == 0)[0][0], :2]  # coincident solver points: rank-0 covariance
        types[k], canvas[k], pos[k], solved[k] = t, c, p, s
        _, out[k], _ = R.model_solve_transform(t, c, p, s, np.zeros((0, 4), dtype=np.int32), np.zeros(0))
    # constraint acceptance + degree -> radian: points 0,1; lines 2,3
    ctypes_ = np.array([0, 0, 1, 1], dtype=np.int32)
    ccanvas = np.array([[0, 0, 0, 0], [10, 0, 0, 0], [0, 0, 10, 0], [0, 0, 0, 10]], dtype=float)
    degs = np.array([0.0, 1.0, 30.0, 45.0, 60.0, 90.0, 120.0, 179.5, 180.0, 33.333333333333336, 1e-9, 359.0])
    con, val = [], []
    for d in degs:
        con.append([2, 3, 1, int(d > 50)]), val.append(d)
    con += [[0, 1, 0, 0], [0, 2, 0, 0], [2, 3, 0, 0], [0, 1, 1, 0], [0, 2, 1, 1]]   # last three are rejected
    val += [5.0, 2.5, 7.0, 30.0, 30.0]
    n_acc, _, stored = R.model_solve_transform(ctypes_, ccanvas, np.zeros((4, 4)), np.zeros(4, dtype=np.uint8),
                                               np.array(con, dtype=np.int32), np.array(val))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "model_transform.npz"), types=types, canvas=canvas, pos=pos,
                        solved=solved, out=out, con_types=ctypes_, con_canvas=ccanvas, con=np.array(con, dtype=np.int32),
                        con_value=np.array(val), con_stored=stored, con_accepted=n_acc)
    moved = [(k, float(np.abs(out[k] - canvas[k]).max())) for k in (0, 3, 4, 6, 7)]
    print("cases", NCASE, "accepted", n_acc, "of", len(val), "stored", stored[:4], "moved", moved)


if __name__ == "__main__":
    main()

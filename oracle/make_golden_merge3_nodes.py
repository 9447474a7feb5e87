#!/usr/bin/env python
"""Golden vectors for the Merge3 enumeration loops and the merge node, generated from
the reference's own solver classes - Merge3PppSolver / Merge3PllSolver / Merge3LppSolver / Merge3LlpSolver /
Merge3FallbackSolver::solve and the case order of bottom_up_plan_solver.cpp:393-431 - compiled into
oracle/_ref/libgcs_ref.so (oracle/build_ref.sh, oracle/ref_merge3_driver.cpp gcs_ref_m3_merge).  Run in the build
container:

    python oracle/make_golden_merge3_nodes.py        ->  tests/golden/merge3_nodes.npz

Stored per scenario: the flattened inputs (element types, canvas rows, the three clusters' members and poses), which
entry point it is for, and what the reference returns (ids and poses of the merged cluster, the deciding case).
The scenarios are those of tests/test_merge3.py (_m3_scenario: three rigid clusters of one sketch under their own
motions, a noisy canvas), every shape of M3_SHAPES / M3_NODE_SHAPES / the PPP shapes, seeded."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import ref_lib as R  # noqa: E402
import test_merge3 as T  # noqa: E402


def main():
    rng = np.random.default_rng(20261019)
    todo = []
    for rep in range(3):
        for case in ("pll", "lpp", "llp"):
            todo += [(case, spec, True) for spec in T.M3_SHAPES[case]]
        todo += [("ppp", dict(ra=(2, 0), rb=(2, 0), f=(2, 0), r=(1, 1), a=(1, 0), b=(0, 1)), True), ("ppp", dict(ra=(3, 0), rb=(2, 0), f=(3, 1)), True)]
        todo += [("node", spec, expect != "fallback") for expect, spec in T.M3_NODE_SHAPES]
        todo += [("fallback", spec, True) for spec in (dict(ra=(2, 0), rb=(2, 0), a=(1, 1)), dict(ra=(1, 1), rb=(0, 2), r=(1, 0), b=(2, 0)))]
    out = {"which": [], "n_el": [], "types": [], "canvas4": [], "counts": [], "ids": [], "pose4": [], "ref_n": [], "ref_ids": [], "ref_pose4": [],
           "ref_by": []}
    with T._quiet_stderr():
        for which, spec, permute in todo:
            types, canvas4, clusters = T._m3_scenario(rng, spec, permute=permute)
            n, ids, pose, by = R.m3_merge(which, types, canvas4, clusters)
            out["which"].append(T.H.M3_CASES[which])
            out["n_el"].append(len(types))
            out["types"].append(types)
            out["canvas4"].append(canvas4)
            out["counts"].append([len(c) for c in clusters])
            out["ids"].append(np.array([i for c in clusters for i, _ in c], dtype=np.int32))
            out["pose4"].append(np.array([p for c in clusters for _, p in c], dtype=np.float64).reshape(-1, 4))
            out["ref_n"].append(n)
            out["ref_ids"].append(ids)
            out["ref_pose4"].append(pose.reshape(-1, 4))
            out["ref_by"].append(by)
    path = os.path.join(ROOT, "tests", "golden", "merge3_nodes.npz")
    np.savez_compressed(path, which=np.array(out["which"], dtype=np.int32), n_el=np.array(out["n_el"], dtype=np.int32),
                        types=np.concatenate(out["types"]).astype(np.int32), canvas4=np.concatenate(out["canvas4"]),
                        counts=np.array(out["counts"], dtype=np.int32), ids=np.concatenate(out["ids"]).astype(np.int32),
                        pose4=np.concatenate(out["pose4"]), ref_n=np.array(out["ref_n"], dtype=np.int32),
                        ref_ids=np.concatenate(out["ref_ids"]).astype(np.int32), ref_pose4=np.concatenate(out["ref_pose4"]),
                        ref_by=np.array(out["ref_by"], dtype=np.int32))
    print(f"{path}: {len(todo)} scenarios, {int(np.sum(np.array(out['ref_n']) > 0))} with a merged pose")


if __name__ == "__main__":
    main()

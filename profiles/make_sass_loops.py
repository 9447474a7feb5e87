"""Rebuilds profiles/r2f_sass_hot_loops.md from the SASS of the library in the tree (profiles/loopstat.py per kernel).
Usage: python profiles/make_sass_loops.py"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "2d_geometry_constraint_solver_b200", "libgcs_b200.so")


def loops(pat, *extra):
    return subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "loopstat.py"), LIB, pat, *extra],
                          capture_output=True, text=True).stdout


def main():
    version = importlib.import_module("2d_geometry_constraint_solver_b200").capi.load().gcs_b200_version().decode()
    out = ["# Hot loops of the final round-2 kernels (line form, linear K4), from the SASS of the shipped library\n",
           "`cuobjdump -sass 2d_geometry_constraint_solver_b200/libgcs_b200.so` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a -O3 "
           f"--fmad=false`), summarised by `profiles/loopstat.py`; library: {version}.\n",
           "What to read off: the contracted update loop in its line form is 19 instructions of which 8 FP64 and one MUFU (27 issue cycles "
           "per update and warp; the Cramer form it replaces - `profiles/r2_sass_hot_loops.md` - was 34 / 21 / 55), the same loop in every "
           "kind that has a line form (K1, K2, K3, K5); the second loop of each kernel is the outer decision loop around it, the others the "
           "careful-mode replay (Cramer form, 39 FP64) and the two inlined literal re-runs (88 FP64: the reference's arithmetic).  "
           "`newton_linear_kernel<2>` (K4 in the contracted class) has no loop on its certified path at all: the backward branches "
           "belong to the out-of-line literal fall-back.\n"]
    for k in (1, 5, 3):
        out.append(loops(f"newton_static_kernelILi{k}ELi2ELb1", "--sass", "1"))
    out.append(loops("newton_linear_kernelILi2E"))
    out.append(loops("newton_seq_kernelILi1ELi8ELb1", "--sass", "1"))
    open(os.path.join(ROOT, "profiles", "r2f_sass_hot_loops.md"), "w").write("\n".join(out))


if __name__ == "__main__":
    main()

"""FP64 pipe characterisation on the GPU box: peak, latency, and DFMA lanes/SM/clk versus
resident warps x per-thread ILP.  Usage: python profiles/probe_fp64.py"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
gcs.capi.init([0])
lib = gcs.capi.load()
print("DFMA peak TFLOP/s (FMA=2):", lib.gcs_b200_fp64_probe(0, 0))
print("DADD/DMUL mix TFLOP/s     :", lib.gcs_b200_fp64_probe(0, 1))
print("dependent DFMA latency clk:", lib.gcs_b200_fp64_probe(0, 2))
print("DFMA lanes per SM per clock (peak 64):")
print("warps/SM " + " ".join(f"ILP{i:>2}" for i in (1, 2, 4, 8)))
for w in (1, 2, 4, 8, 12, 16, 20, 24, 28, 32):
    print(f"{w:8d} " + " ".join(f"{lib.gcs_b200_fp64_probe(0, 1000 + 100 * i + w):5.1f}" for i in (1, 2, 4, 8)))
c = (C.c_uint64 * 8)()
print("selftest rc", lib.gcs_b200_selftest(0, 1, 1 << 26, c), list(c))

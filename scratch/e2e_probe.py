import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
capi, synth = gcs.capi, gcs.synth
capi.init([0])
def pin_like(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True); v = t.numpy(); v[...] = a; return t, v
keep = []
def mk(h):
    t, slab = pin_like(np.stack(h.cols)); keep.append(t)
    t, code = pin_like(h.code); keep.append(t)
    hb = capi.HostBatch(h.kind, h.n_seeds, [slab[c] for c in range(slab.shape[0])], code, None, 0, want_cand=False)
    m = hb.n
    t, oslab = pin_like(np.zeros((capi.OUT_COLS[h.kind], m))); keep.append(t)
    hb.out = [oslab[c] for c in range(oslab.shape[0])]
    t, hb.iters = pin_like(np.zeros((h.n_seeds, m), np.int16)); keep.append(t)
    t, hb.converged = pin_like(np.zeros((h.n_seeds, m), np.uint8)); keep.append(t)
    t, hb.root_index = pin_like(np.zeros(m, np.uint8)); keep.append(t)
    hb.cand = None
    return hb
n = 1 << 19
a, b = mk(synth.make_pp(n)), mk(synth.make_ang(n))
def t(f, reps=10):
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    return (time.perf_counter() - t0) / reps * 1e3
print("K1 sync   %.3f ms" % t(lambda: capi.solve_host(a, 0)))
print("K5 sync   %.3f ms" % t(lambda: capi.solve_host(b, 0)))
def both():
    capi.solve_host_async(a, 0); capi.solve_host_async(b, 0); capi.wait(0)
print("both      %.3f ms" % t(both))
def enq():
    t0 = time.perf_counter(); capi.solve_host_async(a, 0); capi.solve_host_async(b, 0); t1 = time.perf_counter(); capi.wait(0); return t1 - t0
print("enqueue only %.3f ms" % (np.mean([enq() for _ in range(10)]) * 1e3))
ca, cb = a.cbatch(), b.cbatch()
import ctypes as C
lib = capi.load()
def raw():
    lib.gcs_b200_solve_host_async(C.byref(ca), 0); lib.gcs_b200_solve_host_async(C.byref(cb), 0); lib.gcs_b200_wait(0)
print("raw both  %.3f ms" % t(raw))

"""Plan timing of the 100 k-point linkage on the host alone (no GPU needed: the call fails at the
first device call, after the plan has printed its split under GCS_HOST_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
os.environ["GCS_HOST_TRACE"] = "1"
import sketch_gen as S, host_lib as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
el, ed = S.make_linkage_unchecked(n, seed=3)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t = time.time()
    rc, _, st = H.system_solve_ex(el, ed)
    print("rc", rc, st, f"{time.time() - t:.3f}s", H.last_error()[:80])

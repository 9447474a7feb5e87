"""configs[3] through the host mirror, repeated: the split of solveGcs (GCS_HOST_TRACE=1 adds the plan's own)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import sketch_gen as S, host_lib as H
if os.environ.get("GCS_HOST_SO"):
    H.HOST_SO = os.path.join(H.HOST_DIR, os.environ["GCS_HOST_SO"])  # A/B against another build of the host library
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
t = time.time()
el, ed = S.make_linkage(n, seed=4)
print(f"generated in {time.time() - t:.1f}s", flush=True)
H.system_solve_ex(el[:2000], [e for e in ed if max(e["a"], e["b"]) < 2000])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    rc, _, st = H.system_solve_ex(el, ed)
    print("rc", rc, st, H.last_error()[:80] if rc else "", flush=True)

"""Drop-in check at the source level: the reference's own GUI model,
gui/src/constraint_model.cpp, is compiled unmodified against this repo's host headers (no Eigen:
the forwarding headers under host/include/Eigen supply the 2-D slice it uses) and linked with
libgcs_host.so.  The CPU test builds the client and runs the model calls that need no device; the
GPU test runs ConstraintModel::solveConstraintSystem() through the batched device path and checks
the constraints on the returned canvas coordinates (tolerance 1e-9 relative, the north star's)."""
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

CLIENT = os.path.join(ROOT, "build", "dropin", "model_client")


def test_reference_model_builds_against_host_headers():
    if not os.path.isdir(entry.REFERENCE):
        pytest.skip("reference sources not present")
    entry.build_cuda()
    entry.build_host()
    exe = entry.build_dropin(force=True)
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True, timeout=60)
    assert out.returncode == 0, out.stdout
    assert "status Nodes: 4  Edges: 5" in out.stdout


def _dist_to_line(p, a, b):
    dx, dy = b[0] - a[0], b[1] - a[1]
    return abs(dx * (p[1] - a[1]) - dy * (p[0] - a[0])) / math.hypot(dx, dy)


@pytest.mark.gpu
def test_reference_model_solves_on_device():
    if not os.path.exists(CLIENT):
        pytest.skip("build/dropin/model_client not built (needs the reference sources at build time)")
    out = subprocess.run([CLIENT, "solve"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    pts = [tuple(map(float, ln.split()[1:])) for ln in out.stdout.splitlines() if ln.startswith("point ")]
    line = [tuple(map(float, ln.split()[1:])) for ln in out.stdout.splitlines() if ln.startswith("line ")][0]
    a, b, c = pts
    rel = 1e-9
    assert math.dist(a, b) == pytest.approx(3.0, rel=rel)
    assert math.dist(b, c) == pytest.approx(4.0, rel=rel)
    assert math.dist(a, c) == pytest.approx(5.0, rel=rel)
    assert _dist_to_line(a, line[:2], line[2:]) == pytest.approx(2.0, rel=rel)
    assert _dist_to_line(c, line[:2], line[2:]) == pytest.approx(1.0, rel=rel)
    # the transform is a rigid fit onto the drawn sketch: the solved triangle keeps the drawn
    # orientation (counter-clockwise a, b, c) and stays near the drawn centroid
    cross = (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])
    assert cross > 0

#!/usr/bin/env bash
# Builds oracle/_ref/libgcs_ref.so: the reference's own Newton-Raphson path sources, compiled
# from where they lie under /root/reference (never copied), against the stand-in headers in
# oracle/ref_shim for the third-party / newer-compiler pieces this image lacks.
# Only runs where /root/reference exists (the build container); the GPU box uses the .so.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${GCS_REFERENCE:-/root/reference}"
cs="$ref/src/constraint_solver"
[ -d "$cs" ] || { echo "reference not present at $ref" >&2; exit 1; }
mkdir -p "$here/_ref"
# -fstack-reuse=none: ConstraintGraph::getConstraintBetweenNodes (gcs_data_structures.hpp:66-71)
# binds `const auto& edge` to `.value()` of a temporary std::expected, i.e. reads a dangling
# reference.  With stack-slot reuse g++ 13 -O2 clobbers it (std::bad_expected_access on ~25% of
# the partially solved leaves); keeping the slot alive gives the behaviour the author observes.
# -include format/tuple/iostream: headers the solver TUs use but only get transitively on GCC 15.
CXX="${GCS_REF_CXX:-/usr/bin/g++}"
"$CXX" -std=c++23 -O2 -ffp-contract=off -fno-fast-math -fstack-reuse=none -fPIC -fopenmp -shared -w \
  -fvisibility=hidden -include format -include tuple -include iostream \
  -I "$here/ref_shim" -I "$cs/includes" -I "$cs/src" -I "$ref/src/structures/include" \
  -o "$here/_ref/libgcs_ref.so" \
  -I "$ref/gui/src" \
  "$here/ref_driver.cpp" "$here/ref_merge3_driver.cpp" "$here/ref_model_driver.cpp" "$here/ref_graph_members.cpp" \
  "$ref/gui/src/constraint_model.cpp" \
  "$cs/src/solving/bottom_up/merge3_solver_common.cpp" \
  "$cs/src/solving/bottom_up/merge3_ppp_solver.cpp" \
  "$cs/src/solving/bottom_up/merge3_pll_solver.cpp" \
  "$cs/src/solving/bottom_up/merge3_lpp_solver.cpp" \
  "$cs/src/solving/bottom_up/merge3_llp_solver.cpp" \
  "$cs/src/solving/bottom_up/merge3_fallback_solver.cpp" \
  "$cs/src/model/elements.cpp" "$cs/src/model/constraints.cpp" \
  "$cs/src/solving/solvers/point_point_solvers.cpp" \
  "$cs/src/solving/solvers/point_line_solvers.cpp" \
  "$cs/src/solving/solvers/line_angle_solvers.cpp"
echo "built $here/_ref/libgcs_ref.so"

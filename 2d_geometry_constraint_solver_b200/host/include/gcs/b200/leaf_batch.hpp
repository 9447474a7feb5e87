// leaf_batch.hpp — host packer between the reference-shaped solver entry points and the C ABI
// (include/gcs_b200.h).
//
// The reference solves leaves one at a time: classifyAndSolve(leaf) probes eight matches()
// predicates and the winner's solve() assigns roles, anchors, builds two equations, runs solve2D,
// disambiguates and writes back (component_solver.hpp:31-66; point_point_solvers.cpp,
// point_line_solvers.cpp, line_angle_solvers.cpp).  Here the same steps are split so that many
// leaves share one kernel launch:
//   classify()  - the eight predicates in dispatch order (no numerics; can run on PREDICTED
//                 solved flags, which is what lets a whole decomposition be scheduled up front);
//   pack()      - role assignment (ascending NodeId), anchoring, canvas-side signs: the "packer
//                 work" of SURVEY.md Appendix B, producing one row of a kind's SoA batch;
//   KindBatch   - growable structure-of-arrays batch + the gcs_b200_batch descriptor over it;
//   apply()     - write-back through Element::updateElementPosition (sets m_isSet);
//   solveLeaves - the batched replacement of the sequential loop in
//                 DeficitStreeBasedTopDownStrategy::solveGcs (stree_top_down_strategy.cpp:41-45):
//                 a symbolic pass reproduces the reference's left-to-right classification, leaves
//                 are levelled by their read/write footprints, each level is one launch per kind.
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include <gcs/export.hpp>
#include <gcs/model/gcs_data_structures.hpp>
#include <gcs/model/solve_result.hpp>

#include "gcs_b200.h"

namespace Gcs::B200 {

enum class SolverId : int {
    None = 0,
    ZeroFixedPointsTriangle = 1,
    ZeroFixedPPLTriangle = 2,
    ZeroFixedLLPAngleTriangle = 3,
    TwoFixedPointsDistance = 4,
    TwoFixedPointsLine = 5,
    FixedPointAndLineFreePoint = 6,
    TwoFixedLinesFreePoint = 7,
    FixedLineAndPointFreeLine = 8,
};

GCS_API const char* solverName(SolverId id);
GCS_API int kindOf(SolverId id);  // GCS_KIND_* of the equation pair the solver builds

// "is this element already solved?"  Empty = ask the element (Element::isElementSet()).
using SetQuery = std::function<bool(const Element*)>;

GCS_API bool matches(SolverId id, const ConstraintGraph& component, const SetQuery& isSet = {});
// First match in the reference's dispatch order (component_solver.hpp:35-60), or None.
GCS_API SolverId classify(const ConstraintGraph& component, const SetQuery& isSet = {});

// Which elements a solver writes (anchors + the free element) and reads (the fixed ones).
struct Footprint {
    std::array<Element*, 3> writes {};
    int nWrites = 0;
    std::array<Element*, 3> reads {};
    int nReads = 0;
};
GCS_API Footprint footprint(SolverId id, const ConstraintGraph& component, const SetQuery& isSet = {});

// One packed leaf: a row of the kind's batch plus where the result goes.
struct PackedLeaf {
    SolverId solver = SolverId::None;
    int kind = 0;
    double in[GCS_MAX_IN_COLS] = {};
    std::uint8_t code = 0;
    Element* target = nullptr;  // receives a point (x, y) or a line (p1, p2)
};

// Role assignment + anchoring + canvas-side signs for one leaf.  The zero-fixed solvers place
// their anchor elements here (point_point_solvers.cpp:48-50, point_line_solvers.cpp:179-181,
// line_angle_solvers.cpp:249-274).  Throws what the reference throws for malformed leaves
// (std::bad_expected_access for a missing / virtual constraint).
GCS_API PackedLeaf pack(SolverId id, ConstraintGraph& component);
GCS_API void apply(const PackedLeaf& leaf, const double out[GCS_MAX_OUT_COLS]);

// Structure-of-arrays batch of one kind.
class GCS_API KindBatch {
public:
    explicit KindBatch(int kind = 0);
    int kind() const { return m_kind; }
    std::size_t size() const { return m_code.size(); }
    void clear();
    void push(const PackedLeaf& leaf);
    // descriptor over the current contents (host pointers, 2 seeds, default guesses);
    // allocates the output columns
    gcs_b200_batch descriptor();
    // after a solve: hand every row's output to its target element
    void applyAll();
    const std::vector<double>& column(int c) const { return m_in[c]; }
    const std::vector<std::uint8_t>& codes() const { return m_code; }
    std::vector<double>& out(int c) { return m_out[c]; }
    const std::vector<PackedLeaf>& leaves() const { return m_leaves; }

private:
    int m_kind;
    std::array<std::vector<double>, GCS_MAX_IN_COLS> m_in;
    std::vector<std::uint8_t> m_code;
    std::array<std::vector<double>, GCS_MAX_OUT_COLS> m_out;
    std::vector<std::int16_t> m_iters;
    std::vector<std::uint8_t> m_conv, m_root;
    std::vector<PackedLeaf> m_leaves;
};

// One leaf through the CUDA path (batch of one).  Throws std::runtime_error when the CUDA
// library cannot run (no device): there is no CPU fallback.
GCS_API SolveResult solveSingle(SolverId id, ConstraintGraph& component, int device = 0);

struct BatchReport {
    std::size_t leaves = 0, solved = 0, unsupported = 0, waves = 0, launches = 0;
    std::size_t shardedLaunches = 0;  // launches of a kind batch that went over several devices (solveLeavesOnDevices)
    double planSeconds = 0, packSeconds = 0, deviceSeconds = 0, applySeconds = 0;  // where solveLeaves spent its time
    std::vector<SolveResult> results;  // per leaf, in input order
    std::vector<int> level;            // wave of each leaf (-1 = unsupported)
    std::vector<SolverId> solver;      // solver chosen per leaf
};

// Symbolic pass only: per leaf the solver the reference's sequential loop would choose and the
// wave it can run in.  No numerics, no device.
GCS_API BatchReport planLeaves(const std::vector<ConstraintGraph>& leaves);
// Batched solve with the reference's sequential semantics; one launch per kind per wave.
GCS_API BatchReport solveLeaves(std::vector<ConstraintGraph>& leaves, int device = 0);
// The same with one sketch's waves spread over several GPUs: a wave's batch of a kind with at least
// `minRowsPerDevice` rows per device is cut into contiguous index ranges over the first nDevices
// devices of gcs_b200_init (gcs_b200_solve_sharded: one pipeline per device, no collective; the
// solved positions meet again in the host's elements, which is where the next wave is packed from -
// the "separator exchange" of SURVEY.md section 8f rank 1 is that write-back).  Smaller batches stay
// on the first device.  Results do not depend on nDevices.
GCS_API BatchReport solveLeavesOnDevices(std::vector<ConstraintGraph>& leaves, int nDevices, std::size_t minRowsPerDevice = 16384);

// Kernel class of every batch the host mirror launches: GCS_VARIANT_DEFAULT (bit-identical to the
// reference arithmetic; the default) or GCS_VARIANT_CONTRACTED (iteration counts, flags and roots
// of each solve identical for identical inputs, coordinates to 1e-9 relative - in a sketch the
// coordinates of one wave are the inputs of the next, so results then agree with the reference's
// loop to that tolerance, not bit for bit).  The environment variable GCS_B200_HOST_VARIANT sets
// the initial value.  Returns the previous setting.
GCS_API int setKernelVariant(int variant);
GCS_API int kernelVariant();

}  // namespace Gcs::B200

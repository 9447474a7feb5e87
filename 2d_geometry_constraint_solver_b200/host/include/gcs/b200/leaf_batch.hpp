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

// Structure-of-arrays batch of one kind, in ONE block of host memory: input columns, output columns
// and the code column at a constant spacing, so a call moves them with one strided copy each way
// (gcs_b200.h, "host buffers"), page-locked once the batch is large enough for that to matter
// (gcs_b200_host_alloc; ordinary memory when there is no device).  Blocks are recycled through a
// small process-wide pool: a solve of many waves, or many solves, allocate once.
class GCS_API KindBatch {
public:
    explicit KindBatch(int kind = 0);
    ~KindBatch();
    KindBatch(KindBatch&& other) noexcept;
    KindBatch& operator=(KindBatch&& other) noexcept;
    KindBatch(const KindBatch&) = delete;
    KindBatch& operator=(const KindBatch&) = delete;

    int kind() const { return m_kind; }
    std::size_t size() const { return m_size; }
    void clear();                     // keeps the storage
    void reserve(std::size_t rows);   // the kind must be known
    void push(const PackedLeaf& leaf);
    // rows [0, rows) to be filled with set() - from any thread, each row once
    void resize(std::size_t rows);
    void set(std::size_t row, const PackedLeaf& leaf);
    // descriptor over the current contents (host pointers, 2 seeds, default guesses; the flag
    // columns are not asked for: write-back needs the positions only)
    gcs_b200_batch descriptor();
    // after a solve: hand every row's output to its target element
    void applyAll();
    const double* column(int c) const { return doubles() + static_cast<std::size_t>(c) * m_cap; }
    const double* out(int c) const { return doubles() + (static_cast<std::size_t>(m_nin) + static_cast<std::size_t>(c)) * m_cap; }
    const std::uint8_t* codes() const { return m_slab + static_cast<std::size_t>(m_nin + m_nout) * m_cap * sizeof(double); }
    Element* target(std::size_t row) const { return m_target[row]; }
    bool pageLocked() const { return m_pinned; }

private:
    const double* doubles() const { return reinterpret_cast<const double*>(m_slab); }
    void grow(std::size_t rows);
    void release();
    int m_kind;
    int m_nin = 0, m_nout = 0;
    std::size_t m_size = 0, m_cap = 0, m_bytes = 0;
    unsigned char* m_slab = nullptr;
    bool m_pinned = false;
    std::vector<Element*> m_target;
};

// One leaf through the CUDA path (batch of one).  Throws std::runtime_error when the CUDA
// library cannot run (no device): there is no CPU fallback.
GCS_API SolveResult solveSingle(SolverId id, ConstraintGraph& component, int device = 0);

struct BatchReport {
    std::size_t leaves = 0, solved = 0, unsupported = 0, waves = 0, launches = 0;
    std::size_t shardedLaunches = 0;  // launches of a kind batch that went over several devices (solveLeavesOnDevices)
    double planSeconds = 0, packSeconds = 0, deviceSeconds = 0, applySeconds = 0;  // where solveLeaves spent its time
    std::vector<SolveStatus> status;   // per leaf, in input order
    // what classifyAndSolve would have returned for the leaf (the message is made here, on demand)
    SolveResult result(std::size_t leaf) const
    {
        if (status.at(leaf) == SolveStatus::Success) return SolveResult::success();
        return { status[leaf], "No solver matches this component configuration" };
    }
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

// Placeholder for the reference's bottom-up DR-plan strategy (reference:
// includes/gcs/decomposition/bottom_up/bottom_up_strategy.hpp:18-58).  That strategy - cluster
// graph reduction, plan tree, Merge3 / Merge2 plan solving - is outside this repo's path
// (DESIGN.md section 7); only its numeric Merge3 helpers are provided
// (solving/bottom_up/merge3_solver_common.hpp).  The class exists so that client code which can
// SELECT it (the GUI model, gui/src/constraint_model.cpp:366-368) compiles against these headers;
// choosing it fails loudly at run time.
#pragma once

#include <stdexcept>
#include <vector>

#include <gcs/export.hpp>
#include <gcs/orchestration/solving_strategy.hpp>

namespace Gcs {

class GCS_API BottomUpDrPlanStrategy : public GcsSolvingStrategy {
public:
    Constrainedness checkConstraintGraphConstrainedness(const ConstraintGraph&) override { notBuilt(); }
    bool resolve(ConstraintGraph&) override { notBuilt(); }
    std::vector<ConstraintGraph> decomposeConstraintGraph(ConstraintGraph&) override { notBuilt(); }
    void solveGcs(std::vector<ConstraintGraph>&) override { notBuilt(); }
    [[nodiscard]] bool hasReductionResult() const { return false; }
    ~BottomUpDrPlanStrategy() override = default;

private:
    [[noreturn]] static void notBuilt()
    {
        throw std::runtime_error("BottomUpDrPlanStrategy is not part of the B200 path; use DeficitStreeBasedTopDownStrategy");
    }
};

}  // namespace Gcs

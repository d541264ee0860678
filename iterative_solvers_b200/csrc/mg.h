// Internal: the opt-in multigrid-preconditioned CG (mg.cu), called from b200cg_solve.
#pragma once
#include "plan.h"

namespace b200cg {

// Runs the PCG loop on a plan whose init kernel has been enqueued; the result is left in the plan's x and in the
// device-state mirror. launches: kernels enqueued; interrupted: the caller's stop flag ended the loop.
int mg_pcg_solve(b200cg_plan_s* P, const volatile int* stop_flag, int64_t* launches, bool* interrupted);
int mg_prepare(b200cg_plan_s* P);  // builds the level hierarchy on first use (allocations: outside the timed solve)
int mg_levels(const b200cg_plan_s* P);
void mg_free(b200cg_plan_s* P);

}  // namespace b200cg

// engine.hpp -- internal declarations shared by engine.cu and host_model.cpp
#pragma once
#include "../../include/ellp_b200.h"

namespace ellp {
// solve_trivial_problem (reference: src/solvers/trivial/solve_trivial_problem.rs:5-96); pt->N / pt->N_side must have
// room for n entries; pt->nN receives the number written.
int host_solve_trivial(const ellp_std_form* sf, ellp_point* pt, bool minimize);
}  // namespace ellp



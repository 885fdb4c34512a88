// x3d_state.cuh -- sub-system state owned by the context (complete types for Ctx::~Ctx).
#pragma once
#include "x3d_ctx.cuh"

namespace x3d {

struct DecompState {
  virtual ~DecompState() = default;
};
struct PoissonState {
  virtual ~PoissonState() = default;
};
struct SolverState {
  virtual ~SolverState() = default;
};

}  // namespace x3d

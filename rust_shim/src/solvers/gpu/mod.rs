//! GPU solvers: the reference's two-phase drivers with `solve_with_initial` running on a B200
//! through the C ABI of `include/ellp_b200.h` (libellp_b200.so).
pub mod ffi;
pub mod gpu_simplex_solver;

pub use gpu_simplex_solver::{GpuDualSimplexSolver, GpuPrimalSimplexSolver};

// Stand-in for SWMF share/Library/src/linear_solver_wrapper_c.h (un-vendored), OURS: the GMRES of the field solve is not on
// the particle path; calling it in this build is an error.
#pragma once
#include <cstdlib>
template <class... A>
inline void linear_solver_wrapper(A...) { abort(); }
template <class... A>
inline void linear_solver_matvec_c(A...) { abort(); }

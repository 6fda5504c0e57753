// Fixed-grid forward / reverse-sweep kernels for (ode_hidden_dim=25, ode_state_dim=5): the CVS and challenge configs.
#include "slode_fixed.cuh"

SLODE_DEFINE_FIXED_SHAPE(25, 5)

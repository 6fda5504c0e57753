// Fixed-grid kernels for (ode_hidden_dim=25, ode_state_dim=8): the proc config; see slode_fixed.cuh.
#include "slode_fixed.cuh"

SLODE_DEFINE_FIXED_SHAPE(25, 8)

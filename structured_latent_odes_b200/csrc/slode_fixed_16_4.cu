// Fixed-grid forward / reverse-sweep kernels for (ode_hidden_dim=16, ode_state_dim=4); see slode_fixed.cuh.
#include "slode_fixed.cuh"

SLODE_DEFINE_FIXED_SHAPE(16, 4)

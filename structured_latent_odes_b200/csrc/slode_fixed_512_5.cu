// Fixed-grid forward / reverse-sweep kernels for (ode_hidden_dim=512, ode_state_dim=5); see slode_fixed.cuh.
#include "slode_fixed.cuh"

SLODE_DEFINE_FIXED_SHAPE(512, 5)

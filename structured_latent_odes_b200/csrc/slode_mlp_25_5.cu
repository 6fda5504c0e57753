// dopri5 forward / discrete backward kernels for (ode_hidden_dim=25, ode_state_dim=5); see slode_dopri5_kernels.cuh.
#define SLODE_PACK_SYM slode_c_pack_25_5
#include "slode_mlp_kernels.cuh"
#include "slode_dopri5_kernels.cuh"

SLODE_DEFINE_DOPRI5(25, 5)
SLODE_DEFINE_DOPRI5_BWD(25, 5)

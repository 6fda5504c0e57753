// dopri5 forward / discrete backward kernels for (ode_hidden_dim=16, ode_state_dim=4); see slode_dopri5_kernels.cuh.
#define SLODE_PACK_SYM slode_c_pack_16_4
#include "slode_mlp_kernels.cuh"
#include "slode_dopri5_kernels.cuh"

SLODE_DEFINE_DOPRI5(16, 4)
SLODE_DEFINE_DOPRI5_BWD(16, 4)

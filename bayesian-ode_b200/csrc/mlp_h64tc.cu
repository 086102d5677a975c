// Tensor-core instantiation of the MLP solver kernels (H = 64, mma.sync tf32 3x split; see mlp_tc.cuh).
#define BODE_H 64
#define BODE_MF MlpTcField
#define BODE_SFX 64tc
#include "mlp_inst.cuh"

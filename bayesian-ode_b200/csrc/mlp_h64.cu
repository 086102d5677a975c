#define BODE_H 64
#include "mlp_inst.cuh"

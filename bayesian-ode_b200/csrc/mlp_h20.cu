#define BODE_H 20
#include "mlp_inst.cuh"

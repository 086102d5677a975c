#define BODE_M 3
#include "npde_inst.cuh"

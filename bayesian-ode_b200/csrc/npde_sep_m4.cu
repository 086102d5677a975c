#define BODE_M 4
#include "npde_inst.cuh"

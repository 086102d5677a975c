#define BODE_M 6
#include "npde_inst.cuh"

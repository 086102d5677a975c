#define BODE_M 5
#include "npde_inst.cuh"

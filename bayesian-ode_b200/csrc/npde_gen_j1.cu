#define BODE_JPL 1
#include "npde_gen_inst.cuh"

#define BODE_JPL 2
#include "npde_gen_inst.cuh"

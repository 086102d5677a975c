#define BODE_JPL 8
#include "npde_gen_inst.cuh"

#define BODE_JPL 4
#include "npde_gen_inst.cuh"

#define BODE_ROW_MY 8
#include "npde_row_inst.cuh"

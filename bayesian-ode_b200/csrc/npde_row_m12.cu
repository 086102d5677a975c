#define BODE_ROW_MY 12
#include "npde_row_inst.cuh"

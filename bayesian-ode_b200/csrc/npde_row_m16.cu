#define BODE_ROW_MY 16
#include "npde_row_inst.cuh"

// MPPI solve kernels of the `none` cost plugin (see cps_mppi_inst.cuh).
#define CPS_MPPI_COST COST_NONE
#define CPS_MPPI_NAME none
#include "cps_mppi_inst.cuh"

// MPPI solve kernels of the `grad` cost plugin (see cps_mppi_inst.cuh).
#define CPS_MPPI_COST COST_GRAD
#define CPS_MPPI_NAME grad
#include "cps_mppi_inst.cuh"

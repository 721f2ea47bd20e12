// MPPI solve kernels of the `default` cost plugin (see cps_mppi_inst.cuh).
#define CPS_MPPI_COST COST_DEFAULT
#define CPS_MPPI_NAME default
#include "cps_mppi_inst.cuh"

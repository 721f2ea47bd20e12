// MPPI solve kernels of the `qb` cost plugin (see cps_mppi_inst.cuh).
#define CPS_MPPI_COST COST_QB
#define CPS_MPPI_NAME qb
#include "cps_mppi_inst.cuh"

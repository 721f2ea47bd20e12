// MPPI solve kernels of the `gradmin` cost plugin (see cps_mppi_inst.cuh).
#define CPS_MPPI_COST COST_GRADMIN
#define CPS_MPPI_NAME gradmin
#include "cps_mppi_inst.cuh"

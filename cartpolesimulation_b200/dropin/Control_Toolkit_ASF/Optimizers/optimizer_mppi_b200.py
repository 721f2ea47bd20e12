"""Drop-in plugin file: copy to <CartPoleSimulation>/Control_Toolkit_ASF/Optimizers/optimizer_mppi_b200.py.

import_optimizer_by_name("mppi-b200") (Control_Toolkit/others/globals_and_utils.py:89-119) finds this file by its
name and takes the class `optimizer_mppi_b200` from it; controller_mpc then constructs and drives it exactly like
optimizer_mppi (Control_Toolkit/Controllers/controller_mpc.py:56-109).  Needs `cartpolesimulation_b200` on sys.path
and its libcps_b200.so built; there is no CPU fallback.
"""
from cartpolesimulation_b200.optimizer_mppi_b200 import optimizer_mppi_b200  # noqa: F401

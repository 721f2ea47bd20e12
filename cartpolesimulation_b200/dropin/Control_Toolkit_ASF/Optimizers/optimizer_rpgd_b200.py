"""Drop-in plugin file: copy to <CartPoleSimulation>/Control_Toolkit_ASF/Optimizers/optimizer_rpgd_b200.py and give
config_optimizers.yml a `rpgd-b200:` block with the keys of `rpgd:`.  Found by import_optimizer_by_name("rpgd-b200")
(Control_Toolkit/others/globals_and_utils.py:89-119); driven by controller_mpc like optimizer_rpgd_tf."""
from cartpolesimulation_b200.optimizer_forward_b200 import optimizer_rpgd_b200  # noqa: F401

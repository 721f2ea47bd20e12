"""Drop-in plugin file: copy to <CartPoleSimulation>/Control_Toolkit_ASF/Optimizers/optimizer_random_action_b200.py and
give config_optimizers.yml a `random-action-b200:` block with the keys of `random-action-tf:`.  Found by
import_optimizer_by_name("random-action-b200") (Control_Toolkit/others/globals_and_utils.py:89-119)."""
from cartpolesimulation_b200.optimizer_forward_b200 import optimizer_random_action_b200  # noqa: F401

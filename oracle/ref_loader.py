"""Shim loader that imports the LIVE reference (/root/reference) read-only.

TEST INFRASTRUCTURE ONLY.  This module exists solely so that
``oracle/gen_golden.py`` can run the unmodified reference code of the MPPI hot
path in this container and freeze its outputs into ``tests/golden/*.npz``.
It cannot travel to the GPU box (``/root/reference`` does not exist there), so
nothing in ``tests/ -m gpu``, ``bench.py`` or ``__graft_entry__`` imports it.

What is shimmed (all outside the reference tree; nothing is written there):
  * stub modules for packages absent from this image (tensorflow, matplotlib,
    ruamel.yaml, watchdog, gymnasium, engineering_notation, ...), needed because
    hot-path modules import ``others/globals_and_utils.py`` (TF + matplotlib at
    module top, others/globals_and_utils.py:19-21), ``CartPole/cartpole_parameters.py:2``
    (ruamel) and ``Control_Toolkit/Cost_Functions/CostFunctionUpdater.py:3-4`` (watchdog);
  * ``SI_Toolkit.Compile`` alias (superproject imports it, CartPole/cartpole_equations.py:3;
    the pinned submodule only has SI_Toolkit/Functions/TF/Compile.py);
  * ``ComputationLibrary.loop`` (called at CartPole/cartpole_equations.py:255, absent from the
    pinned SI_Toolkit/computation_library.py);
  * ``PyTorchLibrary.zeros(shape=...)`` keyword (SI_Toolkit/Predictors/predictor_ODE.py:39);
  * compilation globally disabled so that Compile.py never imports real TF.
"""
from __future__ import annotations

import importlib
import os
import sys
import tempfile
import types

REF = os.environ.get("CPS_REFERENCE_ROOT", "/root/reference")

_LOADED = False


class _Stub(types.ModuleType):
    """Permissive module: any attribute is a callable stub class."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        obj = _StubObj(f"{self.__name__}.{name}")
        setattr(self, name, obj)
        return obj


class _StubObj:
    def __init__(self, name="stub"):
        self._n = name

    def __call__(self, *a, **k):
        return _StubObj(self._n + "()")

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _StubObj(self._n + "." + name)

    def __mro_entries__(self, bases):  # allows "class X(stub.Thing):"
        return (object,)

    def __iter__(self):
        return iter(())


_STUB_NAMES = [
    "tensorflow", "tensorflow_probability", "matplotlib", "matplotlib.pyplot",
    "matplotlib.widgets", "matplotlib.colors", "matplotlib.cm", "matplotlib.figure",
    "matplotlib.animation", "matplotlib.patches", "matplotlib.transforms", "matplotlib.ticker",
    "engineering_notation", "ruamel", "ruamel.yaml", "watchdog", "watchdog.observers",
    "watchdog.events", "gymnasium", "gymnasium.spaces", "gym", "gym.spaces", "seaborn",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "Control_Toolkit", "Optimizers"))


def load(workdir=None):
    """Make the reference importable; idempotent.  Changes cwd to the reference root (its YAML configs are opened
    cwd-relative at import time) or, if given, to `workdir`: a scratch workspace holding edited COPIES of the
    application-specific folders, which then shadow the reference's."""
    global _LOADED
    if _LOADED:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "cps_numba_cache"))
    os.makedirs(os.environ["NUMBA_CACHE_DIR"], exist_ok=True)
    os.chdir(workdir or REF)
    for p in (REF, os.path.join(REF, "SI_Toolkit", "src")) + ((workdir,) if workdir else ()):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for n in _STUB_NAMES:
        if n not in sys.modules:
            try:
                importlib.import_module(n)
            except Exception:
                m = _Stub(n)
                m.__path__ = []  # behave as a package
                sys.modules[n] = m
    wd = sys.modules["watchdog.events"]
    if isinstance(wd, _Stub):
        wd.FileSystemEventHandler = object

    # --- force compilation off before Compile.py is imported -----------------------------
    import SI_Toolkit.load_and_normalize as lan
    _orig_load_yaml = lan.load_yaml

    def load_yaml_patched(path, *a, **k):
        cfg = _orig_load_yaml(path, *a, **k)
        if str(path).endswith("CONFIG_COMPILATION.yml") and isinstance(cfg, dict):
            cfg = dict(cfg)
            cfg["GLOBALLY_DISABLE_COMPILATION"] = True
        return cfg

    lan.load_yaml = load_yaml_patched
    import SI_Toolkit.Functions.TF.Compile as compile_mod
    sys.modules["SI_Toolkit.Compile"] = compile_mod
    import SI_Toolkit
    SI_Toolkit.Compile = compile_mod

    # --- computation-library shims ------------------------------------------------------
    import SI_Toolkit.computation_library as cl

    def _loop(self, body, state, steps, counter=0):
        carry = (counter,) + tuple(state)
        for _ in range(int(steps)):
            carry = body(*carry)
        return carry

    cl.ComputationLibrary.loop = _loop
    import torch

    def _zeros(*args, shape=None, **kw):
        if shape is not None:
            return torch.zeros(tuple(shape), **kw)
        return torch.zeros(*args, **kw)

    _orig_init = cl.PyTorchLibrary.__init__

    def _init(self, *a, **k):  # `zeros` is an instance attribute set in __init__ (computation_library.py:456)
        _orig_init(self, *a, **k)
        self.zeros = _zeros

    cl.PyTorchLibrary.__init__ = _init
    _LOADED = True


# ------------------------------------------------------------------------------------------
# Convenience constructors around the unmodified reference classes
# ------------------------------------------------------------------------------------------

def torch_lib():
    load()
    from SI_Toolkit.computation_library import PyTorchLibrary
    return PyTorchLibrary()


def variable_parameters(lib, target_position=0.0, target_equilibrium=1.0, L=0.395, m_pole=0.087):
    load()
    from SI_Toolkit.General.variable_parameters import VariableParameters
    vp = VariableParameters(lib)
    vp.set_attributes({
        "target_position": float(target_position), "target_equilibrium": float(target_equilibrium),
        "L": float(L), "m_pole": float(m_pole)})
    return vp


def predictor_ODE_v0(horizon, dt=0.02, intermediate_steps=10, batch_size=1, variable_parameters=None):
    """Unmodified SI_Toolkit/Predictors/predictor_ODE_v0.py (numba explicit Euler + bounce)."""
    load()
    from SI_Toolkit.Predictors.predictor_ODE_v0 import predictor_ODE_v0 as P
    return P(horizon=horizon, dt=dt, intermediate_steps=intermediate_steps, batch_size=batch_size,
             variable_parameters=variable_parameters)


def predictor_ODE(horizon, dt=0.02, intermediate_steps=10, batch_size=1, variable_parameters=None):
    """Unmodified SI_Toolkit/Predictors/predictor_ODE.py under PyTorchLibrary (Euler-Cromer + atan2)."""
    load()
    from SI_Toolkit.Predictors.predictor_ODE import predictor_ODE as P
    return P(horizon=horizon, dt=dt, computation_library=torch_lib(), intermediate_steps=intermediate_steps,
             disable_individual_compilation=True, batch_size=batch_size,
             variable_parameters=variable_parameters)


class ODEv0CoreAdapter:
    """Duck-typed PredictorWrapper giving optimizer_mppi a predict_core over predictor_ODE_v0
    (SURVEY Appendix B.6): torch -> numpy -> predictor_ODE_v0.predict -> torch."""

    def __init__(self, horizon, batch_size, dt=0.02, intermediate_steps=10, variable_parameters=None):
        self.horizon, self.batch_size, self.dt = horizon, batch_size, dt
        self.intermediate_steps = intermediate_steps
        self.variable_parameters = variable_parameters
        self.predictor = predictor_ODE_v0(horizon, dt, intermediate_steps, batch_size, variable_parameters)
        self.num_states, self.num_control_inputs = 6, 1
        self.predictor_type = "ODE_v0"

    def predict_core(self, s, Q):
        import numpy as np
        import torch
        out = self.predictor.predict(s.numpy().astype(np.float32), Q.numpy().astype(np.float32))
        return torch.from_numpy(np.asarray(out).reshape(Q.shape[0], self.horizon + 1, 6))

    def update(self, Q0, s):
        pass

    def copy(self):
        return _NullPredictor()

    def configure(self, **kw):
        pass


class ODECoreAdapter:
    """PredictorWrapper-shaped holder around the unmodified torch predictor_ODE."""

    def __init__(self, horizon, batch_size, dt=0.02, intermediate_steps=10, variable_parameters=None):
        self.horizon, self.batch_size, self.dt = horizon, batch_size, dt
        self.predictor = predictor_ODE(horizon, dt, intermediate_steps, batch_size, variable_parameters)
        self.num_states, self.num_control_inputs = 6, 1
        self.predictor_type = "ODE"

    def predict_core(self, s, Q):
        return self.predictor.predict_core(s, Q)

    def update(self, Q0, s):
        pass

    def copy(self):
        return _NullPredictor()

    def configure(self, **kw):
        pass


class _NullPredictor:
    def configure(self, **kw):
        pass

    def predict_core(self, s, Q):
        raise RuntimeError("single-trajectory predictor not configured in the oracle harness")

    def update(self, **kw):
        pass


def cost_function(name, lib, vp, batch_size, horizon):
    """Unmodified Control_Toolkit/Cost_Functions/cost_function_wrapper.py -> CartPole plugin `name`."""
    load()
    from Control_Toolkit.Cost_Functions.cost_function_wrapper import CostFunctionWrapper
    import contextlib
    import io
    cw = CostFunctionWrapper()
    with contextlib.redirect_stdout(io.StringIO()):
        cw.configure(batch_size=batch_size, horizon=horizon, variable_parameters=vp,
                     environment_name="CartPole", computation_library=lib,
                     cost_function_specification=name)
    return cw


class InjectedNormal:
    """Replaces optimizer.rng: returns pre-supplied N(0,1) draws, one [K, n_ind, 1] tensor per solve
    (the 'identical injected noise' hook, Control_Toolkit/Optimizers/optimizer_mppi.py:172-174)."""

    def __init__(self, draws):
        self.draws = list(draws)
        self.i = 0

    def normal(self, shape, dtype=None):
        d = self.draws[self.i]
        self.i += 1
        assert list(d.shape) == list(shape), (d.shape, shape)
        return d


def optimizer_mppi(predictor, cost, K, T, dt=0.02, seed=1, cc_weight=1.0, R=1.0, LBD=100.0, NU=1000.0,
                   SQRTRHOINV=0.03, period=10, logging=True, lo=-1.0, hi=1.0):
    """Unmodified Control_Toolkit/Optimizers/optimizer_mppi.py under PyTorchLibrary."""
    load()
    import numpy as np
    from Control_Toolkit.Optimizers.optimizer_mppi import optimizer_mppi as O
    opt = O(predictor=predictor, cost_function=cost,
            control_limits=(np.array([lo], dtype=np.float32), np.array([hi], dtype=np.float32)),
            computation_library=torch_lib(), seed=seed, cc_weight=cc_weight, R=R, LBD=LBD,
            mpc_horizon=T, num_rollouts=K, NU=NU, SQRTRHOINV=SQRTRHOINV,
            period_interpolation_inducing_points=period, optimizer_logging=logging,
            calculate_optimal_trajectory=False)
    opt.configure(num_states=6, num_control_inputs=1, dt=dt, predictor_specification="ODE")
    return opt

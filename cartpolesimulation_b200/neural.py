"""predictor_autoregressive_neural backed by the CUDA network kernel (net_kernel, csrc/cps_net.cu) -- same
constructor keywords, methods and bookkeeping as the reference's

  SI_Toolkit/src/SI_Toolkit/Predictors/predictor_autoregressive_neural.py:44-360

Loading mirrors Functions/General/Initialization.py:35-238 (net-info .txt, checkpoint) and
Functions/General/Normalising.py:15-186 (minmax_sym coefficients, float32).  Supported: library `Pytorch`
checkpoints (`ckpt.pt` / `<name>.pt`, a torch state_dict of the `Sequence` network, Functions/Pytorch/Network.py)
of type GRU or Dense, plain or differential (`D_*` outputs).  TensorFlow checkpoints cannot be read without
TensorFlow; `weights_from_keras_gru` converts keras GRU arrays (reset_after=True) if the caller has them as numpy.
The arithmetic always runs on the GPU: there is no CPU path.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib as L
from .core import Engine
from .predictors import CONTROL_INPUTS, STATE_INDICES, STATE_VARIABLES, template_predictor

SUPPORTED_TYPES = ("GRU", "Dense")


# ----------------------------------------------------------------------------------------------------------------
# net-info / normalisation / checkpoint loading (host-side plumbing)
# ----------------------------------------------------------------------------------------------------------------
def load_net_info_from_txt_file(txt_path: str, net_info=None):
    """Headed sections of the net-info file (Initialization.py:35-104)."""
    net_info = SimpleNamespace() if net_info is None else net_info
    with open(txt_path, newline="") as f:
        lines = f.read().splitlines()
    for i, line in enumerate(lines[:-1]):
        nxt = lines[i + 1].rstrip("\n")
        if line == "LIBRARY:":
            net_info.library = nxt
        elif line == "NET NAME:":
            net_info.net_name = nxt
        elif line == "NET FULL NAME:":
            net_info.net_full_name = nxt
        elif line == "INPUTS:":
            net_info.inputs = nxt.split(sep=", ")
        elif line == "OUTPUTS:":
            net_info.outputs = nxt.split(sep=", ")
        elif line == "TYPE:":
            net_info.net_type = nxt.split(sep=", ")
        elif line == "NORMALIZATION:":
            net_info.path_to_normalization_info = nxt
        elif line == "NORMALIZE:":
            net_info.normalize = nxt == "True"
        elif line == "WASH OUT LENGTH:":
            net_info.wash_out_len = int(nxt)
        elif line == "CONSTRUCT NETWORK:":
            net_info.construct_network = nxt
        elif line == "TIMESTEP MEAN [s]:":
            net_info.dt = float(nxt)
    return net_info


def load_normalization_info(path):
    """(columns, table[4, n]) with rows mean, std, max, min (load_and_normalize.py:562-563)."""
    import pandas as pd
    df = pd.read_csv(path, index_col=0, comment="#")
    return [str(c) for c in df.columns], df.loc[["mean", "std", "max", "min"]].values.astype(np.float64)


def _norm_rows(norm, names):
    cols, table = norm
    t = np.full((4, len(names)), np.nan, dtype=np.float32)
    for i, n in enumerate(names):
        if n in cols:
            t[:, i] = np.asarray(table, dtype=np.float32)[:, cols.index(n)]
    return t


def normalization_coeffs(norm, names):
    """minmax_sym a, b with normalized = a*x + b (Normalising.py:44-46), float32 arithmetic as the torch library."""
    t = _norm_rows(norm, names)
    mx, mn = t[2], t[3]
    a = np.float32(2.0) / (mx - mn)
    b = np.float32(-1.0) + np.float32(2.0) * (-mn / (mx - mn))
    return a.astype(np.float32), b.astype(np.float32)


def denormalization_coeffs(norm, names):
    """minmax_sym A, B with x = A*normalized + B (Normalising.py:93-95)."""
    t = _norm_rows(norm, names)
    mx, mn = t[2], t[3]
    A = (mx - mn) / np.float32(2.0)
    B = (mx - mn) / np.float32(2.0) + mn
    return A.astype(np.float32), B.astype(np.float32)


def flatten_state_dict(net_type, sd):
    """torch `Sequence` state_dict (layers.<l>.weight_ih ... / layers.<l>.weight ...) -> (hidden sizes, flat float32
    weight vector in the order include/cps.h documents)."""
    get = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], dtype=np.float32)
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith("layers.")})
    n_hidden = len(idx) - 1
    parts, hsz = [], []
    for l in range(n_hidden):
        if net_type == "GRU":
            w_ih, w_hh = get(f"layers.{l}.weight_ih"), get(f"layers.{l}.weight_hh")
            parts += [w_ih, w_hh, get(f"layers.{l}.bias_ih"), get(f"layers.{l}.bias_hh")]
            hsz.append(w_hh.shape[1])
        else:
            w = get(f"layers.{l}.weight")
            parts += [w, get(f"layers.{l}.bias")]
            hsz.append(w.shape[0])
    parts += [get(f"layers.{n_hidden}.weight"), get(f"layers.{n_hidden}.bias")]
    return hsz, np.concatenate([p.reshape(-1) for p in parts]).astype(np.float32)


def weights_from_keras_gru(kernel, recurrent_kernel, bias):
    """keras GRU (reset_after=True) arrays -> torch GRUCell arrays.  keras: kernel [in, 3H], recurrent_kernel [H, 3H],
    bias [2, 3H], gate order (z, r, h); torch: weight_ih [3H, in], weight_hh [3H, H], gate order (r, z, n).  The math
    is identical (the reference's own restatement: Functions/TF/TF2Numpy.py:21-51)."""
    kernel, rk, bias = (np.asarray(x, dtype=np.float32) for x in (kernel, recurrent_kernel, bias))
    H = rk.shape[0]
    perm = np.concatenate([np.arange(H, 2 * H), np.arange(0, H), np.arange(2 * H, 3 * H)])
    return (np.ascontiguousarray(kernel.T[perm]), np.ascontiguousarray(rk.T[perm]),
            np.ascontiguousarray(bias[0][perm]), np.ascontiguousarray(bias[1][perm]))


def build_net_spec(net_type, inputs, outputs, hsz, weights, norm=None, dt=None):
    """The flat description cps_net_load takes, derived exactly as predictor_autoregressive_neural.__init__ does
    (:157-199): which inputs are controls / state features, output -> state index maps, (de)normalisation vectors,
    differential-network scaling (Normalising.py:111-186, autoregression.py:118-158)."""
    if net_type not in SUPPORTED_TYPES:
        raise NotImplementedError(f"network type {net_type!r} is not supported on the GPU path (supported: {SUPPORTED_TYPES})")
    inputs, outputs = list(inputs), list(outputs)
    ext = [x for x in inputs if x in CONTROL_INPUTS]
    if ext != ["Q"] or inputs[0] != "Q":
        raise NotImplementedError("the network must take the control input Q as its first input")
    state_in = inputs[1:]
    for n in state_in:
        if n not in STATE_INDICES:
            raise ValueError(f"unknown network input {n!r}")
    differential = any("D_" in o for o in outputs)  # :135
    out_names = [(o[2:] if o[:2] == "D_" else o) for o in outputs]
    for n in out_names:
        if n not in STATE_INDICES:
            raise ValueError(f"unknown network output {n!r}")
    spec = dict(net_type=net_type, hsz=[int(h) for h in hsz], weights=np.asarray(weights, dtype=np.float32),
                in_idx=[STATE_INDICES[n] for n in state_in], out_idx=[STATE_INDICES[n] for n in out_names],
                inputs=inputs, outputs=outputs, differential=differential)
    n_in, n_out = len(inputs), len(outputs)
    if norm is not None:
        spec["norm_a"], spec["norm_b"] = normalization_coeffs(norm, inputs)
        spec["denorm_A"], spec["denorm_B"] = denormalization_coeffs(norm, out_names)
    else:
        spec["norm_a"], spec["norm_b"] = np.ones(n_in, np.float32), np.zeros(n_in, np.float32)
        spec["denorm_A"], spec["denorm_B"] = np.ones(n_out, np.float32), np.zeros(n_out, np.float32)
    if differential:
        if dt is None:
            raise ValueError("Differential network was loaded but timestep dt was not provided to the predictor")
        if norm is not None:
            a, b = normalization_coeffs(norm, out_names)       # of the integrated variables
            C, D = denormalization_coeffs(norm, outputs)       # of the derivatives
            spec["diff_p1"] = (a * C * np.float32(dt)).astype(np.float32)
            spec["diff_p2"] = (a * D * np.float32(dt)).astype(np.float32)
            spec["out_norm_a"], spec["out_norm_b"] = a, b
        else:
            spec["diff_p1"] = np.full(n_out, np.float32(dt))
            spec["diff_p2"] = np.zeros(n_out, np.float32)
            spec["out_norm_a"], spec["out_norm_b"] = np.ones(n_out, np.float32), np.zeros(n_out, np.float32)
        spec["out_to_in"] = [out_names.index(n) for n in state_in]
    elif out_names[:len(state_in)] != state_in:
        # autoregression.py:94-98 feeds the output vector back as the next input vector, position by position
        raise ValueError("the network outputs must repeat its state inputs in the same order (autoregressive feedback)")
    return spec


# value ranges of the reference's shipped normalisation table
# (GymlikeCartPole/Dense-7IN-32H1-32H2-1OUT-0/NI_2024-08-17_22-23-01.csv)
DEFAULT_RANGES = {"Q": (1.0, -1.0), "angle": (np.pi, -np.pi), "angleD": (18.38, -18.38), "angle_cos": (1.0, -1.0),
                  "angle_sin": (1.0, -1.0), "position": (0.198, -0.198), "positionD": (1.125, -1.125)}


def synthetic_net_spec(hsz=(64, 64), net_type="GRU", seed=0):
    """A randomly initialised network of the architecture the reference's configs name (GRU-6IN-<H1>-<H2>-5OUT,
    config_predictors.yml:10,32): no dynamics model ships with the reference, so benchmarks and smoke tests use seeded
    weights with torch's default GRUCell / Linear initialisation U(-1/sqrt(H), 1/sqrt(H))."""
    inputs = ["Q", "angleD", "angle_cos", "angle_sin", "position", "positionD"]
    outputs = inputs[1:]
    rng = np.random.default_rng(seed)
    parts, n_in = [], len(inputs)
    G = 3 if net_type == "GRU" else 1
    for H in hsz:
        k = 1.0 / np.sqrt(H)
        parts.append(rng.uniform(-k, k, (G * H, n_in)))
        if net_type == "GRU":
            parts += [rng.uniform(-k, k, (3 * H, H)), rng.uniform(-k, k, 3 * H), rng.uniform(-k, k, 3 * H)]
        else:
            parts.append(rng.uniform(-k, k, H))
        n_in = H
    k = 1.0 / np.sqrt(n_in)
    parts += [rng.uniform(-k, k, (len(outputs), n_in)), rng.uniform(-k, k, len(outputs))]
    w = np.concatenate([q.reshape(-1) for q in parts]).astype(np.float32)
    cols = list(DEFAULT_RANGES)
    table = np.array([[0.0] * len(cols), [1.0] * len(cols), [DEFAULT_RANGES[c][0] for c in cols],
                      [DEFAULT_RANGES[c][1] for c in cols]])
    return build_net_spec(net_type, inputs, outputs, list(hsz), w, (cols, table))


def net_flops_per_step(spec) -> int:
    """FLOPs of one network step for one rollout (multiply-add = 2)."""
    G = 3 if spec["net_type"] == "GRU" else 1
    f, i = 0, 1 + len(spec["in_idx"])
    for H in spec["hsz"]:
        f += 2 * G * H * (i + (H if G == 3 else 0))
        i = H
    return f + 2 * len(spec["out_idx"]) * i


def load_model(model_name, path_to_model=None, dt=None):
    """(spec, net_info) of a stored network: <path_to_models>/<net_name>/{<net_name>.txt, ckpt.pt | <net_name>.pt, NI csv}."""
    model_name = os.path.normpath(model_name)
    if len(model_name.split(os.sep)) > 1:  # predictor_autoregressive_neural.py:67-78
        path_to_models = os.path.join(*model_name.split(os.sep)[:-1]) + os.sep
        if model_name.startswith(os.sep):
            path_to_models = os.sep + path_to_models
        net_name = model_name.split(os.sep)[-1]
    else:
        if path_to_model is None:
            raise ValueError("path_to_model is required when model_name carries no path")
        path_to_models = os.path.normpath(path_to_model) + os.sep
        net_name = model_name
    folder = os.path.join(path_to_models, net_name)
    if not os.path.isdir(folder):
        raise FileNotFoundError("{} not found".format(net_name))
    txt = os.path.join(folder, net_name + ".txt")
    if not os.path.isfile(txt):
        raise FileNotFoundError("The corresponding .txt file is missing (information about inputs and outputs) at the "
                                "location {}".format(txt))
    info = load_net_info_from_txt_file(txt, SimpleNamespace(path_to_models=path_to_models))
    info.parent_net_name = net_name
    info.path_to_net = folder
    if getattr(info, "library", "TF") != "Pytorch":
        raise NotImplementedError("only Pytorch checkpoints can be loaded here (TensorFlow is not available); convert "
                                  "keras GRU arrays with weights_from_keras_gru and use build_net_spec")
    ntype = info.net_name.split("-")[0]
    ckpt = next((p for p in (os.path.join(folder, net_name + ".pt"), os.path.join(folder, "ckpt.pt")) if os.path.isfile(p)), None)
    if ckpt is None:
        raise FileNotFoundError("The corresponding .ckpt file is missing (information about weights and biases) in " + folder)
    sd = torch.load(ckpt, map_location="cpu", weights_only=True)
    if ntype not in SUPPORTED_TYPES:
        raise NotImplementedError(f"network type {ntype!r} is not supported on the GPU path (supported: {SUPPORTED_TYPES})")
    hsz, flat = flatten_state_dict(ntype, sd)
    norm = None
    if getattr(info, "normalize", False):
        p = info.path_to_normalization_info
        if not os.path.isfile(p):  # Initialization.py:71-74: look beside the network
            p = os.path.join(folder, os.path.basename(p))
        norm = load_normalization_info(p)
    if hasattr(info, "dt"):
        dt = info.dt
    spec = build_net_spec(ntype, info.inputs, info.outputs, hsz, flat, norm, dt)
    return spec, info


def spec_from_reference_predictor(pred):
    """Network description from a live REFERENCE predictor_autoregressive_neural object (its torch `net`, `net_info`
    and `normalization_info`), so that the reference's own PredictorWrapper can be handed to optimizer_mppi_b200."""
    info = pred.net_info
    if getattr(info, "library", None) != "Pytorch":
        raise NotImplementedError("only Pytorch networks can be taken over from a reference predictor")
    ntype = info.net_name.split("-")[0]
    if ntype not in SUPPORTED_TYPES:
        raise NotImplementedError(f"network type {ntype!r} is not supported on the GPU path")
    hsz, flat = flatten_state_dict(ntype, pred.net.state_dict())
    norm = None
    ni = getattr(pred, "normalization_info", None)
    if ni is not None:
        norm = ([str(c) for c in ni.columns], ni.loc[["mean", "std", "max", "min"]].values.astype(np.float64))
    return build_net_spec(ntype, info.inputs, info.outputs, hsz, flat, norm, getattr(pred, "dt", None))


# ----------------------------------------------------------------------------------------------------------------
class predictor_autoregressive_neural(template_predictor):
    supported_computation_libraries = ("TF", "Pytorch")

    def __init__(self, model_name=None, path_to_model=None, horizon=None, dt=None, batch_size=None,
                 variable_parameters=None, disable_individual_compilation=False, update_before_predicting=True,
                 mode=None, hls=False, input_quantization="float", device=None, net_spec=None, **kwargs):
        super().__init__(horizon=horizon, batch_size=batch_size)
        if hls or input_quantization != "float" or mode == "simple evaluation":
            raise NotImplementedError("hls / input quantisation / 'simple evaluation' are outside the B200 hot path")
        self.dt = dt
        self.variable_parameters = variable_parameters
        if net_spec is None:
            net_spec, self.net_info = load_model(model_name, path_to_model, dt)
        else:
            self.net_info = SimpleNamespace(inputs=net_spec["inputs"], outputs=net_spec["outputs"],
                                            net_type=net_spec["net_type"], library="Pytorch")
        if hasattr(self.net_info, "dt"):
            self.dt = self.net_info.dt
        self.net_spec = net_spec
        self.differential_network = bool(net_spec["differential"])
        self.engine = Engine(num_rollouts=max(int(batch_size or 1), 1), horizon=int(horizon), dt=float(self.dt or 0.02),
                             integrator="neural", cost=None, device=device)
        self.engine.net_load(net_spec)
        self.device = self.engine.device
        self.update_before_predicting = update_before_predicting
        self.last_initial_state = None
        self.last_optimal_control_input = None
        self.model_input_features = list(net_spec["inputs"])
        self.model_output_features = list(net_spec["outputs"])
        self.output = None

    # -- tensors in, same kind out ------------------------------------------------------------------------------
    def _to_dev(self, x):
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=self.device, dtype=torch.float32).contiguous(), ("torch", x.device)
        return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float32)), device=self.device), ("numpy", None)

    def predict_core(self, initial_state, Q):
        """[B,6] (or [1,6]: tiled), [B,T,1] -> [B,T+1,6]; starts from the stored hidden state (:291), which it leaves
        untouched."""
        s_d, kind = self._to_dev(initial_state)
        Q_d, _ = self._to_dev(Q)
        if Q_d.ndim != 3 or Q_d.shape[2] != 1:
            raise ValueError(f"Q must have shape [batch_size, horizon, 1], got {tuple(Q_d.shape)}")
        self.last_initial_state = s_d.reshape(-1, 6)[0].clone()
        traj, _ = self.engine.net_rollout(s_d, Q_d[:, :, 0], q_layout=L.ROLLOUT_MAJOR, traj_layout=L.ROLLOUT_MAJOR)
        if kind[0] == "torch":
            return traj if kind[1] == self.device else traj.to(kind[1])
        return traj.cpu().numpy()

    def predict(self, initial_state, Q, last_optimal_control_input=None) -> np.ndarray:
        initial_state = np.asarray(initial_state, dtype=np.float32)
        Q = np.asarray(Q, dtype=np.float32)
        if initial_state.ndim == 1:  # check_dimensions (autoregression.py:161-174)
            initial_state = initial_state[np.newaxis, :]
        if Q.ndim == 2:
            Q = Q[np.newaxis, :, :]
        elif Q.ndim == 1:
            Q = Q[np.newaxis, np.newaxis, :]
        if self.update_before_predicting and self.last_initial_state is not None and (
                last_optimal_control_input is not None or self.last_optimal_control_input is not None):
            if last_optimal_control_input is None:
                last_optimal_control_input = self.last_optimal_control_input
            self.update_internal_state_tf(last_optimal_control_input, self.last_initial_state)
        self.output = self.predict_core(initial_state, Q)
        return self.output

    def update_internal_state_tf(self, Q0=None, s=None):
        """One network step on (Q0, s) advancing the stored hidden state (:332-352).  The reference keeps one hidden
        row per batch entry, all identical because its callers tile s and Q0; the first row is used here."""
        if self.net_spec["net_type"] == "Dense":
            return
        s_d, _ = self._to_dev(s)
        q_d, _ = self._to_dev(Q0)
        self.engine.net_update(s_d.reshape(-1, 6)[0].contiguous(), q_d.reshape(-1)[:1].contiguous())

    def update_internal_state(self, Q0=None, s=None):
        if s is None:
            s = self.last_initial_state
        if self.update_before_predicting:
            self.last_optimal_control_input = Q0
            s_d, _ = self._to_dev(s)
            self.last_initial_state = s_d.reshape(-1, 6)[0].clone()
        else:
            self.update_internal_state_tf(Q0, s)

    def reset(self):
        self.last_optimal_control_input = None  # (:355-357); the hidden state is kept, as in the reference

    def reset_internal_states(self):
        """net.reset_internal_states() (:112): zero hidden state."""
        self.engine.net_reset_state()

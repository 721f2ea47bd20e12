"""Builds the flat network description (cps_net_load / cps_oracle_net_rollout arguments) from a net golden file.

Follows the reference's bookkeeping: SI_Toolkit/src/SI_Toolkit/Predictors/predictor_autoregressive_neural.py:163-199
(input/output index maps), Functions/General/Normalising.py:15-108 (minmax_sym coefficients, computed in float32 as
the torch library does)."""
import json

import numpy as np

STATE_VARIABLES = ["angle", "angleD", "angle_cos", "angle_sin", "position", "positionD"]


def minmax_sym_coeffs(table, cols, names):
    """(a, b, A, B) float32 vectors for `names`; table rows: mean, std, max, min (Normalising.py:1-9)."""
    t = np.asarray(table, dtype=np.float32)
    idx = [list(cols).index(n) for n in names]
    mx, mn = t[2, idx], t[3, idx]
    a = np.float32(2.0) / (mx - mn)
    b = np.float32(-1.0) + np.float32(2.0) * (-mn / (mx - mn))
    A = (mx - mn) / np.float32(2.0)
    B = (mx - mn) / np.float32(2.0) + mn
    return a.astype(np.float32), b.astype(np.float32), A.astype(np.float32), B.astype(np.float32)


def net_spec_from_golden(z):
    meta = json.loads(str(z["meta"]))
    inputs, outputs, ntype = meta["inputs"], meta["outputs"], meta["type"]
    keys = [str(k) for k in z["state_dict_keys"]]
    n_lay = max(int(k.split(".")[1]) for k in keys)  # index of the output layer
    layers, hsz = [], []
    for l in range(n_lay):
        if ntype == "GRU":
            lay = tuple(z[f"w__layers.{l}.{n}"] for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))
            hsz.append(lay[1].shape[1])
        else:
            lay = (z[f"w__layers.{l}.weight"], z[f"w__layers.{l}.bias"])
            hsz.append(lay[0].shape[0])
        layers.append(lay)
    out_layer = (z[f"w__layers.{n_lay}.weight"], z[f"w__layers.{n_lay}.bias"])
    cols = [str(c) for c in z["norm_cols"]]
    a, b, _, _ = minmax_sym_coeffs(z["norm_table"], cols, inputs)
    # differential networks (outputs D_*): the integrated variables carry the output names without the prefix
    # (predictor_autoregressive_neural.py:184-190)
    out_names = [(o[2:] if o[:2] == "D_" else o) for o in outputs]
    _, _, A, B = minmax_sym_coeffs(z["norm_table"], cols, out_names)
    diff = None
    if any("D_" in o for o in outputs):
        # Normalising.py:111-186: p1 = a * C * dt, p2 = a * D * dt with a of the integrated variables and C, D of the
        # derivatives; autoregression.py:137-146: state -> output and output -> input index maps
        on_a, on_b, _, _ = minmax_sym_coeffs(z["norm_table"], cols, out_names)
        _, _, Cd, Dd = minmax_sym_coeffs(z["norm_table"], cols, outputs)
        dt = np.float32(meta["dt"])
        diff = dict(p1=(on_a * Cd * dt).astype(np.float32), p2=(on_a * Dd * dt).astype(np.float32), on_a=on_a, on_b=on_b,
                    out_to_in=[out_names.index(n) for n in inputs[1:]])
    return dict(net_type=ntype, hsz=hsz, layers=layers, out_layer=out_layer,
                in_idx=[STATE_VARIABLES.index(n) for n in inputs[1:]],
                out_idx=[STATE_VARIABLES.index(n) for n in out_names],
                norm_a=a, norm_b=b, denorm_A=A, denorm_B=B, diff=diff, meta=meta)


def write_model_dir(root, z):
    """Writes <root>/<net full name>/{<name>.txt, ckpt.pt, NI.csv} from a net golden file, in the reference's formats
    (Functions/General/Initialization.py:35-104; torch state_dict of Functions/Pytorch/Network.py Sequence).
    Returns the model path to use as `predictor_specification` / `model_name`."""
    import os
    import torch
    meta = json.loads(str(z["meta"]))
    full = meta["net"]
    d = os.path.join(root, full)
    os.makedirs(d, exist_ok=True)
    short = "-".join(p for p in full.split("-") if not (p.endswith("IN") or p.endswith("OUT")))[:-2]
    cols = [str(c) for c in z["norm_cols"]]
    ni = os.path.join(d, "NI.csv")
    with open(ni, "w") as f:
        f.write("," + ",".join(cols) + "\n")
        for r, rowname in enumerate(["mean", "std", "max", "min"]):
            f.write(rowname + "," + ",".join(repr(float(z["norm_table"][r, c])) for c in range(len(cols))) + "\n")
    with open(os.path.join(d, full + ".txt"), "w") as f:
        f.write("\n".join([
            "CREATED:", "2026-01-01 00:00:00", "", "LIBRARY:", "Pytorch", "", "NET NAME:", short, "",
            "NET FULL NAME:", full, "", "INPUTS:", ", ".join(meta["inputs"]), "", "OUTPUTS:", ", ".join(meta["outputs"]),
            "", "TYPE:", meta["type"], "", "NORMALIZATION:", ni, "", "NORMALIZE:", "True", "",
            "WASH OUT LENGTH:", "10", "", "CONSTRUCT NETWORK:", "with cells", ""]))
    sd = {str(k): torch.from_numpy(np.array(z["w__" + str(k)])) for k in z["state_dict_keys"]}
    torch.save(sd, os.path.join(d, "ckpt.pt"))
    return os.path.join(root, full)

"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HP_KEYS = ["tau2", "update_tau2", "tau2_a0", "tau2_b0", "sigma2", "update_sigma2", "sigma2_a0",
           "sigma2_b0", "p0", "update_p0", "min_p0", "p0_a0", "p0_b0"]

SAMPLER_CASES = ["linear_c1", "linear_minp0", "linear_fixed", "linear_dense", "linear_odd",
                 "probit_small", "probit_sigma"]
LINEAR_CASES = [c for c in SAMPLER_CASES if c.startswith("linear")]
PROBIT_CASES = [c for c in SAMPLER_CASES if c.startswith("probit")]
COUNTING_CASES = ["hand", "sparse_a", "sparse_b", "wide", "float_vals"]
MULTI_CASES = ["multi_b3", "multi_b5", "multi_b2_even", "multi_b4_wrap", "multi_b5_cap2", "multi_fixed",
               "multi_probit_b3", "multi_probit_b2"]


def load_sampler_case(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    hp = {k: float(d["hp_" + k]) for k in HP_KEYS}
    for k in ("update_tau2", "update_sigma2", "update_p0"):
        hp[k] = bool(hp[k])
    case = dict(name=name, X=d["X"], y=d["y"], probit=bool(d["probit"]), N=int(d["N"]),
                chains=int(d["chains"]), tn_seed=int(d["tn_seed"]), hp=hp, chains_data=[])
    for c in range(case["chains"]):
        pre = f"c{c}_"
        st = {k[len(pre) + 3:]: d[k] for k in d.files if k.startswith(pre + "st_")}
        ref = {k[len(pre) + 4:]: d[k] for k in d.files if k.startswith(pre + "ref_")}
        case["chains_data"].append(dict(streams=st, ref=ref))
    return case


def stacked_streams(case):
    """Per-chain recorded streams -> the (chains, sweeps, ...) arrays the C ABI takes."""
    keys = case["chains_data"][0]["streams"].keys()
    out = {}
    for k in keys:
        arr = np.stack([cd["streams"][k] for cd in case["chains_data"]])
        out["z" if k == "zn" else k] = arr
    return out


def load_multi_case(name):
    """Block-sampler fixture written by oracle/pin_multi.py (reference run + recorded stream)."""
    case = load_sampler_case(name)
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    case["bsize"] = int(d["bsize"])
    case["max_signals"] = int(d["max_signals"])
    return case


def latent_err(a, b):
    """Error of probit latents against their unit scale (values near the truncation point cancel)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0))) if a.size else 0.0


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den))


def load_counting_case(name):
    d = np.load(os.path.join(GOLDEN, f"counting_{name}.npz"))
    out = {"samples": d["samples"]}
    for k in d.files:
        if k == "samples":
            continue
        out[k] = json.loads(str(d[k]))
    return out


def golden_groups(entry):
    """JSON [[group, pep_hex], ...] -> list of (tuple, float)."""
    return [(tuple(g), float.fromhex(h)) for g, h in entry]


def cg_list(cgs):
    return [(tuple(sorted(int(j) for j in cg.group)), float(cg.pep)) for cg in cgs]

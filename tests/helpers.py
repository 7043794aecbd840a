"""Shared test utilities: golden fixtures, synthetic weights, oracle configs."""
import os

import numpy as np
import torch

from nmrf_b200.synthetic import state_dict_fingerprint, synthetic_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def make_cfg(max_disp, K, L):
    import nmrf_b200
    cfg = nmrf_b200.get_cfg()
    cfg.DPN.MAX_DISP, cfg.DPN.NUM_PROPOSALS = int(max_disp), int(K)
    cfg.NMP.NUM_PROP_LAYERS, cfg.NMP.NUM_INFER_LAYERS, cfg.NMP.NUM_REFINE_LAYERS = (int(x) for x in L)
    return cfg


def build_product_model(max_disp, K, L, seed, mode):
    """nmrf_b200.NMRF (CPU, parameters from synthetic_state_dict). Returns (model, state_dict)."""
    import nmrf_b200
    model = nmrf_b200.build_model(make_cfg(max_disp, K, L)).eval()
    sd = synthetic_state_dict(model.state_dict(), seed=seed, mode=mode)
    model.load_state_dict(sd, strict=True)
    return model, sd


def oracle_cfg(max_disp, K, L, taps=True):
    from oracle import nmrf_oracle as O
    return O.OracleConfig(max_disp=int(max_disp), num_proposals=int(K), num_prop_layers=int(L[0]),
                          num_infer_layers=int(L[1]), num_refine_layers=int(L[2]), taps={} if taps else None)


def check_fingerprint(sd, expected):
    got = state_dict_fingerprint(sd)
    assert abs(got - float(expected)) <= 1e-6 * abs(float(expected)), (
        f"synthetic weights differ from the ones the golden fixture was generated with "
        f"(fingerprint {got} vs {float(expected)}): torch RNG drift?")


def epe(a, b):
    return float((a.double() - b.double()).abs().mean())

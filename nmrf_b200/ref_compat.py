"""Stand-ins for the four third-party modules the REFERENCE tree imports but this image does not have (no network):
timm (Mlp, DropPath, to_2tuple, trunc_normal_), yacs.config.CfgNode, omegaconf.DictConfig, imageio.

Only needed to import the reference's own torch modules next to this package -- the Swin-T + DeformNeck encoder of
BASELINE config 5 (`nmrf.models.backbone.SwinAdaptor`, built by `nmrf_b200.config.create_backbone`), which is outside the hot
path and stays the reference's code -- and by the test infrastructure (oracle/ref_shims.py).  `install_missing()` registers a
stand-in ONLY for a module that cannot be imported; an installed timm / yacs is never shadowed.  Each stand-in restates a few
lines of glue (timm 0.9.16 `Mlp.forward` is fc1 -> act -> drop -> fc2 -> drop; `DropPath` is the identity in eval mode).
"""
import importlib.util
import sys
import types

import torch
from torch import nn


class _Mlp(nn.Module):
    """timm 0.9.16 `Mlp` forward restated: fc1 -> act -> drop1 -> fc2 -> drop2."""

    def __init__(self, in_features, hidden_features=None, out_features=None,
                 act_layer=nn.GELU, bias=True, drop=0.0, **_):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class _DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        raise NotImplementedError("DropPath in training mode is out of scope")


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class _CfgNode(dict):
    """Just enough of yacs.config.CfgNode for `nmrf.config` to import."""

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__(init_dict or {})

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        import copy
        return copy.deepcopy(self)



def _missing(name):
    try:
        return name not in sys.modules and importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def install_missing():
    """register stand-ins for timm / yacs / omegaconf / imageio where the real module is absent; returns the names installed"""
    done = []
    if _missing("timm"):
        layers = types.ModuleType("timm.models.layers")
        layers.Mlp, layers.DropPath, layers.to_2tuple = _Mlp, _DropPath, _to_2tuple
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm, timm_models, timm_layers = types.ModuleType("timm"), types.ModuleType("timm.models"), types.ModuleType("timm.layers")
        for k in ("Mlp", "DropPath", "to_2tuple", "trunc_normal_"):
            setattr(timm_layers, k, getattr(layers, k))
        timm.models, timm.layers, timm_models.layers = timm_models, timm_layers, layers
        sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": layers, "timm.layers": timm_layers})
        done.append("timm")
    if _missing("yacs"):
        yacs, yacs_config = types.ModuleType("yacs"), types.ModuleType("yacs.config")
        yacs_config.CfgNode = _CfgNode
        yacs.config = yacs_config
        sys.modules.update({"yacs": yacs, "yacs.config": yacs_config})
        done.append("yacs")
    if _missing("omegaconf"):
        omegaconf = types.ModuleType("omegaconf")
        omegaconf.DictConfig = type("DictConfig", (dict,), {})
        sys.modules["omegaconf"] = omegaconf
        done.append("omegaconf")
    if _missing("imageio"):
        sys.modules["imageio"] = types.ModuleType("imageio")
        done.append("imageio")
    return done

"""Minimal stand-in for the reference's yacs config (nmrf/config/default.py:20-61): only the keys
the model reads, same names, YAML overlays (`configs/*.yaml` of the reference load unchanged;
unknown keys -- solver, datasets ... -- are kept but ignored)."""
import copy

import yaml

_DEFAULTS = {
    "BACKBONE": {"MODEL_TYPE": "resnet", "NORM_FN": "instance", "OUT_CHANNELS": 256, "WEIGHT_URL": "",
                 "DROP_PATH": 0.0, "COMPAT": True},
    "DPN": {"MAX_DISP": 320, "COST_GROUP": 4, "NUM_PROPOSALS": 4, "CONTEXT_DIM": 64},
    "NMP": {"PROP_EMBED_DIM": 128, "INFER_EMBED_DIM": 128, "MLP_RATIO": 4, "SPLIT_SIZE": 1, "WINDOW_SIZE": 6,
            "REFINE_WINDOW_SIZE": 4, "PROP_N_HEADS": 4, "INFER_N_HEADS": 4, "NUM_PROP_LAYERS": 5,
            "NUM_INFER_LAYERS": 5, "NUM_REFINE_LAYERS": 5, "RETURN_INTERMEDIATE": True, "ATTN_DROP": 0.0,
            "PROJ_DROP": 0.0, "DROP_PATH": 0.0, "DROPOUT": 0.0, "NORMALIZE_BEFORE": True, "WITH_REFINEMENT": True},
    "DATASETS": {"DIVIS_BY": 8},
    "SOLVER": {"AUX_LOSS": False},
}


class CfgNode(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    @staticmethod
    def _wrap(d):
        return CfgNode({k: CfgNode._wrap(v) if isinstance(v, dict) else v for k, v in d.items()})

    def merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k].merge(v)
            else:
                self[k] = CfgNode._wrap(v) if isinstance(v, dict) else v
        return self

    def merge_from_file(self, path):
        with open(path) as f:
            return self.merge(yaml.safe_load(f) or {})

    def merge_from_list(self, kv):
        for key, val in zip(kv[0::2], kv[1::2]):
            node = self
            *parents, leaf = key.split(".")
            for p in parents:
                node = node[p]
            node[leaf] = yaml.safe_load(val) if isinstance(val, str) else val
        return self


def get_cfg():
    return CfgNode._wrap(copy.deepcopy(_DEFAULTS))


def create_backbone(cfg):
    """`create_backbone` of the reference (nmrf/models/backbone.py:176-200).  The ResNet-style encoder is built here; the
    Swin-T + DeformNeck encoder is the REFERENCE's own torch module (outside the hot path), constructed from the
    reference tree when it is importable, with `nmrf_b200.msda` installed as its `MultiScaleDeformableAttention` extension --
    the one native op of that encoder (boundary B2)."""
    import torch.nn as nn
    from .backbone import Backbone
    model_type = cfg.BACKBONE.MODEL_TYPE
    if model_type == "resnet":
        if cfg.BACKBONE.NORM_FN == "instance":
            norm_layer = nn.InstanceNorm2d
        elif cfg.BACKBONE.NORM_FN == "batch":
            norm_layer = nn.BatchNorm2d
        else:
            raise ValueError(f"Invalid backbone normalization type: {cfg.BACKBONE.NORM_FN}")
        return Backbone(cfg.BACKBONE.OUT_CHANNELS, norm_layer)
    if model_type == "swin":
        from . import msda, ref_compat
        msda.install_as_reference_extension()                      # the encoder's one native op (boundary B2)
        ref_compat.install_missing()                               # timm / yacs / omegaconf / imageio stand-ins where absent
        try:
            from nmrf.models.backbone import SwinAdaptor          # the reference's module, unmodified
        except Exception as e:                                     # the reference tree is not on sys.path
            raise NotImplementedError(
                "BACKBONE.MODEL_TYPE='swin' uses the reference's SwinAdaptor (nmrf/models/backbone.py:101-158): put the reference "
                f"tree on sys.path (baseline/stage_reference.py stages it under baseline/_ref) -- import failed with: {e!r}") from e
        backbone = SwinAdaptor(out_channels=cfg.BACKBONE.OUT_CHANNELS, drop_path_rate=cfg.BACKBONE.DROP_PATH)
        if cfg.BACKBONE.WEIGHT_URL:
            import torch
            weight = torch.load(cfg.BACKBONE.WEIGHT_URL, map_location="cpu")
            weight = weight.get("model", weight)
            weight = weight.get("state_dict", weight)
            weight = {k: v for k, v in weight.items() if "attn_mask" not in k and not k.startswith(("norm", "head"))}
            backbone.backbone.load_state_dict(weight)
        return backbone
    raise ValueError(f"Do not find {model_type}")


def build_model(cfg):
    """`nmrf.models.build_model` (models/__init__.py:9-10) without the training criterion."""
    from .model import DPN, NMRF
    backbone = create_backbone(cfg)
    dpn = DPN(cost_group=cfg.DPN.COST_GROUP, num_proposals=cfg.DPN.NUM_PROPOSALS, feat_dim=cfg.BACKBONE.OUT_CHANNELS,
              context_dim=cfg.DPN.CONTEXT_DIM, num_prop_layers=cfg.NMP.NUM_PROP_LAYERS,
              prop_embed_dim=cfg.NMP.PROP_EMBED_DIM, mlp_ratio=cfg.NMP.MLP_RATIO, split_size=cfg.NMP.SPLIT_SIZE,
              prop_n_heads=cfg.NMP.PROP_N_HEADS)
    return NMRF(backbone=backbone, dpn=dpn, num_proposals=cfg.DPN.NUM_PROPOSALS, max_disp=cfg.DPN.MAX_DISP,
                num_infer_layers=cfg.NMP.NUM_INFER_LAYERS, num_refine_layers=cfg.NMP.NUM_REFINE_LAYERS,
                infer_embed_dim=cfg.NMP.INFER_EMBED_DIM, infer_n_heads=cfg.NMP.INFER_N_HEADS,
                mlp_ratio=cfg.NMP.MLP_RATIO, window_size=cfg.NMP.WINDOW_SIZE,
                refine_window_size=cfg.NMP.REFINE_WINDOW_SIZE, normalize_before=cfg.NMP.NORMALIZE_BEFORE,
                divis_by=cfg.DATASETS.DIVIS_BY, compat=cfg.BACKBONE.COMPAT)

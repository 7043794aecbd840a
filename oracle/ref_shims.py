"""TEST INFRASTRUCTURE ONLY -- import shims that let the UNMODIFIED reference
(`/root/reference`, aeolusguan/NMRF) be imported and run on a CPU-only box.

Used only by `oracle/make_golden.py` (fixture generation, in the build
container) and by tests that cross-check the oracle restatement against the
real reference when `/root/reference` is present.  Nothing in the product
path (`nmrf_b200/`) may import this file.

The reference needs six modules that are not installed here:
  timm.models.layers / timm.layers : Mlp, DropPath, to_2tuple, trunc_normal_
      (NMP.py:8, NMRF.py:5, DPN.py:4, backbone.py:9, adaptor_modules.py:6)
  yacs.config.CfgNode               (config/config.py:12)
  omegaconf.DictConfig              (config/config.py:340)
  imageio                           (utils/frame_utils.py:8)
  MultiScaleDeformableAttention     (ops/functions/ms_deform_attn_func.py:11)
      -> routed to the reference's own pure-PyTorch
         ms_deform_attn_core_pytorch (ms_deform_attn_func.py:49-70)
"""
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("NMRF_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "nmrf", "models"))


from nmrf_b200.ref_compat import _CfgNode, _DropPath, _Mlp, _to_2tuple, install_missing  # noqa: E402,F401  (shared stand-ins)


def install():
    """Install the stand-ins into sys.modules, put the reference on sys.path, and route its MultiScaleDeformableAttention
    extension to the reference's own pure-PyTorch implementation."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "nmrf_ref_shims_installed" in sys.modules:
        return
    install_missing()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    msda = types.ModuleType("MultiScaleDeformableAttention")
    sys.modules["MultiScaleDeformableAttention"] = msda
    from ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa: E402

    def ms_deform_attn_forward(value, shapes, level_start, loc, w, im2col_step):
        return ms_deform_attn_core_pytorch(value, shapes, loc, w)

    msda.ms_deform_attn_forward = ms_deform_attn_forward
    sys.modules["nmrf_ref_shims_installed"] = types.ModuleType("nmrf_ref_shims_installed")


def build_reference_model(*, feat_dim=256, max_disp=192, num_proposals=4,
                          num_prop_layers=8, num_infer_layers=8, num_refine_layers=8,
                          divis_by=8, seed=0):
    """Construct the reference NMRF (ResNet backbone) from explicit kwargs,
    seeded, with `dpn.prop_head.layers[-1]` re-randomised (SURVEY.md H6:
    DPN.py:68-69 zero-initialises it, which would hide the propagation stack)."""
    install()
    from nmrf.models.NMRF import NMRF
    from nmrf.models.DPN import DPN
    from nmrf.models.backbone import Backbone

    torch.manual_seed(seed)
    backbone = Backbone(feat_dim, nn.InstanceNorm2d)
    dpn = DPN(cost_group=4, num_proposals=num_proposals, feat_dim=feat_dim, context_dim=64,
              num_prop_layers=num_prop_layers, prop_embed_dim=128, mlp_ratio=4, split_size=1,
              prop_n_heads=4, normalize_before=True)
    model = NMRF(backbone=backbone, dpn=dpn, num_proposals=num_proposals, max_disp=max_disp,
                 num_infer_layers=num_infer_layers, num_refine_layers=num_refine_layers,
                 infer_embed_dim=128, infer_n_heads=4, mlp_ratio=4, window_size=6,
                 refine_window_size=4, return_intermediate=False, normalize_before=True,
                 divis_by=divis_by, compat=True)
    g = torch.Generator().manual_seed(seed + 12345)
    with torch.no_grad():
        last = model.dpn.prop_head.layers[-1]
        last.weight.copy_(torch.nn.init.trunc_normal_(torch.empty_like(last.weight), std=0.02, generator=g))
    model.eval()
    return model

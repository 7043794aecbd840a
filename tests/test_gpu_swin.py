"""GPU: BASELINE config 5 -- the reference's Swin-T + DeformNeck encoder (UNMODIFIED, from baseline/_ref) with
`nmrf_b200.msda` as its MultiScaleDeformableAttention extension (operator boundary B2), module-level:
  * the encoder's features with our kernel == with the reference's own pure-PyTorch `ms_deform_attn_core_pytorch`,
  * `nmrf_b200.build_model(cfg)` with BACKBONE.MODEL_TYPE = "swin" constructs (configs/sceneflow_swint.yaml keys) and runs
    the whole forward: the hot path behind a foreign encoder (C = 128, DIVIS_BY = 32).
Skipped when baseline/_ref has not been staged (python baseline/stage_reference.py, in the build container)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nmrf")), reason="baseline/_ref not staged")]


@pytest.fixture(scope="module")
def reference_on_path():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from nmrf_b200 import msda, ref_compat
    ref_compat.install_missing()
    msda.install_as_reference_extension()
    import ops.functions.ms_deform_attn_func as F
    return F


def test_swin_adaptor_with_our_msda_matches_reference_pytorch_core(reference_on_path):
    Fmod = reference_on_path
    import nmrf_b200.msda as msda
    from nmrf.models.backbone import SwinAdaptor
    torch.manual_seed(0)
    enc = SwinAdaptor(out_channels=128, drop_path_rate=0.0).eval().cuda()
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(2, 3, 192, 256, generator=g) * 255).cuda()
    calls = {"n": 0}

    class Counting:
        @staticmethod
        def ms_deform_attn_forward(*a):
            calls["n"] += 1
            return msda.ms_deform_attn_forward(*a)
    with torch.no_grad():
        Fmod.MSDA = Counting
        ours = enc(x.clone())
        core = Fmod.ms_deform_attn_core_pytorch

        class Core:
            @staticmethod
            def ms_deform_attn_forward(value, shapes, level_start, loc, w, step):
                return core(value, shapes, loc, w)
        Fmod.MSDA = Core
        ref = enc(x.clone())
        Fmod.MSDA = msda
    assert calls["n"] == 4                                  # the neck's four Extractors (adaptor_modules.py:145-188)
    assert ours[0].shape == (2, 128, 48, 64) and ours[1].shape == (2, 128, 24, 32)
    for a, b in zip(ours, ref):
        assert float((a - b).abs().max() / b.abs().max()) <= 2e-5


def test_build_model_swin_runs_the_whole_forward(reference_on_path):
    import nmrf_b200
    from nmrf_b200.synthetic import synthetic_pair
    cfg = nmrf_b200.get_cfg()
    cfg.merge_from_file(os.path.join(REF, "configs", "sceneflow_swint.yaml"))
    cfg.BACKBONE.DROP_PATH = 0.0
    cfg.DPN.MAX_DISP, cfg.NMP.NUM_PROP_LAYERS, cfg.NMP.NUM_INFER_LAYERS, cfg.NMP.NUM_REFINE_LAYERS = 128, 1, 1, 1
    torch.manual_seed(0)
    model = nmrf_b200.build_model(cfg).eval().cuda()
    assert type(model.image_encoder).__name__ == "SwinAdaptor" and not model.compat and model.divis_by == 32
    img1, img2 = synthetic_pair(1, 200, 300, 128, index=1)          # padded to 224 x 320 (DIVIS_BY 32)
    out = model({"img1": img1, "img2": img2})
    assert out["disp"].shape == (1, 200, 300) and out["disp_pred"].shape == (1, 224, 320)
    assert out["proposal"].shape == (1, 28 * 40, 4) and out["prob"].shape == (28 * 40, 16)
    assert bool(torch.isfinite(out["disp"]).all()) and float(out["disp"].min()) >= 0.0
    out2 = model({"img1": img1, "img2": img2})
    assert torch.equal(out["initial_proposal"], out2["initial_proposal"])

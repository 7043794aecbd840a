"""CPU: host-side logic -- config, checkpoint contract, weight packing, padding geometry."""
import numpy as np
import pytest
import torch

from helpers import build_product_model, golden, make_cfg


def test_state_dict_contract_matches_reference_checkpoints():
    """keys + shapes of nmrf_b200.NMRF == those of the reference model (recorded in the golden by
    oracle/make_golden.py after `reference.load_state_dict(strict=True)`)."""
    for name in ("e2e_tiny", "e2e_small"):
        g = golden(name)
        model, _ = build_product_model(g["max_disp"], g["K"], g["L"], 0, "reference")
        sd = model.state_dict()
        assert sorted(sd.keys()) == [str(k) for k in g["keys"]]
        assert [str(tuple(sd[k].shape)) for k in sorted(sd.keys())] == [str(s) for s in g["shapes"]]
        assert sd["inference.layers.0.nmp.attn.relative_position_index"].dtype == torch.int64


def test_backbone_prefix_follows_compat_flag():
    import nmrf_b200
    cfg = make_cfg(64, 2, (1, 1, 1))
    cfg.BACKBONE.COMPAT = False
    keys = nmrf_b200.build_model(cfg).state_dict().keys()
    assert any(k.startswith("image_encoder.") for k in keys) and not any(k.startswith("backbone.") for k in keys)


def test_reference_style_yaml_overlay(tmp_path):
    import nmrf_b200
    p = tmp_path / "kitti.yaml"
    p.write_text("DPN:\n  MAX_DISP: 192\nNMP:\n  NUM_INFER_LAYERS: 3\nSOLVER:\n  MAX_ITER: 1000\nDATASETS:\n  TRAIN: ['kitti']\n")
    cfg = nmrf_b200.get_cfg()
    cfg.merge_from_file(str(p))
    cfg.merge_from_list(["DPN.NUM_PROPOSALS", "2"])
    assert cfg.DPN.MAX_DISP == 192 and cfg.NMP.NUM_INFER_LAYERS == 3 and cfg.DPN.NUM_PROPOSALS == 2
    assert cfg.NMP.WINDOW_SIZE == 6 and cfg.SOLVER.MAX_ITER == 1000
    m = nmrf_b200.build_model(cfg)
    assert len(m.inference.layers) == 3 and m.max_disp == 192


def test_unsupported_configs_fail_loudly():
    import nmrf_b200
    cfg = make_cfg(64, 2, (1, 1, 1))
    cfg.NMP.SPLIT_SIZE = 7
    with pytest.raises(NotImplementedError):
        nmrf_b200.build_model(cfg)
    cfg = make_cfg(64, 2, (1, 1, 1))
    cfg.BACKBONE.MODEL_TYPE = "swin"
    with pytest.raises(NotImplementedError):
        nmrf_b200.build_model(cfg)


def test_weight_packing_is_layout_only():
    from nmrf_b200.hotpath import HotPathConfig, PackedWeights
    model, sd = build_product_model(192, 4, (2, 2, 2), 3, "stress")
    pw = PackedWeights(sd, model.hot_path_config())
    L = pw.prop_layers[1]
    q = "dpn.propagation.layers.1.nmp"
    assert L["qkv_w"].shape == (384, 192)
    assert torch.equal(L["qkv_w"][:128], sd[q + ".q.weight"]) and torch.equal(L["qkv_w"][128:256], sd[q + ".k.weight"])
    assert torch.equal(L["qkv_w"][256:, :128], sd[q + ".v.weight"]) and float(L["qkv_w"][256:, 128:].abs().max()) == 0
    assert torch.equal(L["qkv_b"][256:], sd[q + ".v.bias"])
    S = pw.stacks["inference"]["layers"][0]
    s = "inference.layers.0.self_nmp"
    assert S["s_qkv_w"].shape == (384, 160)
    assert torch.equal(S["s_qkv_w"][:128, :159], sd[s + ".q.weight"]) and float(S["s_qkv_w"][:, 159].abs().max()) == 0
    assert torch.equal(S["s_qkv_w"][256:, :128], sd[s + ".v.weight"]) and float(S["s_qkv_w"][256:, 128:].abs().max()) == 0
    assert S["qkv_w"].shape == (384, 160) and torch.equal(S["qkv_w"][:, :159], sd["inference.layers.0.nmp.qkv.weight"])
    assert pw.ce0_w.shape == (128, 48) and float(pw.ce0_w[:, 36:].abs().max()) == 0
    assert pw.pproj_w.shape == (128, 160)
    assert "self" not in "".join(pw.stacks["refinement"]["layers"][0].keys()) or True
    assert "s_qkv_w" not in pw.stacks["refinement"]["layers"][0]


def test_center_pad_matches_reference_formula():
    from nmrf_b200.hotpath import center_pad
    for n in range(1, 40):
        for ws in (4, 6):
            pad = (ws - n % ws) % ws                       # NMP.py:747-754
            assert center_pad(n, ws) == (n + pad, pad // 2)
    assert center_pad(68, 6) == (72, 2) and center_pad(120, 6) == (120, 0) and center_pad(47, 6) == (48, 0)


def test_synthetic_weights_are_deterministic_and_mode_dependent():
    from nmrf_b200.synthetic import state_dict_fingerprint, synthetic_pair, synthetic_state_dict
    model, sd = build_product_model(64, 2, (1, 1, 1), 0, "reference")
    again = synthetic_state_dict(model.state_dict(), 0, "reference")
    assert all(torch.equal(sd[k], again[k]) for k in sd)
    assert float(sd["dpn.prop_head.layers.2.weight"].abs().max()) > 0        # H6: not the zero init
    assert float(sd["inference.layers.0.nmp.attn.relative_position_enc_table"].abs().max()) == 0
    stress = synthetic_state_dict(model.state_dict(), 0, "stress")
    assert float(stress["inference.layers.0.nmp.attn.relative_position_enc_table"].abs().max()) > 0
    assert state_dict_fingerprint(sd) != state_dict_fingerprint(stress)
    a, b = synthetic_pair(1, 32, 64, 64, index=1)
    a2, _ = synthetic_pair(1, 32, 64, 64, index=1)
    assert torch.equal(a, a2) and a.shape == b.shape == (1, 3, 32, 64) and float(a.max()) <= 255


def test_param_containers_do_not_compute():
    model, _ = build_product_model(64, 2, (1, 1, 1), 0, "reference")
    with pytest.raises(RuntimeError, match="libnmrf_b200"):
        model.inference(torch.zeros(1))


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the staged reference, or its oracle port, on the host cores; no GPU involved) prints ONE JSON line with the keys
    the driver reads; same metric / unit / workload as the B200 arm."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("stereo pairs/sec at 960x540") and d["config"]["workload"] == "sceneflow_540x960_D192_K4_L8"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

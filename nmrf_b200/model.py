"""Drop-in `NMRF` model: the reference's module boundary B1 on top of the CUDA hot path.

Mirrors `nmrf.models.NMRF.NMRF` (reference nmrf/models/NMRF.py:21-262): same constructor
arguments, same sub-module attribute names -- hence the same state-dict keys, so released
checkpoints load with `strict=True` -- same `forward(sample) -> dict` contract (eval mode).
The sub-modules below are PARAMETER CONTAINERS: their tensors are packed once
(`hotpath.PackedWeights`) and consumed by libnmrf_b200.so.  Only the feature extractor and
the conv3x3+InstanceNorm+ReLU+conv1x1 heads (`concatconv`, `gw`, `dpn.proj`) run in torch
(cuDNN); they are outside the hot path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .exactconv import ExactConvCache, patch_convs
from .encoder import FusedEncoder
from .backbone import Backbone
from .hotpath import HotPathConfig, HotPathPlan, PackedWeights


# ----------------------------------------------------------------------------------------------
# parameter containers (names dictated by the reference checkpoints)
# ----------------------------------------------------------------------------------------------
class _Params(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f"{type(self).__name__} only holds parameters; the computation runs in libnmrf_b200.so")


class MLP(_Params):
    """NMP.py:54-66: `layers.{i}` Linear stack."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))


class Mlp(_Params):
    """timm Mlp container: fc1, fc2."""

    def __init__(self, in_features, hidden_features, out_features=None):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features or in_features)


class _ProposalAttention(_Params):     # BasicAttention, NMP.py:70-88
    def __init__(self, dim, qk_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.q, self.k, self.v = nn.Linear(qk_dim, dim), nn.Linear(qk_dim, dim), nn.Linear(dim, dim)
        self.proj = nn.Linear(dim, dim)


class _WindowAttention(_Params):       # WindowAttention, NMP.py:155-183
    def __init__(self, dim, ws):
        super().__init__()
        self.relative_position_enc_table = nn.Parameter(torch.zeros((2 * ws - 1) ** 2, dim * 3))
        c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0) + (ws - 1)
        self.register_buffer("relative_position_index", rel[..., 0] * (2 * ws - 1) + rel[..., 1])


class _SwinBlock(_Params):             # SwinNMP, NMP.py:314-341
    def __init__(self, dim, qkv_dim, ws, mlp_ratio):
        super().__init__()
        self.qkv = nn.Linear(qkv_dim, 3 * dim)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _WindowAttention(dim, ws)
        self.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class _StripeAttention(_Params):       # CSWinAttention, NMP.py:402-427
    def __init__(self, dim):
        super().__init__()
        self.get_v = nn.Conv2d(dim, dim, 3, 1, 1, groups=dim, bias=False)


class _StripeBlock(_Params):           # CSWinNMP, NMP.py:510-542
    def __init__(self, dim, qk_dim, mlp_ratio):
        super().__init__()
        self.q, self.k, self.v = nn.Linear(qk_dim, dim), nn.Linear(qk_dim, dim), nn.Linear(dim, dim)
        self.norm1 = nn.LayerNorm(dim)
        self.proj = nn.Linear(dim, dim)
        self.attns = nn.ModuleList(_StripeAttention(dim // 2) for _ in range(2))
        self.mlp = Mlp(dim, int(dim * mlp_ratio), dim)
        self.norm2 = nn.LayerNorm(dim)


class PropagationLayer(_Params):       # NMP.py:903-917
    def __init__(self, dim, context_dim, mlp_ratio):
        super().__init__()
        self.nmp = _StripeBlock(dim, dim + context_dim, mlp_ratio)


class InferenceLayer(_Params):         # NMP.py:932-948
    def __init__(self, dim, ws, mlp_ratio):
        super().__init__()
        self.self_nmp = _ProposalAttention(dim, dim + 31)
        self.nmp = _SwinBlock(dim, dim + 31, ws, mlp_ratio)


class RefinementLayer(_Params):        # NMP.py:961-972
    def __init__(self, dim, ws, mlp_ratio):
        super().__init__()
        self.nmp = _SwinBlock(dim, dim + 31, ws, mlp_ratio)


class Propagation(_Params):            # NMP.py:603-616
    def __init__(self, dim, cost_group, layers):
        super().__init__()
        self.cost_encoder = nn.Sequential(nn.Linear(cost_group * 9, dim), nn.GELU(), nn.Linear(dim, dim))
        self.proj = nn.Linear(dim + 31, dim, bias=False)
        self.layers = layers
        self.norm = nn.LayerNorm(dim)


class MRFStack(_Params):               # Inference / Refinement, NMP.py:670-680
    def __init__(self, cost_group, dim, layers):
        super().__init__()
        self.ffn = Mlp(dim + cost_group, dim, dim)
        self.layers = layers
        self.norm = nn.LayerNorm(dim)


def _conv_head(cin, cout):
    """conv3x3 -> InstanceNorm -> ReLU -> conv1x1, no biases (NMRF.py:56-65, DPN.py:45-49)."""
    return nn.Sequential(nn.Conv2d(cin, 128, 3, 1, 1, bias=False), nn.InstanceNorm2d(128), nn.ReLU(inplace=True),
                         nn.Conv2d(128, cout, 1, 1, 0, bias=False))


class DPN(_Params):
    """Disparity proposal network container (reference nmrf/models/DPN.py:23-69)."""

    def __init__(self, cost_group=4, num_proposals=4, feat_dim=256, context_dim=64, num_prop_layers=5,
                 prop_embed_dim=128, mlp_ratio=4, split_size=1, prop_n_heads=4, **_unused):
        super().__init__()
        if split_size != 1 or prop_n_heads != 4 or prop_embed_dim != 128 or context_dim != 64:
            raise NotImplementedError("libnmrf_b200 supports SPLIT_SIZE=1, 4 heads, embed 128, context 64 (the defaults)")
        self.mlp = nn.Sequential(nn.Conv1d(cost_group, 8, 5, 1, 2), nn.ReLU(inplace=True),
                                 nn.Conv1d(8, 16, 5, 1, 2), nn.ReLU(inplace=True), nn.Conv1d(16, 1, 5, 1, 2))
        self.eps = 1e-3
        self.num_proposals, self.cost_group = num_proposals, cost_group
        self.proj = _conv_head(feat_dim, context_dim)
        layers = nn.ModuleList(PropagationLayer(prop_embed_dim, context_dim, mlp_ratio) for _ in range(num_prop_layers))
        self.propagation = Propagation(prop_embed_dim, cost_group, layers)
        self.prop_head = MLP(prop_embed_dim, prop_embed_dim, 1, 3)
        # DPN initialises ITSELF (DPN.py:67-69): its own _init_weights, then the last prop_head layer zeroed
        self.apply(_init_weights)
        nn.init.constant_(self.prop_head.layers[-1].weight, 0.0)
        nn.init.constant_(self.prop_head.layers[-1].bias, 0.0)


def _init_weights(m):
    """`_init_weights` of the reference (NMRF.py:154-165, DPN.py:90-105): convs kaiming_normal(fan_out) (Conv1d: zero bias),
    Linear trunc_normal(.02) / zero bias, LayerNorm / InstanceNorm 1 / 0."""
    if isinstance(m, (nn.Conv1d, nn.Conv2d)):
        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        if isinstance(m, nn.Conv1d) and m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=0.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, (nn.LayerNorm, nn.InstanceNorm2d)):
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
        if m.weight is not None:
            nn.init.constant_(m.weight, 1.0)


class NMRF(nn.Module):
    """`model(sample)` with sample = {'img1','img2'}: float32 [B,3,H,W] in 0..255 (any device).
    Returns {'disp' [B,H,W], 'disp_pred' [B,Hp,Wp], 'proposal' [B,h8*w8,K], 'initial_proposal',
    'prob' [B*h8*w8, D]}  (reference NMRF.py:189-262).  Inference only."""

    def __init__(self, backbone, dpn, num_proposals, max_disp, num_infer_layers, num_refine_layers,
                 infer_embed_dim=128, infer_n_heads=4, mlp_ratio=4, window_size=6, refine_window_size=4,
                 with_refinement=True, attn_drop=0., proj_drop=0., drop_path=0., dropout=0.,
                 return_intermediate=False, normalize_before=True, activation="gelu", aux_loss=False,
                 divis_by=8, compat=True):
        super().__init__()
        if infer_embed_dim != 128 or infer_n_heads != 4:
            raise NotImplementedError("libnmrf_b200 is built for INFER_EMBED_DIM=128, INFER_N_HEADS=4 (the defaults)")
        if not normalize_before or activation != "gelu" or not with_refinement:
            raise NotImplementedError("only NORMALIZE_BEFORE=True, GELU, WITH_REFINEMENT=True (the released configs)")
        self.num_proposals, self.max_disp, self.divis_by = num_proposals, max_disp, divis_by
        self.window_size, self.refine_window_size = window_size, refine_window_size
        feat_dim = backbone.output_dim
        self.concatconv = _conv_head(feat_dim, 64)
        self.gw = _conv_head(feat_dim, 256)
        dim = infer_embed_dim
        self.inference = MRFStack(32, dim, nn.ModuleList(InferenceLayer(dim, window_size, mlp_ratio)
                                                         for _ in range(num_infer_layers)))
        self.infer_head = MLP(dim, dim, 8 * 8, 3)
        self.infer_score_head = nn.Linear(dim, 8 * 8)
        # Same order as the reference (NMRF.py:86-87): `apply(_init_weights)` covers only what exists at this point --
        # concatconv, gw, inference, infer_head, infer_score_head.  The refinement stack keeps torch's default init, DPN has
        # initialised itself, and a passed-in encoder (e.g. a pretrained Swin backbone) is never touched.
        self.apply(_init_weights)
        self.refinement = MRFStack(32, dim, nn.ModuleList(RefinementLayer(dim, refine_window_size, mlp_ratio)
                                                          for _ in range(num_refine_layers)))
        self.refine_head = MLP(dim, dim, 4 * 4, 3)
        self.dpn = dpn
        self.compat = compat
        if compat:
            self.backbone = backbone
        else:
            self.image_encoder = backbone
        self.register_buffer("device_indicator_tensor", torch.empty(0))
        self._packed = None
        self._plans = {}
        self.cudnn_benchmark = False      # module path only (foreign encoders): let cuDNN autotune its convolutions
        # out-of-path torch convolutions: "3xtf32" = fp32-accurate on TF32 tensor cores (exactconv.py),
        # "fp32" = cuDNN with TF32 disabled (slow on B200), "tf32" = cuDNN default (fast, breaks parity)
        self.conv_mode = "3xtf32"
        self._conv_cache = ExactConvCache()
        # fused NHWC execution of the stock ResNet-style backbone + conv heads (encoder.py); other backbones fall back to
        # running their torch modules with the 3xTF32 convolution wrapper
        self.fused_encoder = True
        self._encoder = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # ---- plumbing ---------------------------------------------------------------------------
    @property
    def device(self):
        return self.device_indicator_tensor.device

    def invalidate(self):
        """Drop packed weights and launch plans (call after changing parameters in place)."""
        self._packed = None
        self._plans = {}
        self._encoder = None
        if hasattr(self, "_conv_cache"):
            self._conv_cache.clear()

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def hot_path_config(self):
        return HotPathConfig(max_disp=self.max_disp, num_proposals=self.num_proposals, cost_group=self.dpn.cost_group,
                             window_size=self.window_size, refine_window_size=self.refine_window_size,
                             num_prop_layers=len(self.dpn.propagation.layers),
                             num_infer_layers=len(self.inference.layers), num_refine_layers=len(self.refinement.layers),
                             eps=self.dpn.eps)

    def plan_for(self, B, C, h8, w8, H, W):
        if self.device.type != "cuda":
            raise RuntimeError("nmrf_b200.NMRF runs on CUDA only (there is no CPU path); call .cuda() first")
        if self._packed is None:
            self._packed = PackedWeights(self.state_dict(), self.hot_path_config())
        key = (B, C, h8, w8, H, W)
        if key not in self._plans:
            self._plans[key] = HotPathPlan(self._packed, self.hot_path_config(), B, C, h8, w8, H, W, self.device)
        return self._plans[key]

    # ---- forward ------------------------------------------------------------------------------
    def extract_feature(self, img1, img2):
        """NMRF.py:172-187: backbone on cat(left,right); returns ([f1@1/8, f1@1/4], [f2@1/8, f2@1/4])."""
        enc = self.backbone if self.compat else self.image_encoder
        feats = enc(torch.cat((img1, img2), dim=0))[::-1]
        return [f.chunk(2, 0)[0] for f in feats], [f.chunk(2, 0)[1] for f in feats]

    @torch.no_grad()
    def forward(self, sample):
        if self.training:
            raise RuntimeError("nmrf_b200.NMRF implements the inference path only; call .eval()")
        img1 = sample["img1"].to(self.device, torch.float32, non_blocking=True)
        img2 = sample["img2"].to(self.device, torch.float32, non_blocking=True)
        return self.forward_device(img1, img2)

    @torch.no_grad()
    def forward_device(self, img1, img2):
        if self.device.type != "cuda":
            raise RuntimeError("nmrf_b200.NMRF runs on CUDA only (there is no CPU path); call .cuda() first")
        # everything below launches on the MODEL's device: weight packing, the encoder and the plan make it current
        # (the library configures its kernels per device and launches on the current one)
        with torch.cuda.device(self.device):
            return self._forward_device(img1, img2)

    def _forward_device(self, img1, img2):
        B, _, H, W = img1.shape
        d = self.divis_by                                          # frame_utils.py:264-269 ('proposal' mode)
        pad_h, pad_w = (((H // d) + 1) * d - H) % d, (((W // d) + 1) * d - W) % d
        if (self.fused_encoder and self.compat and self.conv_mode == "3xtf32" and type(self.backbone) is Backbone
                and isinstance(self.backbone.norm1, nn.InstanceNorm2d)):      # the fused encoder computes InstanceNorm only
            Hp_, Wp_ = H + pad_h, W + pad_w                        # the replicate pad happens inside nmrf_image_prep
            h8, w8 = Hp_ // 8, Wp_ // 8
            plan = self.plan_for(B, self.backbone.output_dim, h8, w8, H, W)
            if self._encoder is None:
                self._encoder = FusedEncoder(self)
            self._encoder.run(img1, img2, plan, (Hp_, Wp_))
            plan.run()
            return self._outputs(plan, B)
        if pad_h or pad_w:
            img1 = F.pad(img1, [0, pad_w, 0, pad_h], mode="replicate")
            img2 = F.pad(img2, [0, pad_w, 0, pad_h], mode="replicate")
        img1 = img1.contiguous(memory_format=torch.channels_last)
        img2 = img2.contiguous(memory_format=torch.channels_last)
        # exact-fp32 convolutions: cuDNN's default TF32 moves the features by ~5e-4 relative, which flips
        # top-K / argmax decisions downstream (EPE 0.2-0.4 px measured) -- same reason as DESIGN.md §3
        self._conv_cache.enabled = self.conv_mode == "3xtf32"
        if self._conv_cache.enabled:
            for m in (self.backbone if self.compat else self.image_encoder, self.concatconv, self.gw, self.dpn.proj):
                patch_convs(m, self._conv_cache)
        with torch.backends.cudnn.flags(enabled=True, benchmark=self.cudnn_benchmark, allow_tf32=(self.conv_mode == "tf32")):
            f1, f2 = self.extract_feature(img1, img2)                  # [1/8, 1/4]
            C, h8, w8 = f1[0].shape[1:]
            plan = self.plan_for(B, C, h8, w8, H, W)
            nhwc = lambda t: t.permute(0, 2, 3, 1)
            plan.f1_8.copy_(nhwc(f1[0]))
            plan.f2_8.copy_(nhwc(f2[0]))
            plan.context.copy_(nhwc(self.dpn.proj(f1[0])))
            for s, (cc, gw) in enumerate(((plan.cc8, plan.gw8), (plan.cc4, plan.gw4))):
                both = torch.cat((f1[s], f2[s]), 0)                    # InstanceNorm is per-sample: batching is exact
                c = nhwc(self.concatconv(both))
                g = nhwc(self.gw(both))
                cc[0].copy_(c[:B]); cc[1].copy_(c[B:])
                gw[0].copy_(g[:B]); gw[1].copy_(g[B:])
        plan.run()
        return self._outputs(plan, B)

    def _outputs(self, plan, B):
        K = self.num_proposals
        return {
            "proposal": plan.labels.reshape(B, -1, K).clone(),
            "prob": plan.prob.clone(),
            "initial_proposal": plan.seeds.reshape(B, -1, K).float(),
            "disp": plan.disp.clone(),
            "disp_pred": plan.disp_pred.clone(),
        }

"""Drop-in mirror of the reference's models/hand_net.py heads over the scat_b200 CUDA kernels.

``EncoderTransformer(opt, mean_params)`` keeps the reference's constructor, option fields
(``vit_heads, pl_reg, iteration, pos_embed, mask_rate``), attribute names, return arity and state_dict keys
(hand_net.py:315-398) so it can replace ``models.hand_net.EncoderTransformer`` behind the existing
train.py / eval.py.  The ResNet-50 backbone stays a PyTorch/cuDNN feature producer (north_star); everything
after it -- 1x1 conv, positional encoding, token masking, the 3-layer narrowing transformer, the
autoregressive regressor, the root-relative step and the path-length VJP -- runs in hand-written sm_100a
kernels through the C ABI.  There is no eager fallback.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch
import torch.nn as nn

from . import functional as SF
from . import vision_transformer, vision_transformer_attn


def _to_gpu(t):
    """mean_params.clone().cuda() of hand_net.py:321.  Construction (state_dict handling, key checks) also
    works on a CPU-only host; every compute entry then raises, there is no CPU path."""
    t = t.clone().float()
    return t.cuda() if torch.cuda.is_available() else t


class PositionalEncoding(nn.Module):
    """hand_net.py:61-77.  Buffer ``pe[1,max_len,d_model]``; the add itself is fused into the conv epilogue."""

    def __init__(self, d_model, dropout=0.0, max_len=5000):
        super().__init__()
        if dropout != 0.0:
            raise ValueError("the reference path uses dropout=0.0 (hand_net.py:343)")
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        raise RuntimeError("PositionalEncoding is applied inside the fused conv+PE+mask kernel; "
                           "call EncoderTransformer instead")


def _resnet50_backbone():
    """ResNet-50 with the reference's output signature (resnet.py:142-162): returns
    (relu(fc1(relu(avgpool))), x1, x2, x3, x4).  Plain PyTorch/cuDNN; out of the kernel scope."""
    import torchvision

    class _ResNet(torchvision.models.ResNet):
        def __init__(self):
            super().__init__(torchvision.models.resnet.Bottleneck, [3, 4, 6, 3])
            del self.fc
            self.fc1 = nn.Linear(2048, 1024)

        def forward(self, x):
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
            x1 = self.layer1(x)
            x2 = self.layer2(x1)
            x3 = self.layer3(x2)
            x4 = self.layer4(x3)
            x = torch.relu(torch.flatten(self.avgpool(x4), 1))
            return torch.relu(self.fc1(x)), x1, x2, x3, x4

    return _ResNet()


def get_model(arch):
    """hand_net.py:20-25 (pretrained weights cannot be downloaded offline; random init as in resnet.py:118-123)."""
    if arch == "resnet50":
        return _resnet50_backbone()
    raise ValueError("Invalid Backbone Architecture")


class EncoderTransformer(nn.Module):
    """reg_transformer head, hand_net.py:315-398.  ``precision`` selects the GEMM arithmetic:
    "tf32" (tcgen05 kind::tf32, default), "bf16", or "fp32" (CUDA-core parity mode)."""

    def __init__(self, opt, mean_params, precision: str = "tf32", backbone: nn.Module | None = None):
        super().__init__()
        self.mean_params = _to_gpu(mean_params)        # :321 plain attribute, not a buffer
        heads = opt.vit_heads
        self.pl = opt.pl_reg
        self.full_content = 21
        self.conv1x1_channel_reduction = nn.Conv2d(512, 21, 1, 1, 0, bias=False)
        self.transformer = vision_transformer.Transformer(dim=784, depth=3, heads=heads, dim_head=64, mlp_dim=392,
                                                          dropout=0.0)
        self.main_encoder = backbone if backbone is not None else get_model("resnet50")
        self.iteration = opt.iteration
        self.pos_embed = opt.pos_embed
        print("Position Encoding open" if self.pos_embed is True else "Position Encoding close")
        self.positionalEncoding = PositionalEncoding(784, max_len=21)
        self.mask_token = nn.Parameter(torch.randn(1, 1, 784))
        self.mask_rate = opt.mask_rate
        self.regressor = nn.Linear(1024 + 66, 66)
        self.precision = precision
        self.last_mask = []          # indices drawn by the most recent forward (for inspection/tests)
        self._mask_dev = None

    # -- C-ABI parameter order (include/scat_b200.h) ---------------------------------------------
    def head_parameters(self):
        return ([self.mask_token, self.conv1x1_channel_reduction.weight] + self.transformer.ordered_parameters()
                + [self.regressor.weight, self.regressor.bias])

    def _draw_mask(self):
        # hand_net.py:369-372: exactly one random.shuffle per forward when 0.1 <= mask_rate <= 0.9,
        # also in eval mode; the same indices for every sample
        if self.mask_rate >= 0.1 and self.mask_rate <= 0.9:
            masked = list(range(self.full_content))
            random.shuffle(masked)
            return masked[: int(self.mask_rate * self.full_content)]
        return []

    def config(self, n_masked: int, x2_dtype: str = "fp32") -> SF.HeadConfig:
        return SF.HeadConfig(heads=self.transformer.heads, iteration=int(self.iteration),
                             pos_embed=bool(self.pos_embed), n_masked=n_masked, pl_reg=bool(self.pl),
                             precision=self.precision, x2_dtype=x2_dtype)

    def forward_features(self, main_feat, x2, mask_idx=None):
        """The head proper, from the backbone seam tensors (main_feat[B,1024] fp32, x2[B,512,28,28] fp32 or -- from a
        bf16 / autocast backbone -- bfloat16, in which case x2.grad is delivered as bfloat16 too)."""
        if self.pl and not torch.is_grad_enabled():
            # same failure as the reference under no_grad (autograd.grad at hand_net.py:396)
            raise RuntimeError("element 0 of tensors does not require grad and does not have a grad_fn")
        masked = self._draw_mask() if mask_idx is None else list(mask_idx)
        self.last_mask = masked
        dev = x2.device
        if len(masked):
            # host-supplied index tensor (hand_net.py:370-372): a small pageable host-to-device copy, stream-ordered with
            # the kernels that read it.  (It synchronises the stream; the graph-replayed HeadTrainStep stages the indices
            # through pinned memory instead.)
            mask_dev = torch.tensor(masked, dtype=torch.int32, device=dev)
        else:
            mask_dev = None
        pe = self.positionalEncoding.pe[0] if self.pos_embed else None
        mean = self.mean_params.reshape(-1)
        if main_feat.dtype != torch.float32:          # an autocast backbone's fc1 output: the regressor is always fp32
            main_feat = main_feat.float()
        return SF.HeadFunction.apply(self.config(len(masked)), mask_dev, mean, pe, x2, main_feat,
                                     *self.head_parameters())

    def forward(self, main_input):
        main_feat, x1, x2, x3, x4 = self.main_encoder(main_input)          # :356 (cuDNN)
        return self.forward_features(main_feat, x2)


class EncoderTransformerCoarse(nn.Module):
    """reg_transformer_coarse head, hand_net.py:216-311: the attention-visualisation variant that eval.py:788-834 runs.
    Same conv + positional encoding + masking front end; the transformer of models/vision_transformer_attn.py (attention
    on the raw tokens, LayerNorm on the attention branch's output, then the residual; heads fixed at 8, :236); camera =
    Linear(1027 -> 3) applied once (:259,298); returns (pred_params[B,66], feat_visual[B,21,28,28], attn[B,8,21,21]).
    Inference path: the call must run under torch.no_grad() -- where the reference itself cannot return the path-length
    term (autograd.grad at :309 fails under no_grad), so ``opt.pl_reg`` must be False."""

    def __init__(self, opt, mean_params, precision: str = "tf32", backbone: nn.Module | None = None):
        super().__init__()
        if precision not in ("fp32", "tf32"):
            raise ValueError("EncoderTransformerCoarse: precision 'fp32' or 'tf32'")
        self.mean_params = _to_gpu(mean_params)
        self.pl = opt.pl_reg
        self.full_content = 21
        self.conv1x1_channel_reduction = nn.Conv2d(512, 21, 1, 1, 0, bias=False)
        self.transformer = vision_transformer_attn.Transformer(dim=784, depth=3, heads=8, dim_head=64, mlp_dim=392,
                                                               dropout=0.0)
        self.main_encoder = backbone if backbone is not None else get_model("resnet50")
        self.iteration = opt.iteration
        self.pos_embed = opt.pos_embed
        print("Position Encoding open" if self.pos_embed is True else "Position Encoding close")
        self.positionalEncoding = PositionalEncoding(784, max_len=21)
        self.mask_token = nn.Parameter(torch.randn(1, 1, 784))
        self.mask_rate = opt.mask_rate
        self.regressor = nn.Linear(1024 + 3, 3)
        self.precision = precision
        self.last_mask = []

    def head_parameters(self):
        return ([self.mask_token, self.conv1x1_channel_reduction.weight] + self.transformer.ordered_parameters()
                + [self.regressor.weight, self.regressor.bias])

    _draw_mask = EncoderTransformer._draw_mask

    def forward_features(self, main_feat, x2, mask_idx=None):
        if self.pl:
            # hand_net.py:308-309: the path-length term is an autograd.grad call; this is the no_grad inference path
            raise RuntimeError("element 0 of tensors does not require grad and does not have a grad_fn")
        if torch.is_grad_enabled() and (x2.requires_grad or main_feat.requires_grad
                                        or any(p.requires_grad for p in self.head_parameters())):
            raise RuntimeError("scat_b200.EncoderTransformerCoarse is an inference path; wrap the call in torch.no_grad()")
        masked = self._draw_mask() if mask_idx is None else list(mask_idx)
        self.last_mask = masked
        mask_dev = torch.tensor(masked, dtype=torch.int32, device=x2.device) if len(masked) else None
        pe = self.positionalEncoding.pe[0] if self.pos_embed else None
        cfg = SF.HeadConfig(heads=8, iteration=0, pos_embed=bool(self.pos_embed), n_masked=len(masked), pl_reg=False,
                            precision=self.precision)
        return SF.coarse_forward(cfg, mask_dev, self.mean_params.reshape(-1), pe, x2, main_feat,
                                 [p.detach() for p in self.head_parameters()])

    def forward(self, main_input):
        main_feat, x1, x2, x3, x4 = self.main_encoder(main_input)          # :262
        return self.forward_features(main_feat, x2)


class H3DWEncoder(nn.Module):
    """FrankMocap baseline head, hand_net.py:28-58: feat = relu(fc2(relu(main_feat))); 3 x (pred += Linear(1085->61)).
    fc2 runs on the GEMM kernel, the loop on the fused regressor kernel (root_relative off).  Inference only."""

    def __init__(self, opt, mean_params, backbone: nn.Module | None = None):
        super().__init__()
        self.mean_params = _to_gpu(mean_params)
        self.total_params_dim = 61
        self.feat_encoder = nn.Sequential(nn.ReLU(inplace=False), nn.Linear(1024, 1024), nn.ReLU(inplace=False))
        self.regressor = nn.Sequential(nn.Linear(1024 + self.total_params_dim, self.total_params_dim))
        self.main_encoder = backbone if backbone is not None else get_model("resnet50")

    def forward_features(self, main_feat):
        if torch.is_grad_enabled() and main_feat.requires_grad:
            raise RuntimeError("scat_b200.H3DWEncoder is an inference path; wrap the call in torch.no_grad()")
        fc2 = self.feat_encoder[1]
        x = torch.relu(main_feat)          # elementwise glue on the seam tensor
        feat = SF.gemm(x, fc2.weight.detach(), epilogue="bias", bias=fc2.bias.detach(), precision="fp32")
        feat = torch.relu_(feat)
        reg = self.regressor[0]
        pred = SF.regressor_fwd(feat, None, self.mean_params.reshape(-1), reg.weight.detach(), reg.bias.detach(),
                                iteration=3, root_relative=False)
        return feat, pred

    def forward(self, main_input):
        main_feat, _, _, _, _ = self.main_encoder(main_input)
        return self.forward_features(main_feat)

"""Parameter containers with the reference's ``state_dict`` layout (nets/layers.py:59-254) and the two free
functions the callers import.  The modules below never run a PyTorch forward: they only hold weights under the
reference's key names so that ``load_state_dict(strict=True)`` of reference checkpoints works; the arithmetic is
done by ``imp_release_b200.engine`` on packed copies of these weights."""
from __future__ import annotations

import torch
import torch.nn as nn

SHARING_LAYERS = [False, False] * 2 + [False, False, True, True] * 21   # nets/gms.py:17, nets/adgm.py:18


def normalize_keypoints(kpts: torch.Tensor, image_shape) -> torch.Tensor:
    """nets/layers.py:49-56.  Boundary helper (2N floats); called by eval/matching.py:24 on the caller's device."""
    _, _, height, width = image_shape
    # same fp32 arithmetic as the reference (centre = size/2, scale = fp32(max(W,H)) * 0.7) but with host scalars, so
    # that no host->device copy happens here (keeps the forward CUDA-graph capturable)
    scale = float(torch.tensor(float(max(width, height)), dtype=torch.float32) * 0.7)
    out = kpts.clone()
    out[..., 0].sub_(float(width) / 2)
    out[..., 1].sub_(float(height) / 2)
    return out.div_(scale)


def _conv(cin: int, cout: int) -> nn.Conv1d:
    return nn.Conv1d(cin, cout, kernel_size=1, bias=True)


def mlp_params(channels) -> nn.Sequential:
    """Same child indices as the reference MLP() (conv at 0, 3, 6, ...; norm/activation own no parameters)."""
    mods = []
    for i in range(1, len(channels)):
        mods.append(_conv(channels[i - 1], channels[i]))
        if i < len(channels) - 1:
            mods += [nn.Identity(), nn.Identity()]
    return nn.Sequential(*mods)


class KeypointEncoderParams(nn.Module):        # keys: encoder.{0,3,6,9,12}.{weight,bias}
    def __init__(self, feature_dim: int, layers):
        super().__init__()
        self.encoder = mlp_params([3] + list(layers) + [feature_dim])
        nn.init.constant_(self.encoder[-1].bias, 0.0)


class AttentionParams(nn.Module):              # keys: merge.*, proj.{0,1,2}.*
    def __init__(self, d_model: int):
        super().__init__()
        self.merge = _conv(d_model, d_model)
        self.proj = nn.ModuleList([_conv(d_model, d_model) for _ in range(3)])


class PropagationParams(nn.Module):
    """Non-sharing: attn.*, mlp.{0,3}.*   Sharing: proj.*, merge.*, mlp.{0,3}.*  (nets/layers.py:182-198)."""

    def __init__(self, feature_dim: int, sharing: bool):
        super().__init__()
        self.sharing_attention = sharing
        if not sharing:
            self.attn = AttentionParams(feature_dim)
        else:
            self.proj = _conv(feature_dim, feature_dim)
            self.merge = _conv(feature_dim, feature_dim)
        self.mlp = mlp_params([2 * feature_dim, 2 * feature_dim, feature_dim])
        nn.init.constant_(self.mlp[-1].bias, 0.0)


class GNNParams(nn.Module):                    # keys: layers.{i}.*
    def __init__(self, feature_dim: int, layer_names, sharing_layers=None):
        super().__init__()
        n = len(layer_names)
        self.sharing_layers = list(sharing_layers[:n]) if sharing_layers is not None else [False] * n
        self.layers = nn.ModuleList([PropagationParams(feature_dim, self.sharing_layers[i]) for i in range(n)])
        self.names = list(layer_names)

"""Weights of the two hot-path networks: seeded synthetic state_dicts in the reference's own
key layout, Lightning-checkpoint loading, and BatchNorm folding.

No checkpoint can be downloaded in this environment, so tests and bench use
`synth_deflowpp_state_dict(seed)`; a real `seflowpp_best.ckpt` goes through
`load_deflowpp_checkpoint` (keys carry a `model.` prefix: OSF/src/models/basic/__init__.py:10-16).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

# (name, cin, cout, stride) of the 16 ConvWithNorms of the shared per-frame encoder
# (OSF/src/models/basic/unet.py:110-127)
ENCODER_LAYERS: List[Tuple[str, int, int, int]] = (
    [("encoder_step_1.0", 32, 64, 2)] + [(f"encoder_step_1.{i}", 64, 64, 1) for i in range(1, 4)] +
    [("encoder_step_2.0", 64, 128, 2)] + [(f"encoder_step_2.{i}", 128, 128, 1) for i in range(1, 6)] +
    [("encoder_step_3.0", 128, 256, 2)] + [(f"encoder_step_3.{i}", 256, 256, 1) for i in range(1, 6)])
# UpsampleSkip(skip, latent, out) blocks (unet.py:18-35, 128-130)
DECODER_BLOCKS = [("decoder_step1", 768, 384, 384), ("decoder_step2", 384, 192, 192),
                  ("decoder_step3", 192, 96, 96)]


def _xavier(rng: np.random.Generator, shape, gain: float = 1.0) -> torch.Tensor:
    receptive = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[1] * receptive, shape[0] * receptive
    a = gain * math.sqrt(6.0 / (fan_in + fan_out))
    return torch.from_numpy(rng.uniform(-a, a, size=shape).astype(np.float32))


def _bn(rng, c, prefix, sd):
    sd[prefix + ".weight"] = torch.from_numpy(rng.uniform(0.8, 1.2, c).astype(np.float32))
    sd[prefix + ".bias"] = torch.from_numpy(rng.uniform(-0.1, 0.1, c).astype(np.float32))
    sd[prefix + ".running_mean"] = torch.from_numpy(rng.normal(0, 0.05, c).astype(np.float32))
    sd[prefix + ".running_var"] = torch.from_numpy(rng.uniform(0.6, 1.4, c).astype(np.float32))
    sd[prefix + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)


def synth_deflowpp_state_dict(seed: int = 0, act_gain: float = 1.5, dec_gain: float = 0.6,
                              flow_gain: float = 0.12) -> Dict[str, torch.Tensor]:
    """Seeded DeFlowPP weights with the 156 state_dict entries of the reference class
    (OSF/src/models/deflow.py:90-113).  Xavier-uniform weights (the reference's weights_init,
    OSF/src/utils/mics.py:98-105) with a gain on the GELU layers that keeps activations O(1) through
    the 16-layer encoder, small non-zero biases and non-trivial BatchNorm running statistics so
    that BN folding is really exercised.  dec_gain / flow_gain keep the backbone output O(1-10) and
    the predicted flow in the physical range of a 10 Hz sweep pair (|flow| up to a few metres)."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    p = "embedder.feature_net.pfn_layers.0"
    sd[p + ".0.weight"] = _xavier(rng, (32, 9), 1.5)
    _bn(rng, 32, p + ".1", sd)
    for name, cin, cout, _ in ENCODER_LAYERS:
        q = "backbone." + name
        sd[q + ".conv.weight"] = _xavier(rng, (cout, cin, 3, 3), act_gain)
        sd[q + ".conv.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, cout).astype(np.float32))
        _bn(rng, cout, q + ".batchnorm", sd)
    for name, skip, latent, out in DECODER_BLOCKS:
        q = "backbone." + name
        for sub, shape in ((".u1_u2.0", (latent, skip, 1, 1)), (".u3", (latent, latent, 1, 1)),
                           (".u4_u5.0", (out, 2 * latent, 3, 3)), (".u4_u5.1", (out, out, 3, 3))):
            sd[q + sub + ".weight"] = _xavier(rng, shape, dec_gain if sub.startswith(".u4") else 1.0)
            sd[q + sub + ".bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, shape[0]).astype(np.float32))
    sd["backbone.decoder_step4.weight"] = _xavier(rng, (96, 96, 3, 3), 1.0)
    sd["backbone.decoder_step4.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, 96).astype(np.float32))
    sd["head.offset_encoder.weight"] = _xavier(rng, (96, 3), 1.0)
    sd["head.offset_encoder.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, 96).astype(np.float32))
    for g in ("convz", "convr", "convq"):
        sd[f"head.gru.{g}.weight"] = _xavier(rng, (192, 288, 1), 1.0)
        sd[f"head.gru.{g}.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, 192).astype(np.float32))
    sd["head.decoder.0.weight"] = _xavier(rng, (48, 288), 1.0)
    sd["head.decoder.0.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, 48).astype(np.float32))
    sd["head.decoder.2.weight"] = _xavier(rng, (3, 48), flow_gain)
    sd["head.decoder.2.bias"] = torch.from_numpy(rng.uniform(-0.05, 0.05, 3).astype(np.float32))
    return sd


def load_deflowpp_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """Lightning `.ckpt` -> DeFlowPP state_dict (strip the `model.` prefix)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
    out = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
    return out if out else dict(sd)


def fold_bn(weight: torch.Tensor, bias, bn_w, bn_b, mean, var, eps: float):
    """Fold eval-mode BatchNorm into the preceding linear map (double precision, rounded once):
    y = (W x + b - mean) / sqrt(var + eps) * g + beta  ==  W' x + b'."""
    w = weight.double()
    b = torch.zeros(w.shape[0], dtype=torch.float64) if bias is None else bias.double()
    s = bn_w.double() / torch.sqrt(var.double() + eps)
    w2 = w * s.view(-1, *([1] * (w.dim() - 1)))
    b2 = (b - mean.double()) * s + bn_b.double()
    return w2.float(), b2.float()


def synth_neural_prior_state_dict(seed: int = 0, filter_size: int = 128, layer_size: int = 8
                                  ) -> Dict[str, torch.Tensor]:
    """Seeded Neural_Prior weights in the reference layout (nsfp_module.py:7-26): hidden layers
    `nn_layers.{0,2,..}.0.{weight,bias}` with torch's default Linear init (kaiming_uniform(a=sqrt 5)
    => U(+-1/sqrt(fan_in)) for both), last layer xavier_uniform + zero bias -- which is what
    `init_weights` really does (it only matches the bare last Linear: nsfp_module.py:35-39)."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    dims = [3] + [filter_size] * layer_size
    for i in range(layer_size):
        bound = 1.0 / math.sqrt(dims[i])
        sd[f"nn_layers.{2 * i}.0.weight"] = torch.from_numpy(
            rng.uniform(-bound, bound, (dims[i + 1], dims[i])).astype(np.float32))
        sd[f"nn_layers.{2 * i}.0.bias"] = torch.from_numpy(
            rng.uniform(-bound, bound, dims[i + 1]).astype(np.float32))
    sd[f"nn_layers.{2 * layer_size}.weight"] = _xavier(rng, (3, filter_size), 1.0)
    sd[f"nn_layers.{2 * layer_size}.bias"] = torch.zeros(3)
    return sd

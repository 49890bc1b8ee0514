"""`--modality_bridger` of predictions_runner.py:182-184,225-227 on the sm_100a kernels.

The reference maps a CLIP image embedding into the text-embedding space with a supervised MLP before `clip_project`
(others/supervised_embedding_bridger.py:87-108: `MLP(640, 640, 640, 8)` = 8 `nn.Linear(640, 640)` with ReLU between them,
none after the last; weights in others/weights_modality_mapper.pt, loaded at :21-24).  `ModalityBridger` holds the same
parameters under the same names (`layers.{i}.weight/bias`, so that checkpoint loads strictly) and runs the stack as 8
tcgen05 GEMM launches with the bias + ReLU epilogue fused.  Inference only, like the reference's use of it.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import CapdecError


class ModalityBridger(nn.Module):
    """others/supervised_embedding_bridger.py:87-108 (`MLP(input_dim, hidden_dim, output_dim, num_layers)`)."""

    def __init__(self, input_dim: int = 640, hidden_dim: int = 640, output_dim: int = 640, num_layers: int = 8):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        for layer in self.layers:                       # :94-97 (`ones = True`): identity weights at construction
            nn.init.eye_(layer.weight)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise CapdecError("ModalityBridger runs on CUDA tensors only (capdec_b200 has no CPU path)")
        lead = x.shape[:-1]
        cur = x.detach().to(torch.float32).reshape(-1, x.shape[-1]).contiguous()
        for i, layer in enumerate(self.layers):
            if layer.weight.shape[1] % 4 or layer.weight.shape[0] % 4:
                raise CapdecError("ModalityBridger: layer widths must be multiples of 4 (16-byte TMA pitch)")
            out = torch.empty(cur.shape[0], layer.weight.shape[0], device=cur.device, dtype=torch.float32)
            act = ops.ACT_RELU if i < self.num_layers - 1 else ops.ACT_NONE          # :105
            ops.linear_fwd(cur, layer.weight, "linear", layer.bias, out, act=act)
            cur = out
        return cur.reshape(*lead, cur.shape[-1])


def get_map_to_text_space_using_modality_bridger(path: str = "others/weights_modality_mapper.pt", device="cuda"):
    """Same name and return value as others/supervised_embedding_bridger.py:19-29: a callable image-embedding ->
    text-space embedding, with the reference's trained weights loaded (strict)."""
    model = ModalityBridger(640, 640, 640, 8).to(device)
    model.load_state_dict(torch.load(path, map_location=device))
    model.eval()

    def map_to_text_space_using_modality_bridger(image_embedding):
        return model(image_embedding)

    return map_to_text_space_using_modality_bridger

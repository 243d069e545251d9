"""SiluAndMul on the F3 kernel (reference: layers/activation.py:6-13)."""
import torch
from torch import nn

from .. import ops


class SiluAndMul(nn.Module):

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.silu_and_mul(x)

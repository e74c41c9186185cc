import torch
from torch import nn


class SequenceINN(nn.Module):
    """Chain of invertible modules applied in order; returns (output, summed log|det J|)."""

    def __init__(self, *dims):
        super().__init__()
        self.shapes = [tuple(dims)]
        self.module_list = nn.ModuleList()

    def append(self, module_class, cond=None, cond_shape=None, **kwargs):
        if cond is not None:
            raise NotImplementedError("conditional blocks are outside the scoring path")
        module = module_class([self.shapes[-1]], **kwargs)
        self.module_list.append(module)
        self.shapes.append(self.shapes[-1])  # AllInOneBlock preserves the shape
        return module

    def forward(self, x_or_z, c=None, rev=False, jac=True):
        if rev:
            raise NotImplementedError("only the forward (density) direction is on the scoring path")
        log_det_jac = torch.zeros(x_or_z.shape[0], device=x_or_z.device, dtype=x_or_z.dtype)
        out = (x_or_z,)
        for module in self.module_list:
            out, j = module(out, jac=jac)
            log_det_jac = log_det_jac + j
        return out[0], log_det_jac

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F


class AllInOneBlock(nn.Module):
    """Affine coupling + global affine (ActNorm-like) + fixed channel permutation.

    forward(x): x1, x2 = split(x, [C - C//2, C//2]); a = 0.1 * subnet(x1); s = clamp * tanh(a[:, :C//2]);
    y = cat(x1, x2 * exp(s) + a[:, C//2:]); out = conv1x1(y * scale + offset, w_perm)
    log|det| = sum(s) + (H*W) * sum(log scale),  scale = 0.1 * softplus_{beta=0.5}(global_scale).
    """

    def __init__(self, dims_in, dims_c=(), subnet_constructor=None, affine_clamping=2.0, gin_block=False,
                 global_affine_init=1.0, global_affine_type="SOFTPLUS", permute_soft=False,
                 learned_householder_permutation=0, reverse_permutation=False):
        super().__init__()
        if dims_c or gin_block or permute_soft or learned_householder_permutation or reverse_permutation:
            raise NotImplementedError("option outside the scoring path")
        if global_affine_type != "SOFTPLUS":
            raise NotImplementedError("only the default SOFTPLUS global affine is used by the reference")
        channels = dims_in[0][0]
        self.input_rank = len(dims_in[0]) - 1
        self.sum_dims = tuple(range(1, 2 + self.input_rank))
        self.splits = [channels - channels // 2, channels // 2]
        self.permute_function = {0: F.linear, 1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[self.input_rank]
        self.in_channels = channels
        self.clamp = affine_clamping

        global_scale = 2.0 * np.log(np.exp(0.5 * 10.0 * global_affine_init) - 1)
        self.softplus = nn.Softplus(beta=0.5)
        ones = [1] * self.input_rank
        self.global_scale = nn.Parameter(torch.ones(1, channels, *ones) * float(global_scale))
        self.global_offset = nn.Parameter(torch.zeros(1, channels, *ones))

        w = np.zeros((channels, channels))
        for i, j in enumerate(np.random.permutation(channels)):
            w[i, j] = 1.0
        self.w_perm = nn.Parameter(torch.FloatTensor(w).view(channels, channels, *ones), requires_grad=False)
        self.w_perm_inv = nn.Parameter(torch.FloatTensor(w.T).view(channels, channels, *ones), requires_grad=False)

        if subnet_constructor is None:
            raise ValueError("subnet_constructor is required")
        self.subnet = subnet_constructor(self.splits[0], 2 * self.splits[1])
        self.last_jac = None

    def global_scale_activation(self, a):
        return 0.1 * self.softplus(a)

    def _permute(self, x):
        scale = self.global_scale_activation(self.global_scale)
        perm_log_jac = torch.sum(torch.log(scale))
        return self.permute_function(x * scale + self.global_offset, self.w_perm), perm_log_jac

    def _affine(self, x, a):
        a = a * 0.1
        ch = x.shape[1]
        sub_jac = self.clamp * torch.tanh(a[:, :ch])
        return x * torch.exp(sub_jac) + a[:, ch:], torch.sum(sub_jac, dim=self.sum_dims)

    def forward(self, x, c=(), rev=False, jac=True):
        if rev:
            raise NotImplementedError("only the forward (density) direction is on the scoring path")
        x1, x2 = torch.split(x[0], self.splits, dim=1)
        a1 = self.subnet(x1)
        x2, j2 = self._affine(x2, a1)
        x_out = torch.cat((x1, x2), 1)
        x_out, global_scaling_jac = self._permute(x_out)
        n_pixels = x_out[0, :1].numel()
        log_jac_det = j2 + n_pixels * global_scaling_jac
        return (x_out,), log_jac_det

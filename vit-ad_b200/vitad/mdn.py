"""Mixture-density ("GMM") head under the reference's names (src/classes/MixtureDensityNetwork.py).

``GaussianMixtureDensityNetwork.forward`` does not materialise the two [B,P,768,K] tensors the reference
returns; it hands back deferred handles in the same ``MdnReturn`` dataclass, and ``get_probability_map`` /
``log_likelihood`` consume them with the fused CUDA kernels.  The Gumbel noise the reference draws inside
``gumbel_softmax`` (:62) from torch's global generator is either an explicit tensor (``gumbel=``: what the parity
tests inject into both sides) or generated inside the mixing-weight kernel by a counter-based generator keyed by
``(seed, batch_index, token, mixture)`` (include/vitad.h: vitad_gmm_log_pi_seeded) — a batch then scores the same
whichever rank scores it and in whichever order.  With neither, a fresh seed is drawn from torch's CPU generator per
call, which reproduces the reference's stochastic validation (and follows ``torch.manual_seed``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import Tensor, nn

from . import _lib, custom_ops
from ._lib import check, lib

BIAS_FILL = 0.001  # src/util/HelperFunctions.py:7


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class DeferredMdnTensor:
    """Stands for pi / sigma / mu of one forward call without materialising them."""

    def __init__(self, head: "GaussianMixtureDensityNetwork", x: Tensor, which: str):
        self.head, self.x, self.which = head, x, which

    @property
    def shape(self):
        B, P, D = self.x.shape
        K = self.head.num_gaussians
        return (B, P, K) if self.which == "pi" else (B, P, D, K)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]


@dataclass
class MdnReturn:
    """Same fields as the reference dataclass (MixtureDensityNetwork.py:26-32)."""

    pi: Tensor | DeferredMdnTensor
    sigma: Tensor | DeferredMdnTensor
    mu: Tensor | DeferredMdnTensor


class GaussianMixtureDensityNetwork(nn.Module):
    """Drop-in for MixtureDensityNetwork.py:105-171: same constructor, parameters and state_dict keys
    (pi/sigma/mu .weight/.bias), same initialisation."""

    def __init__(self, input_dim: int, output_dim: int, num_gaussians: int, cluster_centers: Tensor | None = None):
        super().__init__()
        if input_dim != 768 or output_dim != input_dim:
            raise ValueError("vitad GMM head supports input_dim == output_dim == 768 (DeiT / EsViT features)")
        _lib.gmm_plan(num_gaussians)  # raises for an unsupported K
        self.elu = nn.ELU()
        self.pi = nn.Linear(input_dim, num_gaussians)
        nn.init.xavier_normal_(self.pi.weight)
        self.sigma = nn.Linear(input_dim, input_dim * num_gaussians)
        nn.init.xavier_normal_(self.sigma.weight)
        self.mu = nn.Linear(input_dim, input_dim * num_gaussians)
        nn.init.xavier_normal_(self.mu.weight)
        if cluster_centers is not None:
            with torch.no_grad():
                for i, bias in enumerate(cluster_centers):
                    self.mu.bias[i] = bias
        else:
            self.mu.bias.data.fill_(BIAS_FILL)
        self.out_dim = output_dim
        self.num_gaussians = num_gaussians
        self._packed = None
        self._handle = custom_ops.register_module(self)

    def _apply(self, fn, recurse=True):
        self._packed = None
        return super()._apply(fn, recurse)

    def load_state_dict(self, state_dict, strict=True, **kw):
        self._packed = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _pack(self, device):
        K, D = self.num_gaussians, self.out_dim
        nbytes = lib.vitad_gmm_packed_weight_bytes(D, K)
        packed = torch.empty(nbytes // 2, device=device, dtype=torch.float16)
        f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        sw, sb, mw, mb = f32(self.sigma.weight), f32(self.sigma.bias), f32(self.mu.weight), f32(self.mu.bias)
        check(lib.vitad_gmm_pack_weights(sw.data_ptr(), sb.data_ptr(), mw.data_ptr(), mb.data_ptr(), D, K,
                                         packed.data_ptr(), _stream()))
        torch.cuda.current_stream().synchronize()  # sw..mb may be temporaries
        self._packed = dict(w=packed, pi_w=f32(self.pi.weight), pi_b=f32(self.pi.bias), device=device, key=None)

    def forward(self, x: Tensor) -> MdnReturn:
        """x: [batch, patches, 768] → deferred (pi, sigma, mu)."""
        if not x.is_cuda:
            raise RuntimeError("GaussianMixtureDensityNetwork (vitad): CUDA input required — no CPU path")
        return MdnReturn(pi=DeferredMdnTensor(self, x, "pi"), sigma=DeferredMdnTensor(self, x, "sigma"),
                         mu=DeferredMdnTensor(self, x, "mu"))

    # -- fused scoring -----------------------------------------------------------------------------
    def patch_log_likelihood(self, x: Tensor, gumbel: Tensor | None = None, seed: int | None = None,
                             batch_index: int = 0) -> Tensor:
        """L[b,p] = mean_d logsumexp_k(log pi + log N(x_d; mu_dk, sigma_dk))  (:49-72, :86-88).
        `gumbel` [B,P,K]: explicit noise; else the kernel generates it from (`seed`, `batch_index`).
        = torch.ops.vitad.gmm_patch_loglik."""
        if not x.is_cuda:
            raise RuntimeError("vitad GMM head: CUDA input required — no CPU path")
        if gumbel is None and seed is None:  # fresh noise per call, like the reference; follows torch.manual_seed
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        xaug = None
        side = getattr(x, "_vitad_xaug", None)  # the encoder's fp16 copy of exactly these values (encoders.py)
        if side is not None and side[1] == x._version and side[0].shape[0] == x.shape[0] * x.shape[1]:
            xaug = side[0]
        return torch.ops.vitad.gmm_patch_loglik(x, xaug, gumbel, self._handle, int(seed or 0) & (2 ** 63 - 1),
                                                int(batch_index) & 0xFFFFFFFF)

    def _run(self, x: Tensor, xaug: Tensor | None, gumbel: Tensor | None, seed: int, batch_index: int) -> Tensor:
        """CUDA implementation of torch.ops.vitad.gmm_patch_loglik for this head's weights."""
        key = _param_key(self, x.device)
        if self._packed is None or self._packed.get("key") != key:
            self._pack(x.device)
            self._packed["key"] = key
        pk = self._packed
        B, P, D = x.shape
        K = self.num_gaussians
        M = B * P
        xf = x.reshape(M, D)
        if not xf.is_contiguous() or xf.dtype != torch.float32:
            xf = xf.to(torch.float32).contiguous()
            xaug = None
        if xaug is None or xaug.shape[0] != M:
            xaug = torch.empty((M, _lib.MDN_KA), device=x.device, dtype=torch.float16)
            check(lib.vitad_gmm_make_operand(xf.data_ptr(), xf.stride(0), xaug.data_ptr(), M, D, _stream()))
        n_kc, kc, _ = _lib.gmm_plan(K)
        lp2 = torch.empty((M, n_kc * kc), device=x.device, dtype=torch.float32)
        # mixing weights in fp32 on the CUDA cores.  The split-fp16 tensor-core form (vitad_gmm_log_pi_tc) is 2x
        # faster (88 -> 41 us at batch 32) but the tensor core's truncating fp32 accumulation leaves 1.1e-4 on the
        # log2-probabilities where this kernel leaves 1.9e-5 (tests/test_gmm_gpu.py), and that error enters every
        # feature's logsumexp with the same sign: not worth 1.5% of the step.
        if gumbel is None:
            check(lib.vitad_gmm_log_pi_seeded(xf.data_ptr(), xf.stride(0), pk["pi_w"].data_ptr(), pk["pi_b"].data_ptr(),
                                              seed, batch_index, lp2.data_ptr(), M, D, K, _stream()))
        else:
            g = gumbel.to(device=x.device, dtype=torch.float32).reshape(M, K).contiguous()
            check(lib.vitad_gmm_log_pi(xf.data_ptr(), xf.stride(0), pk["pi_w"].data_ptr(), pk["pi_b"].data_ptr(),
                                       g.data_ptr(), lp2.data_ptr(), M, D, K, _stream()))
        ld_ws = (M + 31) // 32 * 32
        ll_ws = torch.empty((D, ld_ws), device=x.device, dtype=torch.float32)
        L = torch.empty((M,), device=x.device, dtype=torch.float32)
        hook = getattr(self, "timing_hook", None)  # (start, end) CUDA events around the dominant kernel (bench.py)
        if hook is not None:
            hook[0].record()
        check(lib.vitad_gmm_patch_loglik(xaug.data_ptr(), pk["w"].data_ptr(), lp2.data_ptr(), xf.data_ptr(),
                                         xf.stride(0), ll_ws.data_ptr(), ld_ws, L.data_ptr(), M, D, K, _stream()))
        if hook is not None:
            hook[1].record()
        return L.view(B, P)

    def score(self, x: Tensor, gumbel: Tensor | None = None, seed: int | None = None, batch_index: int = 0):
        """→ (probability_map [B,P], image_scores [B] = 1 - min_p prob)."""
        L = self.patch_log_likelihood(x, gumbel, seed, batch_index)
        return torch.ops.vitad.gmm_finish(L)


def _finish(L: Tensor):
    """CUDA implementation of torch.ops.vitad.gmm_finish: batch-global max, exp, per-image min (:90-95, ValidatorMDN.py:133)."""
    L = L.to(torch.float32).contiguous()
    B, P = L.shape
    prob = torch.empty_like(L)
    scores = torch.empty((B,), device=L.device, dtype=torch.float32)
    check(lib.vitad_gmm_finish(L.data_ptr(), prob.data_ptr(), scores.data_ptr(), B, P, _stream()))
    return prob, scores


def _param_key(module: nn.Module, device) -> tuple:
    """See encoders._param_key: device + sum of the parameters' in-place version counters."""
    v = 0
    for t in module.parameters():
        v += t._version
    return (device, v)


def gumbel_noise(seed: int, batch_index: int, shape, device=None) -> Tensor:
    """The noise the seeded kernels add for (seed, batch_index), fp32 [B,P,K] (include/vitad.h: vitad_gumbel_noise)."""
    B, P, K = shape
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    out = torch.empty((B, P, K), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        check(lib.vitad_gumbel_noise(seed & (2 ** 64 - 1), batch_index & 0xFFFFFFFF, out.data_ptr(), B * P, K, _stream()))
    return out


def _head_of(pi, sigma, mu) -> GaussianMixtureDensityNetwork:
    for t in (sigma, mu, pi):
        if isinstance(t, DeferredMdnTensor):
            return t.head
    raise RuntimeError(
        "vitad get_probability_map/log_likelihood expect the deferred MdnReturn of the vitad "
        "GaussianMixtureDensityNetwork (materialised sigma/mu tensors are the reference's CPU/PyTorch path)"
    )


def get_probability_map(x: Tensor, pi, sigma, mu, gumbel: Tensor | None = None, seed: int | None = None,
                        batch_index: int = 0) -> Tensor:
    """MixtureDensityNetwork.py:75-97: exp(L - max over the whole batch), [batch, patches]."""
    prob, _ = _head_of(pi, sigma, mu).score(x, gumbel, seed, batch_index)
    return prob


def log_likelihood(x: Tensor, pi, sigma, mu, gumbel: Tensor | None = None, seed: int | None = None,
                   batch_index: int = 0) -> Tensor:
    """Per-patch mean over features of the reference's log_likelihood (:49-72) — the only reduction of it
    the scoring path uses (:86-88).  The per-feature tensor is never materialised."""
    return _head_of(pi, sigma, mu).patch_log_likelihood(x, gumbel, seed, batch_index)


def mdn_loss(x: Tensor, pi, sigma, mu, gumbel: Tensor | None = None, seed: int | None = None, batch_index: int = 0):
    """mean(-log_likelihood) (:100-102), forward value only (training is outside the scoring path)."""
    return -log_likelihood(x, pi, sigma, mu, gumbel, seed, batch_index).mean()


def log_gaussian_density(x: Tensor, mu: Tensor, sigma: Tensor) -> Tensor:
    """MixtureDensityNetwork.py:35-46 on materialised tensors: -log sigma - 0.5 log 2pi - 0.5 ((x - mu) / sigma)^2.
    The scoring path never materialises mu/sigma (the fused kernel evaluates this per accumulator element,
    csrc/mdn.cu EpiMdn::chunk); this elementwise form exists for callers that import the name and hold real tensors."""
    if isinstance(mu, DeferredMdnTensor) or isinstance(sigma, DeferredMdnTensor):
        raise RuntimeError("log_gaussian_density needs materialised mu/sigma; the vitad head scores through "
                           "log_likelihood / get_probability_map without materialising them")
    return -torch.log(sigma) - 0.5 * math.log(2 * math.pi) - 0.5 * torch.pow((x - mu) / sigma, 2)

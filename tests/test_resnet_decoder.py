"""Reverse-ResNet decoder of `ae_deit`: oracle vs the reference-generated golden, the GEMM
formulation of the packed weights (CPU emulation of the kernel sequence of csrc/decoder.cu), and the CUDA path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import golden
from oracle import vitad_oracle as O
from oracle import weights as W


def _decoder_sd(seed=43):
    return W.make_resnet_decoder_state_dict(seed=seed)


def _latents(B, seed=1):
    return torch.randn(B, 768, generator=torch.Generator().manual_seed(seed)) * 0.7


def _module(sd):
    from vitad.autoencoders import DecoderResNetVariableEmbeddingSize

    dec = DecoderResNetVariableEmbeddingSize(768)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items()})
    return dec.eval()


def emulate_packed(pk, z, half=False):
    """The kernel sequence of vitad_resnet_decoder_forward on the packed matrices, in torch on the CPU (fp32, or with
    operands/activations rounded to fp16 like the CUDA path when half=True).  Test infrastructure only."""
    q = (lambda t: t.half().float()) if half else (lambda t: t)
    B = z.shape[0]

    def gemm(a, w, b):
        return q(a) @ q(w).t() + b

    def im2col(x, g, c, taps, pad):
        t = F.pad(x.view(B, g, g, c), pad)
        return torch.cat([t[:, dy:dy + g, dx:dx + g, :] for dy in range(taps) for dx in range(taps)], -1).reshape(B * g * g, -1)

    h = q(F.relu(gemm(z, pk["fc1_w"], pk["fc1_b"])))
    f = q(F.relu(gemm(h, pk["fc2_w"], pk["fc2_b"])))
    g = pk["grid0"]
    x = f.view(B, 1, -1).expand(B, g * g, f.shape[1]).reshape(B * g * g, -1)
    for b in pk["blocks"]:
        w = b["width"]
        h1 = q(F.relu(gemm(x, b["w3"], b["b3"])))
        if b["stride"] == 1:
            h2 = q(F.relu(gemm(im2col(h1, g, w, 3, (0, 0, 1, 1, 1, 1)), b["w2"], b["b2"])))
            go = g
        else:
            y = q(F.relu(gemm(im2col(h1, g, w, 2, (0, 0, 0, 1, 0, 1)), b["w2"], b["b2"])))  # [(b,i,j), (a,c,co)]
            h2 = y.view(B, g, g, 2, 2, w).permute(0, 1, 3, 2, 4, 5).reshape(B * 4 * g * g, w)
            go = 2 * g
        if b["wup"] is None:
            resid = x
        else:
            up = q(gemm(x, b["wup"], b["bup"]))
            if b["stride"] == 2:
                r = torch.zeros(B, go, go, b["cout"])
                r[:, ::2, ::2, :] = up.view(B, g, g, -1)
                resid = r.reshape(B * go * go, -1)
            else:
                resid = up
        x = q(F.relu(gemm(h2, b["w1"], b["b1"]) + resid))
        g = go
    c = pk["last_c"]
    y = torch.tanh(gemm(im2col(x, g, c, 3, (0, 0, 1, 1, 1, 1)), pk["last_w"], pk["last_b"]))[:, :48]
    return y.view(B, g, g, 3, 4, 4).permute(0, 3, 1, 4, 2, 5).reshape(B, 3, 4 * g, 4 * g)


def test_oracle_matches_reference_golden():
    g = golden("resnet_decoder")
    with torch.no_grad():
        recon = O.resnet_decoder_forward(_decoder_sd(), _latents(2))
    np.testing.assert_allclose(recon.numpy()[:, :, ::4, ::4], g["recon_sub"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(recon.sum(dim=(2, 3)).numpy(), g["recon_sum"], rtol=1e-4, atol=0.05)


def test_state_dict_keys_are_the_reference_layout():
    g = golden("resnet_decoder")
    dec = _module(_decoder_sd())
    assert sorted(dec.state_dict().keys()) == sorted(str(k) for k in g["state_dict_keys"])


def test_packed_gemm_formulation_equals_the_module_stack():
    """BatchNorm folding, the flipped 3x3 kernels, the four-phase stride-2 GEMM, the even-pixel identity path and the
    collapsed image head reproduce the layer-by-layer definition (fp32 on the CPU)."""
    from vitad.autoencoders import pack_resnet_decoder

    sd = _decoder_sd()
    z = _latents(2)
    with torch.no_grad():
        ref = O.resnet_decoder_forward(sd, z)
        got = emulate_packed(pack_resnet_decoder(_module(sd)), z)
    assert got.shape == ref.shape == (2, 3, 224, 224)
    assert (got - ref).abs().max().item() <= 2e-4


def test_get_model_ae_deit_builds_the_resnet_decoder():
    from vitad.model_helper import get_model

    model = get_model("ae_deit", 224)
    assert type(model.decoder).__name__ == "DecoderResNetVariableEmbeddingSize" and model.architecture == "transformer"
    with pytest.raises(RuntimeError):
        model.decoder(torch.zeros(1, 768))  # no CPU path


def _check_against_oracle(got, ref):
    """Tolerances of the fp16-operand CUDA path against the fp32 definition.  53 GEMM layers round their operands to
    fp16 between the latent and the image; emulating exactly that rounding on the CPU (emulate_packed(half=True)) gives
    max 1.1e-2 / rms 7.8e-4 on the tanh-range image and 1.7e-3 of the maximum on a per-pixel L2 map (weights and
    activations contribute equally; bf16 operands would be 8x worse), so these bounds are the arithmetic's floor
    with ~2x margin, not slack for the kernels — kernel errors (a wrong tap, phase or border) are O(1)."""
    diff = got - ref
    assert diff.abs().max().item() <= 2.5e-2, diff.abs().max().item()
    assert diff.pow(2).mean().sqrt().item() <= 1.6e-3, diff.pow(2).mean().sqrt().item()
    x = torch.rand(ref.shape, generator=torch.Generator().manual_seed(5))
    amap_ref, amap_got = ((ref - x) ** 2).mean(1), ((got - x) ** 2).mean(1)
    assert (amap_got - amap_ref).abs().max().item() <= 4e-3 * amap_ref.max().item()
    s_ref, s_got = amap_ref.amax((1, 2)), amap_got.amax((1, 2))
    assert ((s_got - s_ref).abs() / s_ref).max().item() <= 1e-3


def test_fp16_rounding_floor_of_the_formulation():
    """The bound used on the GPU is the CPU-emulated floor of the same arithmetic (see _check_against_oracle)."""
    from vitad.autoencoders import pack_resnet_decoder

    sd = _decoder_sd()
    z = _latents(2)
    with torch.no_grad():
        ref = O.resnet_decoder_forward(sd, z)
        got = emulate_packed(pack_resnet_decoder(_module(sd)), z, half=True)
    _check_against_oracle(got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 3, 32])
def test_resnet_decoder_cuda_matches_oracle(B):
    from vitad.autoencoders import pack_resnet_decoder

    sd = _decoder_sd()
    z = _latents(B, seed=B)
    dec = _module(sd)
    with torch.no_grad():
        ref = O.resnet_decoder_forward(sd, z)
        emu = emulate_packed(pack_resnet_decoder(dec), z[:2], half=True)
        got = dec.cuda()(z.cuda()).cpu()
    assert got.shape == (B, 3, 224, 224)
    _check_against_oracle(got, ref)
    # against the CPU emulation of the same rounding points: differences are only accumulation order and fp16 ties
    n = min(B, 2)
    assert (got[:n] - emu[:n]).pow(2).mean().sqrt().item() <= 1.0e-3


@pytest.mark.gpu
def test_resnet_decoder_batch_invariance():
    """An image's reconstruction does not depend on the batch it is decoded in (tile schedules differ with M)."""
    dec = _module(_decoder_sd()).cuda()
    z = _latents(5, seed=9).cuda()
    with torch.no_grad():
        full = dec(z)
        one = dec(z[3:4])
    assert torch.equal(full[3:4], one)


@pytest.mark.gpu
def test_resnet_decoder_cuda_matches_reference_golden():
    g = golden("resnet_decoder")
    with torch.no_grad():
        got = _module(_decoder_sd()).cuda()(_latents(2).cuda()).cpu().numpy()
    diff = got[:, :, ::4, ::4] - g["recon_sub"]
    assert np.abs(diff).max() <= 2.5e-2 and np.sqrt((diff ** 2).mean()) <= 1.6e-3


@pytest.mark.gpu
def test_c_abi_rejects_bad_decoder_arguments():
    """Error behaviour of the C entry points: negative status + message, nothing launched, no exception from C."""
    import ctypes as C

    from vitad import _lib

    dec = _module(_decoder_sd()).cuda()
    z = _latents(2).cuda()
    with torch.no_grad():
        dec(z)
    w = dec._packed["w"]
    ws = dec._packed["ws"]
    out = torch.empty(2, 3, 224, 224, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    f = _lib.lib.vitad_resnet_decoder_forward
    assert f(C.byref(w), z.data_ptr(), 2, ws.data_ptr(), ws.numel(), None, s) < 0  # null output
    assert f(C.byref(w), z.data_ptr(), 0, ws.data_ptr(), ws.numel(), out.data_ptr(), s) < 0  # empty batch
    assert f(C.byref(w), z.data_ptr(), 2, ws.data_ptr(), 1024, out.data_ptr(), s) < 0  # workspace too small
    assert b"workspace" in _lib.lib.vitad_last_error()
    assert f(C.byref(w), z.data_ptr(), 2, ws.data_ptr() + 16, ws.numel() - 16, out.data_ptr(), s) < 0  # misaligned workspace
    bad = _lib.ResnetDecoderWeights.from_buffer_copy(w)
    bad.blocks[3].cin = 1000  # channels no longer chain
    assert _lib.lib.vitad_resnet_decoder_workspace_bytes(C.byref(bad), 2) == 0
    assert f(C.byref(bad), z.data_ptr(), 2, ws.data_ptr(), ws.numel(), out.data_ptr(), s) < 0
    with pytest.raises(_lib.VitadError):
        _lib.check(f(C.byref(bad), z.data_ptr(), 2, ws.data_ptr(), ws.numel(), out.data_ptr(), s))
    torch.cuda.synchronize()

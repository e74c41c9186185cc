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


def emulate_packed(pk, z, half=False, split=False):
    """The kernel sequence of vitad_resnet_decoder_forward on the packed matrices, in torch on the CPU: fp32; with
    operands/activations rounded to fp16 like the plain CUDA path (half=True); or the split-fp16 arithmetic the decoder
    uses by default (split=True: activations and weights as hi + lo fp16 pairs, products hi*w_hi + lo*w_hi + hi*w_lo).
    Test infrastructure only."""
    h16 = lambda t: t.half().float()
    if split:
        q = lambda t: h16(t) + h16(t - h16(t))  # what a [hi | lo] pair holds

        def gemm(a, w, b):
            a_hi, w_hi = h16(a), h16(w)
            return (a_hi + h16(a - a_hi)) @ w_hi.t() + a_hi @ h16(w - w_hi).t() + b
    else:
        q = h16 if half else (lambda t: t)

        def gemm(a, w, b):
            return q(a) @ q(w).t() + b
    B = z.shape[0]

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


def _check_against_oracle(got, ref, plain_fp16=False):
    """The CUDA path against the fp32 definition: north_star's 1e-3 on the L2 map (per element, helpers.assert_rel) and
    on the image score.  53 GEMM layers sit between the latent and the image; with plain fp16 operands their rounding
    leaves max 1.1e-2 / rms 7.8e-4 on the tanh-range image and ~2e-3 of the maximum on a per-pixel L2 map (CPU emulation
    of exactly those rounding points: emulate_packed(half=True); weights and activations contribute equally, no single
    stage dominates), which is why the decoder runs the split-fp16 arithmetic by default (emulate_packed(split=True):
    max 1.1e-5).  plain_fp16=True keeps the old floor bounds for the optional fast mode."""
    from helpers import assert_map_parity, assert_rel

    diff = got - ref
    x = torch.rand(ref.shape, generator=torch.Generator().manual_seed(5))
    amap_ref, amap_got = ((ref - x) ** 2).mean(1), ((got - x) ** 2).mean(1)
    s_ref, s_got = amap_ref.amax((1, 2)), amap_got.amax((1, 2))
    if plain_fp16:
        assert diff.abs().max().item() <= 2.5e-2, diff.abs().max().item()
        assert diff.pow(2).mean().sqrt().item() <= 1.6e-3, diff.pow(2).mean().sqrt().item()
        assert (amap_got - amap_ref).abs().max().item() <= 4e-3 * amap_ref.max().item()
        assert ((s_got - s_ref).abs() / s_ref).max().item() <= 2e-3
        return
    assert diff.abs().max().item() <= 1e-3, diff.abs().max().item()
    assert diff.pow(2).mean().sqrt().item() <= 1e-4, diff.pow(2).mean().sqrt().item()
    assert_map_parity(amap_got.numpy(), amap_ref.numpy(), what="L2 map")
    assert_rel(s_got.numpy(), s_ref.numpy(), 1e-3, what="image score")


def test_rounding_floor_of_the_two_arithmetics():
    """CPU emulation of the kernel sequence: plain fp16 operands miss the 1e-3 map tolerance (the reason the split
    arithmetic exists), the split-fp16 arithmetic meets it with two orders of magnitude to spare."""
    from vitad.autoencoders import pack_resnet_decoder, split3_weights

    sd = _decoder_sd()
    z = _latents(2)
    pk = pack_resnet_decoder(_module(sd))
    with torch.no_grad():
        ref = O.resnet_decoder_forward(sd, z)
        _check_against_oracle(emulate_packed(pk, z, half=True), ref, plain_fp16=True)
        with pytest.raises(AssertionError):
            _check_against_oracle(emulate_packed(pk, z, half=True), ref)
        _check_against_oracle(emulate_packed(pk, z, split=True), ref)
    # the packed split form: [w_hi | w_hi | w_lo] per tap, hi + lo reproduces the weight to 2^-22
    w = pk["blocks"][0]["w2"]
    s3 = split3_weights(w, 9).view(w.shape[0], 9, 3, -1)
    assert torch.equal(s3[:, :, 0], s3[:, :, 1]) and torch.equal(s3[:, :, 0], w.view(w.shape[0], 9, -1).half().float())
    assert ((s3[:, :, 0] + s3[:, :, 2]).reshape(w.shape) - w).abs().max().item() <= 2.0 ** -21 * w.abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("B", [1, 3, 32])
def test_resnet_decoder_cuda_matches_oracle(B):
    from vitad.autoencoders import pack_resnet_decoder

    sd = _decoder_sd()
    z = _latents(B, seed=B)
    dec = _module(sd)
    with torch.no_grad():
        ref = O.resnet_decoder_forward(sd, z)
        got = dec.cuda()(z.cuda()).cpu()
    assert got.shape == (B, 3, 224, 224)
    _check_against_oracle(got, ref)


@pytest.mark.gpu
def test_resnet_decoder_plain_fp16_mode_sits_on_its_rounding_floor():
    """split_fp16 = False (the optional fast mode): same kernels on plain fp16 operands, bounded by the CPU-emulated
    floor of that arithmetic."""
    from vitad.autoencoders import pack_resnet_decoder

    sd = _decoder_sd()
    z = _latents(3, seed=3)
    dec = _module(sd)
    dec.split_fp16 = False
    with torch.no_grad():
        ref = O.resnet_decoder_forward(sd, z)
        emu = emulate_packed(pack_resnet_decoder(dec), z[:2], half=True)
        got = dec.cuda()(z.cuda()).cpu()
    _check_against_oracle(got, ref, plain_fp16=True)
    # against the CPU emulation of the same rounding points: differences are only accumulation order and fp16 ties
    assert (got[:2] - emu[:2]).pow(2).mean().sqrt().item() <= 1.0e-3


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
    assert np.abs(diff).max() <= 1e-3 and np.sqrt((diff ** 2).mean()) <= 1e-4


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

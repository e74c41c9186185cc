"""GPU parity of the tcgen05 GEMM + fused epilogues against a plain fp32 torch reference
(inputs rounded to fp16 so both sides see identical operands; fp32 accumulate on both)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(m, n, k, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = (torch.randn(m, k, generator=g) * 0.5).to(torch.float16).cuda()
    w = (torch.randn(n, k, generator=g) * 0.05).to(torch.float16).cuda()
    b = (torch.randn(n, generator=g) * 0.1).cuda()
    return a, w, b


def _ref(a, w, b):
    return a.float() @ w.float().t() + b


@pytest.mark.parametrize("block_n", [0, 96, 128, 192, 256])
@pytest.mark.parametrize(
    "m,n,k",
    [(128, 256, 64), (128, 256, 768), (6336, 768, 768), (6336, 2304, 768), (200, 3072, 784), (77, 512, 3072),
     (6336, 800, 128), (1568, 96, 96), (3000, 1152, 384)],
)
def test_linear_bias_fp16(m, n, k, block_n):
    """Covers the single-CTA kernel (M <= 128), the CTA-pair kernel's full-width tiles and its cut tail
    (6336x768: 74 tiles of 256 columns + 8 pieces of 32; 6336x800: pieces of 128 with a ragged last column block)."""
    from vitad import _lib, ops

    # block_n is a hint: 96 exists only in the single-CTA kernel (M <= 128), elsewhere the library picks
    a, w, b = _mk(m, n, k)
    out = ops.linear(a, w, b, _lib.EPI_BIAS_F16, block_n=block_n)
    torch.cuda.synchronize()
    ref = _ref(a, w, b)
    err = (out.float() - ref).abs().max().item()
    tol = 2e-3 * ref.abs().max().item()  # fp16 output rounding (2^-11 relative)
    assert err <= tol, f"max abs err {err} > {tol}"


def test_linear_f32_exactness():
    """fp32 output: only accumulation-order differences remain (<= 1e-4 relative to row scale)."""
    from vitad import _lib, ops

    a, w, b = _mk(6272, 128, 768, seed=1)
    out = ops.linear(a, w, b, _lib.EPI_F32)
    torch.cuda.synchronize()
    ref = _ref(a, w, b)
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_linear_gelu_and_residual():
    from vitad import _lib, ops

    a, w, b = _mk(6336, 3072, 768, seed=2)
    out = ops.linear(a, w, b, _lib.EPI_BIAS_GELU_F16)
    ref = torch.nn.functional.gelu(_ref(a, w, b))  # exact erf GELU (timm Mlp: nn.GELU())
    torch.cuda.synchronize()
    # fp16 output rounding (2^-11 relative) + 3.5e-6 absolute from the sigmoid-polynomial GELU of the epilogue
    assert ((out.float() - ref).abs() <= 6e-4 * ref.abs() + 1e-5).all()

    a, w, b = _mk(6336, 768, 3072, seed=3)
    resid = torch.randn(6336, 768, device="cuda")
    ref = resid + _ref(a, w, b)
    out = ops.linear(a, w, b, _lib.EPI_RESIDUAL_F32, out=resid, resid=resid)
    torch.cuda.synchronize()
    assert out.data_ptr() == resid.data_ptr()
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()


def test_linear_qkv_layout():
    from vitad import ops

    B, T, H, Tpad = 3, 198, 12, 256
    a, w, b = _mk(B * T, 3 * H * 64, 768, seed=4)
    q = torch.zeros(B, H, T, 64, device="cuda", dtype=torch.float16)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, H, 64, Tpad, device="cuda", dtype=torch.float16)
    ops.linear_qkv(a, w, b, B, T, H, Tpad, q, k, vt, 0.125)
    torch.cuda.synchronize()
    ref = _ref(a, w, b).reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)  # [3,B,H,T,64]
    tol = 1e-2 * ref.abs().max().item()
    assert (q.float() - ref[0] * 0.125).abs().max().item() <= tol
    assert (k.float() - ref[1]).abs().max().item() <= tol
    assert (vt[..., :T].float() - ref[2].transpose(-1, -2)).abs().max().item() <= tol
    assert vt[..., T:].abs().max().item() == 0


def test_linear_patch_embed():
    from vitad import ops

    B, P, prefix, Cdim = 4, 196, 2, 768
    a, w, b = _mk(B * P, Cdim, 768, seed=5)
    pos = torch.randn(prefix + P, Cdim, device="cuda")
    out = torch.zeros(B, prefix + P, Cdim, device="cuda")
    ops.linear_patch_embed(a, w, b, pos, out, P, prefix)
    torch.cuda.synchronize()
    ref = _ref(a, w, b).reshape(B, P, Cdim) + pos[prefix:]
    assert (out[:, prefix:] - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    assert out[:, :prefix].abs().max().item() == 0


def test_linear_rejects_bad_arguments():
    from vitad import _lib, ops

    a, w, b = _mk(128, 256, 64)
    with pytest.raises(_lib.VitadError):
        ops.linear(a[:, :40], w[:, :40], b, _lib.EPI_BIAS_F16)  # K not a multiple of 16
    with pytest.raises(RuntimeError):
        ops.linear(a.cpu(), w.cpu(), b.cpu())  # no CPU path


@pytest.mark.parametrize("pair", [0, 1])
def test_cta_pair_and_single_cta_kernels_agree(pair):
    """Both GEMM variants (cta_group::2 pairs / single CTA) against the fp32 reference, ragged M."""
    from vitad import _lib, ops

    _lib.lib.vitad_set_cta_pair(pair)
    try:
        for m, n, k in [(6336, 768, 768), (300, 256, 3072), (129, 2304, 768), (6272, 512, 784)]:
            a, w, b = _mk(m, n, k, seed=m)
            out = ops.linear(a, w, b, _lib.EPI_F32)
            torch.cuda.synchronize()
            ref = _ref(a, w, b)
            assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item(), (pair, m, n, k)
    finally:
        _lib.lib.vitad_set_cta_pair(1)


@pytest.mark.parametrize("B,g,c,n", [(32, 28, 128, 128), (3, 14, 256, 256), (32, 14, 384, 64), (2, 56, 64, 64), (1, 14, 64, 32)])
def test_implicit_conv3x3_matches_conv2d(B, g, c, n):
    """conv_grid mode of vitad_linear_f16 (tap-shifted TMA loads of a zero-bordered activation) against F.conv2d."""
    from vitad import ops

    gen = torch.Generator().manual_seed(g * c)
    x = (torch.randn(B, c, g, g, generator=gen) * 0.5).half()
    w = (torch.randn(n, c, 3, 3, generator=gen) * 0.03).half()
    b = torch.randn(n, generator=gen) * 0.1
    ref = torch.nn.functional.conv2d(x.float().cuda(), w.float().cuda(), b.cuda(), padding=1)  # [B, n, g, g]
    rows = x.permute(0, 2, 3, 1).reshape(B * g * g, c).contiguous().cuda()
    wk = w.permute(0, 2, 3, 1).reshape(n, 9 * c).contiguous().cuda()
    for relu in (False, True):
        out = ops.conv3x3(ops.pad_pixels(rows, B, g), wk, b.cuda(), B, g, relu=relu)
        torch.cuda.synchronize()
        r = ref.relu() if relu else ref
        r = r.permute(0, 2, 3, 1).reshape(B * g * g, n)
        assert (out.float() - r).abs().max().item() <= 2e-3 * r.abs().max().item()


def test_out_pad_grid_scatters_into_the_bordered_layout():
    from vitad import _lib, ops
    import ctypes as C

    B, g, k, n = 3, 14, 256, 128
    a, w, b = _mk(B * g * g, n, k, seed=7)
    out = torch.zeros(B * (g + 2) ** 2, n, device="cuda", dtype=torch.float16)
    args = _lib.LinearArgs()
    args.a, args.w, args.bias = a.data_ptr(), w.data_ptr(), b.data_ptr()
    args.m, args.n, args.k, args.lda, args.ldw = a.shape[0], n, k, k, k
    args.epilogue, args.out, args.ldo, args.out_pad_grid = _lib.EPI_BIAS_RELU_F16, out.data_ptr(), n, g
    _lib.check(_lib.lib.vitad_linear_f16(C.byref(args), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = ops.pad_pixels(_ref(a, w, b).relu().half(), B, g)
    assert (out.float() - ref.float()).abs().max().item() <= 2e-3 * ref.float().abs().max().item()
    assert torch.equal(out.view(B, g + 2, g + 2, n)[:, 0], torch.zeros_like(out.view(B, g + 2, g + 2, n)[:, 0]))

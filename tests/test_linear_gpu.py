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
    # v_natural: V is stored like K (the attention kernel reads it as an MN-major operand): same values, other layout
    q2, k2, vn = torch.zeros_like(q), torch.zeros_like(q), torch.zeros_like(q)
    ops.linear_qkv(a, w, b, B, T, H, Tpad, q2, k2, vn, 0.125, v_natural=True)
    torch.cuda.synchronize()
    assert torch.equal(q2, q) and torch.equal(k2, k)
    assert torch.equal(vn, vt[..., :T].transpose(-1, -2).contiguous())


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


@pytest.mark.parametrize("M,K", [(1, 768), (100, 768), (256, 3072), (300, 768), (1000, 3072), (6336, 768), (6336, 3072), (8500, 768)])
def test_residual_gemm_layernorm_fused_matches_torch(M, K):
    """vitad_linear_resid_ln_f16 (csrc/gemm_ln.cuh): x += a @ w.T + bias in place (fp32), h = LayerNorm(x) (fp16) — against
    torch on the same fp16-rounded operands; rows beyond M in the last 256-row block must not be touched; 8500 rows need
    more clusters than one wave holds.  The separate launches (residual GEMM epilogue + vitad_layernorm768_tree) produce
    the same bits, and a row's bits do not depend on the rows around it."""
    import ctypes as C

    from vitad import _lib, ops

    g = torch.Generator().manual_seed(M + K)
    a = (torch.randn(M, K, generator=g) * 0.5).half().cuda()
    w = (torch.randn(768, K, generator=g) * (K ** -0.5)).half().cuda()
    bias = (torch.randn(768, generator=g) * 0.1).cuda()
    gamma = (1.0 + 0.2 * torch.randn(768, generator=g)).cuda()
    beta = (0.1 * torch.randn(768, generator=g)).cuda()
    x0 = (torch.randn(M, 768, generator=g) * 2.0 + 0.3).cuda()
    pad = 40  # guard rows behind the tensors: a kernel that writes past M corrupts them
    xbuf = torch.full((M + pad, 768), 7.0, device="cuda")
    xbuf[:M] = x0
    hbuf = torch.full((M + pad, 768), 3.0, device="cuda", dtype=torch.float16)
    args = _lib.LinearLnArgs()
    args.a, args.w, args.bias, args.m, args.k, args.lda, args.ldw = a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, K, K, K
    args.x, args.gamma, args.beta, args.eps, args.h, args.ldh = xbuf.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, hbuf.data_ptr(), 768
    s = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.vitad_linear_resid_ln_f16(C.byref(args), s))
    torch.cuda.synchronize()
    ref_x = x0.double() + a.double() @ w.double().t() + bias.double()
    ref_h = torch.nn.functional.layer_norm(ref_x, (768,), gamma.double(), beta.double(), 1e-6)
    assert torch.equal(xbuf[M:], torch.full((pad, 768), 7.0, device="cuda")) and torch.equal(
        hbuf[M:], torch.full((pad, 768), 3.0, device="cuda", dtype=torch.float16))
    ex = (xbuf[:M].double() - ref_x).abs().max().item()
    eh = (hbuf[:M].double() - ref_h).abs().max().item()
    assert ex <= 2e-5 * max(1.0, ref_x.abs().max().item()), ex  # fp32 accumulation of exact fp16 products
    assert eh <= 2e-3 * max(1.0, ref_h.abs().max().item()), eh  # fp16 output rounding (2^-11 relative)
    # the separate launches: residual-GEMM epilogue, then the standalone LayerNorm with the same expression tree
    x_sep = x0.clone()
    ops.linear(a, w, bias, _lib.EPI_RESIDUAL_F32, out=x_sep, resid=x_sep)
    h_sep = torch.empty(M, 768, device="cuda", dtype=torch.float16)
    _lib.check(_lib.lib.vitad_layernorm768_tree(x_sep.data_ptr(), gamma.data_ptr(), beta.data_ptr(), h_sep.data_ptr(), M, 768, 768,
                                                1e-6, s))
    torch.cuda.synchronize()
    assert torch.equal(x_sep, xbuf[:M]), (x_sep - xbuf[:M]).abs().max().item()
    assert torch.equal(h_sep, hbuf[:M]), (h_sep.float() - hbuf[:M].float()).abs().max().item()
    if M == 300:  # the same rows inside a larger problem give the same bits (batch invariance)
        x2 = torch.empty(M + 512, 768, device="cuda")
        x2[:M] = x0
        x2[M:] = 1.0
        a2 = torch.zeros(M + 512, K, device="cuda", dtype=torch.float16)
        a2[:M] = a
        h2 = torch.empty(M + 512, 768, device="cuda", dtype=torch.float16)
        args2 = _lib.LinearLnArgs.from_buffer_copy(args)
        args2.a, args2.m, args2.x, args2.h = a2.data_ptr(), M + 512, x2.data_ptr(), h2.data_ptr()
        _lib.check(_lib.lib.vitad_linear_resid_ln_f16(C.byref(args2), s))
        torch.cuda.synchronize()
        assert torch.equal(x2[:M], xbuf[:M]) and torch.equal(h2[:M], hbuf[:M])


def test_fused_and_unfused_encoder_are_bit_identical():
    """vitad_set_fused_ln(0) keeps the separate residual-GEMM + LayerNorm launches at every batch size; with fusion on, the
    encoder switches to the fused kernel from 4096 token rows (batch 21).  Both forms evaluate the same expression tree
    (csrc/ln_tree.cuh): the whole forward is bit-identical, at a batch on either side of the switch."""
    from oracle import weights as W
    from vitad import _lib
    from vitad.encoders import EncoderDeit

    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    enc = enc.cuda().eval()
    for B in (24, 5):
        imgs = W.synthetic_images(seed=3, batch=B).cuda()
        with torch.no_grad():
            fused = enc(imgs)
            _lib.lib.vitad_set_fused_ln(0)
            try:
                plain = enc(imgs)
            finally:
                _lib.lib.vitad_set_fused_ln(1)
        torch.cuda.synchronize()
        assert torch.equal(fused.patch_embedding, plain.patch_embedding), B
        assert torch.equal(fused.latent_space, plain.latent_space)
        assert torch.isfinite(fused.patch_embedding).all()

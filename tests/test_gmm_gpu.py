"""GPU parity of the fused MDN/GMM head and its score tail against the CPU oracle and the
reference-generated golden fixtures.  Parity metric for scores/maps: per element |a-b| <= 1e-3 * max(|b|, floor)
(helpers.assert_rel: north_star's 1e-3 relative; the floor, 5 % of the largest element, covers the exact zeros — the
batch's arg-max patch scores exactly 0)."""
import numpy as np
import pytest
import torch

from helpers import assert_map_parity, assert_rel, golden, gumbel

pytestmark = pytest.mark.gpu


def _head(sd, K):
    from vitad.mdn import GaussianMixtureDensityNetwork

    m = GaussianMixtureDensityNetwork(768, 768, K)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_gmm_head_k130_matches_reference_golden(tag, stress):
    from oracle import weights as W
    from vitad.mdn import get_probability_map, log_likelihood

    g = golden("gmm_head_k130_p49")
    head = _head(W.make_mdn_state_dict(seed=22, num_gaussians=130, stress=stress), 130)
    x = torch.randn(3, 49, 768, generator=torch.Generator().manual_seed(9)).cuda()
    gn = gumbel((3, 49, 130), 800).cuda()
    with torch.no_grad():
        r = head(x)
        L = log_likelihood(x, r.pi, r.sigma, r.mu, gumbel=gn).cpu().numpy()
        prob = get_probability_map(x, r.pi, r.sigma, r.mu, gumbel=gn).cpu().numpy()
    spread = g[f"{tag}_L"].max() - g[f"{tag}_L"].min()
    assert np.abs(L - g[f"{tag}_L"]).max() <= max(1e-3 * spread, 5e-5), (np.abs(L - g[f"{tag}_L"]).max(), spread)
    assert np.abs(prob - g[f"{tag}_prob"]).max() <= 1e-3


@pytest.mark.parametrize("K", [150, 170])
def test_gmm_head_reference_default_k150_matches_reference_golden(K):
    """The reference's default mixture count is 150 (startTraining_mdn.py:37) and its result tables go up to 170
    (csv_results_gmm): two chunks of 80 / 88 accumulator slots (vitad_gmm_plan)."""
    from oracle import weights as W
    from vitad.mdn import get_probability_map, log_likelihood

    g = golden("gmm_head_k150_k170_p49")
    head = _head(W.make_mdn_state_dict(seed=20 + K, num_gaussians=K, stress=True), K)
    x = torch.randn(2, 49, 768, generator=torch.Generator().manual_seed(K)).cuda()
    gn = gumbel((2, 49, K), 900 + K).cuda()
    with torch.no_grad():
        r = head(x)
        L = log_likelihood(x, r.pi, r.sigma, r.mu, gumbel=gn).cpu().numpy()
        prob = get_probability_map(x, r.pi, r.sigma, r.mu, gumbel=gn).cpu().numpy()
    ref = g[f"k{K}_L"]
    spread = ref.max() - ref.min()
    assert np.abs(L - ref).max() <= max(1e-3 * spread, 5e-5), (np.abs(L - ref).max(), spread)
    assert np.abs(prob - g[f"k{K}_prob"]).max() <= 1e-3


def test_gmm_plan_covers_the_reference_range():
    from vitad import _lib

    for K in (1, 37, 100, 104, 105, 112, 113, 130, 144, 145, 150, 160, 161, 170, 176, 177, 208):
        n_kc, kc, kcv = _lib.gmm_plan(K)
        assert n_kc * kcv >= K and kcv <= kc and kc % 8 == 0 and 2 * kc <= 256, (K, n_kc, kc, kcv)
    with pytest.raises(_lib.VitadError):
        _lib.gmm_plan(209)


@pytest.mark.parametrize("K", [100, 130, 150, 200])
def test_seeded_gumbel_noise_is_the_noise_the_kernel_adds(K):
    """vitad_gmm_log_pi_seeded (noise generated in the kernel, Philox keyed by (seed, batch_index, token, mixture)) equals
    vitad_gmm_log_pi fed with vitad_gumbel_noise of the same key, bit for bit; the noise is standard Gumbel."""
    from oracle import weights as W
    from vitad.mdn import gumbel_noise

    head = _head(W.make_mdn_state_dict(seed=K, num_gaussians=K, stress=True), K)
    B, P = 3, 196
    x = torch.randn(B, P, 768, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        for seed, bi in ((1234, 0), (1234, 5), (2 ** 40 + 7, 5)):
            gn = gumbel_noise(seed, bi, (B, P, K))
            a = head.patch_log_likelihood(x, seed=seed, batch_index=bi)
            b = head.patch_log_likelihood(x, gumbel=gn)
            assert torch.equal(a, b), (seed, bi)
        g0, g1, g2 = gumbel_noise(1, 0, (B, P, K)), gumbel_noise(1, 1, (B, P, K)), gumbel_noise(2, 0, (B, P, K))
        assert torch.equal(g0, gumbel_noise(1, 0, (B, P, K)))
        assert not torch.equal(g0, g1) and not torch.equal(g0, g2)
        # rows of a smaller batch are the first rows of a larger one (the key is the token index, not the batch size)
        assert torch.equal(gumbel_noise(1, 0, (1, P, K)), g0[:1])
    big = gumbel_noise(99, 3, (32, 196, K)).double().flatten()
    assert torch.isfinite(big).all()
    if K == 100:  # 2^26 draws: a 24-bit uniform rounded to 1.0 once in 2^24 draws (g = +inf, found by the NaN-propagating score tail)
        many = torch.cat([gumbel_noise(7, bi, (64, 196, 130)).flatten() for bi in range(42)])
        assert many.numel() > 2 ** 26 and torch.isfinite(many).all() and many.max().item() < 20.0 and many.min().item() > -4.0
    assert abs(big.mean().item() - 0.5772157) < 5e-3 and abs(big.var().item() - np.pi ** 2 / 6) < 2e-2
    assert abs(torch.corrcoef(torch.stack([big[:-1], big[1:]]))[0, 1].item()) < 5e-3


def test_sharded_validators_reproduce_the_unsharded_results_bit_for_bit():
    """batch i goes to rank i % W; the stitched per-rank results of the GMM and NF validators equal the
    one-process results exactly (the GMM noise is keyed by the global batch index, not by call order)."""
    from oracle import weights as W
    from vitad.encoders import EncoderDeit
    from vitad.mdn import GaussianMixtureDensityNetwork
    from vitad.nf import NormalizingFlow
    from vitad.synthetic import batches, make_category
    from vitad.validators import ValidatorMdn, ValidatorNF

    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    props = {"dataset": "s", "dataclass": "x", "num_gaussians": 100, "fp_thres": 0.3}
    mdn_sd = W.make_mdn_state_dict(21, 100, stress=True)
    np.random.seed(0)
    nf = NormalizingFlow(768, 224, 196, 0.16, 20)
    nf_sd = W.make_nf_state_dict(31, stress=True)
    bl = batches(*make_category("cable", 75, seed=501), batch_size=16)  # 5 batches, the last one short (11)

    def run(kind, rank, world):
        if kind == "gmm":
            v = ValidatorMdn([GaussianMixtureDensityNetwork(768, 768, 100)], enc, None, props, weights_object=[mdn_sd],
                             rank=rank, world_size=world, gumbel_seed=77)
            return v.valid_loop_transformer(bl, keep_origs=False)
        v = ValidatorNF([nf], enc, None, props, weights_object=[nf_sd], rank=rank, world_size=world)
        return v.valid_loop_transformer_nf(bl, keep_origs=False)

    for kind in ("gmm", "nf"):
        full = run(kind, 0, 1)
        for world in (2, 3):
            parts = [run(kind, r, world) for r in range(world)]
            stitched = {}
            for key in ("image_scores", "pixel_scores", "image_labels"):
                pieces = {}
                for part in parts:
                    off = 0
                    for b, n in zip(part["batch_index"], part["batch_sizes"]):
                        pieces[int(b)] = part[key][off: off + n]
                        off += int(n)
                stitched[key] = np.concatenate([pieces[b] for b in sorted(pieces)])
            for key, v in stitched.items():
                assert np.array_equal(v, full[key]), (kind, world, key)
        assert np.isfinite(full["image_scores"]).all() and full["image_scores"].std() > 0


@pytest.mark.parametrize("K", [100, 130, 37, 110, 150, 176, 200])
@pytest.mark.parametrize("B,P", [(2, 196), (1, 49), (5, 196)])
def test_gmm_patch_loglik_matches_oracle(K, B, P):
    from oracle import vitad_oracle as O
    from oracle import weights as W

    sd = W.make_mdn_state_dict(seed=K, num_gaussians=K, stress=True)
    head = _head(sd, K)
    x = torch.randn(B, P, 768, generator=torch.Generator().manual_seed(B * 1000 + P))
    gn = gumbel((B, P, K), 5)
    with torch.no_grad():
        ref = O.mdn_patch_loglik(x, sd, gn)
        L = head.patch_log_likelihood(x.cuda(), gn.cuda()).cpu()
    spread = (ref.max() - ref.min()).item()
    err = (L - ref).abs().max().item()
    assert err <= max(1e-3 * spread, 5e-5), (err, spread)


@pytest.mark.parametrize("tag,stress", [("default", False), ("stress", True)])
def test_gmm_validator_path_matches_reference_golden(tag, stress):
    """DeiT (CUDA) + GMM head (CUDA) + score tail (CUDA) vs ValidatorMdn.valid_loop_transformer run by the
    reference: two batches (2 + 1 images)."""
    from oracle import weights as W
    from vitad import ops
    from vitad.encoders import EncoderDeit

    g = golden("gmm_validator_k100")
    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    enc = enc.cuda().eval()
    head = _head(W.make_mdn_state_dict(seed=21, num_gaussians=100, stress=stress), 100)
    imgs = W.synthetic_images(seed=5, batch=3)
    scores, maps = [], []
    with torch.no_grad():
        for call, (s, e) in enumerate(((0, 2), (2, 3))):
            f = enc(imgs[s:e].cuda())
            prob, sc = head.score(f.patch_embedding, gumbel((e - s, 196, 100), 700 + call).cuda())
            if call == 0:
                L = head.patch_log_likelihood(f.patch_embedding, gumbel((2, 196, 100), 700).cuda()).cpu().numpy()
                ref = g[f"{tag}_L_batch0"]
                assert np.abs(L - ref).max() <= max(2e-3 * (ref.max() - ref.min()), 1e-4)
            mp, _ = ops.bilinear_up(prob.view(-1, 14, 14), 224, align_corners=True, post_one_minus=True)
            scores.append(sc.cpu())
            maps.append(mp.cpu())
    scores, maps = torch.cat(scores).numpy(), torch.cat(maps).numpy()
    ref_s, ref_m = g[f"{tag}_image_scores"], g[f"{tag}_pixel_scores_sub"]
    assert_rel(scores, ref_s, 1e-3, what="image scores")
    assert_map_parity(maps[:, :, ::8, ::8], ref_m, what="anomaly maps")
    np.testing.assert_allclose(maps.sum(axis=(1, 2, 3)), g[f"{tag}_pixel_scores_sum"], rtol=2e-3)


def test_gmm_finish_and_upsample_match_oracle():
    from oracle import vitad_oracle as O
    from vitad import _lib, ops

    L = torch.randn(7, 196, generator=torch.Generator().manual_seed(3)) * 0.3 - 900.0
    prob = torch.empty(7, 196, device="cuda")
    scores = torch.empty(7, device="cuda")
    Lc = L.cuda()
    _lib.check(_lib.lib.vitad_gmm_finish(Lc.data_ptr(), prob.data_ptr(), scores.data_ptr(), 7, 196,
                                         torch.cuda.current_stream().cuda_stream))
    rp = O.mdn_probability_map(L)
    rs, rm = O.mdn_scores(rp, 224, 16)
    mp, _ = ops.bilinear_up(prob.view(-1, 14, 14), 224, align_corners=True, post_one_minus=True)
    torch.cuda.synchronize()
    assert (prob.cpu() - rp).abs().max().item() <= 2e-6
    assert (scores.cpu() - rs).abs().max().item() <= 2e-6
    assert (mp.cpu() - rm).abs().max().item() <= 2e-6


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("g_in", [14, 7])
def test_bilinear_matches_torch_interpolate(align, g_in):
    from vitad import ops

    x = torch.rand(5, g_in, g_in, generator=torch.Generator().manual_seed(1)).cuda()
    out, mx = ops.bilinear_up(x, 224, align_corners=align, want_max=True)
    ref = torch.nn.functional.interpolate(x.unsqueeze(1), size=(224, 224), mode="bilinear", align_corners=align)
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() <= 2e-6
    assert (mx - ref.amax(dim=(1, 2, 3))).abs().max().item() <= 2e-6


def test_recon_l2_map_matches_reference_golden():
    from vitad import ops

    g = golden("recon_l2")
    gen = torch.Generator().manual_seed(4)
    images = torch.rand(3, 3, 224, 224, generator=gen)
    recon = torch.tanh(torch.randn(3, 3, 224, 224, generator=gen))
    amap, score = ops.l2_map_score(recon.cuda(), images.cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(score.cpu().numpy(), g["image_scores"], rtol=1e-6)
    np.testing.assert_allclose(amap.cpu().numpy()[:, :, ::8, ::8], g["map_sub"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(amap.sum(dim=(1, 2, 3)).cpu().numpy(), g["map_sum"], rtol=1e-5)


@pytest.mark.parametrize("K,M", [(100, 6272), (130, 49), (37, 196)])
def test_log_pi_tensor_core_path_matches_fp32_kernel(K, M):
    """Mixing weights: the split-fp16 tensor-core GEMM (x = xh + xl, W = Wh + Wl, three exact products; an alternative
    the library offers) against the fp32 CUDA-core kernel the head uses, both against float64.  Measured on B200:
    the tensor core's truncating fp32 accumulation leaves ~1.1e-4 on log2(softmax(pi(x)+g)+1e-15), the fp32 kernel
    ~1.9e-5 — which is why the head stays on the fp32 kernel."""
    import ctypes as C

    from vitad import _lib
    from vitad._lib import check, lib

    g = torch.Generator().manual_seed(K + M)
    x = (torch.randn(M, 768, generator=g) * 1.5).cuda()
    w = (torch.randn(K, 768, generator=g) * 0.05).cuda()
    b = (torch.randn(K, generator=g) * 0.1).cuda()
    gn = gumbel((M, 1, K), 99).reshape(M, K).cuda().contiguous()
    n_kc, kc, _ = _lib.gmm_plan(K)
    s = torch.cuda.current_stream().cuda_stream
    ref = torch.empty(M, n_kc * kc, device="cuda")
    check(lib.vitad_gmm_log_pi(x.data_ptr(), 768, w.data_ptr(), b.data_ptr(), gn.data_ptr(), ref.data_ptr(), M, 768, K, s))
    packed = torch.empty(lib.vitad_gmm_pi_packed_bytes(768, K) // 2, device="cuda", dtype=torch.float16)
    check(lib.vitad_gmm_pack_pi(w.data_ptr(), 768, K, packed.data_ptr(), s))
    ws = torch.empty(lib.vitad_gmm_log_pi_workspace_bytes(M, 768, K), device="cuda", dtype=torch.uint8)
    out = torch.empty_like(ref)
    check(lib.vitad_gmm_log_pi_tc(x.data_ptr(), 768, packed.data_ptr(), b.data_ptr(), gn.data_ptr(), out.data_ptr(), M, 768,
                                  K, ws.data_ptr(), ws.numel(), s))
    torch.cuda.synchronize()
    pad = ref < -1e29
    assert torch.equal(pad, out < -1e29)
    exact = torch.log2(torch.softmax(x.double() @ w.double().t() + b.double() + gn.double(), -1) + 1e-15).float()
    kcv = kc if n_kc == 1 else (K + 1) // 2
    cols = [(k // kcv) * kc + k % kcv for k in range(K)]
    err_tc = (out[:, cols] - exact).abs().max().item()
    err_f32 = (ref[:, cols] - exact).abs().max().item()
    print(f"log2-probability max error vs float64: tensor-core split-fp16 {err_tc:.2e}, fp32 CUDA-core kernel {err_f32:.2e}")
    # log2-probabilities reach -20: 1e-4 absolute is a few fp32 ulps of summation-order noise in the 768-term logits
    assert err_f32 <= 5e-5 and err_tc <= 3e-4, (err_tc, err_f32)


def test_nan_is_not_swallowed_by_the_score_reductions():
    """torch.max / torch.amax propagate NaN in the reference (MixtureDensityNetwork.py:90-92, ValidatorNF.py:137-142,
    ValidatorRecon.py:116): a numerical failure upstream must surface as a NaN score (and make the metrics raise), not as
    a plausible number."""
    from vitad import ops

    L = torch.randn(3, 196, device="cuda") - 900.0
    L[1, 17] = float("nan")
    prob, scores = torch.ops.vitad.gmm_finish(L)
    assert torch.isnan(scores).all() and torch.isnan(prob).all()  # the batch-global max couples every image
    x = torch.rand(4, 14, 14, device="cuda")
    x[2, 3, 3] = float("nan")
    _, mx = ops.bilinear_up(x, 224, align_corners=False, want_max=True)
    assert torch.isnan(mx[2]) and torch.isfinite(mx[[0, 1, 3]]).all()
    recon, img = torch.rand(2, 3, 224, 224, device="cuda"), torch.rand(2, 3, 224, 224, device="cuda")
    recon[0, 1, 100, 7] = float("nan")
    _, sc = ops.l2_map_score(recon, img)
    assert torch.isnan(sc[0]) and torch.isfinite(sc[1])

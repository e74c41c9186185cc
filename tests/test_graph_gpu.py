"""CUDA-graph replay of a scoring step (vitad.graphed.GraphedStep): the captured chain — programmatic-dependent-launch
edges, the independent side grid of the fused GMM kernel, the memsets of the implicit convolutions — reproduces the eager
results bit for bit."""
import pytest
import torch

from oracle import weights as W

pytestmark = pytest.mark.gpu


def test_graphed_deit_gmm_step_equals_eager():
    from vitad import _lib, ops
    from vitad.encoders import EncoderDeit
    from vitad.graphed import GraphedStep
    from vitad.mdn import GaussianMixtureDensityNetwork

    B, K = 8, 100  # 1568 token rows: the fused GMM kernel runs its 4-CTA grid plus the side CTA-pair grid
    enc = EncoderDeit(224)
    enc.load_state_dict(W.make_deit_state_dict(seed=11, stress=True))
    head = GaussianMixtureDensityNetwork(768, 768, K)
    head.load_state_dict(W.make_mdn_state_dict(seed=21, num_gaussians=K, stress=True))
    enc, head = enc.cuda().eval(), head.cuda().eval()
    g = torch.randn(B, 196, K, device="cuda")

    def step(imgs):
        f = enc(imgs)
        prob, scores = head.score(f.patch_embedding, g)
        maps, _ = ops.bilinear_up(prob.view(-1, 14, 14), 224, align_corners=True, post_one_minus=True)
        return scores, maps

    a = W.synthetic_images(seed=3, batch=B).cuda()
    b = W.synthetic_images(seed=4, batch=B).cuda()
    with torch.no_grad():
        eager = [tuple(t.clone() for t in step(x)) for x in (a, b)]
        _lib.lib.vitad_set_gmm_split(0)
        unsplit = tuple(t.clone() for t in step(a))
        _lib.lib.vitad_set_gmm_split(-1)
    gstep = GraphedStep(step, a)
    for x, ref in ((a, eager[0]), (b, eager[1]), (a, eager[0])):
        out = gstep(x)
        torch.cuda.synchronize()
        assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    assert torch.equal(unsplit[0], eager[0][0]) and torch.equal(unsplit[1], eager[0][1])


def test_graphed_resnet_decoder_equals_eager():
    from vitad.autoencoders import DecoderResNetVariableEmbeddingSize
    from vitad.graphed import GraphedStep

    dec = DecoderResNetVariableEmbeddingSize(768)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in W.make_resnet_decoder_state_dict(seed=43).items()})
    dec = dec.cuda().eval()
    z = torch.randn(4, 768, device="cuda") * 0.7
    z2 = torch.randn(4, 768, device="cuda") * 0.7
    with torch.no_grad():
        ref, ref2 = dec(z).clone(), dec(z2).clone()
    gstep = GraphedStep(lambda lat: (dec(lat),), z)
    assert torch.equal(gstep(z2)[0], ref2)
    assert torch.equal(gstep(z)[0], ref)

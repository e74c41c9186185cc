"""CPU study: predicted parity of the bf16-operand / fp32-accumulate design vs the fp32 oracle."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import vitad_oracle as O, weights as W

DT = torch.float16 if os.environ.get("EMU","bf16")=="fp16" else torch.bfloat16
bf = lambda t: t.to(DT).float()

def lin(x, w, b): return bf(x) @ bf(w).t() + b

def deit_emu(sd, images, p="deit."):
    w = sd[p + "patch_embed.proj.weight"]; B = images.shape[0]
    patches = images.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, 196, 768)
    x = lin(patches, w.reshape(768, -1), sd[p + "patch_embed.proj.bias"])
    x = torch.cat((sd[p + "cls_token"].expand(B, -1, -1), sd[p + "dist_token"].expand(B, -1, -1), x), 1) + sd[p + "pos_embed"]
    for i in range(12):
        q_ = f"{p}blocks.{i}."
        h = O.layer_norm(x, sd[q_ + "norm1.weight"], sd[q_ + "norm1.bias"], 1e-6)
        qkv = lin(h, sd[q_ + "attn.qkv.weight"], sd[q_ + "attn.qkv.bias"]).reshape(B, 198, 3, 12, 64).permute(2, 0, 3, 1, 4)
        q, k, v = bf(qkv[0] * 0.125), bf(qkv[1]), bf(qkv[2])
        s = q @ k.transpose(-2, -1)
        pmat = torch.exp(s - s.amax(-1, keepdim=True)); l = pmat.sum(-1, keepdim=True)
        o = (bf(pmat) @ v) / l
        o = o.transpose(1, 2).reshape(B, 198, 768)
        x = x + lin(o, sd[q_ + "attn.proj.weight"], sd[q_ + "attn.proj.bias"])
        h = O.layer_norm(x, sd[q_ + "norm2.weight"], sd[q_ + "norm2.bias"], 1e-6)
        h = O.gelu_erf(lin(h, sd[q_ + "mlp.fc1.weight"], sd[q_ + "mlp.fc1.bias"]))
        x = x + lin(h, sd[q_ + "mlp.fc2.weight"], sd[q_ + "mlp.fc2.bias"])
    x = O.layer_norm(x, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    return x[:, 2:], x[:, 0]

def mdn_emu(x, sd, g):
    sd2 = dict(sd)
    for k in ("sigma.weight", "mu.weight"): sd2[k] = bf(sd[k])
    for k in ("sigma.bias", "mu.bias"):
        hi = bf(sd[k]); sd2[k] = hi + bf(sd[k] - hi)
    # GEMM sees bf16(x); epilogue uses fp32 x for (x - mu): emulate by patching
    B, P, D = x.shape; K = sd["pi.weight"].shape[0]
    log_pi = O.mdn_log_pi(x, sd, g).reshape(B * P, K)   # pi path: computed in (tf32/bf16x?) -> assume fp32-accurate here
    xf = x.reshape(B * P, D); xb = bf(xf)
    out = torch.empty(B * P)
    for s in range(0, B * P, 256):
        xc, xcb = xf[s:s+256], xb[s:s+256]
        sigma = (torch.nn.functional.elu(xcb @ sd2["sigma.weight"].t() + sd2["sigma.bias"]) + 1 + 1e-15).view(-1, D, K)
        mu = (xcb @ sd2["mu.weight"].t() + sd2["mu.bias"]).view(-1, D, K)
        dens = -torch.log(sigma) - O.LOG_SQRT_2PI - 0.5 * ((xc.unsqueeze(-1) - mu) / sigma) ** 2
        out[s:s+256] = torch.logsumexp(log_pi[s:s+256].unsqueeze(1) + dens, -1).mean(1)
    return out.view(B, P)

torch.set_num_threads(8)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for stress_enc, stress_mdn in ((False, False), (True, False), (True, True)):
    sd = W.make_deit_state_dict(11, stress=stress_enc); mdn = W.make_mdn_state_dict(21, 100, stress=stress_mdn)
    imgs = W.synthetic_images(5, B); g = O.gumbel_noise((B, 196, 100), torch.Generator().manual_seed(700))
    with torch.no_grad():
        t0, _ = O.deit_forward(sd, imgs); t1, _ = deit_emu(sd, imgs)
        print(f"enc stress={stress_enc}: token max abs err {(t0 - t1).abs().max():.3e}  rms {((t0-t1)**2).mean().sqrt():.3e} (token rms {t0.pow(2).mean().sqrt():.3f})")
        L0 = O.mdn_patch_loglik(t0, mdn, g)
        L_mdn_only = mdn_emu(t0, mdn, g)       # exact encoder, bf16 MDN
        L_all = mdn_emu(t1, mdn, g)            # bf16 encoder + bf16 MDN
        for nm, L in (("mdn-bf16 only", L_mdn_only), ("enc+mdn bf16", L_all)):
            s0, m0 = O.mdn_scores(O.mdn_probability_map(L0), 224, 16); s1, m1 = O.mdn_scores(O.mdn_probability_map(L), 224, 16)
            print(f"  mdn stress={stress_mdn} {nm}: L max abs err {(L - L0).abs().max():.3e}; L spread {L0.max()-L0.min():.3f}; "
                  f"score {s0.tolist()[:3]} rel err {((s1 - s0).abs() / s0.abs().clamp_min(1e-3)).max():.3e}; map rel err {((m1-m0).abs()/m0.abs().clamp_min(1e-3)).max():.3e}")

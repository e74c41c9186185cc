#!/usr/bin/env python
"""Designs the fixed synthetic anomaly sets of the AUROC parity tests (north_star: "image-level AUROC identical to 4
decimals on a fixed synthetic anomaly set"; SURVEY.md §7: a set designed for separation with an asserted minimum gap).

With random-init weights the image scores of arbitrary synthetic images cluster, and a near-tie can swap under a 1e-3
numerical difference without any kernel being wrong.  Each set is therefore a list of per-image seeds (vitad.synthetic.
make_designed_set) picked from a candidate pool so that the ORACLE's image scores of the whole set are pairwise separated
by >= GAP x the allowed numerical noise (1e-3 of the largest score), with both labels interleaved along the score axis.
CPU only (the oracle); prints the seed lists that tests/helpers.py commits.  Re-run only if the synthetic weights or the
image generator change:   python tools/design_anomaly_sets.py [gmm|nf|recon]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "vit-ad_b200"))

from oracle import vitad_oracle as O  # noqa: E402
from oracle import weights as W  # noqa: E402
from vitad.synthetic import make_designed_set  # noqa: E402

GAP = 11.0  # x (1e-3 * max score); the tests assert >= 10
POOL = 160
NF_GAIN = 3.0  # subnet output gain of the NF test weights: keeps the NF image scores mid-range (4 saturates them near 1)


def pick(scores, labels, n, noise):
    """Greedy walk up the score axis taking images >= GAP*noise apart, alternating labels where possible."""
    order = np.argsort(scores)
    chosen, last = [], -np.inf
    want = None
    for rnd in range(2):  # second round relaxes the label alternation
        for i in order:
            if i in chosen or scores[i] - last < GAP * noise:
                continue
            if rnd == 0 and want is not None and labels[i] != want:
                continue
            chosen.append(i)
            last = scores[i]
            want = 1 - labels[i]
        if len(chosen) >= n:
            break
        last = -np.inf if not chosen else max(scores[c] for c in chosen)
    return chosen


def design_gmm(n=24, K=100):
    seeds = list(range(9000, 9000 + POOL))
    images, labels, _ = make_designed_set(seeds)
    enc_sd = W.make_deit_state_dict(seed=11, stress=True)
    mdn_sd = W.make_mdn_state_dict(seed=21, num_gaussians=K, stress=True)
    Ls = []
    with torch.no_grad():
        for s in range(0, POOL, 16):
            tok, _ = O.deit_forward(enc_sd, images[s:s + 16])
            g = torch.stack([O.gumbel_noise((196, K), torch.Generator().manual_seed(4242 + seeds[s + j])) for j in range(tok.shape[0])])
            Ls.append(O.mdn_patch_loglik(tok, mdn_sd, g))
    L = torch.cat(Ls).numpy()  # [POOL, 196]
    # the batch-global max couples the set: the image holding the pool's largest L is always part of it
    top = int(np.argmax(L.max(1)))
    scores = 1.0 - np.exp(L.min(1) - L.max())
    noise = 1e-3 * scores.max()
    lab = labels.numpy()
    best = None
    chosen = [top]
    last_sorted = sorted(range(POOL), key=lambda i: scores[i])
    sel = [top]
    for i in last_sorted:
        if i == top:
            continue
        if all(abs(scores[i] - scores[j]) >= GAP * noise for j in sel):
            sel.append(i)
    # thin to n keeping labels balanced and interleaved
    sel = sorted(sel, key=lambda i: scores[i])
    if len(sel) > n:
        keep = {top}
        idx = np.linspace(0, len(sel) - 1, n).round().astype(int)
        keep |= {sel[j] for j in idx}
        sel = sorted(keep, key=lambda i: scores[i])[:n] if top in sorted(keep, key=lambda i: scores[i])[:n] else sorted(keep, key=lambda i: scores[i])[-n:]
    print("gmm", len(sel), "labels", lab[sel].tolist(), "scores", np.round(scores[sel], 4).tolist())
    print("DESIGNED_GMM =", [seeds[i] for i in sel])


def design_independent(name, score_fn, n):
    seeds = list(range(9500, 9500 + POOL)) if name == "nf" else list(range(9800, 9800 + POOL))
    images, labels, _ = make_designed_set(seeds)
    scores = score_fn(images)
    noise = 1e-3 * np.abs(scores).max()
    lab = labels.numpy()
    sel = []
    for i in np.argsort(scores):
        if all(abs(scores[i] - scores[j]) >= GAP * noise for j in sel):
            sel.append(int(i))
    if len(sel) > n:
        idx = np.linspace(0, len(sel) - 1, n).round().astype(int)
        sel = [sel[j] for j in idx]
    print(name, len(sel), "labels", lab[sel].tolist(), "scores", np.round(scores[sel], 5).tolist())
    print(f"DESIGNED_{name.upper()} =", [seeds[i] for i in sel])


def nf_scores(images):
    enc_sd = W.make_deit_state_dict(seed=11, stress=True)
    nf_sd = W.make_nf_state_dict(seed=31, stress=True, subnet_gain=NF_GAIN)
    out = []
    with torch.no_grad():
        for s in range(0, images.shape[0], 16):
            tok, _ = O.deit_forward(enc_sd, images[s:s + 16], block_index=0)
            _, amap, _, _ = O.nf_forward(nf_sd, O.tokens_to_nchw(tok), flow_steps=20, img_size=224)
            out.append(O.nf_scores(amap))
    return torch.cat(out).numpy()


def recon_scores(images):
    sd = {("encoder." + k): v for k, v in W.make_deit_state_dict(seed=11, stress=True).items()}
    sd.update(W.make_resnet_decoder_state_dict(seed=43))
    out = []
    with torch.no_grad():
        for s in range(0, images.shape[0], 16):
            _, cls = O.deit_forward(sd, images[s:s + 16], prefix="encoder.deit.")
            sc, _ = O.recon_l2_scores(O.resnet_decoder_forward(sd, cls), images[s:s + 16])
            out.append(sc)
    return torch.cat(out).numpy()


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    which = sys.argv[1:] or ["gmm", "nf", "recon"]
    if "gmm" in which:
        design_gmm()
    if "nf" in which:
        design_independent("nf", nf_scores, 16)
    if "recon" in which:
        design_independent("recon", recon_scores, 16)

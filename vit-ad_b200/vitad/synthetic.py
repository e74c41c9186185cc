"""Synthetic MVTecAD/BTAD-shaped validation data: there is no network for the real datasets, so
the sweep and the AUROC parity test use seeded images with the loader's contract (fp32 NCHW in [0,1], image
labels, [B,1,S,S] masks; src/data_loader/GeneralDataset.py:38-59,114) and pasted high-contrast patches as
anomalies.  Test-set sizes per category follow the published datasets."""
from __future__ import annotations

import torch

MVTEC_TEST_SIZES = {
    "bottle": 83, "cable": 150, "capsule": 132, "carpet": 117, "grid": 78, "hazelnut": 110, "leather": 124,
    "metal_nut": 115, "pill": 167, "screw": 160, "tile": 117, "toothbrush": 42, "transistor": 100, "wood": 79,
    "zipper": 151,
}  # 1725 images
BTAD_TEST_SIZES = {"01": 70, "02": 230, "03": 441}


def make_category(name: str, n_images: int, seed: int, size: int = 224, anomaly_fraction: float = 0.5):
    """→ images [n,3,S,S] fp32, image_labels [n] int64, pixel_labels [n,1,S,S] fp32.  A smooth per-category
    texture plus noise; anomalous images get 1-3 pasted 32x32 high-contrast patches (mask = patch support)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(1, 3, 8, 8, generator=g)  # blocky 8x8 texture, exact replication (thread-count independent)
    base = base.repeat_interleave(size // 8, dim=2).repeat_interleave(size // 8, dim=3)
    images = (0.7 * base + 0.3 * torch.rand(n_images, 3, size, size, generator=g)).clamp(0, 1)
    labels = (torch.rand(n_images, generator=g) < anomaly_fraction).long()
    masks = torch.zeros(n_images, 1, size, size)
    for i in range(n_images):
        if labels[i]:
            for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
                y, x = (int(v) for v in torch.randint(0, size - 32, (2,), generator=g))
                images[i, :, y:y + 32, x:x + 32] = (torch.rand(3, 1, 1, generator=g) > 0.5).float()
                masks[i, :, y:y + 32, x:x + 32] = 1.0
    return images, labels, masks


def make_diverse_image(seed: int, size: int = 224):
    """One image of a deliberately heterogeneous family (texture block size, noise share, contrast and brightness all
    drawn from the seed; anomalous images get 1-3 pasted flat patches of 16-64 px): with random-init weights the image
    scores of a homogeneous family cluster within a few 1e-3 of each other, this family spreads them.
    → image [1,3,S,S] fp32 in [0,1], label [1] int64, mask [1,1,S,S] fp32."""
    g = torch.Generator().manual_seed(int(seed))
    blk = (2, 4, 8, 16, 32)[int(torch.randint(0, 5, (1,), generator=g))]
    base = torch.rand(1, 3, size // blk, size // blk, generator=g)
    base = base.repeat_interleave(blk, dim=2).repeat_interleave(blk, dim=3)
    alpha = 0.05 + 0.9 * float(torch.rand(1, generator=g))        # noise share
    contrast = 0.2 + 0.8 * float(torch.rand(1, generator=g))
    bright = 0.5 + 0.3 * (float(torch.rand(1, generator=g)) - 0.5)
    img = (1 - alpha) * base + alpha * torch.rand(1, 3, size, size, generator=g)
    img = ((img - 0.5) * contrast + bright).clamp(0, 1)
    label = (torch.rand(1, generator=g) < 0.5).long()
    mask = torch.zeros(1, 1, size, size)
    if label[0]:
        for _ in range(int(torch.randint(1, 4, (1,), generator=g))):
            ps = int(torch.randint(16, 65, (1,), generator=g))
            y, x = (int(v) for v in torch.randint(0, size - ps, (2,), generator=g))
            img[0, :, y:y + ps, x:x + ps] = (torch.rand(3, 1, 1, generator=g) > 0.5).float()
            mask[0, :, y:y + ps, x:x + ps] = 1.0
    return img, label, mask


def make_designed_set(seeds, size: int = 224):
    """A fixed anomaly set given as one seed per image (tools/design_anomaly_sets.py picks the seeds so that the oracle's
    image scores of the whole set are pairwise well separated): image i = make_diverse_image(seeds[i])."""
    parts = [make_diverse_image(int(s), size=size) for s in seeds]
    return (torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts]), torch.cat([p[2] for p in parts]))


def batches(images, labels, masks, batch_size: int = 32):
    """The reference DataLoader: shuffle=False, no drop_last (GeneralDataLoader.py:152-156) → short tail batch."""
    return [(images[s:s + batch_size], masks[s:s + batch_size], labels[s:s + batch_size])
            for s in range(0, images.shape[0], batch_size)]

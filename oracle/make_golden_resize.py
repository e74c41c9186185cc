"""tests/golden/resize_u8.npz: Pillow's own BILINEAR resize (what torchvision transforms.Resize runs on the PIL images of
src/data_loader/GeneralDataset.py:38-59) on the seeded images of tests/test_resize.py.  Run in the build container (Pillow
12.2.0); the GPU box only reads the committed outputs."""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from test_resize import GOLDEN_SIZES, SIZES, _image  # noqa: E402

out = {}
for h, w in GOLDEN_SIZES:
    out[f"out_{h}x{w}"] = np.asarray(Image.fromarray(_image(h, w, SIZES.index((h, w)))).resize((224, 224), Image.BILINEAR))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_u8.npz"), **out)
print({k: v.shape for k, v in out.items()})

"""Stand-in for torchmetrics (missing here).  The reference only constructs an SSIM object in
src/classes/CnnAutoEncoder.py:48; SSIM is a training loss and is never evaluated on the scoring path."""
from torch import nn


class StructuralSimilarityIndexMeasure(nn.Module):
    def __init__(self, data_range=1.0, reduction="elementwise_mean", **kwargs):
        super().__init__()
        self.data_range = data_range
        self.reduction = reduction

    def forward(self, preds, target):
        raise NotImplementedError("SSIM is a training-only loss; outside the scoring path")

"""Stand-in for timm==0.6.13 (requirements.txt:176 of the reference; not vendored, not installable here).

TEST INFRASTRUCTURE ONLY.  It restates, from the published timm 0.6.13 source as remembered, exactly the
pieces the reference touches on the scoring path:
  * timm.models.deit_base_distilled_patch16_224  (src/classes/transformer/TransformerEncoder.py:134-136)
  * timm.models.layers.{DropPath, to_2tuple, trunc_normal_} (src/classes/transformer/SwinTransformerModule.py:20)
The DeiT arithmetic is cross-checked against transformers.DeiTModel in tests/test_oracle_cpu.py.
"""
from . import models  # noqa: F401

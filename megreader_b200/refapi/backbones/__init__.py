# mirrors backbones/__init__.py (hot-path factories; resnet families arrive with SURVEY.md §8 rows A10)
from .crnn import crnn_backbone  # noqa: F401

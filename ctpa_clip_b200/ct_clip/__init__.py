from .ct_clip import CTCLIP  # noqa: F401
from .ctvit import CTViT  # noqa: F401

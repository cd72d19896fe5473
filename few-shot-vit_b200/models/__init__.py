"""Drop-in `models` package: same registry names as the reference for the SUN / Visformer hot path
('meta-baseline', 'visformer_micro_80' (+ alias 'visformer'), 'classifier', 'linear-classifier', 'nn-classifier',
'token-label').  Put `few-shot-vit_b200/` on PYTHONPATH ahead of the reference sub-project to switch its scripts over."""
from .models import make, load, register, models  # noqa: F401
from . import visformer  # noqa: F401
from . import meta_baseline  # noqa: F401
from . import classifier  # noqa: F401
from . import token_label  # noqa: F401

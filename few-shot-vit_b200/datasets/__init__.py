"""Drop-in pieces of the reference's `datasets` package that sit next to the hot path: the episode sampler (index work,
bit-exact with the reference under the same numpy seed) and the device-resident image store of the on-device input path
(sunb200/input.py).  The file-backed datasets themselves (pickles, image folders, timm augmentation) are out of scope."""
from .samplers import CategoriesSampler  # noqa: F401
from sunb200.input import DeviceImageStore, preprocess_u8  # noqa: F401

"""Same module name as the reference's model_v2.py: `from model_v2 import MusicAttrRegVAE`."""
from .models import MusicAttrRegVAE  # noqa: F401
from .siblings import MusicAttrCVAE, MusicAttrFaderNets, MusicAttrSingleVAE  # noqa: F401  (model_v2.py:174-586)

"""Same module name as the reference's gmm_model.py: `from gmm_model import MusicAttrRegGMVAE`."""
from .models import MusicAttrRegGMVAE  # noqa: F401

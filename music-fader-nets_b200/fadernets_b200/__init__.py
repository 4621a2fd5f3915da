"""fadernets_b200 -- B200-native (sm_100a) GM-VAE / VAE hot path of Music FaderNets.

Host-side mirror of the reference interface (model classes + trainer step functions) over the
C-ABI CUDA library lib/libfadernets_b200.so.  There is no CPU fallback."""
from ._lib import LIB, LIB_PATH, FaderNetsError, symbols  # noqa: F401
from .models import MusicAttrRegGMVAE, MusicAttrRegVAE  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .siblings import MusicAttrCVAE, MusicAttrFaderNets, MusicAttrSingleVAE  # noqa: F401

EVENT_DIMS, RHYTHM_DIMS, NOTE_DIMS, CHROMA_DIMS = 342, 3, 16, 24      # trainer_gmm.py:35-38

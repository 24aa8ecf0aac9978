"""Import glue for the API mirror: puts the repository root on sys.path (so `multimodal_ssl_avmnist_b200` resolves when the
scripts are run from inside AVMNIST_Experiments/, like the reference's) and picks the Lightning implementation."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

try:                                     # the real package when it is installed
    import lightning.pytorch as pl       # noqa: F401
    from lightning.pytorch.callbacks import ModelCheckpoint, EarlyStopping  # noqa: F401
    from lightning.pytorch.loggers import CSVLogger  # noqa: F401
    HAVE_LIGHTNING = True
except Exception:                        # built-in stand-in (multimodal_ssl_avmnist_b200/pl_shim.py)
    from multimodal_ssl_avmnist_b200 import pl_shim as pl  # noqa: F401
    from multimodal_ssl_avmnist_b200.pl_shim import ModelCheckpoint, EarlyStopping, CSVLogger  # noqa: F401
    HAVE_LIGHTNING = False

"""gtc_b200 -- B200-native (sm_100a) feature front-end of Guitar-Tablature-Classification.

Host side (Python + ctypes) of libgtc.so: CQT segment operator design, device ops, synthetic data, sharding.
The directory that contains this package (``guitar-tablature-classification_b200/``) also holds the drop-in
modules named like the reference's (``cqt``, ``new_cqt``, ``jam_to_tablature``, ``my_dataloader``,
``ViT_dataloader``); put that directory on ``sys.path`` to use them.
"""
from .cqt_design import CqtRecipe, build_operator, get_operator, n_frames_of  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["CqtRecipe", "build_operator", "get_operator", "n_frames_of", "_lib"]

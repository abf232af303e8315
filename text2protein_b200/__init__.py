"""text2protein_b200 -- B200-native drop-in for the PC-sampling hot path of szhan227/text2protein.

Host side: ``text2protein_b200.score_sde_pytorch`` mirrors the reference modules of the same names
(``sampling``, ``sde_lib``, ``utils``, ``models.utils``, ``models.ncsnpp``, ``models.ema``).
Device side: ``libt2p.so`` (C ABI in include/t2p.h; sources in csrc/), hand-written for sm_100a.
Put this directory on ``sys.path`` to let existing ``from score_sde_pytorch import sampling`` imports resolve
to the native implementation.
"""
from .config import AttrDict, load_config  # noqa: F401

__all__ = ["AttrDict", "load_config"]

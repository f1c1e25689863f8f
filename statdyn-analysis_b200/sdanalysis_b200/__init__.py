"""sdanalysis_b200 -- B200-native drop-in for the per-frame dynamics hot path of ``sdanalysis``.

Module names mirror the reference package (``dynamics``, ``order``, ``read``, ``StepSize``,
``frame``, ``molecules``) so that ``from sdanalysis_b200 import dynamics, order, read`` replaces
``from sdanalysis import dynamics, order, read`` for the hot path.

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI library
``libsdb200.so`` (``include/sdb200.h``); there is no CPU fallback.  The library is loaded on first
use and a missing library raises ``ImportError``.
"""

__version__ = "0.1.0"

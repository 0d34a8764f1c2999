"""pyblip_b200 -- B200-native implementation of pyblip's data-parallel hot path.

Same Python API as amspector100/pyblip for the multi-chain Gibbs samplers and the
candidate-group PIP counting; everything numeric runs in libblipgpu.so
(hand-written CUDA for sm_100a).  There is no CPU fallback.
"""
from . import linear, probit, nprior, create_groups, create_groups_cts, changepoint, weight_fns, blip
from .linear import LinearSpikeSlab
from .probit import ProbitSpikeSlab
from .nprior import NPrior
from .blip import BLiP, BLiP_cts

name = "pyblip_b200"
__version__ = "0.1.0"

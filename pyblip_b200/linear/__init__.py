""" Submodule of classes implementing the linear spike slab Gibbs sampler (B200). """
from . import linear
from .linear import LinearSpikeSlab

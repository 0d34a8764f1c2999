""" Submodule of classes implementing the probit spike slab Gibbs sampler (B200). """
from . import probit
from .probit import ProbitSpikeSlab

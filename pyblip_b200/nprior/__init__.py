""" Submodule implementing the neuronized-prior sampler (B200). """
from . import nprior
from .nprior import NPrior

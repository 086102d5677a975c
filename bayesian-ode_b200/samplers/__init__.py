"""Samplers of the reference's ``samplers/`` package rebuilt as fused CUDA updates over particle-batched chains."""
from .sampler import ChainStore, Sampler
from .langevin import HAMCMC, SGLD, pSGLD
from .hamiltonian import aSGHMC
from .stein import RBFKernel, SVGD

__all__ = ["Sampler", "ChainStore", "SGLD", "pSGLD", "HAMCMC", "aSGHMC", "SVGD", "RBFKernel"]

"""Samplers of the reference's ``samplers/`` package rebuilt as fused CUDA updates over particle-batched chains."""
from .sampler import ChainStore, Sampler
from .langevin import HAMCMC, MALA, SGLD, cSGLD, pSGLD
from .hamiltonian import aSGHMC, acSGHMC
from .stein import RBFKernel, SVGD

__all__ = ["Sampler", "ChainStore", "SGLD", "pSGLD", "cSGLD", "MALA", "HAMCMC", "aSGHMC", "acSGHMC", "SVGD", "RBFKernel"]

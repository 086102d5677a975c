"""Samplers of the reference's ``samplers/`` package rebuilt as fused CUDA updates over particle-batched chains."""
from .sampler import ChainStore, Sampler
from .langevin import HAMCMC, HAMCMC2, HAMCMC3, HAMCMC4, MALA, SGLD, cSGLD, pSGLD
from .hamiltonian import aSGHMC, acSGHMC
from .stein import RBFKernel, SVGD

__all__ = ["Sampler", "ChainStore", "SGLD", "pSGLD", "cSGLD", "MALA", "HAMCMC", "HAMCMC2", "HAMCMC3", "HAMCMC4", "aSGHMC", "acSGHMC", "SVGD", "RBFKernel"]

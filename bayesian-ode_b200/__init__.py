"""bayesian_ode_b200 -- the B200-native (sm_100a) hot path of jaivardhankapoor/bayesian-ode.

Public surface mirrors the reference:
  odeint, odeint_adjoint          (torchdiffeq/__init__.py:1-2)
  NPDEField / KernelRegression    (scripts/vanderpol/gp.py:56-71)
  NPDEPosterior                   (loss_closure, gp.py:342-353)
  samplers.*                      (samplers/{langevin,hamiltonian,stein}.py)
  posterior_predictive            (ensemble re-integration of a chain, gp.py:440-464)
  driver.run_sampler / worker     (JSON configuration + data pickle driver, gp.py:290-391, 504-564)
All compute runs in hand-written CUDA behind the C ABI of include/bode_b200.h; there is no CPU path.
"""
from . import _lib
from .fields import KernelRegression, MLPField, NPDEField, rbf_kernel
from .odeint import TupleField, last_dopri5_stats, odeint, odeint_adjoint
from .posterior import MLPPosterior, NPDEPosterior
from . import samplers
from .predictive import ensemble_trajectories, posterior_predictive
from . import driver

__all__ = ["odeint", "odeint_adjoint", "TupleField", "NPDEField", "KernelRegression", "NPDEPosterior", "MLPField", "MLPPosterior", "rbf_kernel", "samplers", "ensemble_trajectories", "posterior_predictive", "driver", "_lib"]

"""classpp_public_b200: B200-native (sm_100a) implementation of the CLASS++ hot path
(perturbations -> transfer -> spectra) behind the reference's module interface.
See DESIGN.md / INTEGRATION.md."""
from .modules import (AnalyticPrimordial, BackgroundModule, Context, CosmoComputationError,  # noqa: F401
                      CosmoSevereError, Inputs, PerturbationsModule, SpectraModule, TabulatedPrimordial,
                      ThermodynamicsModule, TransferModule)

__all__ = ["Inputs", "Context", "BackgroundModule", "ThermodynamicsModule", "PerturbationsModule",
           "TransferModule", "SpectraModule", "AnalyticPrimordial", "TabulatedPrimordial",
           "CosmoComputationError", "CosmoSevereError"]

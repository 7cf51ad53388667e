"""composablestatespacemodels_b200 -- the B200 (sm_100a) particle-filter hot path of
jonnylaw/ComposableStateSpaceModels behind the reference's own API names.

Host side (this package, Python because no JVM exists in the build image -- see DESIGN.md):
model/parameter composition, the ParticleFilter / Resampling / PMMH interfaces.
Device side (csrc/, hand-written CUDA behind the C ABI of include/cssm.h): everything that
touches a particle.  There is no CPU fallback.
"""
from . import _abi
from .tree import Tree, Leaf, Branch
from .sde import Sde, SdeParameter, BrownianParameter, GenBrownianParameter, OuParameter
from .parameters import ParamNode, Parameters, flattenParams, perturb, perturbMvn
from . import model as Model
from .model import UnparamModel
from .resampling import Resampling
from .filter import (Data, TimedObservation, StateSpace, PfState, PfOut, ForecastOut, ObservationWithState, CredibleInterval,
                     Filter, FilterLgcp, FilterInit, FilterInterpolate, PfStateInterpolate, ParticleFilter,
                     GpuFilterHandle, ShardedGroup, StaleStateError)
from .pmmh import (MetropolisHastings, ParticleMetropolisHastings, ApproxPMMH, approxPmmh, pmmhStep, MetropState,
                   GpuBootstrapFilter, runChains)
from . import streaming as Streaming

F32, F64 = _abi.F32, _abi.F64
__all__ = ["Tree", "Leaf", "Branch", "Sde", "SdeParameter", "BrownianParameter", "GenBrownianParameter", "OuParameter",
           "ParamNode", "Parameters", "flattenParams", "perturb", "perturbMvn", "Model", "UnparamModel", "Resampling",
           "Data", "TimedObservation", "StateSpace", "PfState", "PfOut", "ForecastOut", "ObservationWithState", "CredibleInterval", "Filter", "FilterLgcp", "FilterInit", "FilterInterpolate", "PfStateInterpolate", "ParticleFilter",
           "GpuFilterHandle", "ShardedGroup", "MetropolisHastings", "ParticleMetropolisHastings", "ApproxPMMH", "approxPmmh", "pmmhStep", "MetropState",
           "GpuBootstrapFilter", "runChains", "StaleStateError", "Streaming",
           "F32", "F64"]

"""rmsb200 -- Python host mirror of the ReactionMechanismSimulator.jl API surface for the batched
stiff-kinetics hot path, over the C ABI of librmsb200.so (include/rmsb200.h).

The reference is a Julia package; Julia is not available in the build image, so this package is the
executable host side (same names, argument meaning and error behaviour), and `julia/RMSB200.jl`
holds the `ccall` shim a Julia user would load.  There is no CPU fallback: every compute call goes
to the CUDA library and raises if it is missing.
"""
from ._lib import (DOMAIN_CONSTANT_P, DOMAIN_CONSTANT_TAPHI, DOMAIN_CONSTANT_TP, DOMAIN_CONSTANT_TV, DOMAIN_CONSTANT_V, LIB_PATH, RMSBError,
                   lib)
from .coupled import CoupledEnsemble, CoupledPhase, SystemSimulation, coupled_domain
from .coupled import make_interface as ReactiveInternalInterfaceConstantTPhi
from .domain import (ConstantPDomain, ConstantPEnsemble, ConstantTAPhiDomain, ConstantTPDomain, ConstantTPEnsemble, ConstantTVDomain, ConstantTVEnsemble,
                     ConstantVDomain, ConstantVEnsemble, Domain, ParametrizedPDomain, ParametrizedTConstantVDomain, ParametrizedTPDomain,
                     ParametrizedVDomain)
from .interfaces import Inlet, Outlet, VolumeMaintainingOutlet, VolumetricFlowRateInlet, VolumetricFlowRateOutlet, pack_interfaces
from .ensemble import balanced_shards, shard_bounds, shard_domain, solve_sharded
from .mechanism import (ElementaryReaction, NASA, NASApolynomial, R, Solvent, Species, getmechdict, getspeciesradius, readinput, writeyml)
from .profile import spline_table
from .phase import IdealDiluteSolution, IdealGas, IdealSurface, Phase
from .reactor import RMSB200, Reactor, Solution, solve, threadedsensitivities
from .simulation import Simulation, idt_grid, ignition_delay
from .transitory import getdividingtimescale, transitorysensitivitiesfulltrapezoidal

__all__ = [n for n in dir() if not n.startswith("_")]

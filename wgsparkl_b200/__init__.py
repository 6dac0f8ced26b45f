"""wgsparkl_b200 — B200-native (sm_100a) implementation of wgsparkl's MPM substep hot path.

The package holds only what that path needs: the CUDA kernels + C ABI (csrc/, built into
libb200mpm.so) and a host-side mirror of the reference's Rust interface for the path
(`MpmPipeline`, `MpmData`, `Particle`, `SimulationParams`, ...). There is no CPU fallback:
constructing an `MpmPipeline` without the compiled library or without a CUDA device raises.
"""
from . import abi  # noqa: F401
from .models import DruckerPrager, DruckerPragerPlasticState, ElasticCoefficients  # noqa: F401
from .solver import (  # noqa: F401
    Cdf,
    Particle,
    ParticleDynamics,
    ParticlePhase,
    SimulationParams,
    make_particles,
    particles_to_abi,
)

__all__ = [
    "abi",
    "DruckerPrager",
    "DruckerPragerPlasticState",
    "ElasticCoefficients",
    "Cdf",
    "Particle",
    "ParticleDynamics",
    "ParticlePhase",
    "SimulationParams",
    "make_particles",
    "particles_to_abi",
]

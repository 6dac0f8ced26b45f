"""Host-side mirror of wgsparkl's `solver` data structs (src/solver/particle{2,3}d.rs,
params.rs, particle_update.rs) and vectorised builders for large synthetic scenes."""
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import abi
from .models import DruckerPrager, DruckerPragerPlasticState, ElasticCoefficients

f32 = np.float32
F32_MAX = float(np.finfo(np.float32).max)


@dataclass
class SimulationParams:
    """solver/params.rs:6-16."""

    gravity: Sequence[float]
    dt: float

    def to_abi(self):
        out = np.zeros((), dtype=abi.sim_params_dtype)
        g = np.zeros(3, dtype=f32)
        g[: len(self.gravity)] = np.asarray(self.gravity, dtype=f32)
        out["gravity"] = g
        out["dt"] = f32(self.dt)
        return out


@dataclass(frozen=True)
class ParticlePhase:
    """solver/particle_update.rs:37-42."""

    phase: float
    max_stretch: float


@dataclass
class Cdf:
    """solver/particle3d.rs:43-50."""

    normal: Sequence[float] = (0.0, 0.0, 0.0)
    rigid_vel: Sequence[float] = (0.0, 0.0, 0.0)
    signed_distance: float = 0.0
    affinity: int = 0


@dataclass
class ParticleDynamics:
    """solver/particle3d.rs:16-42."""

    velocity: np.ndarray
    def_grad: np.ndarray
    affine: np.ndarray
    cdf: Cdf
    init_volume: float
    init_radius: float
    mass: float

    @staticmethod
    def with_density(radius, density, dim=3):
        init_volume = f32(f32(radius) * f32(2.0)) ** dim  # powi(exponent)
        init_volume = f32(init_volume)
        return ParticleDynamics(
            velocity=np.zeros(dim, dtype=f32),
            def_grad=np.eye(dim, dtype=f32),
            affine=np.zeros((dim, dim), dtype=f32),
            cdf=Cdf(),
            init_volume=float(init_volume),
            init_radius=float(f32(radius)),
            mass=float(f32(init_volume * f32(density))),
        )


@dataclass
class Particle:
    """solver/particle3d.rs:53-60."""

    position: Sequence[float]
    dynamics: ParticleDynamics
    model: ElasticCoefficients
    plasticity: Optional[DruckerPrager] = None
    phase: Optional[ParticlePhase] = None
    model_kind: int = abi.MODEL_COROTATED  # additive (see include/b200mpm.h)


def particles_to_abi(particles: Sequence[Particle], dim: int) -> np.ndarray:
    """Flatten `&[Particle]` the way GpuParticles::from_particles / GpuModels::from_particles do
    (particle3d.rs:192-210, models/mod.rs:20-49), resolving the `Option`s identically."""
    out = np.zeros(len(particles), dtype=abi.particle_dtype)
    default_dp = DruckerPrager.new(-1.0, -1.0)
    default_state = DruckerPragerPlasticState()
    for i, p in enumerate(particles):
        o = out[i]
        o["position"][:dim] = np.asarray(p.position, dtype=f32)[:dim]
        d = p.dynamics
        o["velocity"][:dim] = np.asarray(d.velocity, dtype=f32)[:dim]
        # nalgebra matrices are column-major: flatten by columns.
        o["def_grad"][: dim * dim] = np.asarray(d.def_grad, dtype=f32)[:dim, :dim].T.reshape(-1)
        o["affine"][: dim * dim] = np.asarray(d.affine, dtype=f32)[:dim, :dim].T.reshape(-1)
        o["cdf_normal"][:dim] = np.asarray(d.cdf.normal, dtype=f32)[:dim]
        o["cdf_rigid_vel"][:dim] = np.asarray(d.cdf.rigid_vel, dtype=f32)[:dim]
        o["cdf_signed_distance"] = d.cdf.signed_distance
        o["cdf_affinity"] = d.cdf.affinity
        o["init_volume"] = d.init_volume
        o["init_radius"] = d.init_radius
        o["mass"] = d.mass
        o["lambda"] = p.model.lambda_
        o["mu"] = p.model.mu
        dp = p.plasticity if p.plasticity is not None else default_dp
        o["dp_h0"], o["dp_h1"], o["dp_h2"], o["dp_h3"] = dp.h0, dp.h1, dp.h2, dp.h3
        o["dp_lambda"], o["dp_mu"] = dp.lambda_, dp.mu
        o["plastic_det"] = default_state.plastic_deformation_gradient_det
        o["plastic_hardening"] = default_state.plastic_hardening
        o["plastic_log_vol_gain"] = default_state.log_vol_gain
        ph = p.phase if p.phase is not None else ParticlePhase(0.0, -1.0)
        o["phase"], o["max_stretch"] = ph.phase, ph.max_stretch
        o["model"] = p.model_kind
    return out


def make_particles(
    positions: np.ndarray,
    dim: int,
    radius: float,
    density: float,
    model: ElasticCoefficients,
    plasticity: Optional[DruckerPrager] = None,
    phase: Optional[ParticlePhase] = None,
    model_kind: int = abi.MODEL_COROTATED,
    velocity: Optional[np.ndarray] = None,
) -> np.ndarray:
    """Vectorised equivalent of building `Vec<Particle>` with one material
    (the pattern of every reference scene, e.g. crates/wgsparkl3d/examples/sand3.rs:34-52)."""
    n = positions.shape[0]
    out = np.zeros(n, dtype=abi.particle_dtype)
    out["position"][:, :dim] = positions[:, :dim].astype(f32)
    if velocity is not None:
        out["velocity"][:, :dim] = velocity[:, :dim].astype(f32)
    dyn = ParticleDynamics.with_density(radius, density, dim)
    eye = np.zeros(9, dtype=f32)
    eye[: dim * dim] = np.eye(dim, dtype=f32).reshape(-1)
    out["def_grad"][:] = eye
    out["init_volume"] = dyn.init_volume
    out["init_radius"] = dyn.init_radius
    out["mass"] = dyn.mass
    out["lambda"] = model.lambda_
    out["mu"] = model.mu
    dp = plasticity if plasticity is not None else DruckerPrager.new(-1.0, -1.0)
    out["dp_h0"], out["dp_h1"], out["dp_h2"], out["dp_h3"] = dp.h0, dp.h1, dp.h2, dp.h3
    out["dp_lambda"], out["dp_mu"] = dp.lambda_, dp.mu
    st = DruckerPragerPlasticState()
    out["plastic_det"] = st.plastic_deformation_gradient_det
    out["plastic_hardening"] = st.plastic_hardening
    out["plastic_log_vol_gain"] = st.log_vol_gain
    ph = phase if phase is not None else ParticlePhase(0.0, -1.0)
    out["phase"], out["max_stretch"] = ph.phase, ph.max_stretch
    out["model"] = model_kind
    return out

"""Host-side mirror of wgsparkl's `models` module (src/models/mod.rs, drucker_prager.rs).

All arithmetic is done in float32 like the Rust original so that the Lamé parameters uploaded
to the device are bit-identical to what the reference would upload.
"""
from dataclasses import dataclass

import numpy as np

f32 = np.float32


def lame_lambda_mu(young_modulus, poisson_ratio):
    """models/mod.rs:52-61."""
    e, nu = f32(young_modulus), f32(poisson_ratio)
    one, two = f32(1.0), f32(2.0)
    lam = e * nu / ((one + nu) * (one - two * nu))
    mu = e / (two * (one + nu))
    return f32(lam), f32(mu)


@dataclass(frozen=True)
class ElasticCoefficients:
    """models/mod.rs:63-75."""

    lambda_: float
    mu: float

    @staticmethod
    def from_young_modulus(young_modulus, poisson_ratio):
        lam, mu = lame_lambda_mu(young_modulus, poisson_ratio)
        return ElasticCoefficients(float(lam), float(mu))


@dataclass(frozen=True)
class DruckerPrager:
    """models/drucker_prager.rs:6-33."""

    h0: float
    h1: float
    h2: float
    h3: float
    lambda_: float
    mu: float

    @staticmethod
    def new(young_modulus, poisson_ratio):
        if young_modulus > 0.0:
            lam, mu = lame_lambda_mu(young_modulus, poisson_ratio)
        else:
            lam, mu = f32(-1.0), f32(-1.0)
        rad = lambda deg: float(f32(np.deg2rad(np.float64(deg))))  # f32::to_radians
        return DruckerPrager(rad(35.0), rad(9.0), float(f32(0.2)), rad(10.0), float(lam), float(mu))


@dataclass(frozen=True)
class DruckerPragerPlasticState:
    """models/drucker_prager.rs:36-52 (Default)."""

    plastic_deformation_gradient_det: float = 1.0
    plastic_hardening: float = 1.0
    log_vol_gain: float = 0.0

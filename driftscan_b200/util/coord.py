"""Spherical-geometry helpers the boundary needs on the host (the reference takes
them from ``cora.util.coord``, an external package)."""

import numpy as np


def sph_to_cart(sph):
    """(theta, phi) -> Cartesian unit vectors (last axis)."""
    sph = np.asarray(sph, dtype=np.float64)
    theta, phi = sph[..., -2], sph[..., -1]
    sin_t = np.sin(theta)
    return np.stack([sin_t * np.cos(phi), sin_t * np.sin(phi), np.cos(theta)], axis=-1)


def thetaphi_plane_cart(sph):
    """Unit vectors (theta-hat, phi-hat) of the tangent plane at each position."""
    sph = np.asarray(sph, dtype=np.float64)
    theta, phi = sph[..., -2], sph[..., -1]
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    that = np.stack([ct * cp, ct * sp, -st], axis=-1)
    phat = np.stack([-sp, cp, np.zeros_like(sp)], axis=-1)
    return that, phat


def sph_dot(a, b):
    """Dot product of directions given in spherical polars."""
    return np.inner(sph_to_cart(a), sph_to_cart(b))


def norm_vec2(vec):
    """Normalise 2-vectors (last axis) in place; zero vectors are left alone."""
    length = np.hypot(vec[..., 0], vec[..., 1])
    length = np.where(length == 0.0, 1.0, length)
    vec /= length[..., np.newaxis]
    return vec

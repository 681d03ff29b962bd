"""Initial particle states with the reference's sampling laws (host side of the model, numpy only).

    random_positions   triangulatedMeshSpace::randomPosition      (src/models/triangulatedMeshSpace.cpp:108-116)
    random_velocities  triangulatedMeshSpace::randomVectorAtPosition (:118-138) x sqrt(T)
                       (simpleModel::setMaxwellBoltzmannVelocities, src/models/simpleModel.cpp:223-230)
    interaction_range  curvedSpaceSimulation.cpp:75-76

The random stream is numpy's, not the reference's mt19937 wrapper (src/utility/noiseSource.cpp), so states are
statistically, not bitwise, those of the reference mains."""
from __future__ import annotations

import math

import numpy as np

from . import meshes


def random_positions(nF, N, rng):
    """u~U(0,1), v~U(0,1-u), w=1-u-v, face ~ U{0..F-1} (uniform over faces, not over area)."""
    u = rng.random(N)
    v = rng.random(N) * (1 - u)
    face = rng.integers(0, nF, N).astype(np.int32)
    return face, np.stack([u, v, 1 - u - v], axis=1)


def random_velocities(V, corners, face, T, rng):
    """In-plane Gaussian vector at every particle, times sqrt(T)."""
    p0, p1, p2 = V[corners[face, 0]], V[corners[face, 1]], V[corners[face, 2]]
    n = np.cross(p1 - p0, p2 - p0)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    special = (n[:, 0] == 0) & (n[:, 1] == 0)
    o = np.where(special[:, None], np.stack([n[:, 1] - n[:, 2], n[:, 2] - n[:, 0], n[:, 0] - n[:, 1]], 1),
                 np.stack([n[:, 1], -n[:, 0], np.zeros(len(n))], 1))
    o /= np.linalg.norm(o, axis=1, keepdims=True)
    t2 = np.cross(n, o)
    g1 = rng.standard_normal(len(n))[:, None]
    g2 = rng.standard_normal(len(n))[:, None]
    return (g1 * o + g2 * t2) * math.sqrt(T)


def interaction_range(area, N, area_fraction=0.9):
    return 2 * math.sqrt(area_fraction * area / (N * math.pi))


def make_state(V, F, N, seed=13377, T=0.2):
    """(corners in reference order, face, bary, velocities) for N particles on the OFF mesh (V, F)."""
    corners = meshes.reference_corners(F)
    rng = np.random.default_rng(seed)
    face, bary = random_positions(len(F), N, rng)
    vel = random_velocities(V, corners, face, T, rng)
    return corners, face, bary, vel

"""Seeded synthetic initial conditions restating the *distributions* of physim's generators.

These are NOT bit-exact with the reference's ChaCha8Rng streams (rand 0.9.1 / rand_chacha 0.9.0 are
not available offline and no reference test pins generator output); parity runs feed the same
arrays to the oracle and to the GPU, so only the distribution matters.

  cube  — astro/src/initialisers.rs:82-106 over Entity::random (physim-core/src/lib.rs:115-128)
  star  — astro/src/initialisers.rs:154-174
  solar — astro/src/initialisers.rs:387-516
"""
import numpy as np

from .entity import entities


def cube(n, seed=0, spin=0.0, mass=1.0, size=1.0, centre=(0.0, 0.0, 0.0), id=0):
    rng = np.random.default_rng(seed)
    e = entities(n)
    xyz = rng.random((n, 3))
    x = (xyz[:, 0] * 2.0 - 1.0) * size      # random_range(-1.0..1.0)
    y = (xyz[:, 1] * 2.0 - 1.0) * size
    z = xyz[:, 2] * size                    # random_range(0.0..1.0)
    e["vx"] = y * spin
    e["vy"] = -x * spin
    e["x"] = x + centre[0]
    e["y"] = y + centre[1]
    e["z"] = z + centre[2]
    e["mass"] = mass / n if n else 0.0
    e["radius"] = 0.02
    e["id"] = id
    return e


def star(x=0.0, y=0.0, z=0.0, vx=0.0, vy=0.0, vz=0.0, mass=0.0, radius=0.1, id=0, fixed=False):
    e = entities(1)
    e["x"], e["y"], e["z"] = x, y, z
    e["vx"], e["vy"], e["vz"] = vx, vy, vz
    e["mass"], e["radius"], e["id"], e["fixed"] = mass, radius, id, fixed
    return e


def solar(planets=8, asteroids=50, seed=2):
    rng = np.random.default_rng(seed)
    sun = star(z=0.5, mass=1.0, radius=0.1, fixed=True)
    m = rng.lognormal(1.1, 0.1, planets)
    r = rng.lognormal(0.2, 0.7, planets)
    th = rng.uniform(0.0, 2.0 * np.pi, planets)
    p = entities(planets)
    p["x"], p["y"], p["z"] = r * np.sin(th), r * np.cos(th), 0.5
    p["vx"], p["vy"] = -np.cos(th) / np.sqrt(r), np.sin(th) / np.sqrt(r)
    p["radius"], p["mass"] = 0.05, m * 1e-5
    moons = p.copy()
    moons["mass"] /= 10.0
    moons["x"] += 0.01
    moons["vy"] += np.sqrt(moons["mass"] / 0.01)
    moons["radius"] = 0.005
    a = rng.lognormal(0.0, 0.9, asteroids)
    th = rng.uniform(0.0, 2.0 * np.pi, asteroids)
    ecc = rng.uniform(0.0, 0.6, asteroids)
    phi = rng.uniform(0.0, 2.0 * np.pi, asteroids)
    b = a * np.sqrt(1.0 - ecc * ecc)
    x0, y0 = a * np.cos(th), b * np.sin(th)
    r_cur = np.sqrt(x0 * x0 + y0 * y0)
    v_mag = np.sqrt(2.0 / r_cur - 1.0 / a)
    dx, dy = -a * np.sin(th), b * np.cos(th)
    nrm = np.sqrt(dx * dx + dy * dy)
    vx0, vy0 = v_mag * dx / nrm, v_mag * dy / nrm
    c, s = np.cos(phi), np.sin(phi)
    ast = entities(asteroids)
    ast["x"], ast["y"], ast["z"] = x0 * c - y0 * s, x0 * s + y0 * c, 0.5
    ast["vx"], ast["vy"] = vx0 * c - vy0 * s, vx0 * s + vy0 * c
    ast["radius"], ast["mass"] = 0.01, 1e-7
    return np.concatenate([sun, p, moons, ast])


def readme_pipeline(n=100_000, seed=1, spin=1000.0):
    """BASELINE config 1: cube + 2 stars (readme.md:39)."""
    return np.concatenate([
        cube(n, seed=seed, spin=spin),
        star(x=0.2, y=0.2, z=0.5, mass=1e5, radius=0.1),
        star(x=-0.2, y=-0.2, z=0.5, mass=1e5, radius=0.1),
    ])


def headline_pipeline(n=1_000_000, seed=1, spin=500.0):
    """BASELINE config 3: cube + 4 stars (readme.md:79)."""
    return np.concatenate([
        cube(n, seed=seed, spin=spin),
        star(x=0.1, y=0.1, z=0.5, mass=1e5, radius=0.1),
        star(x=-0.1, y=-0.1, z=0.5, mass=1e5, radius=0.1),
        star(x=-0.1, y=0.1, z=0.5, mass=1e5),
        star(x=0.1, y=-0.1, z=0.5, mass=1e5),
    ])

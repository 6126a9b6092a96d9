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


# ---- `cube` on the reference's own random stream ------------------------------
# ChaCha8Rng::seed_from_u64(seed) (astro/src/initialisers.rs:85) -> Entity::random (physim-core/src/
# lib.rs:115-128): x, y = random_range(-1.0..1.0), z = random_range(0.0..1.0).  Restated from the
# published algorithms of the pinned dependencies (Cargo.lock:1285-1317; their sources are not under
# /root/reference):
#   rand_chacha 0.9.0   ChaCha, 8 rounds, 256-bit key = seed, 64-bit block counter from 0 (words 12-13),
#                       64-bit stream id 0 (words 14-15); output words in block order
#   rand_core 0.9.0     SeedableRng::seed_from_u64: the 32-byte seed is eight PCG32 outputs
#                       (state = state * 6364136223846793005 + 11634580027462260723, then XSH-RR)
#                       BlockRng::next_u64: two consecutive u32 words, low word first
#   rand 0.9.1          UniformFloat::sample_single: (next_u64 >> 12) as the mantissa of a double in
#                       [1, 2), minus 1, times (high - low), plus low  (never rejects for these ranges)
# PINNING: the ChaCha block function is checked against the published ChaCha8 / ChaCha20 known-answer
# vectors (tests/test_generators.py).  Seeding and the float conversion have no vector available
# offline and no reference test pins generator output: "parity unpinned" for those two steps until a
# `physim cube ... ! csvsink` file from a real build is added under tests/golden/.
_CHACHA_CONST = (0x61707865, 0x3320646E, 0x79622D32, 0x6B206574)


def chacha_blocks(key_words, counters, rounds=8):
    """ChaCha blocks for 64-bit block counters `counters`, zero stream id: uint32 [len(counters), 16]."""
    c = np.asarray(counters, dtype=np.uint64)
    st = np.zeros((16, len(c)), dtype=np.uint32)
    for i in range(4):
        st[i] = _CHACHA_CONST[i]
    for i in range(8):
        st[4 + i] = key_words[i]
    st[12] = (c & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    st[13] = (c >> np.uint64(32)).astype(np.uint32)
    x = st.copy()

    def rotl(v, r):
        return (v << np.uint32(r)) | (v >> np.uint32(32 - r))

    def qr(a, b, cc, d):
        x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 16)
        x[cc] += x[d]; x[b] ^= x[cc]; x[b] = rotl(x[b], 12)
        x[a] += x[b]; x[d] ^= x[a]; x[d] = rotl(x[d], 8)
        x[cc] += x[d]; x[b] ^= x[cc]; x[b] = rotl(x[b], 7)

    with np.errstate(over="ignore"):
        for _ in range(rounds // 2):
            qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
            qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
        return (x + st).T.copy()


def seed_from_u64(seed):
    """rand_core 0.9 SeedableRng::seed_from_u64: eight PCG32 outputs = the 256-bit ChaCha key (LE words)."""
    mask = (1 << 64) - 1
    state = int(seed) & mask
    words = []
    for _ in range(8):
        state = (state * 6364136223846793005 + 11634580027462260723) & mask
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF)
    return words


def chacha8_u64(seed, first, count):
    """next_u64() outputs number first .. first+count of ChaCha8Rng::seed_from_u64(seed)."""
    key = seed_from_u64(seed)
    w0, w1 = 2 * first, 2 * (first + count)
    b0, b1 = w0 // 16, (w1 + 15) // 16
    words = chacha_blocks(key, np.arange(b0, b1, dtype=np.uint64)).reshape(-1)[w0 - 16 * b0: w1 - 16 * b0]
    return words[0::2].astype(np.uint64) | (words[1::2].astype(np.uint64) << np.uint64(32))


def _unit_from_u64(u):
    """rand 0.9.1 UniformFloat: 52 random mantissa bits -> [1, 2) -> [0, 1)."""
    return ((u >> np.uint64(12)) | np.uint64(1023 << 52)).view(np.float64) - 1.0


def cube_chacha8(n, seed=0, spin=0.0, mass=1.0, size=1.0, centre=(0.0, 0.0, 0.0), id=0, chunk=1 << 20):
    """`cube n=.. seed=.. spin=..` with the reference's random stream and operation order
    (initialisers.rs:82-106); see the pinning note above."""
    e = entities(n)
    for i0 in range(0, n, chunk):
        m = min(chunk, n - i0)
        u = _unit_from_u64(chacha8_u64(seed, 3 * i0, 3 * m)).reshape(m, 3)
        x = (u[:, 0] * 2.0 + -1.0) * size          # value0_1 * scale + low, then e.x *= size
        y = (u[:, 1] * 2.0 + -1.0) * size
        z = (u[:, 2] * 1.0 + 0.0) * size
        sl = slice(i0, i0 + m)
        e["vx"][sl] = y * spin
        e["vy"][sl] = -x * spin
        e["x"][sl] = x + centre[0]
        e["y"][sl] = y + centre[1]
        e["z"][sl] = z + centre[2]
    e["mass"] = mass / n if n else 0.0
    e["radius"] = 0.02
    e["id"] = id
    return e


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

"""Input generators: the ChaCha8 core of `cube_chacha8` against the published known-answer vectors,
and the properties of the restated `cube` stream (host code, no GPU)."""
import numpy as np

from physim_b200 import generators as gen

# draft-strombergson-chacha-test-vectors TC1 (all-zero 256-bit key and IV), first block
CHACHA8_TC1 = ("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
               "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42")
CHACHA20_TC1 = ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")


def test_chacha_block_function_known_answers():
    assert gen.chacha_blocks([0] * 8, [0], 8)[0].astype("<u4").tobytes().hex() == CHACHA8_TC1
    assert gen.chacha_blocks([0] * 8, [0], 20)[0].astype("<u4").tobytes().hex() == CHACHA20_TC1
    # the block counter advances word 12 (and carries into word 13)
    a = gen.chacha_blocks([1, 2, 3, 4, 5, 6, 7, 8], [0, 1, 2 ** 32 - 1, 2 ** 32], 8)
    assert len({r.tobytes() for r in a}) == 4


def test_stream_is_position_addressable():
    whole = gen.chacha8_u64(7, 0, 1000)
    assert np.array_equal(gen.chacha8_u64(7, 123, 456), whole[123:579])
    assert not np.array_equal(gen.chacha8_u64(8, 0, 1000), whole)
    assert len(set(gen.seed_from_u64(0))) == 8 and gen.seed_from_u64(0) != gen.seed_from_u64(1)


def test_cube_chacha8_follows_the_reference_operations():
    n = 100_000
    e = gen.cube_chacha8(n, seed=1, spin=1000.0)
    assert (e["x"] >= -1).all() and (e["x"] < 1).all() and (e["y"] >= -1).all() and (e["y"] < 1).all()
    assert (e["z"] >= 0).all() and (e["z"] < 1).all()
    assert np.array_equal(e["vx"], e["y"] * 1000.0) and np.array_equal(e["vy"], -e["x"] * 1000.0)
    assert (e["vz"] == 0).all() and (e["mass"] == 1.0 / n).all() and (e["radius"] == 0.02).all()
    for k, lo, hi in (("x", -1, 1), ("y", -1, 1), ("z", 0, 1)):
        assert abs(e[k].mean() - (lo + hi) / 2) < 0.01 and abs(e[k].std() - (hi - lo) / np.sqrt(12)) < 0.01
    # values are multiples of 2^-51 (52 mantissa bits mapped to [-1, 1)): the float conversion is exact
    assert np.array_equal(e["x"] * 2.0 ** 51, np.round(e["x"] * 2.0 ** 51))
    # chunked generation == one shot, and a prefix of a longer cube is the shorter cube
    assert gen.cube_chacha8(5000, seed=3, chunk=777).tobytes() == gen.cube_chacha8(5000, seed=3).tobytes()
    a, b = gen.cube_chacha8(1000, seed=3, mass=1000.0), gen.cube_chacha8(2000, seed=3, mass=2000.0)
    assert a.tobytes() == b[:1000].tobytes()

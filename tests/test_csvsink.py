"""csvsink (utilities/src/csvsink.rs:44-80): the host-side C++ sink against the Python restatement
of the reference (oracle/binding.py) — number format, line layout, print_n selection.  Host code
only: no GPU needed."""
import math
import struct

import numpy as np
import pytest

from oracle import binding as ob
from physim_b200 import api, generators as gen

SPECIAL = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1e-7, 1.5e-7, 123456789.125, 1e21, 1e22, 1.7976931348623157e308,
           5e-324, 2.2250738585072014e-308, 1 / 3, 2 / 3, 1e15, 1e16, 1e17, 0.30000000000000004,
           float("inf"), float("-inf"), float("nan"), 4.35, 1e-5, 9007199254740993.0, -2.5e-9]


def test_number_format_is_rust_display():
    # known Rust outputs: println!("{}", x)
    known = {1.0: "1", 0.1: "0.1", 1e-7: "0.0000001", 1e21: "1000000000000000000000", -0.0: "-0",
             1.5: "1.5", 1e16: "10000000000000000", 0.30000000000000004: "0.30000000000000004",
             float("inf"): "inf", float("-inf"): "-inf"}
    for v, want in known.items():
        assert api.csv_format_f64(v) == want, v
        assert ob.rust_display_f64(v) == want, v
    assert api.csv_format_f64(float("nan")) == "NaN" == ob.rust_display_f64(float("nan"))
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2 ** 63, 3000, dtype=np.uint64) | (rng.integers(0, 2, 3000, dtype=np.uint64) << np.uint64(63))
    vals = SPECIAL + [struct.unpack("<d", struct.pack("<Q", int(b)))[0] for b in bits] + list(rng.normal(size=500))
    for v in vals:
        got = api.csv_format_f64(v)
        assert got == ob.rust_display_f64(v), (v, got)
        assert "e" not in got and "E" not in got
        if math.isfinite(v):
            assert float(got) == v                      # round-trips


@pytest.mark.parametrize("print_n", [1, 3, 5])
def test_file_layout_and_print_n(tmp_path, print_n):
    rng = np.random.default_rng(print_n)
    states = []
    for k in range(11):
        s = gen.cube(37, seed=k)
        s["x"] += rng.normal(size=len(s)) * 10.0 ** rng.integers(-8, 8)
        states.append(s)
    path = tmp_path / "out.csv"
    sink = api.CsvSink(str(path), print_n)
    for s in states:
        sink.push(s)
    assert sink.count() == len(states)
    sink.close()
    text = path.read_text()
    assert text == ob.csvsink_lines(states, print_n)
    lines = text.split("\n")
    assert lines[-1] == "" and len(lines) - 1 == 1 + (len(states) - 1) // print_n
    row = lines[0].split(",")
    assert row[-1] == "" and len(row) == 3 * 37 + 1     # trailing comma after every value
    assert float(row[0]) == states[0]["x"][0] and float(row[2]) == states[0]["z"][0]


def test_empty_state_and_bad_arguments(tmp_path):
    sink = api.CsvSink(str(tmp_path / "e.csv"))
    sink.push(gen.cube(0, seed=1))
    sink.close()
    assert (tmp_path / "e.csv").read_text() == "\n"
    with pytest.raises(api.Pb200Error):
        api.CsvSink(str(tmp_path / "z.csv"), 0)          # the reference panics on print_n = 0
    with pytest.raises(api.Pb200Error):
        api.CsvSink(str(tmp_path / "no_such_dir" / "z.csv"))

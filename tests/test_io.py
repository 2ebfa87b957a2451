"""Native XDATCAR ingestion: bit-identical to Python float() parsing, constant- and
variable-cell layouts, malformed input.  Host code only -- runs without a GPU.  The parser is pinned against the
reference's own data files by oracle/validate_xdatcar.py (build container); here the same trajectories come from the
committed frames_* fixtures (int32 = round(frac * 1e8), exact for the files' 8 decimals)."""
import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN


def _xdatcar_text(species, lattice, frac, variable=False, scale=1.0, float_fmt="%.8f"):
    symbols, counts = [], []
    for s in species:
        if symbols and symbols[-1] == s:
            counts[-1] += 1
        else:
            symbols.append(s); counts.append(1)

    def header(lat):
        lines = ["generated", f"   {scale!r}"]
        lines += ["  " + "  ".join(repr(float(v / scale)) for v in row) for row in lat]
        lines += ["  " + "  ".join(symbols), "  " + "  ".join(str(c) for c in counts)]
        return lines
    lat0 = lattice if lattice.ndim == 2 else lattice[0]
    out = header(lat0)
    for f in range(frac.shape[0]):
        if variable and f > 0:
            out += header(lattice[f])
        out.append(f"Direct configuration= {f + 1:5d}")
        out += ["  " + "  ".join(float_fmt % v for v in row) for row in frac[f]]
    return ("\n".join(out) + "\n").encode()


@pytest.mark.parametrize("name", ["frames_readme", "frames_simple_cubic", "frames_argyrodite_0p"])
def test_constant_cell_text_is_parsed_bit_exactly(name):
    from site_analysis_b200.io import parse_xdatcar_text
    fx = np.load(GOLDEN / f"{name}.npz")
    frac = fx["frac_i32"] / 1e8
    species = [str(s) for s in fx["species"]]
    text = _xdatcar_text(species, fx["lattice"], frac)
    for threads in (1, 5):
        sp, lat, got = parse_xdatcar_text(text, n_threads=threads)
        assert sp == species
        assert lat.shape == (3, 3) and np.array_equal(lat.view(np.int64), fx["lattice"].view(np.int64))
        assert got.shape == frac.shape and np.array_equal(got.view(np.int64), frac.view(np.int64))


def test_variable_cell_and_scale():
    from site_analysis_b200.io import parse_xdatcar_text
    rng = np.random.default_rng(3)
    F, species = 7, ["Li"] * 3 + ["O"] * 2
    frac = rng.random((F, 5, 3)) * 2 - 0.5
    lattice = np.eye(3)[None] * 8.0 + rng.normal(0, 0.05, size=(F, 3, 3))
    text = _xdatcar_text(species, lattice, frac, variable=True, float_fmt="%.17g")
    sp, lat, got = parse_xdatcar_text(text)
    assert sp == species and lat.shape == (F, 3, 3)
    assert np.array_equal(got.view(np.int64), frac.view(np.int64))
    assert np.array_equal(lat.view(np.int64), lattice.view(np.int64))
    # scale factor: lattice = scale * vectors, the same float multiplication numpy does
    text = _xdatcar_text(species, lattice[0], frac[:2], scale=2.5, float_fmt="%.17g")
    _, lat, _ = parse_xdatcar_text(text)
    want = np.array([[float(repr(float(v / 2.5))) for v in row] for row in lattice[0]]) * 2.5
    assert np.array_equal(lat.view(np.int64), want.view(np.int64))
    # frames separated by blank lines, Windows line ends, lower-case keyword
    text = _xdatcar_text(species, lattice[0], frac[:3]).replace(b"Direct", b"direct").replace(b"\n", b"\r\n")
    _, _, got = parse_xdatcar_text(text)
    assert got.shape == (3, 5, 3)


def test_number_conversion_equals_python_float():
    from site_analysis_b200 import _native as nat
    L = nat.lib()
    rng = np.random.default_rng(9)
    tokens = ["0", "-0.0", "1", "0.1", ".5", "5.", "1e22", "1e23", "1e-22", "1e-23", "123456789012345", "1234567890123456",
              "9007199254740993", "0.000000000000000000000001234", "1.7976931348623157e308", "4.9e-324", "2.2250738585072011e-308",
              "0.30000000000000004", "123456789012345678901234567890", "1.5d3", "2.5E-3", "+7.25", "00012.500"]
    for _ in range(4000):
        digits = int(rng.integers(1, 22))
        mant = "".join(rng.choice(list("0123456789"), size=digits))
        point = int(rng.integers(0, digits + 1))
        tok = mant[:point] + "." + mant[point:] if rng.random() < 0.8 else mant
        if rng.random() < 0.4:
            tok += f"e{int(rng.integers(-30, 31))}"
        if rng.random() < 0.3:
            tok = "-" + tok
        tokens.append(tok)
    for tok in tokens:
        out, used = C.c_double(), C.c_int64()
        raw = (" " + tok + " \n").encode()
        assert L.sab_parse_double(raw, len(raw), C.byref(out), C.byref(used)) == 0, tok
        want = float(tok.replace("d", "e"))
        assert np.float64(out.value).view(np.int64) == np.float64(want).view(np.int64), tok
        assert used.value == len(tok) + 1
    for bad in (b"abc", b"", b"1.2.3", b"--1", b"1e5x"):
        out = C.c_double()
        assert L.sab_parse_double(bad, len(bad), C.byref(out), None) != 0


def test_malformed_files_raise():
    from site_analysis_b200.io import parse_xdatcar_text
    species, lat = ["Li", "Li"], np.eye(3) * 5
    frac = np.zeros((3, 2, 3)) + 0.25
    good = _xdatcar_text(species, lat, frac)
    assert parse_xdatcar_text(good)[2].shape == (3, 2, 3)
    with pytest.raises(ValueError):
        parse_xdatcar_text(good[:-12])                       # last coordinate line cut in the middle of a number pair
    with pytest.raises(ValueError):
        parse_xdatcar_text(good.replace(b"0.25000000  0.25000000  0.25000000\n", b"0.25000000  0.25000000\n", 1))
    with pytest.raises(ValueError):
        parse_xdatcar_text(b"comment\n1.0\n")
    with pytest.raises(ValueError):
        parse_xdatcar_text(good.replace(b"  2\n", b"  x\n", 1))
    empty = b"\n".join(good.split(b"\n")[:7]) + b"\n"
    sp, _, fr = parse_xdatcar_text(empty)
    assert sp == species and fr.shape == (0, 2, 3)


@pytest.mark.gpu
@pytest.mark.timeout(120)
def test_trajectory_from_xdatcar_matches_reference_golden(oracle_mod, tmp_path):
    """README quick start through the file reader: [[0], [0], [None], [None], [1]] (reference README.md:33-41)."""
    from helpers import trajectory_from_case
    case = oracle_mod.load_case("readme_spherical")
    fx = np.load(GOLDEN / "frames_readme.npz")
    path = tmp_path / "XDATCAR"
    path.write_bytes(_xdatcar_text([str(s) for s in fx["species"]], fx["lattice"], fx["frac_i32"] / 1e8))
    t = trajectory_from_case(case)
    rows = t.trajectory_from_xdatcar(path)
    assert np.array_equal(rows, case["atoms_traj"])
    assert t.atoms_trajectory == [[0], [0], [None], [None], [1]]

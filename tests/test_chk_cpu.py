"""The .chk reader / writer (micmec_b200/chk.py; format restated in SURVEY.md Appendix B) and System.from_file /
to_file: column layout of the records, all kinds, bit-exact round trips - including, in the build container, every
``data/*_micmec.chk`` file the reference ships."""
import glob
import os

import numpy as np
import pytest

import goldenio as gio
from test_force_gpu import make_system


def test_record_layout_and_kinds(tmp_path):
    from micmec_b200.chk import dump_chk, load_chk

    data = {
        "pos": np.array([[1.0, -2.5e-7, 3.0], [4.0, 5.0, 6.125]]),
        "grid": np.arange(8).reshape(2, 2, 2),
        "pbc": np.array([True, False, True]),
        "names": np.array(["fcu", "reo"]),
        "type1/effective_temp": 300.0,
        "count": 7,
        "flag": True,
        "type1/name": "fcu",
        "nothing": None,
    }
    fn = str(tmp_path / "t.chk")
    dump_chk(fn, data)
    lines = open(fn).read().split("\n")
    headers = [line for line in lines if line[42:47] == "kind="]
    assert len(headers) == len(data)
    assert [h[:40].strip() for h in headers] == sorted(data)  # one record per key, keys sorted
    for h in headers:
        assert h[40:42] == "  " and len(h[47:52]) == 5 and h[52] == " "
    pos_header = [h for h in headers if h.startswith("pos ")][0]
    assert pos_header[47:52] == "fltar" and pos_header[53:].strip() == "2,3"
    body = lines[lines.index(pos_header) + 1]
    # four values per line, each "% 22.15e", separated by one blank - as in data/3x3x3_test_micmec.chk:10
    assert len(body) == 4 * 22 + 3 and body[:22] == "% 22.15e" % 1.0 and body[22] == " "
    back = load_chk(fn)
    assert set(back) == set(data)
    for key, val in data.items():
        if isinstance(val, np.ndarray):
            assert back[key].shape == val.shape and np.array_equal(back[key], val), key
        else:
            assert back[key] == val, key
    assert back["grid"].dtype == np.int64 and back["pbc"].dtype == np.bool_


def test_malformed_files_are_rejected(tmp_path):
    from micmec_b200.chk import load_chk

    fn = str(tmp_path / "bad.chk")
    open(fn, "w").write("this is not a chk header\n")
    with pytest.raises(IOError):
        load_chk(fn)
    open(fn, "w").write("%-40s  kind=%-5s %s\n%22s\n" % ("x", "fltar", "3", "1.0"))
    with pytest.raises(IOError):
        load_chk(fn)


@pytest.mark.parametrize("name", ["3x3x3_conf0", "5x5x5_fcu_hollow"])
def test_system_round_trip_is_bit_exact(tmp_path, name):
    from micmec_b200.system import System

    system = make_system(gio.load("force_" + name))
    fn = str(tmp_path / "sys.chk")
    system.to_file(fn)
    back = System.from_file(fn)
    for attr in ("pos", "masses", "surrounding_cells", "surrounding_nodes", "boundary_nodes", "grid", "types"):
        assert np.array_equal(np.asarray(getattr(back, attr)), np.asarray(getattr(system, attr))), attr
    assert np.array_equal(np.array(back.domain.rvecs), np.array(system.domain.rvecs))
    assert set(back.params) == set(system.params)
    for key, val in system.params.items():
        assert np.array_equal(np.asarray(back.params[key]), np.asarray(val)), key
    with pytest.raises(IOError):
        system.to_file(str(tmp_path / "sys.h5"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference data files only exist in the build container")
def test_every_reference_data_file_loads_and_round_trips(tmp_path):
    from micmec_b200.system import System

    files = sorted(glob.glob("/root/reference/data/*_micmec.chk"))
    assert len(files) >= 10
    for fn in files:
        system = System.from_file(fn)
        assert system.pos.shape == (system.nnodes, 3) and np.asarray(system.surrounding_nodes).shape == (system.ncells, 8)
        out = str(tmp_path / os.path.basename(fn))
        system.to_file(out)
        back = System.from_file(out)
        assert np.array_equal(back.pos, system.pos) and np.array_equal(back.masses, system.masses)
        assert np.array_equal(np.asarray(back.types), np.asarray(system.types)) and set(back.params) == set(system.params)

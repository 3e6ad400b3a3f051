"""Waiwera's HDF5 output / restart files without an HDF5 library (none in this image):
waiwera_b200/h5lite.py parses and writes the file format itself, waiwera_b200/output.py lays out the reference's
datasets.  The READER is pinned by two files the reference's own stack wrote (tests/golden/h5/, copied by
tools/make_golden.py::h5_fixtures): the golden cell balances of flow_simulation_test.F90 and a Waiwera output file with
chunked time-sequence datasets.  The WRITER is checked through the reader and structure by structure against what the
library wrote into those files; it cannot be checked against libhdf5 itself here."""
import json
import os

import numpy as np
import pytest

from waiwera_b200 import h5lite, output
from waiwera_b200 import mesh as wmesh

HERE = os.path.dirname(os.path.abspath(__file__))
H5 = os.path.join(HERE, "golden", "h5")


def test_reader_on_files_written_by_the_reference_stack():
    h = h5lite.H5File(os.path.join(H5, "lhs.h5"))
    assert h.groups() == ["cell_fields", "fields"]
    assert h.datasets() == ["cell_fields/lhs_Primary", "cell_index", "fields/lhs"]
    ref = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))
    assert np.array_equal(h["fields/lhs"], np.array(ref["lhs"]["values"]))    # 12 x 99.82512244887545
    assert np.array_equal(h["cell_fields/lhs_Primary"], h["fields/lhs"])
    ci = h["cell_index"]
    assert ci.dtype == np.int32 and ci.shape == (12, 1) and sorted(ci.reshape(-1)) == list(range(12))
    # a Waiwera output file: chunked [time, cell] datasets, int32 index sets, source fields
    h = h5lite.H5File(os.path.join(H5, "oned_two_phase_ss.h5"))
    assert "cell_fields/fluid_pressure" in h and h.shape("cell_fields/fluid_pressure") == (1, 10)
    assert h["time"].reshape(-1).tolist() == [1.0e15]
    assert h["source_fields/source_natural_cell_index"].dtype == np.int32
    g = json.load(open(os.path.join(HERE, "golden", "tracer_oned.json")))["two"]["initial"]
    assert np.array_equal(h["cell_fields/fluid_pressure"][0], g["pressure"])
    assert np.array_equal(h["cell_fields/fluid_temperature"][0], g["temperature"])
    assert np.array_equal(h["cell_fields/fluid_vapour_saturation"][0], g["vapour_saturation"])
    assert np.allclose(h["cell_fields/cell_geometry_volume"], 10.0)
    with pytest.raises(h5lite.H5Error):
        h5lite.H5File(__file__)


def test_restart_from_a_waiwera_output_file():
    """initial.filename (src/initial.F90:421-507): primaries and regions of the last time in natural cell order"""
    prim, region, t = output.read_restart(os.path.join(H5, "oned_two_phase_ss.h5"), "we")
    g = json.load(open(os.path.join(HERE, "golden", "tracer_oned.json")))["two"]["initial"]
    assert t == 1.0e15 and (region == 4).all()
    assert np.array_equal(prim[:, 0], g["pressure"]) and np.array_equal(prim[:, 1], g["vapour_saturation"])
    with pytest.raises(IndexError):
        output.read_restart(os.path.join(H5, "oned_two_phase_ss.h5"), "we", index=3)
    with pytest.raises(KeyError):
        output.read_restart(os.path.join(H5, "oned_two_phase_ss.h5"), "wce")       # no CO2 partial pressure in the file


def _messages_of(h, path):
    ent = h.root_entry
    for p in path.split("/"):
        msgs = h._messages(ent["header"])
        if ent.get("cache") == 1:
            bt, hp = ent["btree"], ent["heap"]
        else:
            d = [m for t, m in msgs if t == 0x11][0]
            bt, hp = int.from_bytes(d[:8], "little"), int.from_bytes(d[8:16], "little")
        ent = [e for e in h._group_entries(bt, hp) if e["name"] == p][0]
    return dict((t, bytes(m)) for t, m in h._messages(ent["header"]))


def test_writer_round_trip_and_structures(tmp_path):
    rng = np.random.default_rng(1)
    data = {"time": rng.uniform(size=(5, 1)), "cell_index": np.arange(7, dtype=np.int32).reshape(7, 1),
            "cell_fields/fluid_pressure": rng.uniform(size=(5, 7)), "cell_fields/fluid_region": np.ones((5, 7)),
            "source_fields/source_rate": rng.uniform(size=(5, 2)), "deep/er/group/x": np.arange(3.0),
            "empty": np.zeros((0, 4)), "i64": np.arange(4, dtype=np.int64), "f32": np.arange(4, dtype=np.float32)}
    for k in range(40):                                   # more entries than one symbol node holds
        data["many/f%02d" % k] = np.full(3, float(k))
    path = str(tmp_path / "t.h5")
    h5lite.write(path, data)
    h = h5lite.H5File(path)
    assert h.datasets() == sorted(data) and h.groups() == ["cell_fields", "deep", "deep/er", "deep/er/group", "many", "source_fields"]
    for k, v in data.items():
        a = h[k]
        assert a.shape == v.shape and a.dtype == v.dtype and np.array_equal(a, v), k
    raw = open(path, "rb").read()
    lib = h5lite.H5File(os.path.join(H5, "oned_two_phase_ss.h5"))
    # superblock: the same versions, offset / length sizes and B-tree ranks as the library's file; end-of-file address
    assert raw[:24] == lib.buf[:24]
    assert int.from_bytes(raw[40:48], "little") == len(raw)
    # messages of a float64 and an int32 dataset: dataspace, datatype and fill value byte for byte as the library
    # writes them (the layout differs by design: contiguous here, chunked there)
    mine, theirs = _messages_of(h, "cell_fields/fluid_region"), _messages_of(lib, "cell_fields/cell_geometry_volume")
    assert mine[0x03] == theirs[0x03] and mine[0x05] == theirs[0x05]
    assert _messages_of(h, "cell_index")[0x03] == _messages_of(lib, "cell_index")[0x03]
    space = _messages_of(h, "cell_index")[0x01]           # version 1, rank 2, no maximum sizes, then the sizes
    assert space[:24] == b"\x01\x02\x00" + bytes(5) + (7).to_bytes(8, "little") + (1).to_bytes(8, "little")
    # every structure is 8-byte aligned and carries its signature; node sizes follow the superblock's ranks
    for sig, size in ((b"TREE", 24 + 33 * 8 + 32 * 8), (b"SNOD", 8 + 8 * 40)):
        pos = raw.find(sig)
        n = 0
        while pos >= 0:
            assert pos % 8 == 0
            n += 1
            pos = raw.find(sig, pos + size)
        assert n >= 6
    assert raw.count(b"HEAP") == 7                        # one local heap per group (root + 6)


def test_output_file_and_restart_round_trip(wo, tmp_path):
    """write_output in the reference's layout from fluid records, read_restart gives back the primaries (also with a
    storage order that is not the natural one, as in a parallel run)"""
    from util import make_problem_wce, oracle_flow
    m, y, region, prm = make_problem_wce(wo, dims=(4, 3, 5), two_phase_layers=2)
    f = oracle_flow(wo, m, prm, y, region)
    fl = f.fluid()
    n = m.ninterior
    order = np.random.default_rng(4).permutation(n).astype(np.int32)
    path = str(tmp_path / "out.h5")
    output.write_output(path, m, "wce", [0.0, 1.0e6], [fl, fl], source_cells=[3, 5],
                        source_history=[[[0, -1.0, 2.0e5], [1, 2.0, 1.0e5]]] * 2, cell_index=order)
    h = h5lite.H5File(path)
    assert set(h.datasets()) >= {"time", "cell_index", "cell_fields/cell_geometry_centroid", "cell_fields/cell_geometry_volume",
                                 "cell_fields/fluid_pressure", "cell_fields/fluid_temperature", "cell_fields/fluid_region",
                                 "cell_fields/fluid_CO2_partial_pressure", "cell_fields/fluid_vapour_saturation",
                                 "source_index", "source_fields/source_rate", "source_fields/source_natural_cell_index"}
    assert h.shape("cell_fields/fluid_pressure") == (2, n) and h["source_fields/source_rate"].tolist() == [[-1.0, 2.0]] * 2
    prim, reg, t = output.read_restart(path, "wce", index=-1)
    assert t == 1.0e6 and np.array_equal(reg, region[:n])
    fl = np.asarray(fl)[:n]
    # cell_index is "natural to global" (src/dm_utils.F90:974-1037): position i of the file holds natural cell order[i],
    # and cell_index[order[i]] = i
    assert np.array_equal(h["cell_fields/fluid_pressure"][1], fl[order, 0])
    assert np.array_equal(h["cell_index"].reshape(-1)[order], np.arange(n))
    assert np.array_equal(prim[:, 0], fl[:, 0]) and np.array_equal(prim[:, 2], fl[:, 7])
    two = reg == 4
    assert two.any() and np.array_equal(prim[two, 1], fl[two, 8 + 9 + 2]) and np.array_equal(prim[~two, 1], fl[~two, 1])
    # the restart state is the state the run started from
    back = wmesh.scale_primaries(prim, reg).reshape(-1)
    assert np.allclose(back, y[:3 * n], rtol=1e-12, atol=1e-14)


def test_reader_on_every_hdf5_file_of_the_reference():
    import glob
    files = sorted(glob.glob("/root/reference/**/*.h5", recursive=True))
    if not files:
        pytest.skip("the reference tree is not here")
    for f in files:
        h = h5lite.H5File(f)
        for d in h.datasets():
            assert h[d].shape == h.shape(d)
    assert len(files) >= 10


def test_reader_on_a_netcdf4_exodus_mesh():
    """the version-2 structures (netCDF-4 writes them): "OHDR" object headers with continuation chunks, link messages in a
    fractal heap indexed by a version-2 B-tree, attribute messages, a chunked unlimited dataset -- on the ExodusII mesh
    of the reference's 3-D MINC benchmark; the mesh read from it is the one the benchmark's ASCII fixture holds"""
    from waiwera_b200 import ingest
    path = os.path.join(H5, "gminc_3d_refined.exo")
    h = h5lite.H5File(path)
    assert {"connect1", "connect2", "coord", "eb_prop1", "time_whole", "num_nodes"} <= set(h.datasets())
    assert h.shape("coord") == (3, 732) and h.shape("connect1") == (75, 6) and h.shape("connect2") == (465, 8)
    assert h.attrs("connect1")["elem_type"] == "WEDGE" and h.attrs("connect2")["elem_type"] == "HEXAHEDRON"
    root = h.attrs("/")
    assert root["title"].startswith("Created by meshio") and root["floating_point_word_size"][0] == 8
    assert abs(float(root["version"][0]) - 5.1) < 1e-6
    con = h["connect2"]
    assert con.dtype.kind == "i" and con.min() >= 1 and con.max() <= 732
    assert h["time_whole"].shape == (1,) and h.shape("time_step") == (0,)
    xyz, elems = ingest.read_exodus(path)
    fx, fe = ingest.read_gmsh(os.path.join(HERE, "golden", "inputs", "gminc_3d_refined.ascii.msh"))
    assert np.array_equal(xyz, fx) and elems == fe
    assert [t for t, _ in elems[:465]] == [5] * 465 and [t for t, _ in elems[465:]] == [6] * 75      # wedges last
    m, _ = ingest.build_mesh(xyz, elems)
    assert np.isclose(m.cell_geom[:540, 3].sum(), 6.0e10)


def test_reader_on_every_exodus_file_of_the_reference():
    import glob
    from waiwera_b200 import ingest
    files = sorted(glob.glob("/root/reference/**/*.exo", recursive=True))
    if not files:
        pytest.skip("the reference tree is not here")
    kinds = set()
    for f in files:
        kinds.add(open(f, "rb").read(4))
        xyz, elems = ingest.read_exodus(f)
        m, _ = ingest.build_mesh(xyz, elems)
        assert m.ninterior == len(elems) > 0 and (m.cell_geom[:, 3] > 0).all()
    assert len(files) >= 16 and kinds == {b"CDF\x02", b"\x89HDF"}


def test_ingest_restarts_from_the_file_the_input_names():
    """the tracer doublet input starts from "initial": {"filename": "doublet_ss.h5"}: ingest reads the Waiwera output
    file itself; the state equals the golden steady state (tests/golden/tracer_doublet.json, the same file read by a
    byte scan in round 1)"""
    from waiwera_b200 import ingest
    p = ingest.load(os.path.join(HERE, "golden", "inputs", "doublet.input.json"))
    g = json.load(open(os.path.join(HERE, "golden", "tracer_doublet.json")))
    assert p.primary.shape == (100, 2) and (p.region == 1).all() and p.restart_time == 1.0e15
    assert np.array_equal(p.primary[:, 0], g["steady_pressure"]) and np.array_equal(p.primary[:, 1], g["steady_temperature"])
    assert np.allclose(p.y.reshape(-1, 2), p.primary / [1e6, 1e2])


def test_filter_pipeline_is_undone_in_reverse_order():
    """chunks of a filtered dataset: shuffle (2) then deflate (1) then fletcher32 (3) on write -> undone back to front; a
    filter whose bit is set in the chunk's mask was skipped on write.  (No file of the reference uses filters: this
    checks the pipeline on bytes made here, and the parsing of both versions of the filter pipeline message.)"""
    import struct
    import zlib
    data = np.arange(1000, dtype="<f8") * 1.5 - 7.0
    raw = data.tobytes()
    shuffled = np.frombuffer(raw, np.uint8).reshape(-1, 8).T.tobytes()
    packed = zlib.compress(shuffled) + b"\0\0\0\0"          # + checksum bytes, which the reader drops
    filters = [(2, [8]), (1, [6]), (3, [])]
    assert h5lite.H5File._unfilter(packed, filters, 0, 8) == raw
    assert h5lite.H5File._unfilter(zlib.compress(raw) + b"\0\0\0\0", filters, 1, 8) == raw           # shuffle skipped for this chunk
    assert h5lite.H5File._unfilter(shuffled, [(2, [8]), (1, [6])], 2, 8) == raw                      # deflate skipped
    # version 2 message: ids < 256 carry no name; version 1: name length field and padding of odd client data
    v2 = bytes([2, 2]) + struct.pack("<HHHI", 2, 0, 1, 8) + struct.pack("<HHHI", 1, 1, 1, 6)
    assert h5lite.H5File._filters(v2) == [(2, [8]), (1, [6])]
    v1 = bytes([1, 2]) + b"\0" * 6 + struct.pack("<HHHH", 2, 8, 0, 1) + b"shuffle\0" + struct.pack("<II", 8, 0) \
        + struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<II", 6, 0)
    assert h5lite.H5File._filters(v1) == [(2, [8]), (1, [6])]
    with pytest.raises(h5lite.H5Error):
        h5lite.H5File._filters(bytes([2, 1]) + struct.pack("<HHHH", 32001, 0, 0, 0))               # a third-party filter

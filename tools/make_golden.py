#!/usr/bin/env python3
"""Extracts the reference's own golden vectors for the hot path into tests/golden/reference_vectors.json.

Run in the build container only (needs /root/reference; the GPU box has neither it nor h5py):
    python tools/make_golden.py

Sources (PETSc HDF5 viewer files; the Vec payloads are contiguous little-endian f64, found by a byte scan
because h5py is not installed):
  test/unit/data/flow_simulation/lhs/lhs.h5      <- test_flow_simulation_lhs   (flow_simulation_test.F90:126-158)
  test/unit/data/flow_simulation/init/primary.h5 <- test_flow_simulation_init  (:90-121), scaled primaries
  test/unit/data/flow_simulation/init/rock.h5    <- same test, 8-double rock records
"""
import json
import os

import numpy as np

REF = "/root/reference/test/unit/data/flow_simulation"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_vectors.json")


def runs(path, lo, hi, minlen):
    """runs of >= minlen consecutive 8-byte-aligned doubles with lo < |x| < hi"""
    b = open(path, "rb").read()
    a = np.frombuffer(b[:len(b) // 8 * 8], dtype="<f8")
    ok = np.isfinite(a) & (np.abs(a) > lo) & (np.abs(a) < hi)
    out, s = [], None
    for i, f in enumerate(ok):
        if f and s is None:
            s = i
        elif not f and s is not None:
            if i - s >= minlen:
                out.append(a[s:i].copy())
            s = None
    if s is not None and len(a) - s >= minlen:
        out.append(a[s:].copy())
    return out


def listing_tables(path):
    """ELEMENT TABLE blocks of an AUTOUGH2 listing: [(time_s, {name: [P, T, Sv, Sl, tracer]})] and the
    GENERATION TABLE rows [(time_s, rate, enthalpy, tracer_flow)]"""
    import re
    lines = open(path).read().splitlines()
    elems, gens, t = [], [], None
    i = 0
    while i < len(lines):
        m = re.search(r"OUTPUT AFTER\s+\d+ TIME STEPS\s+([0-9.E+-]+) SECONDS", lines[i])
        if m:
            t = float(m.group(1))
        if "ELEMENT TABLE" in lines[i]:
            i += 4
            rows = []
            while i < len(lines) and lines[i].strip() and not lines[i].startswith(" EEEE"):
                f = lines[i].split()
                if len(f) >= 9:
                    rows.append([float(v) for v in f[-7:-2]])
                i += 1
            elems.append((t, rows))
        if "GENERATION TABLE" in lines[i]:
            i += 4
            while i < len(lines) and lines[i].strip() and not lines[i].startswith(" GGGG"):
                f = lines[i].split()
                gens.append((t, float(f[-7]), float(f[-6]), float(f[-5])))
                i += 1
        i += 1
    return elems, gens


def tracer_oned():
    """test/benchmark/tracer/oned: AUTOUGH2 listings of the 1-D liquid tracer problems (test_tracer_1d.py compares
    pressure and tracer mass fraction at the last output with tolerance 1e-3 and the tracer production history)"""
    base = "/root/reference/test/benchmark/tracer/oned/run"
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/tracer/oned/run/"
                            "oned_{single,two}_phase.listing (AUTOUGH2), columns P, T, Sv, Sl, tracer mass fraction; "
                            "the last row of every table is the Dirichlet boundary block"}
    for case in ("single", "two"):
        elems, gens = listing_tables(os.path.join(base, "oned_%s_phase.listing" % case))
        # initial state of the transient run = the Waiwera output file the benchmark ships (final state of the
        # *_ss.json run; PETSc HDF5 viewer, contiguous f64 found by a byte scan): pressure and, in the two-phase
        # case, vapour saturation of the 10 cells (the single-phase file holds the uniform 3 MPa / 20 degC state)
        h5 = os.path.join(base, "oned_%s_phase_ss.h5" % case)
        init = {"pressure": [float(v) for v in runs(h5, 1e4, 1e7, 10)[0][:10]]}
        if case == "two":
            init["vapour_saturation"] = [float(v) for v in runs(h5, 1e-3, 1.0, 10)[0][:10]]
            init["temperature"] = [float(v) for v in runs(h5, 10, 200, 10)[0][:10]]
        else:
            init["temperature"] = [float(v) for v in runs(h5, 10, 200, 10)[0][:10]]
        doc[case] = {"times": [t for t, _ in elems], "tables": [rows for _, rows in elems],
                     "source": [list(g) for g in gens], "initial": init}
    out = os.path.join(os.path.dirname(OUT), "tracer_oned.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def main():
    lhs = runs(os.path.join(REF, "lhs", "lhs.h5"), 1.0, 1e4, 12)[0][:12]
    primary = runs(os.path.join(REF, "init", "primary.h5"), 1e-3, 1e3, 12)[0][:12]
    rock = runs(os.path.join(REF, "init", "rock.h5"), 1e-20, 1e12, 96)[0][:96]
    assert len(lhs) == 12 and len(primary) == 12 and len(rock) == 96
    doc = {
        "_generated_by": "tools/make_golden.py from /root/reference/test/unit/data/flow_simulation/{lhs,init}/*.h5",
        "lhs": {
            "source": "lhs/lhs.h5 (test_lhs.json: eos w, T=20 degC, primary 2.0e5 Pa, porosity 0.1, IAPWS-97, 12 cells)",
            "eos": "w", "temperature": 20.0, "pressure": 2.0e5, "porosity": 0.1,
            "values": [float(v) for v in lhs]},
        "primary_scaled": {
            "source": "init/primary.h5 (test_init.json: primary 2.0e5 Pa stored scaled by eos%scale, 1e6 Pa)",
            "pressure": 2.0e5, "values": [float(v) for v in primary]},
        "rock": {
            "source": "init/rock.h5: 12 records x 8 doubles (permeability(3), wet, dry conductivity, porosity, density, specific heat)",
            "values": [float(v) for v in rock]},
    }
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
    tracer_oned()

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


def listing_generic(path):
    """every ELEMENT / GENERATION table of an AUTOUGH2 listing as rows of floats (the numeric tail of each line)"""
    import re
    lines = open(path).read().splitlines()
    out, t, i = [], None, 0
    num = re.compile(r"^[-+]?\d\.\d+E[-+]\d+$")
    while i < len(lines):
        m = re.search(r"OUTPUT AFTER\s+\d+ TIME STEPS\s+([0-9.E+-]+) SECONDS", lines[i])
        if m:
            t = float(m.group(1))
        for kind in ("ELEMENT TABLE", "GENERATION TABLE"):
            if kind in lines[i]:
                i += 4
                rows = []
                while i < len(lines) and lines[i].strip() and lines[i][1:5] not in ("EEEE", "GGGG"):
                    vals = [float(v) for v in lines[i].split() if num.match(v)]
                    if vals:
                        rows.append(vals)
                    i += 1
                out.append((kind[0], t, rows))
        i += 1
    return out


def co2_one_cell():
    """test/benchmark/ncg/co2_one_cell (O'Sullivan et al. 1985, fig. 5): AUTOUGH2 history of the single cell --
    test_co2_one_cell.py compares pressure, temperature, vapour saturation and the production enthalpy at 1e-3"""
    path = "/root/reference/test/benchmark/ncg/co2_one_cell/run/co2_one_cell.listing"
    el = [(t, r[0]) for k, t, r in listing_generic(path) if k == "E"]
    ge = [(t, r[0]) for k, t, r in listing_generic(path) if k == "G"]
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/ncg/co2_one_cell/run/"
                            "co2_one_cell.listing (AUTOUGH2)",
           "times": [t for t, _ in el],
           "columns": ["pressure", "temperature", "gas_saturation", "co2_partial_pressure", "co2_mass_fraction"],
           "element": [r[:5] for _, r in el],
           "source_times": [t for t, _ in ge], "source_enthalpy": [r[1] for _, r in ge]}
    out = os.path.join(os.path.dirname(OUT), "co2_one_cell.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def co2_column():
    """test/benchmark/ncg/co2_column (O'Sullivan et al. 1985, figs 10-11): AUTOUGH2 steady state of the 30-layer
    column for 0 / 0.1 / 1 / 5 % CO2 in the injected fluid -- test_co2_column.py compares pressure, temperature,
    vapour saturation and total CO2 mass fraction of the last output at 1e-3"""
    base = "/root/reference/test/benchmark/ncg/co2_column/run"
    doc = {"_generated_by": "tools/make_golden.py: last ELEMENT table of test/benchmark/ncg/co2_column/run/"
                            "co2_column_{0,0.1,1,5}.listing (AUTOUGH2), atmosphere block dropped",
           "columns": ["pressure", "temperature", "gas_saturation", "co2_partial_pressure", "co2_mass_fraction"]}
    for case in ("0", "0.1", "1", "5"):
        tabs = [(t, r) for k, t, r in listing_generic(os.path.join(base, "co2_column_%s.listing" % case)) if k == "E"]
        t, rows = tabs[-1]
        assert len(rows) == 31
        src = json.load(open(os.path.join(base, "co2_column_%s.json" % case)))
        doc[case] = {"time": t, "element": [r[:5] for r in rows[1:]],
                     "initial_pressure": [p[0] for p in src["initial"]["primary"]],
                     "sources": [[s_.get("component", 1), s_["rate"], s_.get("enthalpy", 0.0)] for s_ in src["source"]]}
    out = os.path.join(os.path.dirname(OUT), "co2_column.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def minc_column():
    """test/benchmark/minc/column: AUTOUGH2 listings of the 11-layer production column, single porosity and MINC
    (2 matrix levels in layers 3-8) -- test_minc_column.py compares P, T, Sv of the last output (2.5e-2), their
    history in the production cell (2e-2) and the production enthalpy history (1e-2)"""
    base = "/root/reference/test/benchmark/minc/column/run"
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/minc/column/run/"
                            "minc_column_{single,minc}.listing (AUTOUGH2); atmosphere block dropped; MINC blocks "
                            "reordered level by level as Waiwera numbers them (test_minc_column.py:57-67)",
           "columns": ["pressure", "temperature", "vapour_saturation"]}
    for case in ("single", "minc"):
        tabs = listing_generic(os.path.join(base, "minc_column_%s.listing" % case))
        el = [(t, r) for k, t, r in tabs if k == "E"]
        ge = [(t, r) for k, t, r in tabs if k == "G"]
        tables = []
        for t, rows in el:
            rows = [r[:3] for r in rows[1:]]
            frac, minc = rows[:11], rows[11:]
            nlev = 2 if case == "minc" else 0
            for l in range(nlev):
                frac = frac + minc[l::nlev]
            tables.append(frac)
        src = json.load(open(os.path.join(base, "minc_column_%s.json" % case)))
        doc[case] = {"times": [t for t, _ in el], "tables": tables,
                     "production_enthalpy": [r[1][1] for _, r in ge], "source_times": [t for t, _ in ge],
                     "initial_primary": src["initial"]["primary"], "initial_region": src["initial"]["region"]}
    out = os.path.join(os.path.dirname(OUT), "minc_column.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def mis_problems():
    """test/benchmark/model_intercomparison_study problems 2a-c, 4, 5a-b, 6: AUTOUGH2 listings.  Kept: P, T, Sv of all
    cells at ~12 evenly spaced output times, the full history of a few cells (production cell first) and the
    production enthalpy history (the input decks: input_fixtures)"""
    import shutil
    base = "/root/reference/test/benchmark/model_intercomparison_study"
    inputs = os.path.join(os.path.dirname(OUT), "inputs")
    os.makedirs(inputs, exist_ok=True)
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/"
                            "model_intercomparison_study/problem{2,4,5,6}/run/*.listing (AUTOUGH2); boundary / atmosphere "
                            "blocks dropped", "columns": ["pressure", "temperature", "vapour_saturation"]}
    for prob, cases in (("problem2", ["problem2a", "problem2b", "problem2c"]), ("problem4", ["problem4"]),
                        ("problem5", ["problem5a", "problem5b"]), ("problem6", ["problem6"])):
        run = os.path.join(base, prob, "run")
        for case in cases:
            inp = json.load(open(os.path.join(run, case + ".json")))
            tabs = listing_generic(os.path.join(run, case + ".listing"))
            el = [(t, r) for k, t, r in tabs if k == "E"]
            ge = [(t, r) for k, t, r in tabs if k == "G"]
            # number of interior cells from the initial conditions / rock types of the input
            ncell = max(max(rt.get("cells", [0]) or [0]) for rt in inp["rock"]["types"]) + 1
            nrow = len(el[-1][1])
            # atmosphere blocks first (problem 4: one; problem 6: one per column), other boundary blocks last
            first = nrow - ncell if prob == "problem4" else 25 if prob == "problem6" else 0
            rows = lambda r: [x[:3] for x in r[first:first + ncell]]
            sel = sorted(set(np.linspace(1, len(el) - 1, 12).round().astype(int).tolist()))
            prod = inp["source"][0]["cell"]
            hist_cells = sorted(set([prod, 0, ncell // 2, ncell - 1]), key=lambda c: (c != prod, c))
            doc[case] = {"ncell": ncell, "times": [t for t, _ in el], "table_index": sel,
                         "tables": [rows(el[i][1]) for i in sel], "history_cells": hist_cells,
                         "history": [[rows(r)[c] for c in hist_cells] for _, r in el],
                         "source_times": [t for t, _ in ge], "production_enthalpy": [r[0][1] for _, r in ge]}
    out = os.path.join(os.path.dirname(OUT), "mis_problems.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def deliverability():
    """test/benchmark/source/deliverability: AUTOUGH2 listings of the 10-cell production problems on deliverability
    (delv: fixed productivity index; delt: with a total-flow limiter; delg_flow: productivity index from the initial
    rate; delg_pi_table: productivity index from a table in time) -- test_deliverability.py compares P, T, Sv of the last output (5e-3), their history in the production cell
    and the generation rate / enthalpy history (1e-2)"""
    base = "/root/reference/test/benchmark/source/deliverability/run"
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/source/deliverability/"
                            "run/deliv_{delv,delt,delg_flow,delg_pi_table}.listing (AUTOUGH2); boundary block dropped",
           "columns": ["pressure", "temperature", "vapour_saturation"]}
    for case in ("delv", "delt", "delg_flow", "delg_pi_table"):
        tabs = listing_generic(os.path.join(base, "deliv_%s.listing" % case))
        el = [(t, r) for k, t, r in tabs if k == "E"]
        ge = [(t, r) for k, t, r in tabs if k == "G"]
        src = json.load(open(os.path.join(base, "deliv_%s.json" % case)))
        doc[case] = {"times": [t for t, _ in el], "tables": [[x[:3] for x in r[:10]] for _, r in el],
                     "source_times": [t for t, _ in ge], "rate": [r[0][0] for _, r in ge],
                     "enthalpy": [r[0][1] for _, r in ge], "step_sizes": src["time"]["step"]["size"],
                     "stop": src["time"]["stop"], "source": src["source"], "boundary": src["boundaries"][0]["primary"],
                     "initial": src["initial"]["primary"]}
    out = os.path.join(os.path.dirname(OUT), "deliverability.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def recharge():
    """test/benchmark/source/recharge: AUTOUGH2 listing of the 10-cell outflow problem with a recharge source
    (rate = -coefficient (P - reference pressure), production only) -- tests/test_recharge.py"""
    base = "/root/reference/test/benchmark/source/recharge/run"
    tabs = listing_generic(os.path.join(base, "recharge_outflow.listing"))
    el = [(t, r) for k, t, r in tabs if k == "E"]
    ge = [(t, r) for k, t, r in tabs if k == "G"]
    src = json.load(open(os.path.join(base, "recharge_outflow.json")))
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/source/recharge/run/"
                            "recharge_outflow.listing (AUTOUGH2)",
           "columns": ["pressure", "temperature", "vapour_saturation"],
           "outflow": {"times": [t for t, _ in el], "tables": [[x[:3] for x in r[:10]] for _, r in el],
                       "source_times": [t for t, _ in ge], "rate": [r[0][0] for _, r in ge],
                       "enthalpy": [r[0][1] for _, r in ge], "step_sizes": src["time"]["step"]["size"],
                       "stop": src["time"]["stop"], "source": src["source"], "initial": src["initial"]["primary"],
                       "rock": src["rock"]["types"][0]}}
    out = os.path.join(os.path.dirname(OUT), "recharge.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def h5_fixtures():
    """two small HDF5 files written by the reference's own stack (PETSc HDF5 viewer), kept as binary fixtures for the
    reader of waiwera_b200/h5lite.py: test/unit/data/flow_simulation/lhs/lhs.h5 (11 KB: the golden cell balances, a
    time-independent Vec) and test/benchmark/tracer/oned/run/oned_two_phase_ss.h5 (42 KB: a Waiwera output file with
    chunked time-sequence datasets, the restart file of the tracer benchmark)"""
    import shutil
    dst = os.path.join(os.path.dirname(OUT), "h5")
    os.makedirs(dst, exist_ok=True)
    for src in ("/root/reference/test/unit/data/flow_simulation/lhs/lhs.h5",
                "/root/reference/test/benchmark/tracer/oned/run/oned_two_phase_ss.h5",
                # an ExodusII mesh in the netCDF-4 container (HDF5 with version-2 object headers, links in a fractal
                # heap, attributes, a chunked dataset): the hexahedra + wedges mesh of the 3-D MINC benchmark (76 KB)
                "/root/reference/test/benchmark/minc/production3d/run/gminc_3d_refined.exo"):
        shutil.copy(src, os.path.join(dst, os.path.basename(src)))
        print("copied", src)
    # test/unit/src/initial_test.F90: the column meshes and the restart files (full, minimal, minimal on a MINC mesh)
    ini = os.path.join(os.path.dirname(OUT), "initial")
    os.makedirs(ini, exist_ok=True)
    for src in ("mesh/3D.exo", "mesh/block3.exo", "mesh/col100.exo", "mesh/col10.exo", "initial/fluid.h5", "initial/fluid_minimal.h5", "initial/fluid_minimal_minc.h5",
                "flow_simulation/mesh/4x3_2d.exo"):          # the last one: the mesh of source_setup_test.F90
        shutil.copy(os.path.join("/root/reference/test/unit/data", src), os.path.join(ini, os.path.basename(src)))
    # the restart file the tracer doublet input names ("initial": {"filename": "doublet_ss.h5"}), next to that input
    shutil.copy("/root/reference/test/benchmark/tracer/doublet/run/doublet_ss.h5",
                os.path.join(os.path.dirname(OUT), "inputs", "doublet_ss.h5"))
    # likewise for the 1-D tracer deck (tests/test_run.py runs it as a whole: restart, flow, tracer, output file)
    shutil.copy("/root/reference/test/benchmark/tracer/oned/run/oned_single_phase_ss.h5",
                os.path.join(os.path.dirname(OUT), "inputs", "oned_single_phase_ss.h5"))


def wae_benchmarks():
    """test/benchmark/ncg/{infiltration,heat_pipe} (eos wae: water, air, energy): AUTOUGH2 ELEMENT tables --
    test_infiltration.py compares the liquid saturation profiles (1e-4), test_heat_pipe.py P, T, Sv and the air
    mass fractions of the last output (5e-3) (the input decks: input_fixtures)"""
    import shutil
    base = "/root/reference/test/benchmark/ncg"
    inputs = os.path.join(os.path.dirname(OUT), "inputs")
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT tables of test/benchmark/ncg/{infiltration,heat_pipe}/run/"
                            "*.listing (AUTOUGH2); boundary blocks dropped",
           "columns": ["pressure", "temperature", "gas_saturation", "air_gas_mass_fraction", "air_liquid_mass_fraction",
                       "air_partial_pressure"]}
    for case, ncell in (("infiltration", 40), ("heat_pipe", 120)):
        run = os.path.join(base, case, "run")
        el = [(t, r) for k, t, r in listing_generic(os.path.join(run, case + ".listing")) if k == "E"]
        doc[case] = {"times": [t for t, _ in el], "tables": [[x[:6] for x in r[:ncell]] for _, r in el]}
    out = os.path.join(os.path.dirname(OUT), "wae_benchmarks.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def tracer_doublet():
    """test/benchmark/tracer/doublet: AUTOUGH2 listing of the injection / production doublet with a tracer pulse
    (100-cell row, producer on deliverability with a total-flow limiter, tracer diffusion) -- test_doublet.py compares
    the tracer mass fraction at every output (1e-3, absolute 1e-6) and the tracer production rate history (1e-3);
    also the steady state the Waiwera output file doublet_ss.h5 holds (P, T), the run's initial condition"""
    base = "/root/reference/test/benchmark/tracer/doublet/run"
    tabs = listing_generic(os.path.join(base, "doublet.listing"))
    el = [(t, r) for k, t, r in tabs if k == "E"]
    ge = [(t, r) for k, t, r in tabs if k == "G"]
    h5 = os.path.join(base, "doublet_ss.h5")
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/tracer/doublet/run/"
                            "doublet.listing (AUTOUGH2); doublet_ss.h5 by byte scan",
           "times": [t for t, _ in el], "pressure": [[x[0] for x in r[:100]] for _, r in el],
           "tracer": [[x[4] for x in r[:100]] for _, r in el],
           "source_times": [t for t, _ in ge], "production_rate": [r[1][0] for _, r in ge],
           "tracer_production": [r[1][2] for _, r in ge],
           "steady_pressure": [float(v) for v in runs(h5, 1e5, 1e8, 100)[0][:100]],
           "steady_temperature": [float(v) for v in runs(h5, 10, 400, 100)[1][:100]]}
    out = os.path.join(os.path.dirname(OUT), "tracer_doublet.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def minc_doublet():
    """test/benchmark/minc/doublet_1d: AUTOUGH2 final states of the 10-cell injection / production doublet, single
    porosity and MINC with one matrix level at fracture spacings 50 / 100 / 200 m -- test_minc_1d.py compares P, T, Sv
    of the last output at 2e-3"""
    base = "/root/reference/test/benchmark/minc/doublet_1d/run"
    doc = {"_generated_by": "tools/make_golden.py: last ELEMENT table of test/benchmark/minc/doublet_1d/run/"
                            "minc_1d_{single,50,100,200}.listing (AUTOUGH2): 10 fracture blocks, then 10 matrix blocks",
           "columns": ["pressure", "temperature", "vapour_saturation"]}
    for case in ("single", "50", "100", "200"):
        el = [(t, r) for k, t, r in listing_generic(os.path.join(base, "minc_1d_%s.listing" % case)) if k == "E"]
        src = json.load(open(os.path.join(base, "minc_1d_%s.json" % case)))
        doc[case] = {"time": el[-1][0], "element": [x[:3] for x in el[-1][1]], "stop": src["time"]["stop"],
                     "step": src["time"]["step"]}
    out = os.path.join(os.path.dirname(OUT), "minc_doublet.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


def minc_production3d():
    """test/benchmark/minc/production3d (3-D MINC production: well on deliverability with a productivity index
    stepped in time, base and refined (hexahedra + wedges) meshes): AUTOUGH2 listings.  Kept, in Waiwera's cell order
    (original cells, then the matrix cells level by level -- test_minc_3d.py::minc_level_map): P, T, Sv of every cell at
    the last output; their history in the cell that contains (10, 10, -1000); rate and enthalpy history of the well."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from waiwera_b200 import ingest
    base = "/root/reference/test/benchmark/minc/production3d/run"
    doc = {"_generated_by": "tools/make_golden.py: ELEMENT / GENERATION tables of test/benchmark/minc/production3d/run/"
                            "minc_3d_{base,refined}.listing (AUTOUGH2); atmosphere blocks dropped, matrix blocks reordered by level",
           "columns": ["pressure", "temperature", "vapour_saturation"]}
    for case in ("base", "refined"):
        inp = json.load(open(os.path.join(base, "minc_3d_%s.json" % case)))
        nodes, elems = ingest.read_exodus(os.path.join(base, inp["mesh"]["filename"]))
        ncell = len(elems)
        natm = len(inp["boundaries"])
        nlev = len(inp["mesh"]["minc"]["geometry"]["matrix"]["volume"])
        tabs = listing_generic(os.path.join(base, "minc_3d_%s.listing" % case))
        el = [(t, r) for k, t, r in tabs if k == "E"]
        ge = [(t, r) for k, t, r in tabs if k == "G"]
        nrow = len(el[-1][1])
        nmatrix = nrow - natm - ncell
        order = list(range(natm, natm + ncell))
        for l in range(nlev):
            order += list(range(natm + ncell + l, natm + ncell + nmatrix, nlev))
        pts = np.array(nodes)
        obs = [c for c, (_, ns) in enumerate(elems)
               if np.all(pts[ns].min(0) <= [10.0, 10.0, -1000.0]) and np.all(pts[ns].max(0) >= [10.0, 10.0, -1000.0])]
        assert len(obs) == 1
        doc[case] = {"ncell": ncell, "nminc": nmatrix // nlev, "times": [t for t, _ in el],
                     "final": [el[-1][1][i][:3] for i in order], "obs_cell": obs[0],
                     "history": [r[natm + obs[0]][:3] for _, r in el],
                     "source_times": [t for t, _ in ge], "source_rate": [r[-1][0] for _, r in ge],
                     "source_enthalpy": [r[-1][1] for _, r in ge]}
    out = os.path.join(os.path.dirname(OUT), "minc_production3d.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


FROM_INPUT = ["source/deliverability/run/deliv_delv.json", "source/deliverability/run/deliv_delt.json",
              "source/deliverability/run/deliv_delw.json", "source/deliverability/run/deliv_delg_flow.json",
              "source/deliverability/run/deliv_delg_limit.json", "source/deliverability/run/deliv_delg_pi_table.json",
              "source/deliverability/run/deliv_delg_pwb_table.json",
              "source/recharge/run/recharge_outflow.json",
              "minc/column/run/minc_column_single.json", "minc/column/run/minc_column_minc.json",
              "minc/doublet_1d/run/minc_1d_single.json", "minc/doublet_1d/run/minc_1d_50.json",
              "minc/doublet_1d/run/minc_1d_100.json", "minc/doublet_1d/run/minc_1d_200.json"]


def benchmarks_from_input():
    """the reference's remaining single-well benchmark decks, for runs from their own input files
    (tests/test_benchmarks_from_input.py): per deck the AUTOUGH2 listing's output times, P / T / Sv of every cell at the
    last output in Waiwera's cell order (atmosphere blocks dropped, MINC matrix blocks reordered level by level), their
    history in the first source's cell, and the rate / enthalpy history of every source"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from waiwera_b200 import ingest
    base = "/root/reference/test/benchmark"
    doc = {"_generated_by": "tools/make_golden.py::benchmarks_from_input from the *.listing files next to " + ", ".join(FROM_INPUT),
           "columns": ["pressure", "temperature", "vapour_saturation"]}
    for rel in FROM_INPUT:
        path = os.path.join(base, rel)
        name = os.path.splitext(os.path.basename(rel))[0]
        p = ingest.load(path)
        m = p.mesh
        ncell = getattr(m, "minc_cells", m.ninterior)
        nlev = m.minc_levels
        nmatrix = m.ninterior - ncell
        tabs = listing_generic(os.path.splitext(path)[0] + ".listing")
        el = [(t, r) for k, t, r in tabs if k == "E"]
        ge = [(t, r) for k, t, r in tabs if k == "G"]
        nrow = len(el[-1][1])
        # atmosphere blocks (boundaries on top faces) come first in the listing, other boundary blocks last
        inp = json.load(open(path))
        tops = [b for b in inp.get("boundaries") or [] if b["faces"]["normal"][-1] > 0.5 and len(b["faces"]["normal"]) == 3]
        natm = sum(len(b["faces"]["cells"]) for b in tops)
        nlast = sum(len(b["faces"]["cells"]) for b in inp.get("boundaries") or []) - natm
        assert nrow == natm + ncell + nmatrix + nlast and not (nmatrix and nlast), (name, nrow, natm, ncell, nmatrix, nlast)
        order = list(range(natm, natm + ncell))
        for l in range(nlev):
            order += list(range(natm + ncell + l, natm + ncell + nmatrix, nlev))
        cell = int(p.source_cells[0])
        nsrc = len(p.source_cells)
        assert all(len(r) == nsrc for _, r in ge), name
        doc[name] = {"ncell": ncell, "ninterior": m.ninterior, "times": [t for t, _ in el],
                     "final": [el[-1][1][i][:3] for i in order], "history_cell": cell,
                     "history": [r[natm + cell][:3] for _, r in el],
                     "source_times": [t for t, _ in ge], "rate": [[q[0] for q in r] for _, r in ge],
                     "enthalpy": [[q[1] for q in r] for _, r in ge]}
    out = os.path.join(os.path.dirname(OUT), "benchmarks_from_input.json")
    with open(out, "w") as f:
        json.dump(doc, f)
    print("wrote", out)


INPUT_KEYS = ("boundaries", "eos", "gravity", "initial", "mesh", "rock", "source", "thermodynamics", "time", "tracer")


def write_ascii_msh(path, nodes, elems):
    lines = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat", "$Nodes", str(len(nodes))]
    lines += ["%d %.17g %.17g %.17g" % (i + 1, x[0], x[1], x[2]) for i, x in enumerate(nodes)]
    lines += ["$EndNodes", "$Elements", str(len(elems))]
    lines += ["%d %d 2 0 0 %s" % (i + 1, t, " ".join(str(n + 1) for n in ns)) for i, (t, ns) in enumerate(elems)]
    lines += ["$EndElements", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


def convert_input(src_json, dst_dir, name=None):
    """Fixture of one reference input deck for the ingest tests: the keys of the JSON input the Newton-step path
    reads (title / output / logfile dropped), re-serialised compactly as <name>.input.json, and its gmsh mesh
    (binary MSH 2.2 in the reference tree) rewritten as ASCII MSH 2.2 <mesh>.ascii.msh with the same node and element
    numbering (coordinates printed with 17 significant digits: the doubles round-trip exactly)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from waiwera_b200 import ingest
    name = name or os.path.splitext(os.path.basename(src_json))[0]
    doc = json.load(open(src_json))
    out = {k: doc[k] for k in INPUT_KEYS if k in doc}
    mesh = out["mesh"] if isinstance(out["mesh"], dict) else {"filename": out["mesh"]}
    src_mesh = os.path.join(os.path.dirname(src_json), mesh["filename"])
    mesh_name = os.path.splitext(os.path.basename(src_mesh))[0] + ".ascii.msh"
    try:
        nodes, elems = ingest.read_mesh(src_mesh)
    except ValueError:
        # HDF5-based ExodusII: rebuild the mesh from the MULgraph geometry file it was generated from (same stem)
        nodes, elems = ingest.read_mulgraph(os.path.splitext(src_mesh)[0] + ".dat")
    write_ascii_msh(os.path.join(dst_dir, mesh_name), nodes, elems)
    out["mesh"] = dict(mesh, filename=mesh_name)
    if isinstance(out.get("initial"), dict) and "filename" in out["initial"]:
        out["initial"] = {"filename": out["initial"]["filename"]}      # HDF5 restart: the tests pass the arrays
    with open(os.path.join(dst_dir, name + ".input.json"), "w") as f:
        json.dump(out, f, sort_keys=True, separators=(",", ":"))
    return name


def input_fixtures():
    """tests/golden/inputs/: the reference's benchmark decks the ingest tests read (see convert_input)"""
    base = "/root/reference/test/benchmark"
    dst = os.path.join(os.path.dirname(OUT), "inputs")
    os.makedirs(dst, exist_ok=True)
    for rel in ("tracer/oned/run/oned_single_phase.json", "ncg/co2_column/run/co2_column_1.json",
                "model_intercomparison_study/problem1/run/problem1.json",
                "model_intercomparison_study/problem2/run/problem2a.json",
                "model_intercomparison_study/problem2/run/problem2b.json",
                "model_intercomparison_study/problem2/run/problem2c.json",
                "model_intercomparison_study/problem4/run/problem4.json",
                "model_intercomparison_study/problem5/run/problem5a.json",
                "model_intercomparison_study/problem5/run/problem5b.json",
                "model_intercomparison_study/problem6/run/problem6.json",
                "ncg/infiltration/run/infiltration.json", "ncg/heat_pipe/run/heat_pipe.json",
                "tracer/doublet/run/doublet.json", "tracer/doublet/run/doublet_ss.json",
                "minc/production3d/run/minc_3d_base.json", "minc/production3d/run/minc_3d_refined.json") + tuple(FROM_INPUT):
        convert_input(os.path.join(base, rel), dst)
    # mesh only: the reference's 3-D hybrid mesh (hexahedra + prisms) of its flow_simulation / initial unit tests
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from waiwera_b200 import ingest
    nodes, elems = ingest.read_gmsh("/root/reference/test/unit/data/mesh/hybrid10.msh", dmplex_order=False)      # the file's own order
    write_ascii_msh(os.path.join(dst, "hybrid10.ascii.msh"), nodes, elems)
    print("wrote", dst)


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
    co2_one_cell()
    co2_column()
    minc_column()
    mis_problems()
    deliverability()
    recharge()
    h5_fixtures()
    wae_benchmarks()
    tracer_doublet()
    minc_production3d()
    benchmarks_from_input()
    minc_doublet()
    input_fixtures()

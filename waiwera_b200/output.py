"""Waiwera's HDF5 output and restart files, on top of the library-free reader / writer of
h5lite.py.

Layout (src/flow_simulation.F90 output routines, PETSc HDF5 viewer with an output sequence):
    time                                  [ntimes, 1]
    cell_index                            [ncells, 1] int32: for every natural cell index the position its values are stored
                                          at (mesh%cell_index, "natural to global": dm_get_cell_index,
                                          src/dm_utils.F90:974-1037; pinned by the reference's 4-process restart file
                                          fluid_minimal_minc.h5, tests/test_ingest.py::test_initial_conditions_known_answers)
    cell_fields/cell_geometry_centroid    [ncells, dim]      cell_fields/cell_geometry_volume [ncells]
    cell_fields/fluid_<field>             [ntimes, ncells]   field names as create_fluid_vector builds them
                                          (src/fluid.F90): pressure, temperature, region, <component>_partial_pressure,
                                          <phase>_<variable>, e.g. vapour_saturation, liquid_density
    cell_fields/tracer_<name>             [ntimes, ncells] tracer mass fractions
    minc/level, minc/parent               [ncells, 1] int32 (MINC meshes): level of every cell, natural index of its original cell
    source_index                          [nsources, 1] int32
    source_fields/source_<field>          [ntimes, nsources]: natural_cell_index (int32), component, rate, enthalpy
                                          (default_output_source_fields, src/source.F90:60-62)
A restart (src/initial.F90:421-507, "initial": {"filename": ..., "index": ...}) loads the EOS's
required_output_fluid_fields at one time index, reorders them with cell_index, and takes the primary variables from
them (eos%primary_variables: pressure, then temperature or -- two-phase -- vapour saturation, then the gas partial
pressure)."""
import numpy as np

from . import h5lite

# bulk entries of a fluid record (src/fluid.F90:236-262)
_BULK = {"pressure": 0, "temperature": 1, "region": 2, "old_region": 3, "phase_composition": 4, "permeability_factor": 5}
_PHASE_VARS = ["density", "viscosity", "saturation", "relative_permeability", "capillary_pressure", "specific_enthalpy",
               "internal_energy"]
_EOS = {"w": (["water"], ["liquid"]), "we": (["water"], ["liquid", "vapour"]),
        "wce": (["water", "CO2"], ["liquid", "vapour"]), "wae": (["water", "air"], ["liquid", "vapour"])}
# required_output_fluid_fields = default_output_fluid_fields of each EOS (src/eos_w.F90:83, eos_we.F90:93-98,
# eos_wce.F90:42-48, eos_wae.F90:45-51)
REQUIRED = {"w": ["pressure", "region"],
            "we": ["pressure", "temperature", "region", "vapour_saturation"],
            "wce": ["pressure", "temperature", "region", "CO2_partial_pressure", "vapour_saturation"],
            "wae": ["pressure", "temperature", "region", "air_partial_pressure", "vapour_saturation"]}
SOURCE_FIELDS = ["natural_cell_index", "component", "rate", "enthalpy"]


def fluid_field_column(eos, name):
    """column of the fluid record (7 + nc - 1 + nphase (8 + nc - 1) doubles) that holds the named field"""
    comps, phases = _EOS[eos]
    nc = len(comps)
    if name in _BULK:
        return _BULK[name]
    for j, c in enumerate(comps):
        if name.lower() == (c + "_partial_pressure").lower():
            return 6 + j
    base = 7 + nc - 1
    for p, ph in enumerate(phases):
        for k, v in enumerate(_PHASE_VARS):
            if name == ph + "_" + v:
                return base + p * (8 + nc - 1) + k
        for j, c in enumerate(comps):
            if name.lower() == (ph + "_" + c + "_mass_fraction").lower():
                return base + p * (8 + nc - 1) + 7 + j
    raise KeyError("fluid field %r of eos %s" % (name, eos))


def write_output(path, mesh, eos, times, fluids, source_cells=None, source_history=None, fields=None, cell_index=None,
                 tracer_names=None, tracer_history=None):
    """times: [nt]; fluids: nt arrays [>= ninterior, dof] of fluid records (wb_get_fluid) in natural cell order;
    source_history: nt arrays [nsources, 3] of (component, rate, enthalpy); fields: fluid fields to write (default: the
    EOS's default output fields); cell_index: storage order, the natural index of the cell stored at each position
    (default: natural order) -- the file's cell_index dataset is its inverse."""
    n = mesh.ninterior
    fields = list(fields or REQUIRED[eos])
    order = np.arange(n, dtype=np.int32) if cell_index is None else np.asarray(cell_index, np.int32)
    position = np.zeros(n, np.int32)
    position[order] = np.arange(n, dtype=np.int32)
    d = {"time": np.asarray(times, float).reshape(-1, 1), "cell_index": position.reshape(-1, 1),
         "cell_fields/cell_geometry_centroid": np.asarray(mesh.cell_geom, float)[:n, :3][order],
         "cell_fields/cell_geometry_volume": np.asarray(mesh.cell_geom, float)[:n, 3][order]}
    for name in fields:
        col = fluid_field_column(eos, name)
        d["cell_fields/fluid_" + name] = np.array([np.asarray(fl)[:n, col][order] for fl in fluids], float).reshape(len(times), n)
    if tracer_names:
        # tracer mass fractions: cell_fields/tracer_<name> (create_tracer_vector field names, src/tracer.F90:152-191)
        hist = [np.asarray(x, float).reshape(-1, len(tracer_names)) for x in tracer_history]
        for j, name in enumerate(tracer_names):
            d["cell_fields/tracer_" + name] = np.array([x[:n, j][order] for x in hist], float).reshape(len(times), n)
    if getattr(mesh, "minc_level", None) is not None and mesh.minc_levels > 0:
        # flow_simulation_output_minc_data (src/flow_simulation.F90:2625-2691): MINC level and natural index of the
        # original single-porosity cell of every cell, in storage order
        d["minc/level"] = np.asarray(mesh.minc_level, np.int32)[:n][order].reshape(-1, 1)
        d["minc/parent"] = np.asarray(mesh.minc_parent, np.int32)[:n][order].reshape(-1, 1)
    if source_cells is not None and len(source_cells):
        ns = len(source_cells)
        hist = np.asarray(source_history, float).reshape(len(times), ns, 3)
        d["source_index"] = np.arange(ns, dtype=np.int32).reshape(-1, 1)
        d["source_fields/source_natural_cell_index"] = np.tile(np.asarray(source_cells, np.int32), (len(times), 1))
        d["source_fields/source_component"] = hist[:, :, 0]
        d["source_fields/source_rate"] = hist[:, :, 1]
        d["source_fields/source_enthalpy"] = hist[:, :, 2]
    h5lite.write(path, d)


def read_restart(path, eos, index=-1):
    """-> (primary [ncells, np], region [ncells] int32, time) in natural cell order from a Waiwera output file"""
    h = h5lite.H5File(path)
    t = h["time"].reshape(-1)
    k = index if index >= 0 else len(t) + index
    if not 0 <= k < len(t):
        raise IndexError("time index %d of %d" % (index, len(t)))
    cell_index = h["cell_index"].reshape(-1).astype(np.int64)
    n = len(cell_index)

    def field(name):
        for key in ("cell_fields/fluid_" + name, "cell_fields/fluid_" + name.lower()):
            if key in h:
                a = h[key]
                v = a[k] if a.ndim == 2 and a.shape[0] == len(t) else a.reshape(-1)
                return np.asarray(v, float)[cell_index]          # natural cell i is stored at position cell_index[i]
        raise KeyError("%s has no field %r (required for a restart of eos %s)" % (path, name, eos))
    f = {name: field(name) for name in REQUIRED[eos]}
    region = np.rint(f["region"]).astype(np.int32)
    cols = [f["pressure"]]
    if eos != "w":
        cols.append(np.where(region == 4, f["vapour_saturation"], f["temperature"]))
    if eos in ("wce", "wae"):
        cols.append(f[REQUIRED[eos][3]])
    return np.stack(cols, 1), region, float(t[k])

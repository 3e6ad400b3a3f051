"""Test infrastructure: the reference's 10-cell deliverability deck (deliv_delw) with its rock type split into two types
whose permeability and porosity are tables in time (src/rock_setup.F90:383-463), for the CPU and GPU tests of
wb_set_rock / ingest.rock_at / run.py."""
import json
import os
import shutil

from test_benchmarks_from_input import INP


def write_deck(tmp_path, nsteps=12):
    doc = json.load(open(os.path.join(INP, "deliv_delw.input.json")))
    shutil.copy(os.path.join(INP, doc["mesh"]["filename"]), str(tmp_path / doc["mesh"]["filename"]))
    base = dict(doc["rock"]["types"][0])
    near = dict(base, name="near", cells=[0, 1, 2, 3, 4], interpolation="step",
                permeability=[[0.0, 1e-13], [3.0e4, 6e-14], [1.2e5, 3e-14]])
    far = dict(base, name="far", cells=[5, 6, 7, 8, 9],
               permeability=[[0.0, 1e-13, 1e-13, 1e-13], [1.0e5, 2e-13, 1e-13, 1e-13], [3.0e5, 5e-14, 1e-13, 1e-13]],
               porosity=[[0.0, 0.1], [1.0e5, 0.09998], [3.0e5, 0.0999]])
    doc["rock"]["types"] = [near, far]
    doc["time"]["step"]["maximum"]["number"] = nsteps
    path = str(tmp_path / "rock_tables.json")
    json.dump(doc, open(path, "w"))
    return path

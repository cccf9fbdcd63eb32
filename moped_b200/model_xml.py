"""Writer of `.moped.xml` model files in the layout the reference's modelling tools emit
(moped-modeling-py/src/MopedModeling.py:900-1000; moped3d/modeling/sfm_export_xml.m): used to turn synthetic
databases into files for the loader (mc_model_db_*) and its tests. Inputs only — parsing lives in
moped_b200/csrc/model_db.cpp.

    <Model name="..." version="...">
      <Points>
        <Point p3d="x y z" nviews="n" avg_err="e" color="R G B" desc_type="SIFT" desc="a b c ... ">
          <Observation camera_id="n" desc_type="SIFT" loc="x y scale ori " desc="a b c ... "/>     (full export only)
        </Point>
      </Points>
      <Cameras> <Camera id="n" rot_type="quat" rot="w x y z" tx="x y z"/> </Cameras>               (full export only)
    </Model>
"""
from __future__ import annotations

import numpy as np


def _fmt(values, exact: bool) -> str:
    # the tools print '{0:6f} ' per value (6 decimals); exact=True keeps every float32 bit ('%.9g')
    if exact:
        return "".join("%.9g " % float(v) for v in values)
    return "".join("{0:6f} ".format(float(v)) for v in values)


def write_model_xml(path, name, xyz, desc, desc_type="SIFT", exact=True, full_export=False, rng=None, version="Bundler v0.4"):
    """xyz [n,3], desc [n,D]; desc_type a string or a list of n strings (one model may mix descriptor types)."""
    xyz = np.asarray(xyz, dtype=np.float32)
    n = len(xyz)
    types = [desc_type] * n if isinstance(desc_type, str) else list(desc_type)
    rng = rng or np.random.default_rng(0)
    with open(path, "w") as f:
        f.write('<Model name="{0}" version="{1}">\n'.format(name, version))
        f.write("  <Points>\n")
        for i in range(n):
            d = np.asarray(desc[i], dtype=np.float32)
            f.write('    <Point p3d="{0}" nviews="{1:d}" avg_err="{2:6f}" color="{3} {4} {5}" desc_type="{6}" desc="{7}">\n'.format(
                _fmt(xyz[i], exact).rstrip(), 3, 0.25, 128, 64, 32, types[i], _fmt(d, exact)))
            if full_export:
                for cam in range(2):
                    f.write('      <Observation camera_id="{0}" desc_type="{1}" loc="{2}" desc="{3}"/>\n'.format(
                        cam, types[i], _fmt(rng.uniform(0, 100, 4), False), _fmt(d + np.float32(0.001) * (cam + 1), False)))
            f.write("</Point>\n")
        f.write("  </Points>\n")
        if full_export:
            f.write("  <Cameras>\n")
            for cam in range(2):
                f.write('    <Camera id="{0}" rot_type="quat" rot="1 0 0 0" tx="0 0 {0}"/>\n'.format(cam))
            f.write("  </Cameras>\n")
        f.write("</Model>\n")


def write_db_xml(directory, db, exact=True, prefix="obj"):
    """One file per object of a synth.make_db database; returns the list of paths in model order."""
    import os
    paths = []
    starts = np.concatenate([[0], np.cumsum(db["n_pts"])])
    for m in range(len(db["n_pts"])):
        p = os.path.join(directory, "%s%06d.moped.xml" % (prefix, m))
        write_model_xml(p, "%s%06d" % (prefix, m), db["xyz"][starts[m]:starts[m + 1]], db["desc"][starts[m]:starts[m + 1]], exact=exact)
        paths.append(p)
    return paths

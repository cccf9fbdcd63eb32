"""Model-database loader (mc_model_db_*) against the reference's own sXML reader and addModel
loop (oracle/_ref: ref_add_model_xml) on files in the modelling tools' format, plus the edge cases of the XML subset."""
import os

import numpy as np
import pytest

from moped_b200 import capi, model_xml, synth


def _same_model(db, r, i, types=("SIFT",)):
    assert db.names()[i] == r.model_names()[i]
    for t in types:
        xyz, ln, vals = r.model_points(i, t)
        raw = _points(db, i, t)
        assert np.array_equal(raw[0], xyz), (i, t)
        assert np.array_equal(raw[1], ln), (i, t)
        assert np.array_equal(raw[2], vals), (i, t)     # bit-exact: both sides convert with correctly rounded strtof
    if sum(len(r.model_points(i, t)[0]) for t in types):
        assert np.array_equal(db.bbox(i), r.model_bbox(i))


def _points(db, i, t):
    """(xyz, desc_len, values) of model i / type t out of the packed view (only model i present -> use a fresh db)."""
    return db._raw[i][t]


class _DB(capi.ModelDB):
    """ModelDB plus a per-model raw view obtained by packing one-model databases (test helper)."""

    def __init__(self):
        super().__init__()
        self._files = []

    def add(self, path):
        self.add_xml_file(path)
        self._files.append(path)

    def raw(self, types, sizes):
        self._raw = []
        for p in self._files_by_model():
            one = capi.ModelDB()
            one.add_xml_file(p)
            per = {}
            for t in types:
                got = None
                for D in sizes:
                    try:
                        pk = one.pack(t, D, normalise=False)
                    except capi.MopedCudaError:
                        continue
                    got = (pk["xyz"], np.full(len(pk["xyz"]), D, np.int32), pk["desc"].reshape(-1))
                    break
                per[t] = got if got is not None else (np.zeros((0, 3), np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32))
            self._raw.append(per)

    def _files_by_model(self):
        # replay of addModel's list update on (name, file) pairs
        lst = []
        for p in self._files:
            one = capi.ModelDB()
            one.add_xml_file(p)
            nm = one.names()[0]
            found = False
            for k in range(len(lst)):
                found = lst[k][0] == nm
                if found:
                    lst[k] = (nm, p)
            if not found:
                lst.append((nm, p))
        return [p for _, p in lst]


def test_tool_format_files_equal_reference(tmp_path, ref_mod):
    db3 = synth.make_db(3, 40, d=128, seed=5, ragged=True)
    paths = model_xml.write_db_xml(str(tmp_path), db3, exact=True)
    # the tools' own 6-decimal formatting, with Observation children and Cameras (full export)
    starts = np.concatenate([[0], np.cumsum(db3["n_pts"])])
    p6 = str(tmp_path / "six.moped.xml")
    model_xml.write_model_xml(p6, "six decimals", db3["xyz"][:30], db3["desc"][:30], exact=False, full_export=True)
    # a model with two descriptor types, SURF-64 interleaved with SIFT-128
    rng = np.random.default_rng(3)
    mixed_desc = [rng.random(64 if k % 3 == 0 else 128).astype(np.float32) for k in range(20)]
    pm = str(tmp_path / "mixed.moped.xml")
    model_xml.write_model_xml(pm, "mixed", rng.normal(size=(20, 3)), mixed_desc, desc_type=["SURF" if k % 3 == 0 else "SIFT" for k in range(20)])
    # same name as the first object: replaces it in place AND is appended again, because the reference's `found` flag
    # only remembers the comparison with the last list entry (moped.cpp:140-146); same name as the last entry: replaced only
    pr = str(tmp_path / "replacement.moped.xml")
    model_xml.write_model_xml(pr, "obj000000", db3["xyz"][50:60], db3["desc"][50:60])
    pr2 = str(tmp_path / "replacement2.moped.xml")
    model_xml.write_model_xml(pr2, "obj000000", db3["xyz"][50:60], db3["desc"][50:60])
    files = paths + [p6, pm, pr, pr2]
    r = ref_mod.Ref(1)
    db = _DB()
    for p in files:
        assert r.add_model_xml(p) == 1
        db.add(p)
    assert db.names() == r.model_names() == ["obj000000", "obj000001", "obj000002", "six decimals", "mixed", "obj000000"]
    db.raw(("SIFT", "SURF"), (128, 64))
    for i in range(6):
        _same_model(db, r, i, ("SIFT", "SURF"))
    # exact=True files reproduce the synthetic arrays bit for bit
    pk = db.pack("SIFT", 128, normalise=False)
    assert np.array_equal(pk["n_pts"][1:3], db3["n_pts"][1:3]) and pk["n_pts"][0] == 10
    assert np.array_equal(pk["desc"][10:10 + db3["n_pts"][1]], db3["desc"][starts[1]:starts[2]])
    assert np.array_equal(pk["xyz"][10:10 + db3["n_pts"][1]], db3["xyz"][starts[1]:starts[2]])
    assert np.array_equal(pk["model_of_row"], np.repeat(np.arange(6), pk["n_pts"]).astype(np.int32))


def test_parallel_load_equals_serial_and_cache_roundtrip(tmp_path):
    db8 = synth.make_db(8, 25, seed=9, ragged=True)
    paths = model_xml.write_db_xml(str(tmp_path), db8)
    a, b = capi.ModelDB(), capi.ModelDB()
    for p in paths:
        a.add_xml_file(p)
    b.add_xml_files(paths, n_threads=4)
    pa, pb = a.pack("SIFT", 128, True), b.pack("SIFT", 128, True)
    for k in pa:
        assert np.array_equal(pa[k], pb[k]), k
    assert np.array_equal(a.pack("SIFT", 128, False)["desc"], db8["desc"])
    cache = str(tmp_path / "models.mopedb")
    a.save(cache)
    c = capi.ModelDB()
    c.load(cache)
    assert c.names() == a.names()
    pc = c.pack("SIFT", 128, True)
    for k in pa:
        assert np.array_equal(pa[k], pc[k]), k
    for i in range(8):
        assert np.array_equal(a.bbox(i), c.bbox(i))
    # removal (Moped::removeModel)
    c.remove("obj000003")
    assert len(c.names()) == 7 and c.pack("SIFT", 128, False)["n_pts"].sum() == db8["n_pts"].sum() - db8["n_pts"][3]
    # normalisation = the stage class's expression
    d = db8["desc"]
    inv = (np.float32(1.0) / np.sqrt((d * d).sum(axis=1, dtype=np.float32))).astype(np.float32)
    assert np.allclose(pa["desc"], d * inv[:, None], rtol=0, atol=1e-7)


XML_EDGE = b"""<!-- leading comment -- with dashes -->
<Model   name="edge \\"case\\"\\nline2" version="1">
  <Openrave><name>x</name><xml>y.xml</xml><transf>1 0 0</transf></Openrave>
  <Points><Point p3d="9 9 9" desc_type="SIFT" desc="9 9"/></Points>
  <!-- the LAST Points element wins -->
  <Points>
    <Point p3d="1.5 -2.25 3e-1" desc_type="SIFT" desc="1 2 3 4"/>
    <Pt p3d="4 5" desc_type="SIFT" desc="5 6 7 8 x 9"></Pt>
    <Point desc="1e 2" p3d="+.5 -.5 5." desc_type="SIFT"  />
    <Point p3d="7 8 9" desc="10 11 12 13"/>
    <Point p3d="1e39 0 0" desc_type="SIFT" desc="1 1e-46 3 4"/>
  </Points>
  <Cameras K="1;2;3;4"><Camera id="0"/></Cameras>
</Model>
"""


def test_xml_subset_edge_cases(tmp_path, ref_mod):
    p = str(tmp_path / "edge.moped.xml")
    with open(p, "wb") as f:
        f.write(XML_EDGE)
    r = ref_mod.Ref(1)
    assert r.add_model_xml(p) == 1
    db = capi.ModelDB()
    db.add_xml_file(p)
    assert db.names() == r.model_names() == ['edge "case"\nline2']
    # SIFT points as the reference sees them: lengths 4, 4 ("x" stops the reading), 0 ("1e" fails), 4 (1e-46 underflows to 0)
    xyz, ln, vals = r.model_points(0, "SIFT")
    assert ln.tolist() == [4, 4, 0, 4]
    # the point without desc_type is filed under "" by both
    xyz0, ln0, vals0 = r.model_points(0, "")
    assert ln0.tolist() == [4] and vals0.tolist() == [10, 11, 12, 13]
    pk0 = db.pack("", 4, normalise=False)
    assert np.array_equal(pk0["desc"].reshape(-1), vals0) and np.array_equal(pk0["xyz"], xyz0)
    # a descriptor shorter than DescriptorSize is an error here (the reference would read past the vector)
    with pytest.raises(capi.MopedCudaError, match="fewer than DescriptorSize"):
        db.pack("SIFT", 4, normalise=False)
    # ... so compare through a copy of the file without the broken point
    p2 = str(tmp_path / "edge2.moped.xml")
    with open(p2, "wb") as f:
        f.write(XML_EDGE.replace(b'<Point desc="1e 2" p3d="+.5 -.5 5." desc_type="SIFT"  />', b'<Point desc="1 2 3 4e0" p3d="+.5 -.5 5." desc_type="SIFT"  />'))
    r2 = ref_mod.Ref(1)
    r2.add_model_xml(p2)
    db2 = capi.ModelDB()
    db2.add_xml_file(p2)
    xyz, ln, vals = r2.model_points(0, "SIFT")
    pk = db2.pack("SIFT", 4, normalise=False)
    assert ln.tolist() == [4, 4, 4, 4]
    assert np.array_equal(pk["desc"].reshape(-1), vals)
    # p3d="4 5" leaves z unset in the reference; p3d="1e39 0 0" fails on the first value: compare the defined ones
    assert np.array_equal(pk["xyz"][[0, 2]], xyz[[0, 2]]) and np.array_equal(pk["xyz"][1, :2], xyz[1, :2])


def test_number_conversion_equals_istream(tmp_path, ref_mod):
    """The loader's decimal -> float conversion (exact fast path + strtof fallback) against `istream >> float` of the
    reference's load on numbers of every shape, including exact float midpoints (ties), long mantissas, subnormals and
    values that need the slow path."""
    rng = np.random.default_rng(11)
    toks = ["16777217", "16777219", "16777217.0000001", "16777216.9999999", "33554434", "33554438", "0.1", "1e-45", "1.4e-45",
            "1.17549435e-38", "1.17549428e-38", "3.4028234e38", "-0", "+5.", "-.5e-3", "123456789012345678", "0.000000000000000000001234567890123",
            "1e22", "1e23", "9007199254740993", "1.00000005960464477539", "1.00000017881393432617"]
    for _ in range(4000):
        k = rng.integers(0, 6)
        if k == 0:
            toks.append("%.9g" % np.float32(rng.normal() * 10.0 ** rng.integers(-8, 8)))
        elif k == 1:
            toks.append("{0:6f}".format(rng.normal()))
        elif k == 2:
            toks.append("%.17g" % (rng.normal() * 10.0 ** rng.integers(-30, 30)))
        elif k == 3:
            toks.append(str(int(rng.integers(0, 2 ** 26)) * 2 + 1))                      # odd integers up to 2^27: many float midpoints
        elif k == 4:
            toks.append("%d.%s" % (rng.integers(0, 1000), "".join(str(d) for d in rng.integers(0, 10, size=rng.integers(1, 25)))))
        else:
            toks.append("%.6e" % (rng.normal() * 10.0 ** rng.integers(-44, 38)))
    rows = [toks[i:i + 8] for i in range(0, len(toks) - 7, 8)]
    p = str(tmp_path / "numbers.moped.xml")
    with open(p, "w") as f:
        f.write('<Model name="numbers"><Points>\n')
        for r in rows:
            f.write('<Point p3d="%s %s %s" desc_type="N" desc="%s"/>\n' % (r[0], r[1], r[2], " ".join(r)))
        f.write("</Points></Model>\n")
    r = ref_mod.Ref(1)
    assert r.add_model_xml(p) == 1
    xyz, ln, vals = r.model_points(0, "N")
    assert (ln == 8).all() and len(ln) == len(rows)
    db = capi.ModelDB()
    db.add_xml_file(p)
    pk = db.pack("N", 8, normalise=False)
    assert np.array_equal(pk["desc"].reshape(-1).view(np.uint32), vals.view(np.uint32))
    assert np.array_equal(pk["xyz"].view(np.uint32), xyz.view(np.uint32))


def test_loader_errors_are_loud(tmp_path):
    db = capi.ModelDB()
    with pytest.raises(capi.MopedCudaError, match="cannot open"):
        db.add_xml_file(str(tmp_path / "missing.xml"))
    with pytest.raises(capi.MopedCudaError, match="no <Points>"):
        db.add_xml_buffer(b'<Model name="a"><Cameras/></Model>')
    with pytest.raises(capi.MopedCudaError):
        db.add_xml_buffer(b'<Model name="a"><Points><Point p3d="1 2 3" desc="1 2')       # truncated
    bad = str(tmp_path / "bad.mopedb")
    with open(bad, "wb") as f:
        f.write(b"not a cache")
    with pytest.raises(capi.MopedCudaError, match="not a moped model cache"):
        db.load(bad)
    assert db.names() == []


@pytest.mark.gpu
def test_upload_from_files_matches_like_direct_upload(tmp_path, oracle_mod):
    """files -> mc_model_db_upload -> MATCH gives the rows mc_db_upload of the same arrays gives (and the oracle's)."""
    db = synth.make_db(6, 200, seed=21)
    fr = synth.make_frame(db, 500, n_visible=3, pts_visible=40, seed=21)
    paths = model_xml.write_db_xml(str(tmp_path), db)
    mdb = capi.ModelDB()
    mdb.add_xml_files(paths)
    ctx = capi.Context(0)
    mdb.upload(ctx, "SIFT", 128)
    pk = mdb.pack("SIFT", 128, True)
    qn = oracle_mod.norm_rows(fr["desc"])
    rows, dist, acc, _ = ctx.match(qn, 0.8, capi.MATCH_TENSOR)
    oidx, odist = oracle_mod.match_2nn(pk["desc"], qn)
    assert np.array_equal(rows, oidx) and np.array_equal(dist, odist)
    ctx2 = capi.Context(0)
    ctx2.db_upload(pk["desc"], pk["xyz"], pk["model_of_row"], 6)
    rows2, dist2, acc2, _ = ctx2.match(qn, 0.8, capi.MATCH_TENSOR)
    assert np.array_equal(rows, rows2) and np.array_equal(dist, dist2) and np.array_equal(acc, acc2)

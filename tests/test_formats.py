"""On-disk formats: CSV / pickle round trips, checked against the reference's own util/io.py and
util/meta.py when /root/reference is present (CPU container); behaviour restated below otherwise."""
import importlib.util
import os
import pickle

import numpy as np
import pytest

from soft_contrastive_learning_b200 import formats

REF_IO = "/root/reference/util/io.py"
REF_META = "/root/reference/util/meta.py"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref_io():
    if not os.path.exists(REF_IO):
        pytest.skip("reference sources not present on this box")
    pytest.importorskip("cv2")                       # util/io.py imports it at module level
    return _load(REF_IO, "ref_util_io"), _load(REF_META, "ref_util_meta")


def _meta(n=7, seed=0):
    rng = np.random.default_rng(seed)
    return {"date": ["2015-0{}-1{}".format(i % 9 + 1, i % 9) for i in range(n)], "folder": ["stereo"] * n,
            "t": [int(t) for t in rng.integers(1e15, 2e15, n)],
            "easting": [float(x) for x in rng.uniform(6e5, 7e5, n)], "northing": [float(x) for x in rng.uniform(5e6, 6e6, n)]}


def test_csv_round_trip_and_xy(tmp_path):
    meta = _meta()
    f = tmp_path / "m.csv"
    formats.save_csv(meta, f)
    back = formats.load_csv(f)
    assert list(back.keys()) == list(meta.keys())
    assert all(back[k] == [str(v) for v in meta[k]] for k in meta)      # values come back as strings (util/io.py:79)
    xy = formats.get_xy(back)
    assert xy.dtype == np.float64 and xy.shape == (7, 2)
    assert np.array_equal(xy, np.array([meta["easting"], meta["northing"]]).T)
    text = open(f).read()
    assert not text.endswith("\n") and text.splitlines()[0] == "date,folder,t,easting,northing"


def test_csv_edge_cases(tmp_path):
    f = tmp_path / "h.csv"
    f.write_text("a,b,c")
    assert formats.load_csv(f) == ["a", "b", "c"]                      # header only -> the keys (util/io.py:80-83)
    g = tmp_path / "n.csv"
    g.write_text("1;2\n3;4")
    assert formats.load_csv(g, delimiter=";", has_header=False, keys=["x", "y"]) == {"x": ["1", "3"], "y": ["2", "4"]}
    got = formats.load_csv(g, delimiter=";", has_header=False)        # no usable keys -> 0..n-1
    assert [int(k) for k in got] == [0, 1] and list(got.values()) == [["1", "3"], ["2", "4"]]
    s = tmp_path / "s.csv"
    formats.save_csv({"loss": 0.5, "step": 3}, s)                     # scalar form (util/io.py:102-104)
    assert open(s).read() == "loss,step\n0.5,3"


def test_feature_pickle_layout(tmp_path):
    feats = np.random.default_rng(1).standard_normal((5, 32)).astype(np.float32)
    f = tmp_path / "lv.pickle"
    formats.save_features(feats, f)
    raw = pickle.load(open(f, "rb"))
    assert isinstance(raw, list) and len(raw) == 5 and raw[0].dtype == np.float32 and raw[0].shape == (32,)
    back = formats.load_features(f)
    assert back.dtype == np.float32 and np.array_equal(back, feats)


def test_same_bytes_and_values_as_reference_io(tmp_path, ref_io):
    rio, rmeta = ref_io
    meta = _meta(11, seed=3)
    a, b = tmp_path / "a.csv", tmp_path / "b.csv"
    formats.save_csv(meta, a)
    rio.save_csv(meta, b)
    assert open(a).read() == open(b).read()
    assert formats.load_csv(a) == rio.load_csv(b)
    assert np.array_equal(formats.get_xy(formats.load_csv(a)), rmeta.get_xy(rio.load_csv(b)))
    h = tmp_path / "h.csv"
    h.write_text("x,y")
    assert formats.load_csv(h) == rio.load_csv(h)
    n = tmp_path / "n.csv"
    n.write_text("1,2\n3,4")
    got, ref = formats.load_csv(n, has_header=False), rio.load_csv(n, has_header=False)
    assert [int(k) for k in got] == [int(k) for k in ref] and list(got.values()) == list(ref.values())
    payload = [[[1, 2]], [[0.5, 1.5]], np.ones((1, 2)), [3], np.array([0.25]), [0, 1, 2]]
    p, q = tmp_path / "p.pickle", tmp_path / "q.pickle"
    formats.save_pickle(payload, p)
    rio.save_pickle(payload, q)
    assert open(p, "rb").read() == open(q, "rb").read()
    assert repr(formats.load_pickle(q)) == repr(rio.load_pickle(p))

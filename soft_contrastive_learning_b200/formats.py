"""On-disk formats either side of the retrieval path.

Host-side readers/writers with the behaviour of the reference's ``util/io.py`` and ``util/meta.py`` for the three
artefacts ``evaluation/top-n.py`` touches, so its producers and consumers (``evaluation/inference.py:192``,
``evaluation/roc.py``) keep working unchanged around the B200 retrieval:

* feature pickles: a Python list of ``float32[D]`` arrays (``evaluation/inference.py:192``; read at ``top-n.py:65-67``),
* meta CSVs with ``easting`` / ``northing`` columns (``util/io.py:46-83``, ``util/meta.py:4``),
* the top-N pickle ``[top_i, top_g_dists, top_f_dists, gt_i, gt_g_dist, ref_idx]`` (``top-n.py:119``).
"""
from __future__ import annotations

import csv
import pickle

import numpy as np


def load_pickle(in_file):
    """util/io.py:35-37."""
    with open(in_file, "rb") as f:
        return pickle.load(f)


def save_pickle(data, out_file):
    """util/io.py:40-42 (default pickle protocol, like the reference)."""
    with open(out_file, "wb") as f:
        pickle.dump(data, f)


def load_csv(in_file, delimiter=",", has_header=True, keys=()):
    """util/io.py:46-83: a dict ``column -> list of str`` (values stay strings).  Without a header the columns are
    named by ``keys`` when their number matches the first row, else 0..n-1.  A file that holds nothing but its header
    row yields the list of column names, as the reference does."""
    table = {}
    names = list(keys)
    body_rows = 0
    with open(in_file) as f:
        for n, row in enumerate(csv.reader(f, delimiter=delimiter)):
            if n == 0:
                if has_header:
                    names = list(row)
                elif len(names) != len(row):
                    names = list(np.arange(len(row)))
                table = {name: [] for name in names}
                if has_header:
                    continue
            body_rows += 1
            for name, value in zip(names, row):
                table[name].append(value)
            if len(row) < len(names):
                raise IndexError("list index out of range")       # the reference indexes row[i] for every key
    return table if body_rows else names


def save_csv(data, out_file, delimiter=","):
    """util/io.py:86-105: header line, then one line per row (or a single line of scalars); no trailing newline."""
    cols = list(data.keys())
    out = [delimiter.join(str(c) for c in cols)]
    if isinstance(data[cols[0]], list):
        out.extend(delimiter.join(str(data[c][r]) for c in cols) for r in range(len(data[cols[0]])))
    else:
        out.append(delimiter.join(str(data[c]) for c in cols))
    with open(out_file, "w") as f:
        f.write("\n".join(out))


def get_xy(meta):
    """util/meta.py:4: float64 [n, 2] of (easting, northing)."""
    return np.array([[e, n] for e, n in zip(meta["easting"], meta["northing"])], dtype=float)


def save_features(features, out_file):
    """The feature pickle of evaluation/inference.py:192: a list with one float32 vector per image."""
    save_pickle([np.asarray(f, dtype=np.float32) for f in features], out_file)


def load_features(in_file):
    """top-n.py:65-67: ``np.array(load_pickle(...))`` -> float32 [n, D]."""
    return np.array(load_pickle(in_file))

import json
import math
from pathlib import Path

import numpy as np

KATS = json.loads((Path(__file__).parent / "golden" / "reference_kats.json").read_text())


def num(v):
    if isinstance(v, str):
        return {"nan": math.nan, "inf": math.inf, "-inf": -math.inf}[v]
    return float(v)


def arr(spec):
    data = np.array([num(v) for v in spec["data"]], dtype=np.float64)
    return data.reshape(spec["shape"], order="F")


def assert_same(got, want, tol=0.0):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), (got, want)
    g, w = got[~nan_g], want[~nan_w]
    if tol == 0.0:
        assert np.array_equal(g, w), (got, want)
    else:
        assert np.all(np.abs(g - w) <= tol * np.maximum(1.0, np.abs(w))), (got, want, np.max(np.abs(g - w)))

"""Synthetic stand-ins for the data files the reference's configs load at import time and this repo cannot ship
(SURVEY.md 8f-1): the UCI sets are git-ignored in the reference (.gitignore:132-136) and the Fourier-shape sets need shapely.

    python tools/make_synthetic_data.py <directory> [--rows N]

writes, with the column layouts the reference's loaders expect:
    uci_data/power/data.npy            [N, 8]  float64   (data.py:304-332 deletes columns 3 and 1 -> d = 6)
    uci_data/gas/ethylene_CO.pickle    DataFrame Time, Meth, Eth + 8 sensor columns (data.py:366-392 -> d = 8)
    uci_data/miniboone/data.npy        [N, 43] float64   (data.py:426-431 drops the last column -> d = 42)
    data/{lens-shape1,plus-shape}_{x,y}_{train,test}.npy   (data.py:466-483; x: 20 / 100 Fourier coefficients, y: 2 / 4)
Every set is an 8-component diagonal Gaussian mixture: right shapes and scales, no real-world content."""
import argparse
import os

import numpy as np


def gmm(rng, n, d, k=8):
    means = 3.0 * rng.standard_normal((k, d))
    stds = 0.3 + rng.random((k, d))
    comp = rng.integers(0, k, n)
    return means[comp] + stds[comp] * rng.standard_normal((n, d))


def main(root, rows=20000, seed=0):
    import pandas as pd
    rng = np.random.default_rng(seed)
    for sub in ("uci_data/power", "uci_data/gas", "uci_data/miniboone", "data", "output/samples"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    np.save(os.path.join(root, "uci_data/power/data.npy"), gmm(rng, rows, 8))
    np.save(os.path.join(root, "uci_data/miniboone/data.npy"), gmm(rng, rows, 43))
    cols = {"Time": np.arange(rows, dtype=np.float64), "Meth": rng.random(rows), "Eth": rng.random(rows)}
    sensors = gmm(rng, rows, 8)
    for i in range(8):
        cols[f"S{i}"] = sensors[:, i]
    pd.DataFrame(cols).to_pickle(os.path.join(root, "uci_data/gas/ethylene_CO.pickle"))
    for name, nx, ny in (("lens-shape1", 20, 2), ("plus-shape", 100, 4)):
        for split, n in (("train", rows), ("test", max(rows // 10, 1))):
            x = 0.5 * gmm(rng, n, nx)
            y = x[:, :ny] * 0.3 + 0.1 * rng.standard_normal((n, ny))     # an observation correlated with the parameters
            np.save(os.path.join(root, f"data/{name}_x_{split}.npy"), x.astype(np.float32))
            np.save(os.path.join(root, f"data/{name}_y_{split}.npy"), y.astype(np.float32))
    return root


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("directory")
    ap.add_argument("--rows", type=int, default=20000)
    a = ap.parse_args()
    print(main(a.directory, a.rows))

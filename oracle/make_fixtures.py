#!/usr/bin/env python
"""oracle/make_fixtures.py -- regenerate tests/golden/ from the read-only reference checkout.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference); the outputs are
committed so that nothing at test/bench time reads /root/reference.

Writes
  tests/golden/tiles/<tile>.i16.xz + tiles.json
        the five GEBCO int16 tiles that ship with the reference (GEBCO-Data/**.nc, NetCDF-3 classic,
        read with scipy because netCDF4 is absent), row-flipped exactly as
        code/subset_bathymetry.py:16-17 does, stored as little-endian int16 column-deltas + xz.
        Bounds are the N/S/W/E numbers in each file name (how code/test_gebco.cpp:132-133 was filled).
  tests/golden/golden_metrics.json
        (a) the reference AUTHOR's published MAE/RMSE/Max rows (results/TestingResults1.csv) keyed by
            tile + removal fraction -- the golden vectors of SURVEY.md section 4;
        (b) the same metrics + NaN counts at 10/50/90 % computed HERE with the unmodified reference
            (oracle/_ref/libgridh_ref.so).
  tests/golden/points_<case>.npz
        per-point outputs of the unmodified reference on a deterministic sub-sample of each case:
        query points, three method outputs, and the four selected (i,j) of the floor- and
        round-centred searches.
"""
import csv
import glob
import hashlib
import json
import lzma
import os
import re
import sys

import numpy as np
from scipy.io import netcdf_file

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob  # noqa: E402

REF = os.environ.get("AUVI_REFERENCE", "/root/reference")
OUT = ob.GOLDEN

TILES = {  # short name -> glob under GEBCO-Data
    "mariana": "Mariana Trench/*/gebco_2024_n13.0188_*.nc",
    "east_pacific": "East-Pacific Rise/*/gebco_2024_n12.085_*.nc",
    "mid_atlantic": "Mid-Atlantic Ridge/*/gebco_2024_n1.0071_*.nc",
    "us_east": "GEBCO_28_Feb_2025_5615bda1e072/gebco_2024_n38.2361_*.nc",
    "mini": "GEBCO_14_Mar_2025_9471568999c2/gebco_2024_n39.6304_*.nc",
}


def write_tiles():
    os.makedirs(os.path.join(OUT, "tiles"), exist_ok=True)
    manifest = {}
    for name, pat in TILES.items():
        (fn,) = glob.glob(os.path.join(REF, "GEBCO-Data", pat))
        ds = netcdf_file(fn, "r", mmap=False)
        elev = np.array(ds.variables["elevation"].data, dtype=np.int16)
        elev = elev[::-1].copy()                       # subset_bathymetry.py:16-17
        m = re.search(r"_n(-?[\d.]+)_s(-?[\d.]+)_w(-?[\d.]+)_e(-?[\d.]+)\.nc$", fn)
        n, s, w, e = map(float, m.groups())
        delta = np.diff(elev.astype(np.int32), axis=1, prepend=0)
        assert np.abs(delta).max() < 32768
        blob = lzma.compress(delta.astype("<i2").tobytes(), preset=9)
        out = f"{name}.i16.xz"
        with open(os.path.join(OUT, "tiles", out), "wb") as f:
            f.write(blob)
        manifest[name] = dict(file=out, n_lat=int(elev.shape[0]), n_lon=int(elev.shape[1]),
                              min_lon=w, max_lon=e, min_lat=s, max_lat=n,
                              source=os.path.relpath(fn, REF),
                              sha256_int16=hashlib.sha256(elev.astype("<i2").tobytes()).hexdigest())
        print(f"tile {name}: {elev.shape} -> {len(blob)} bytes")
    with open(os.path.join(OUT, "tiles", "tiles.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def published_rows(manifest):
    """Map results/TestingResults1.csv Grid-B rows onto (tile, fraction) through the batch size."""
    size_to_case = {}
    for name, m in manifest.items():
        total = m["n_lat"] * m["n_lon"]
        for frac in (0.01, 0.05, 0.10, 0.15, 0.20):
            size_to_case[(int(total * frac), frac)] = name
    out = {}
    with open(os.path.join(REF, "results", "TestingResults1.csv")) as f:
        for row in csv.DictReader(f):
            if row["GridType"] != "B":
                continue
            key = (int(row["BatchSize"]), float(row["RemovalFraction"]))
            if key not in size_to_case:
                continue                                # Kerguelen: input tile missing from checkout
            case = f"{size_to_case[key]}@{key[1]:.2f}"
            rec = out.setdefault(case, {}).setdefault(row["InterpolationType"].lower(), {})
            vals = [row["MAE"], row["RMSE"], row["Max Error"]]
            prev = rec.get(row["Machine"])
            assert prev is None or prev == vals, (case, row)
            rec[row["Machine"]] = vals                  # keep the printed 6-significant-digit text
    return out


def computed_rows(manifest):
    out = {}
    for name in ("mariana", "east_pacific", "mid_atlantic"):
        for frac in (0.10, 0.50, 0.90):
            case = ob.masked_case(name, frac)
            ref = ob.Reference(case["z"], *case["bounds"])
            rec = {}
            for meth in (ob.BILINEAR, ob.CUBIC, ob.KRIGING):
                est = ref.batch(meth, case["pts"], threads=8)
                mae, rmse, mx = ref.metrics(case["truth"], est)
                rec[ob.METHOD_NAMES[meth]] = dict(mae=mae, rmse=rmse, max=mx,
                                                  n_nan=int(np.isnan(est).sum()), n=int(est.size))
            out[f"{name}@{frac:.2f}"] = rec
            print(name, frac, {k: (round(v["mae"], 4), round(v["rmse"], 4), v["n_nan"]) for k, v in rec.items()})
    return out


def write_points(manifest):
    rng = np.random.RandomState(7)
    for name, frac, n_keep in (("mid_atlantic", 0.10, 4096), ("mid_atlantic", 0.50, 4096),
                               ("mid_atlantic", 0.90, 4096), ("mariana", 0.50, 8192),
                               ("mini", 0.50, 1890)):
        case = ob.masked_case(name, frac)
        ref = ob.Reference(case["z"], *case["bounds"])
        n = case["pts"].shape[0]
        keep = np.sort(rng.choice(n, size=min(n_keep, n), replace=False))
        pts = case["pts"][keep]
        rec = dict(index=keep.astype(np.int64), pts=pts, truth=case["truth"][keep])
        for meth in (ob.BILINEAR, ob.CUBIC, ob.KRIGING):
            rec[ob.METHOD_NAMES[meth]] = ref.batch(meth, pts)
        for rule, tag in ((0, "floor"), (1, "round")):
            found, sel = ref.select4(rule, pts)
            rec[f"found_{tag}"] = found
            rec[f"sel_{tag}"] = sel
        fn = os.path.join(OUT, f"points_{name}_{int(frac * 100):02d}.npz")
        np.savez_compressed(fn, **rec)
        print("wrote", fn, os.path.getsize(fn))

    # Grid A: the "small" 10x8 synthetic grid (generate_csv_grids.cpp:100) and a 40x32 one with a
    # few holes punched in, on the driver's bounds (test_interpolation.cpp:143-144), full 2x lattice.
    for tag, n_lon, n_lat, holes in (("small", 10, 8, 0), ("holes", 40, 32, 37)):
        z = ob.synth_grid(n_lat, n_lon)
        if holes:
            idx = np.random.RandomState(3).choice(z.size, size=holes, replace=False)
            z.ravel()[idx] = np.nan
        bounds = (-180.0, -160.0, 20.0, 30.0)
        pts, nn_lat, nn_lon = ob.lattice_queries(n_lat, n_lon, *bounds)
        ref = ob.Reference(z, *bounds)
        rec = dict(z=z, pts=pts, nn=np.array([nn_lat, nn_lon]))
        for meth in (ob.BILINEAR, ob.CUBIC, ob.KRIGING):
            rec[ob.METHOD_NAMES[meth]] = ref.batch(meth, pts)
        for rule, t2 in ((0, "floor"), (1, "round")):
            found, sel = ref.select4(rule, pts)
            rec[f"found_{t2}"] = found
            rec[f"sel_{t2}"] = sel
        fn = os.path.join(OUT, f"lattice_{tag}.npz")
        np.savez_compressed(fn, **rec)
        print("wrote", fn, os.path.getsize(fn))


def main():
    ob.build(ref=True)
    manifest = write_tiles()
    gold = dict(published=published_rows(manifest), computed=computed_rows(manifest),
                note="published = reference author's rows (results/TestingResults1.csv, text as printed); "
                     "computed = unmodified GridH.cpp run in the build container, seed-42 masks")
    with open(os.path.join(OUT, "golden_metrics.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    write_points(manifest)


if __name__ == "__main__":
    main()

"""CPU suite, part 4: the host-only half of the Grid-B data preparation (csrc/prep.cpp).

* auvi_legacy_choice == numpy.random.seed(s); numpy.random.choice(total, n, replace=False)
  (the reference's mask tool, code/subset_bathymetry.py:32-39);
* auvi_netcdf3_find / auvi_netcdf3_read_f64 against scipy.io.netcdf_file on files written here (CDF-1 and CDF-2)
  and, when the read-only checkout is present, on the reference's own GEBCO tiles.
No GPU is needed and nothing under oracle/ is involved.
"""
import glob
import os
import sys

import numpy as np
import pytest
from scipy.io import netcdf_file

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python"))

import auvi  # noqa: E402


@pytest.mark.parametrize("total,n,seed", [(1, 1, 42), (10, 3, 42), (10, 10, 0), (1000, 100, 42), (4097, 4000, 7),
                                          (65158 * 2, 65158, 42), (932184, 93218, 42), (300000, 0, 5)])
def test_legacy_choice_matches_numpy(total, n, seed):
    got = auvi.legacy_choice(total, n, seed)
    np.random.seed(seed)
    want = np.random.choice(total, size=n, replace=False)
    assert np.array_equal(got, want)


def test_legacy_choice_rejects_bad_sizes():
    with pytest.raises(auvi.AuviError, match="0 <= n <= total"):
        auvi.legacy_choice(5, 6, 1)


def _write_nc(path, version, elev, lat, lon):
    f = netcdf_file(path, "w", version=version)
    f.createDimension("lat", lat.size)
    f.createDimension("lon", lon.size)
    f.title = "synthetic tile"
    v = f.createVariable("lat", "d", ("lat",)); v[:] = lat; v.units = "degrees_north"
    v = f.createVariable("lon", "d", ("lon",)); v[:] = lon
    v = f.createVariable("elevation", elev.dtype.char, ("lat", "lon")); v[:] = elev; v.long_name = "Elevation"
    f.close()


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("dt", ["i2", "i4", "f4", "f8"])
def test_netcdf3_header_and_axes(tmp_path, version, dt):
    rng = np.random.RandomState(3)
    elev = (rng.randint(-11000, 100, size=(37, 53))).astype(dt)
    lat, lon = np.linspace(-3.0, 2.5, 37), np.linspace(140.0, 144.0, 53)
    path = str(tmp_path / "t.nc")
    _write_nc(path, version, elev, lat, lon)
    img = open(path, "rb").read()
    v = auvi.netcdf3_find(img, "elevation")
    assert (v.ndims, v.shape[0], v.shape[1], v.n_elems) == (2, 37, 53, 37 * 53)
    assert v.nc_type == {"i2": 3, "i4": 4, "f4": 5, "f8": 6}[dt] and v.elem_bytes == np.dtype(dt).itemsize
    raw = np.frombuffer(img, dtype=">" + dt, count=v.n_elems, offset=v.data_offset).reshape(37, 53)
    assert np.array_equal(raw, elev)
    assert np.array_equal(auvi.netcdf3_read_f64(img, "lat"), lat)
    assert np.array_equal(auvi.netcdf3_read_f64(img, "lon"), lon)
    assert np.array_equal(auvi.netcdf3_read_f64(img, "elevation"), elev.astype(np.float64))


def test_netcdf3_errors(tmp_path):
    with pytest.raises(auvi.AuviError, match="not a NetCDF-3"):
        auvi.netcdf3_find(b"\x89HDF\r\n\x1a\n" + b"\0" * 64, "elevation")
    path = str(tmp_path / "t.nc")
    _write_nc(path, 1, np.zeros((2, 2), "i2"), np.zeros(2), np.zeros(2))
    img = open(path, "rb").read()
    with pytest.raises(auvi.AuviError, match="no variable named"):
        auvi.netcdf3_find(img, "depth")
    with pytest.raises(auvi.AuviError):
        auvi.netcdf3_find(img[:40], "elevation")               # truncated header


@pytest.mark.skipif(not os.path.isdir("/root/reference/GEBCO-Data"), reason="reference checkout not present")
def test_netcdf3_on_the_reference_tiles():
    files = glob.glob("/root/reference/GEBCO-Data/**/gebco_2024_n*.nc", recursive=True)
    assert files
    for fn in files:
        img = open(fn, "rb").read()
        nc = netcdf_file(fn, "r", mmap=False)
        v = auvi.netcdf3_find(img, "elevation")
        e = nc.variables["elevation"][:]
        assert (v.shape[0], v.shape[1]) == e.shape and v.nc_type == 3
        assert np.array_equal(np.frombuffer(img, dtype=">i2", count=v.n_elems, offset=v.data_offset).reshape(e.shape), e)
        assert np.array_equal(auvi.netcdf3_read_f64(img, "lat"), nc.variables["lat"][:])


def test_csv_dims_host_only():
    import ctypes as C
    lib = auvi.load()
    def dims(text):
        r, c = C.c_int64(), C.c_int64()
        rc = lib.auvi_csv_dims(text, len(text), C.byref(r), C.byref(c))
        return rc, r.value, c.value
    assert dims(b"1,2,3\n4,5,6\n") == (0, 2, 3)
    assert dims(b"1,2,3\r\n4,5,6") == (0, 2, 3)                 # CRLF, no trailing newline
    assert dims(b"-5559.0\n") == (0, 1, 1)
    assert dims(b"1,nan\n2,3\n\n\n") == (0, 2, 2)               # trailing blank lines are not rows
    rc, _, _ = dims(b"\n\n")
    assert rc != 0 and b"Grid data is empty" in lib.auvi_last_error()


def test_netcdf3_header_counts_are_bounded(tmp_path):
    """The header's counts come from an untrusted file: truncated images and crafted counts (2^32 - 1 dimensions,
    variables or attribute elements; a shape whose product overflows int64) fail cleanly instead of looping or allocating."""
    import struct
    elev = np.arange(6, dtype="i2").reshape(2, 3)
    path = str(tmp_path / "t.nc")
    _write_nc(path, 1, elev, np.array([0.0, 1.0]), np.array([0.0, 1.0, 2.0]))
    img = open(path, "rb").read()
    v = auvi.netcdf3_find(img, "elevation")
    assert v.n_elems == 6
    for cut in range(8, v.data_offset + 12, 7):                          # every truncation inside the header or the data
        with pytest.raises(auvi.AuviError):
            auvi.netcdf3_find(img[:cut], "elevation")
    huge = struct.pack(">I", 0xFFFFFFFF)
    bad_dims = img[:12] + huge + img[16:]                           # dimension count
    with pytest.raises(auvi.AuviError, match="corrupt NetCDF header"):
        auvi.netcdf3_find(bad_dims, "elevation")
    # both dimension lengths 2^31: the element count must not wrap into something that passes the size check
    at = img.index(b"lat\x00") + 4
    big = img[:at] + struct.pack(">I", 0x80000000) + img[at + 4:]
    at = big.index(b"lon\x00") + 4
    big = big[:at] + struct.pack(">I", 0x80000000) + big[at + 4:]
    with pytest.raises(auvi.AuviError):
        auvi.netcdf3_find(big, "elevation")

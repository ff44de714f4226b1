"""Deterministic synthetic Rossby-wave-like PV fields (host/numpy version).

The reference ships no data that survives in this checkout (its only fixture,
``tests/data/demo_data.nc``, is missing), so the parity tests and the benchmark use
this frozen recipe: a zonal PV gradient whose 2-PVU line is
undulated by a tilted, amplitude-modulated wave (the tilt makes the contour overturn
and fold into filaments) and a faster short wave, plus drifting Gaussian PV anomalies that cut off closed
contours.  Both hemispheres are filled (PV is negative in the south).

``pv_field`` is the host mirror of the device generator ``wbk_synth_pv`` (same
formula; the device version is used by ``bench.py`` to fill HBM directly and its
output is copied back for the CPU baseline, so both legs see identical bytes).
"""

import numpy as np

SEED = 20260101
# frozen recipe (round 1): tuned so that the 1-degree, level-2 event rates resemble the reference's demo pins
# (tests/test_wavebreaking.py:207,224,249: 13 streamers / 3 overturnings / 7 cutoffs per step):
# ~11.8 streamers, 6 overturnings, 5.3 cutoffs per step at 181 x 360.
PARAMS = dict(A=12.0, k=7.0, tilt=2.6, c=0.5, env0=0.6, env1=0.4, env_speed=1.5,
              n_blob=16, blob_amp=3.0, sh_shift=37.0, A2=4.5, k2=19.0, c2=1.1, tilt2=2.0)


def grid_coords(nlat, nlon):
    """lat ascending -90..90 (nlat points), lon 0..360 (exclusive)."""
    lat = np.linspace(-90.0, 90.0, nlat)
    lon = np.arange(nlon) * (360.0 / nlon)
    return lat, lon


def time_axis(ntime, step_hours=1):
    t0 = np.datetime64("2000-01-01T00", "ns")
    return t0 + (np.arange(ntime) * step_hours * 3600 * 10**9).astype("timedelta64[ns]")


def blob_table(seed=SEED, n_blob=PARAMS["n_blob"]):
    """Blob parameters per hemisphere: lat0, lon0, radius (degrees); frozen by the seed."""
    rng = np.random.default_rng(seed)
    tab = np.empty((2, n_blob, 3))
    for h in range(2):
        tab[h, :, 0] = rng.uniform(25.0, 65.0, n_blob)
        tab[h, :, 1] = rng.uniform(0.0, 360.0, n_blob)
        tab[h, :, 2] = rng.uniform(2.5, 5.0, n_blob)
    return tab


def _background(alat, lam_deg, hours, p):
    """|PV| of one hemisphere; alat = |lat| in degrees, lam_deg = longitude in degrees."""
    lam = np.radians(lam_deg)
    env = p["env0"] + p["env1"] * np.cos(lam - np.radians(p["env_speed"] * hours))
    phase = p["k"] * (lam - np.radians(p["c"] * hours)) + p["tilt"] * (alat - 45.0) / 10.0 \
        + 0.7 * np.sin(2.0 * lam)
    phase2 = p["k2"] * (lam - np.radians(p["c2"] * hours)) + p["tilt2"] * (alat - 45.0) / 10.0
    phi = alat - p["A"] * env * np.sin(phase) - p["A2"] * np.sin(phase2)
    s = np.sin(np.radians(phi)) / np.sin(np.radians(45.0))
    return 2.0 * np.sign(s) * np.abs(s) ** 3


def pv_field(nlat, nlon, hours, seed=SEED, dtype=np.float32, params=None):
    """Synthetic PV at the given hours since 2000-01-01T00; returns [len(hours), nlat, nlon]."""
    p = dict(PARAMS)
    if params:
        p.update(params)
    lat, lon = grid_coords(nlat, nlon)
    blobs = blob_table(seed, p["n_blob"])
    hours = np.atleast_1d(np.asarray(hours, dtype=np.float64))
    out = np.empty((len(hours), nlat, nlon), dtype=dtype)
    LAT, LON = np.meshgrid(lat, lon, indexing="ij")
    alat = np.abs(LAT)
    south = LAT < 0
    for ti, h in enumerate(hours):
        lam = np.where(south, LON + p["sh_shift"], LON)
        pv = _background(alat, lam, h, p)
        for hemi in range(2):
            mask = south if hemi == 1 else ~south
            for lat0, lon0, rad in blobs[hemi]:
                lonc = (lon0 + p["c"] * h) % 360.0
                bg = _background(np.float64(lat0), np.float64(lonc + (p["sh_shift"] if hemi else 0.0)), h, p)
                sign = -1.0 if bg > 2.0 else 1.0
                dlon = (LON - lonc + 180.0) % 360.0 - 180.0
                d2 = (alat - lat0) ** 2 + (dlon * np.cos(np.radians(lat0))) ** 2
                pv = pv + np.where(mask, sign * p["blob_amp"] * np.exp(-d2 / (2.0 * rad * rad)), 0.0)
        out[ti] = np.where(south, -pv, pv).astype(dtype)
    return out

#!/usr/bin/env python3
"""Generate tests/golden/geo_golden.npz with oracle/geo_port.c (the C restatement of the reference's WGS84 <-> ENU
transforms, /root/reference/uavPathPlanning.cpp:894-1108, pinned bit for bit on the reference's recorded run in
/root/reference/readme.md:11-28 -- tests/test_geo_oracle.py).

    make -C oracle && python tests/golden/make_geo_golden.py

Per reference point r (rows of `refs`, [lon_deg, lat_deg, alt_m]):
    enu[r]       n ENU rows                                   (inputs)
    lla[r]       enuToWGS84_Batch(enu[r], refs[r])            (cpp:1098-1108)
    steps[r]     fixed-point steps ecefToWGS84 took per point (cpp:939-949)
    enu_back[r]  wgs84ToENU_Batch(lla[r], refs[r])            (cpp:1085-1095)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import geo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "geo_golden.npz")
N = 256

REFS = np.array([
    geo.README_ORIGIN,                   # the uav31_0 case (Inner Mongolia, 40.9 N)
    [0.0, 0.0, 0.0],                     # equator / prime meridian
    [-70.5, -33.4, 520.0],               # southern and western hemisphere
    [25.0, 78.0, 0.0],                   # high latitude
    [179.9999, 10.0, 0.0],               # 11 m west of the antimeridian: eastward points wrap to negative longitude
    [-122.3, 47.6, -30.0],               # negative reference altitude
])


def inputs(r: int) -> np.ndarray:
    rng = np.random.default_rng(4100 + r)
    e = np.empty((N, 3))
    e[:, :2] = rng.normal(0.0, 30e3, (N, 2))       # a mission area of a few tens of km
    e[:, 2] = rng.uniform(-500.0, 12000.0, N)
    e[0] = 0.0                                     # the origin itself
    e[1] = [1e-9, -1e-9, 0.0]
    e[2:10, :2] = rng.normal(0.0, 1.0e6, (8, 2))   # 1 000 km away: the tangent plane is far off the ellipsoid
    e[10] = [0.0, 0.0, 4.0e5]                      # orbital altitude
    e[11] = [12.5, -7.25, -6.0e3]                  # below the ellipsoid
    return e


def main():
    enu = np.stack([inputs(r) for r in range(len(REFS))])
    lla = np.empty_like(enu)
    back = np.empty_like(enu)
    steps = np.zeros(enu.shape[:2], dtype=np.int32)
    for r, ref in enumerate(REFS):
        lla[r], steps[r] = geo.enu_to_wgs84_batch(enu[r], ref, return_steps=True)
        back[r] = geo.wgs84_to_enu_batch(lla[r], ref)
    np.savez_compressed(OUT, refs=REFS, enu=enu, lla=lla, steps=steps, enu_back=back)
    print(OUT, enu.shape, "steps", np.bincount(steps.ravel()), "round trip max |d| m", np.abs(back - enu).max())


if __name__ == "__main__":
    main()

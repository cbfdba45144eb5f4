"""CPU tests of the WGS84 <-> ENU oracle (oracle/geo_port.c): pinned on the reference's own recorded run
(/root/reference/readme.md:11-28, copied into oracle/geo.py as README_*), plus the committed fixture and the analytic
properties of the maps."""
import os

import numpy as np
import pytest

from oracle import geo

HERE = os.path.dirname(os.path.abspath(__file__))


def fmt15(a):
    return [["%.15f" % v for v in row] for row in np.asarray(a)]


def test_port_reproduces_readme_enu_digits():
    """readme.md:14-20: Enu_waypoint printed with 15 decimals.  Rows 1..6 must print identically; row 0 is the origin
    itself, whose east/north are 5e-11 / 5e-10 m of cancellation noise (same digits here)."""
    enu = geo.wgs84_to_enu_batch(geo.README_WGS84, geo.README_ORIGIN)
    assert fmt15(enu) == fmt15(geo.README_ENU)


def test_port_reproduces_readme_wgs84_digits():
    """readme.md:22-28: WGS84Point = enuToWGS84_Batch(Enu_waypoint, origin) printed with 15 decimals."""
    enu = geo.wgs84_to_enu_batch(geo.README_WGS84, geo.README_ORIGIN)
    lla, steps = geo.enu_to_wgs84_batch(enu, geo.README_ORIGIN, return_steps=True)
    assert fmt15(lla) == fmt15(geo.README_WGS84_BACK)
    assert np.all((steps >= 1) & (steps <= 10))
    # and from the printed ENU values as well
    assert fmt15(geo.enu_to_wgs84_batch(geo.README_ENU, geo.README_ORIGIN)) == fmt15(geo.README_WGS84_BACK)


def test_fixture_is_what_the_port_produces():
    z = np.load(os.path.join(HERE, "golden", "geo_golden.npz"))
    for r, ref in enumerate(z["refs"]):
        lla, steps = geo.enu_to_wgs84_batch(z["enu"][r], ref, return_steps=True)
        assert np.array_equal(lla, z["lla"][r]) and np.array_equal(steps, z["steps"][r])
        assert np.array_equal(geo.wgs84_to_enu_batch(z["lla"][r], ref), z["enu_back"][r])


def test_round_trip_and_thread_independence():
    rng = np.random.default_rng(7)
    ref = np.array([109.56, 40.867, 0.0])
    enu = np.column_stack([rng.normal(0, 2e4, 5000), rng.normal(0, 2e4, 5000), rng.uniform(0, 5000, 5000)])
    lla = geo.enu_to_wgs84_batch(enu, ref)
    back = geo.wgs84_to_enu_batch(lla, ref)
    assert np.abs(back - enu).max() < 1e-7      # the 1e-12 rad stopping rule leaves < 1e-7 m
    assert np.array_equal(geo.enu_to_wgs84_batch(enu, ref, threads=4), lla)
    assert np.array_equal(geo.wgs84_to_enu_batch(lla, ref, threads=4), back)


def test_known_ecef_values():
    a, e2 = 6378137.0, 0.006694379990141
    assert np.allclose(geo.wgs84_to_ecef([0.0, 0.0, 0.0]), [a, 0.0, 0.0], atol=1e-9)
    assert np.allclose(geo.wgs84_to_ecef([90.0, 0.0, 100.0]), [0.0, a + 100.0, 0.0], atol=1e-6)
    b = a * np.sqrt(1 - e2)
    assert np.allclose(geo.wgs84_to_ecef([0.0, 90.0, 0.0])[2], b, atol=1e-6)
    # exactly on the axis (p = 0 < 1e-12): the altitude comes from the pole branch (cpp:956-957), but the fixed-point
    # step evaluates p * (1 - e2 N / (N + alt)) = 0 * -inf, so the reference's latitude is NaN there -- reproduced
    lla = geo.ecef_to_wgs84([0.0, 0.0, b + 25.0])
    assert np.isnan(lla[1]) and abs(lla[2] - 25.0) < 1e-6


def test_up_axis_is_the_ellipsoid_normal():
    ref = np.array([12.0, 55.0, 40.0])
    lla = geo.enu_to_wgs84_batch(np.array([[0.0, 0.0, 1000.0]]), ref)[0]
    assert abs(lla[0] - 12.0) < 1e-10 and abs(lla[1] - 55.0) < 1e-10 and abs(lla[2] - 1040.0) < 1e-6


@pytest.mark.parametrize("bad", [np.zeros((3,)), np.zeros((2, 2))])
def test_shape_errors(bad):
    with pytest.raises(ValueError):
        geo.wgs84_to_enu_batch(bad, [0, 0, 0])


def test_device_atan2_polynomial_on_the_host(tmp_path):
    """geo_atan2 (msnap_geo.cuh) restated on the host with the same generated coefficients and the same fma sequence
    (tests/cpp/geo_atan_host.cpp): at most ~1 ulp of pi away from libm's atan2l over 2 M directions and 12 decades."""
    import subprocess

    root = os.path.dirname(HERE)
    exe = str(tmp_path / "geo_atan_host")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-mfma", "-I", os.path.join(root, "cs_pathplan_b200", "csrc"),
                           os.path.join(HERE, "cpp", "geo_atan_host.cpp"), "-o", exe])
    worst = float(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert worst <= 6.0          # units of 2^-53 rad; ulp(pi) = 4 units


def test_generated_atan_header_is_current():
    import subprocess
    import sys

    root = os.path.dirname(HERE)
    out = subprocess.run([sys.executable, os.path.join(root, "cs_pathplan_b200", "csrc", "gen_geo_atan.py")],
                         capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(root, "cs_pathplan_b200", "csrc", "msnap_geo_atan.h")).read()


def test_device_sincos_on_the_host(tmp_path):
    """geo_sincos (msnap_geo.cuh) restated on the host with the same generated constants and fma sequence
    (tests/cpp/geo_sincos_host.cpp): within 2 units of 2^-53 of libm's sinl / cosl on [-7, 7] and up to |x| = 1e5."""
    import subprocess

    root = os.path.dirname(HERE)
    exe = str(tmp_path / "geo_sincos_host")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-mfma", "-I", os.path.join(root, "cs_pathplan_b200", "csrc"),
                           os.path.join(HERE, "cpp", "geo_sincos_host.cpp"), "-o", exe])
    worst = float(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert worst <= 2.5


def test_generated_sincos_header_is_current():
    import subprocess
    import sys

    root = os.path.dirname(HERE)
    out = subprocess.run([sys.executable, os.path.join(root, "cs_pathplan_b200", "csrc", "gen_geo_sincos.py")],
                         capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(root, "cs_pathplan_b200", "csrc", "msnap_geo_sincos.h")).read()

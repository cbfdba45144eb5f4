"""CPU tests: the oracle (oracle/msnap_oracle.py, the line-by-line port) and Oracle B (structured, multi-precision)
against every golden vector generated from the reference's own source, plus analytic known answers."""
import math

import numpy as np
import pytest

from helpers import COEFF_TOL, SAMPLE_TOL, decisions_equivalent, load_golden, oracle_cfg, scaled_coeff_err
from oracle import msnap_oracle as mo
from oracle import msnap_structured as st

CASES = load_golden()
SMALL = [c for c in CASES if c.ns <= 64]


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c.name)
def test_port_matches_reference_golden(case):
    """The port reproduces the reference's samples, coefficients and loop decisions on every fixture."""
    samples, info = mo.generate_trajectory_matrix(case.path, oracle_cfg(case.cfg), case.sdo, case.vo)
    assert np.array_equal(info.Time, case.time)                       # bit-exact time allocation
    assert info.iters == case.iters and info.vw_final == case.vw_final
    noise = case.ref_noise()
    assert scaled_coeff_err(info.PolyCoeff.reshape(case.coeff.shape), case.coeff, case.time) <= COEFF_TOL + 4 * noise
    assert samples.shape == case.samples.shape
    assert np.max(np.abs(samples - case.samples)) <= SAMPLE_TOL
    assert abs(info.max_dev - case.max_dev) <= 1e-8 + 4 * noise


@pytest.mark.parametrize("case", [c for c in CASES if c.ns <= 16], ids=lambda c: c.name)
def test_structured_form_is_the_same_problem(case):
    """Oracle B in plain double agrees with the reference to the reference's own noise level, decisions included."""
    cfg = case.cfg
    Vel = np.array([cfg.get("start_vel", (0, 0, 0)), cfg.get("end_vel", (0, 0, 0))], dtype=float)
    Acc = np.array([cfg.get("start_acc", (0, 0, 0)), cfg.get("end_acc", (0, 0, 0))], dtype=float)
    out = st.reweighted_structured(cfg["order"], case.path, Vel, Acc, case.time, cfg["path_weight"],
                                   cfg["vel_zero_weight"])
    assert out["iters"] == case.iters
    if cfg["path_weight"] > 0:
        # the arg-max decisions agree, or the competing deviations tie to 1e-9 relative (e.g. the mirror-symmetric
        # single segment with zero boundary derivatives, where s and 16-s are the same deviation)
        assert decisions_equivalent(out["best_s"], case.best_s, np.array(out["dist2"], dtype=float))
    c = np.array(out["coeff"], dtype=float)
    assert scaled_coeff_err(c, case.truth_coeff, case.time) <= 1e-9
    assert scaled_coeff_err(c, case.coeff, case.time) <= COEFF_TOL + 4 * case.ref_noise()


def test_golden_covers_config1():
    names = {c.name for c in CASES}
    assert {"uav31_0_v30", "uav31_0_v200"} <= names
    c = next(c for c in CASES if c.name == "uav31_0_v30")
    assert c.samples.shape == (168, 3) and c.iters == 0
    # first sample is the first waypoint, last sample the last waypoint (ms.cpp:132-137, 157-160)
    assert np.allclose(c.samples[0], c.path[0], atol=1e-8) and np.allclose(c.samples[-1], c.path[-1], atol=1e-8)


def test_q_integer_table():
    """the snap Hessian blocks are these integers times T^e (ms.cpp:321-324)."""
    expect = {2: {12, 6, 4}, 3: {720, 360, 120, 192, 72, 36},
              4: {100800, 50400, 20160, 5040, 25920, 10800, 2880, 4800, 1440, 576}}
    for o, vals in expect.items():
        p_order, m = 2 * o - 1, 2 * o
        got = set()
        for i in range(m):
            for l in range(m):
                if m - i <= o or m - l <= o:
                    continue
                ai = mo.factorial(p_order - i) // mo.factorial(p_order - o - i)
                al = mo.factorial(p_order - l) // mo.factorial(p_order - o - l)
                e = p_order - i + p_order - l - (2 * o - 1)
                assert (ai * al) % e == 0          # the reference's integer division is exact for o <= 5
                got.add(ai * al // e)
        assert got == vals


def test_cubic_hermite_closed_form():
    """order 2, one segment, no penalties: the cubic Hermite polynomial."""
    p0, p1, v0, v1, T = 1.5, -2.0, 0.7, -0.3, 2.5
    Path = np.array([[p0, 0, 0], [p1, 0, 0]])
    Vel = np.array([[v0, 0, 0], [v1, 0, 0]])
    c, _ = mo.solve_qp_closed_form(2, Path, Vel, np.zeros((2, 3)), np.array([T]))
    want = [(2 * (p0 - p1) + T * (v0 + v1)) / T ** 3, (3 * (p1 - p0) - T * (2 * v0 + v1)) / T ** 2, v0, p0]
    assert np.allclose(c[0, :4], want, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("order,poly", [(3, [6, -15, 10, 0, 0, 0]), (4, [-20, 70, -84, 35, 0, 0, 0, 0])])
def test_smoothstep_single_segment(order, poly):
    """zero boundary derivatives, one segment: p(t) = p0 + (p1 - p0) h(t/T) with the degree 2o-1 smoothstep."""
    T, p0, p1 = 3.0, 2.0, 11.0
    Path = np.array([[p0, 0, 0], [p1, 0, 0]])
    c, _ = mo.solve_qp_closed_form(order, Path, np.zeros((2, 3)), np.zeros((2, 3)), np.array([T]))
    m = 2 * order
    want = np.array(poly, dtype=float) * (p1 - p0) / T ** np.arange(m - 1, -1, -1)
    want[-1] += p0
    assert np.allclose(c[0, :m], want, rtol=1e-9, atol=1e-9)


def test_invariants_interpolation_and_continuity():
    """Waypoint interpolation and C^(o-1) continuity of the port's output (SURVEY section 8c invariants)."""
    case = next(c for c in CASES if c.name == "rw_o4_ns8_shipped")
    o, m = 4, 8
    co = case.coeff
    for k in range(case.ns):
        for a in range(3):
            c = co[k, a]
            assert abs(np.polyval(c, 0.0) - case.path[k, a]) < 1e-7
            assert abs(np.polyval(c, case.time[k]) - case.path[k + 1, a]) < 1e-7
            if k + 1 < case.ns:
                for r in range(1, o):
                    left = np.polyval(np.polyder(c, r), case.time[k])
                    right = np.polyval(np.polyder(co[k + 1, a], r), 0.0)
                    assert abs(left - right) < 1e-6 * max(1.0, abs(left))


def test_empty_and_short_inputs():
    s, _ = mo.generate_trajectory_matrix(np.zeros((1, 3)), mo.MinimumSnapConfig())
    assert s.shape == (0, 0)                                            # ms.cpp:54-57
    s, _ = mo.generate_trajectory_matrix(np.zeros((4, 2)), mo.MinimumSnapConfig())
    assert s.shape == (0, 0)

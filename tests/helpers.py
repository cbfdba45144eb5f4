"""Shared test utilities: golden-vector access and the parity metrics (hard parts 1-2)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "msnap_golden.npz")

COEFF_TOL = 1e-8     # north_star: coefficients within 1e-8 relative (position-scaled, see scaled_coeff_err)
SAMPLE_TOL = 1e-6    # north_star: every sampled ENU position within 1e-6 m
CFG_KEYS = ("order", "path_weight", "vel_zero_weight", "V_avg", "min_time_s", "sample_distance",
            "start_vel", "end_vel", "start_acc", "end_acc")


class Case:
    def __init__(self, meta, z):
        self.name = meta["name"]
        self.cfg = meta["cfg"]
        self.sdo = meta["sample_distance_override"]
        self.vo = meta["v_avg_override"]
        self.max_dev = meta["max_dev"]
        self.iters = meta["iters"]
        self.vw_final = meta["vw_final"]
        self.truth_max_dev = meta["truth_max_dev"]
        for k in ("path", "time", "coeff", "samples", "best_s", "truth_coeff"):
            setattr(self, k, z[f"{self.name}/{k}"])

    @property
    def ns(self):
        return self.time.shape[0]

    def ref_noise(self):
        """The reference's own rounding error on this case, in the parity metric."""
        return scaled_coeff_err(self.coeff, self.truth_coeff, self.time)


def load_golden():
    z = np.load(GOLDEN)
    manifest = json.loads(bytes(z["manifest"]).decode())
    return [Case(m, z) for m in manifest]


def scaled_coeff_err(c, c_ref, T):
    """max over (segment, axis, power) of |dc_i| T^deg_i / max_i(|c_ref_i| T^deg_i): the error of a coefficient
    relative to what that segment's coefficients can contribute to position.  Plain element-wise relative error is
    not attainable even by the reference against exact arithmetic for the tiny high-order coefficients."""
    c = np.asarray(c, dtype=np.float64)
    c_ref = np.asarray(c_ref, dtype=np.float64).reshape(c.shape)
    m = c.shape[2]
    pw = np.asarray(T, dtype=np.float64)[:, None, None] ** np.arange(m - 1, -1, -1)[None, None, :]
    den = np.max(np.abs(c_ref) * pw, axis=2, keepdims=True)
    den[den == 0] = 1.0
    return float(np.max(np.abs(c - c_ref) * pw / den))


def make_cfg(cls, d):
    return cls(**{k: d[k] for k in CFG_KEYS if k in d})


def oracle_cfg(d):
    from oracle import msnap_oracle as mo

    return make_cfg(mo.MinimumSnapConfig, d)


def decisions_equivalent(best_s, best_s_ref, dist2_table, rel=1e-9):
    """True if the per-segment arg-max indices agree, or differ only between samples whose squared deviations tie
    within `rel` (hard part 2: a 1-ulp change may legitimately flip a tie)."""
    best_s, best_s_ref = np.asarray(best_s), np.asarray(best_s_ref)
    for k in np.nonzero(best_s != best_s_ref)[0]:
        a, b = dist2_table[k][best_s[k]], dist2_table[k][best_s_ref[k]]
        if abs(a - b) > rel * max(abs(a), abs(b), 1e-300):
            return False
    return True

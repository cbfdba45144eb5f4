"""Algorithmic byte and flop model of the batched minimum-snap path.

One file holds the model so that bench.py, DESIGN.md and the profiles agree on the numerators:

bytes  -- what any implementation must move through HBM per batch: waypoints in; segment times, coefficients,
          samples, sample offsets and the three per-trajectory scalars out.  Workspace traffic is NOT algorithmic.
flops  -- fp64 operations of the O(ns) structured algorithm the kernels implement (DFMA = 2, DADD/DMUL/DDIV/
          DSQRT/compare = 1), counted from the device code in cs_pathplan_b200/csrc/msnap_device.cuh.  The sampler is
          counted ONCE per candidate (the count pass + write pass of the generic path evaluates twice; that is
          overhead, not algorithm).
"""
from __future__ import annotations


def algorithmic_bytes(B: int, n_seg: int, order: int, total_samples: int) -> int:
    wp_in = 24 * (n_seg + B)
    times_out = 8 * n_seg
    coeff_out = 48 * order * n_seg            # 3 axes x 2*order doubles
    samples_out = 24 * total_samples
    per_traj = (8 + 4 + 8) * B + 8 * (B + 1)  # max_dev, iters, vw_final, sample_offset
    return wp_in + times_out + coeff_out + samples_out + per_traj


def kernel_bytes(kernel: str, B: int, n_seg: int, order: int, total_samples: int) -> int:
    """Algorithmic bytes of ONE kernel of the uniform-batch pipeline per launch: what that kernel must read and write
    whatever its implementation.  The coefficients and segment times cross HBM twice in the two-kernel pipeline (written
    by the solve kernel, read by the sampler), which is why the per-kernel figures add up to more than
    ``algorithmic_bytes`` (the bytes of the whole path)."""
    wp_in = 24 * (n_seg + B)
    times = 8 * n_seg
    coeff = 48 * order * n_seg
    if kernel == "k_fused_solve":
        return wp_in + times + coeff + (8 + 4 + 8 + 4) * B       # + max_dev, iters, vw_final, flags
    if kernel == "k_sample_scan":
        return coeff + times + 24 * total_samples + 8 * (B + 1)  # + sample_offset
    raise KeyError(kernel)


def kernel_flops(kernel: str, B: int, n_seg: int, order: int, use_pw: bool, total_solves: int,
                 total_candidates: int) -> float:
    """Share of ``algorithmic_flops`` that belongs to one kernel of the uniform-batch pipeline."""
    sampler = total_candidates * sample_candidate_flops(order)
    if kernel == "k_sample_scan":
        return sampler
    if kernel == "k_fused_solve":
        return algorithmic_flops(B, n_seg, order, use_pw, total_solves, total_candidates) - sampler
    raise KeyError(kernel)


def thomas_row_flops(order: int) -> float:
    """One block row of the block-tridiagonal Cholesky for three right-hand sides: forward + backward."""
    b = order - 1
    chol_fma = (b - 1) * b * (b + 1) / 6
    fwd_fma = b * b * (b + 1) / 2 + 3 * b * b + chol_fma + 3 * b * (b - 1) / 2 + b * b * (b - 1) / 2
    fwd_other = b + b * (b - 1) / 2 + 3 * b + b * b + 1
    back_fma = 3 * b * b + 2 * 3 * b * (b - 1) / 2
    back_other = 2 * 3 * b + 3 * b
    return 2 * (fwd_fma + back_fma) + fwd_other + back_other


def deviation_flops(order: int) -> float:
    return 2 * (6 * order + 6) + 6


def assemble_row_flops(order: int, use_pw: bool) -> float:
    o, b = order, order - 1
    f = 2 * (3 * o) + 3 * b * (b + 1) / 2 + b * b + 3 * b * 4
    if use_pw:
        f += 4 * o + 30 + 4 * b * (b + 1) / 2 + 2 * b * b + 3 * b * 4
    return f


def search_flops(order: int) -> float:
    return 17 * 3 * (4 * order + 7)


def coeff_flops(order: int) -> float:
    return 3 * (4 * order + 8 * order * order)


def sample_candidate_flops(order: int) -> float:
    return 3 * 2 * (2 * order - 1) + 10


def algorithmic_flops(B: int, n_seg: int, order: int, use_pw: bool, total_solves: int, total_candidates: int) -> float:
    """total_solves = sum over trajectories of (1 + reweighting iterations); pass 1 is added when use_pw."""
    n_rows = n_seg - B
    rows_per_traj = n_rows / max(B, 1)
    segs_per_traj = n_seg / max(B, 1)
    f = 0.0
    if use_pw:  # pass 1: assemble + one solve + the 17-sample search
        f += n_rows * (assemble_row_flops(order, False) + thomas_row_flops(order)) + n_seg * search_flops(order)
    f += n_rows * assemble_row_flops(order, use_pw)
    per_solve = rows_per_traj * thomas_row_flops(order) + (segs_per_traj * deviation_flops(order) if use_pw else 0.0)
    f += total_solves * per_solve
    f += n_seg * coeff_flops(order)
    f += total_candidates * sample_candidate_flops(order)
    return f

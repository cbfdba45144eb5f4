"""The C++ drop-in headers include/minimum_snap_gpu.hpp, include/bezier_gpu.hpp and include/geo_transform_gpu.hpp: compiles against an Eigen API (the oracle shim here), links
with the C-ABI library, refuses to run without a GPU (CPU test) and reproduces the golden vectors on one (GPU test)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_demo")


def build_demo():
    lib_dir = os.path.join(ROOT, "cs_pathplan_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle", "shim"),
           os.path.join(ROOT, "tests", "cpp", "adapter_demo.cpp"), "-o", EXE, "-L", lib_dir, "-lmsnap_b200",
           f"-Wl,-rpath,{lib_dir}"]
    subprocess.check_call(cmd)
    return EXE


def test_adapter_compiles_links_and_fails_loudly_without_gpu():
    import torch

    exe = build_demo()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; see the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "msnap_create failed" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("speed,name", [(30.0, "uav31_0_v30"), (200.0, "uav31_0_v200")])
def test_adapter_reproduces_golden(golden, speed, name):
    exe = build_demo()
    r = subprocess.run([exe, str(speed)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    n = int(lines[0].split()[1])
    samples = np.array([[float(v) for v in ln.split()] for ln in lines[1:1 + n]])
    case = next(c for c in golden if c.name == name)
    assert samples.shape == case.samples.shape and np.max(np.abs(samples - case.samples)) <= 1e-6
    head = lines[1 + n].split()
    assert head[0] == "coeff" and (int(head[1]), int(head[2])) == (6, 12)
    coeff = np.array([[float(v) for v in ln.split()] for ln in lines[2 + n:2 + n + 6]])
    # the same bare solve through the oracle port
    from oracle import msnap_oracle as mo

    c_ref, info = mo.solve_qp_closed_form(2, case.path, np.zeros((2, 3)), np.zeros((2, 3)),
                                          np.array([110.0, 15.0, 35.0, 70.0, 15.0, 1.0]), 1e-7, 0.01)
    from helpers import scaled_coeff_err

    T = np.array([110.0, 15.0, 35.0, 70.0, 15.0, 1.0])
    assert scaled_coeff_err(coeff.reshape(6, 3, 4), c_ref.reshape(6, 3, 4), T) <= 1e-8
    assert abs(float(head[4]) - info.max_dev) <= 1e-8
    assert lines[2 + n + 6].split() == ["short", "0"]
    # the planner's coordinate transforms through include/geo_transform_gpu.hpp against the reference's recorded run
    from oracle import geo

    assert lines[3 + n + 6].split() == ["geo", "7"]
    rows = np.array([[float(v) for v in ln.split()] for ln in lines[4 + n + 6:4 + n + 6 + 7]])
    assert np.abs(rows[:, :3] - geo.README_ENU).max() <= 1e-7
    assert np.abs(rows[:, 3:5] - geo.README_WGS84_BACK[:, :2]).max() <= 1e-12
    assert np.abs(rows[:, 5] - geo.README_WGS84_BACK[:, 2]).max() <= 1e-6
    # the Bezier drop-in (include/bezier_gpu.hpp) against the compiled reference's golden rows
    k = 4 + n + 6 + 7
    assert lines[k].split()[0] == "bezier"
    nb = int(lines[k].split()[1])
    bz = np.array([[float(v) for v in ln.split()] for ln in lines[k + 1:k + 1 + nb]])
    z = np.load(os.path.join(ROOT, "tests", "golden", "bezier_golden.npz"))
    want = z["uav31_0_d300/rows"]
    assert bz.shape == want.shape and np.abs(bz - want).max() <= 1e-9
    assert lines[k + 1 + nb].split() == ["bezier_short", "0", "3"]

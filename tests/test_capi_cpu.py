"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/msnap.h declares, refuses
to pretend there is a GPU, and its host-only helpers (config defaults, YAML subset) behave like the reference's."""
import ctypes as C
import os
import re

import pytest

from cs_pathplan_b200 import MinimumSnapConfig, _lib, load_minimum_snap_config, shipped_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "msnap.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(msnap_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libmsnap_b200.so does not export {n}"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signatures and include/msnap.h disagree"
    assert L.msnap_version() == 100


def test_no_device_no_handle():
    """Without a GPU the library must fail loudly -- there is no CPU fallback behind the C ABI."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    h = C.c_void_p()
    assert L.msnap_create(0, C.byref(h)) == _lib.ERR_NO_DEVICE
    assert not h.value
    from cs_pathplan_b200 import TrajectoryGeneratorTool

    with pytest.raises(_lib.MsnapError):
        TrajectoryGeneratorTool(0)


def test_config_defaults_match_reference_struct():
    L = _lib.lib()
    c = _lib.msnap_config()
    L.msnap_config_default(C.byref(c))
    d = MinimumSnapConfig.from_c(c)
    assert d == MinimumSnapConfig()                     # minimum_snap.hpp:11-32
    assert (d.order, d.V_avg, d.min_time_s, d.sample_distance) == (3, 5.0, 0.1, 1.0)


SHIPPED_YAML = """# minimum_snap_config.yaml
# comment line
order: 2
#0.01
vel_zero_weight: 0.01
path_weight: 0.0000001
V_avg: 200.0

min_time_s: 1.0
sample_distance: 300.0   # trailing comment
start_vel: [0.0, 0.0, 0.0]
end_vel: [0.0, 0.0, 0.0]
start_acc: [0.0, 0.0, 0.0]
end_acc: [0.0, 0.0, 0.0]
"""


def test_yaml_flat_shipped_values(tmp_path):
    p = tmp_path / "ms.yaml"
    p.write_text(SHIPPED_YAML)
    assert load_minimum_snap_config(str(p)) == shipped_config()


def test_yaml_wrapper_partial_and_malformed(tmp_path):
    p = tmp_path / "wrapped.yaml"
    p.write_text(
        "other: 1\n"
        "minimum_snap:\n"
        "  order: 4\n"
        "  V_avg: fast\n"                 # malformed scalar -> keeps the previous value (yamlAssignIfPresent)
        "  path_weight: 0.25\n"
        "  start_vel: [1.0, 2.0]\n"        # too short -> ignored (yamlAssignVec3IfPresent)
        "  end_acc:\n"
        "    - 0.5\n"
        "    - 0.25\n"
        "    - -1\n"
        "trailing:\n"
        "  order: 9\n"
    )
    c = load_minimum_snap_config(str(p))
    assert c.order == 4 and c.path_weight == 0.25
    assert c.V_avg == 5.0 and tuple(c.start_vel) == (0.0, 0.0, 0.0)
    assert tuple(c.end_acc) == (0.5, 0.25, -1.0)
    # order "2.0" is not an int for yaml-cpp's as<int>() either: ignored
    q = tmp_path / "o.yaml"
    q.write_text("order: 2.0\nmin_time_s: 3\n")
    c2 = load_minimum_snap_config(str(q))
    assert c2.order == 3 and c2.min_time_s == 3.0


def test_yaml_missing_file():
    with pytest.raises(_lib.MsnapError) as e:
        load_minimum_snap_config("/nonexistent/ms.yaml")
    assert e.value.status == _lib.ERR_IO


def test_product_does_not_import_the_oracle():
    """The product path must not route through oracle/ (or any CPU restatement)."""
    pkg = os.path.join(ROOT, "cs_pathplan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "msnap_oracle" not in text and "msnap_ref" not in text, f

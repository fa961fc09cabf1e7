"""CPU: the C-ABI library builds, loads and exports every symbol include/reface_b200.h declares."""
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "reface_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rfb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_all_symbols():
    from reface_b200 import build, runtime
    path = build.build()
    assert os.path.exists(path)
    lib = runtime.load_library(path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/reface_b200.h but not exported"
    assert sorted(runtime.SIGNATURES) == names, "runtime.SIGNATURES must mirror the header one to one"


def test_engine_fails_loudly_without_gpu():
    import pytest
    import torch
    from reface_b200.runtime import Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        Engine(0)


def test_host_schedule_matches_reference_tables(oracle):
    import numpy as np
    from reface_b200.runtime import ddim_schedule
    g = np.load(os.path.join(ROOT, "tests", "golden", "schedule.npz"))
    for S in (5, 30, 50):
        sch = ddim_schedule(S)
        assert np.array_equal(sch["timesteps"], g[f"ts{S}"])
        assert np.array_equal(sch["a_t"], g[f"a{S}"]) and np.array_equal(sch["a_prev"], g[f"ap{S}"])
        o = oracle.ddim_schedule(S)
        for k in ("a_t", "a_prev", "sigma", "sqrt_one_minus_a"):
            assert np.array_equal(sch[k], o[k]), k
    assert len(ddim_schedule(30)["timesteps"]) == 31      # util.py:48-49, assert commented out at :55
    assert np.array_equal(ddim_schedule(50)["alphas_cumprod"], g["alphas_cumprod"])


def test_option_names_used_by_scripts_and_tests_exist():
    """Engine options are strings parsed in capi.cu (rfb_set_option).  Every name the scripts' defaults, the GPU tests and
    the documentation set must be one the library knows -- an unknown name raises on the GPU box only, which costs a
    GPU run to find out."""
    import glob
    capi = open(os.path.join(ROOT, "reface_b200", "csrc", "capi.cu")).read()
    known = set(re.findall(r'k == "([a-z0-9_]+)"', capi))
    assert {"use_graph", "pdl", "gemm_lean", "gemm_mcast", "gemm_mcast_big", "gemm_splitk", "attn_flash"} <= known
    used = set()
    for path in glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "scripts", "*.py")):
        src = open(path).read()
        used |= set(re.findall(r'set_option\("([a-z0-9_]+)"', src))
    defaults = re.search(r"DEFAULTS = \{(.*?)\}", open(os.path.join(ROOT, "scripts", "unet_ab.py")).read(), re.S).group(1)
    used |= set(re.findall(r'"([a-z0-9_]+)":', defaults))
    assert used and not (used - known), f"unknown engine options: {sorted(used - known)}"

"""The drop-in boundary without a GPU: libnavgpu.so builds for sm_100a, loads, exports every symbol that
include/navgpu.h declares (and the ctypes binding declares the same set), fails loudly when no CUDA device is present,
and nothing under navigation_b200/ reaches into oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "navgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(navgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_a_c_abi():
    text = open(os.path.join(ROOT, "include", "navgpu.h")).read()
    assert 'extern "C"' in text and "#include <stdint.h>" in text
    assert "torch" not in text and "std::" not in text  # plain pointers and sizes only
    assert len(declared_symbols()) >= 60


def test_library_exports_every_declared_symbol():
    from navigation_b200 import api, build
    build.build()
    lib = ctypes.CDLL(api.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in navgpu.h but not exported: {missing}"
    assert sorted(api.SIGNATURES) == declared_symbols(), (
        set(api.SIGNATURES) ^ set(declared_symbols()))


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU every computing entry point must refuse with NAVGPU_ERR_CUDA, never compute on the CPU."""
    import navigation_b200
    a = navigation_b200.load()
    if a.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(navigation_b200.api.NavGpuError, match="no CUDA device"):
        a.costmap(10, 10, 0.05)
    with pytest.raises(navigation_b200.api.NavGpuError, match="no CUDA device"):
        a.dwa(10, 10, 0.05)
    with pytest.raises(navigation_b200.api.NavGpuError, match="no CUDA device"):
        a.fleet(2, 10, 10, 0.05, [(0.1, 0.1), (0.1, -0.1), (-0.1, -0.1), (-0.1, 0.1)])
    with pytest.raises(navigation_b200.api.NavGpuError, match="no CUDA device"):
        a.trajectory_planner(10, 10, 0.05, [(0.1, 0.1), (0.1, -0.1), (-0.1, -0.1), (-0.1, 0.1)])
    with pytest.raises(navigation_b200.api.NavGpuError, match="no CUDA device"):
        a.inflate_host(np.zeros((8, 8), np.uint8), 0, 0, 8, 8, np.zeros((4, 4), np.uint8), 2)
    # the host-side table builder is pure arithmetic (the reference's computeCost) and works everywhere
    R, costs, dists = a.build_cost_table(0.05, 0.325, 0.55, 10.0)
    assert R == 11 and costs[0, 0] == 254 and costs.shape == (13, 13)


def test_product_package_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "navigation_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(l for l in text.split("\n") if "never imports or loads anything from oracle/" not in l)
                if re.search(r"import oracle|from oracle|pyoracle|libnavoracle|libnavref|navo_|oracle_api", code):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files referring to oracle/: {bad}"

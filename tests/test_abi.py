"""The C-ABI library loads and exports every symbol include/wavetool_b200.h declares; without a GPU
the device entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from tests.common import ROOT
from wave_tool_b200 import api, build


def header_functions():
    text = open(os.path.join(ROOT, "include", "wavetool_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wt_[a-z0-9_A-Z]+)\s*\(", text)))


def test_library_builds_in_tree_and_exports_header_symbols():
    path = build.build()
    assert os.path.dirname(path) == os.path.join(ROOT, "wave_tool_b200")       # in-tree, not a JIT cache
    L = api.lib()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(api.EXPORTED_SYMBOLS) == names, "api.EXPORTED_SYMBOLS out of sync with the header"
    assert b"sm_100a" in L.wt_version()


def test_struct_layouts_match_header():
    assert C.sizeof(api.Wave) == 24 and C.sizeof(api.Camera) == 36 and C.sizeof(api.Params) == 32
    assert C.sizeof(api.Frame) == 4 * (16 + 32 + 4 + 4 + 4) and C.sizeof(api.Checksum) == 40


def test_binary_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


@pytest.mark.skipif(api.lib().wt_device_count() > 0, reason="a GPU is visible")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(api.WaveToolError) as e:
        api.Context(0)
    assert e.value.status == api.WT_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def build_example(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "wt_headless")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           os.path.join(ROOT, "examples", "wt_headless.cpp"), "-L" + os.path.dirname(build.LIB_PATH), "-lwavetool_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + os.path.dirname(build.LIB_PATH), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


@pytest.mark.skipif(api.lib().wt_device_count() > 0, reason="a GPU is visible")
def test_cpp_host_program_builds_against_the_c_abi_and_fails_loudly_without_gpu(tmp_path):
    """examples/wt_headless.cpp is plain C++ on include/wavetool_b200.h (no CUDA compiler needed for the host side)."""
    import subprocess
    build.build()
    exe = build_example(tmp_path)
    res = subprocess.run([exe, "64", "2"], capture_output=True, text=True)
    assert res.returncode == 2 and "no CPU fallback" in res.stderr

"""CPU: the C-ABI library loads without a GPU driver, exports every symbol include/walkgpt_b200.h declares, its struct
layouts agree with a C compiler, and every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest
import torch

from walkgpt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "walkgpt_b200.h")


def test_library_exports_every_declared_symbol():
    declared = set(re.findall(r"WG_API[^;(]*?(wg_\w+)\s*\(", open(HEADER).read()))
    assert declared == set(_lib.exported_symbols()) and len(declared) >= 20
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported by the .so"
    assert handle.wg_version() == _lib.CONSTANTS["WG_ABI_VERSION"]


def test_library_has_no_libcuda_link_dependency():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out  # cudart is linked statically, the driver is resolved at run time


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs a C++ compiler")
def test_struct_layouts_match_the_c_compiler(tmp_path):
    names = sorted(_lib.STRUCTS)
    src = '#include "walkgpt_b200.h"\n#include <stdio.h>\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    f = tmp_path / "sz.cpp"
    f.write_text(src)
    exe = tmp_path / "sz"
    subprocess.run(["g++", "-I", os.path.join(ROOT, "include"), str(f), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n in names:
        assert int(out[n]) == C.sizeof(_lib.STRUCTS[n]), n


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_no_cpu_fallback_without_a_gpu():
    handle = _lib.lib()
    assert handle.wg_device_check(0) == _lib.CONSTANTS["WG_ERR_UNSUPPORTED"]
    assert b"no CPU fallback" in handle.wg_last_error()
    with pytest.raises(_lib.WalkGPTB200Error):
        _lib.require_device(0)
    # argument validation happens before any device work
    assert handle.wg_gemm(None, None) == _lib.CONSTANTS["WG_ERR_INVALID"]
    a = _lib.GemmArgs()
    a.M, a.N, a.K = 128, 128, 60
    assert handle.wg_gemm(C.byref(a), None) == _lib.CONSTANTS["WG_ERR_INVALID"]
    assert b"multiples of 8" in handle.wg_last_error()


def test_modules_refuse_cpu_tensors():
    from walkgpt_b200.modules import CalibratedTextProjector, CLIPVisionTower, MultiScaleQFormerProjector, postprocess_masks

    with pytest.raises(_lib.WalkGPTB200Error):
        CalibratedTextProjector(256, 256)(torch.zeros(2, 256))
    with pytest.raises(_lib.WalkGPTB200Error):
        MultiScaleQFormerProjector(256, 64)(torch.zeros(1, 64, 256))
    with pytest.raises(_lib.WalkGPTB200Error):
        CLIPVisionTower(layers=1)(torch.zeros(1, 3, 448, 448))
    with pytest.raises(_lib.WalkGPTB200Error):
        postprocess_masks(torch.zeros(1, 1, 64, 64), (448, 448), (448, 448))


def test_launch_counter_and_profiler_entry_points():
    handle = _lib.lib()
    handle.wg_launch_count(1)
    assert handle.wg_launch_count(0) == 0
    assert handle.wg_profile_enable(0) == 0
    buf = C.create_string_buffer(64)
    assert handle.wg_profile_collect(buf, 64) >= 1

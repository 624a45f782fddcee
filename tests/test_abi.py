"""CPU tests of the drop-in boundary: the C ABI library loads, exports exactly what include/b200zk.h
declares, and fails loudly (no CPU fallback) without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from ethrex_b200 import _ffi
    syms = _header_symbols()
    assert len(syms) >= 30
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200zk.h but not exported"
    # the ctypes table binds the same set (no stale or missing prototypes)
    assert sorted(_ffi.SIGNATURES) == syms


def test_rust_sys_crate_declares_the_same_symbols():
    txt = open(os.path.join(ROOT, "rust", "b200zk-sys", "src", "lib.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (b200zk_[a-z0-9_]+)\s*\(", txt)))
    assert rust == _header_symbols()


def test_status_strings_and_version():
    from ethrex_b200 import _ffi
    assert _ffi.lib.b200zk_abi_version() == 2
    assert b"no CPU fallback" in _ffi.lib.b200zk_strerror(_ffi.ERR_NO_DEVICE)
    assert _ffi.lib.b200zk_strerror(_ffi.ERR_NOT_ON_CURVE) == b"input point not on curve"
    # null-context calls are rejected, not crashed
    assert _ffi.lib.b200zk_set_msm_window(None, 4) == _ffi.ERR_INVALID_ARG
    assert _ffi.lib.b200zk_launch_count(None) == 0


def test_no_device_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import ethrex_b200
    with pytest.raises(ethrex_b200.NoDeviceError, match="no CPU fallback"):
        ethrex_b200.Context(0)


def test_error_type_mirrors_backend_error():
    """Display strings of /root/reference/crates/prover/src/backend/error.rs:4-20."""
    from ethrex_b200.errors import B200Error, status_to_error
    assert str(B200Error.proving("x")) == "Proving error: x"
    assert str(B200Error.serialization("x")) == "Serialization error: x"
    assert str(B200Error.verify_not_supported()) == "Not implemented: Verify not implemented for this backend"
    assert status_to_error(3, "m").kind == "Serialization"
    assert status_to_error(5, "m").kind == "Proving"
    assert status_to_error(8, "m").kind == "NotImplemented"


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under ethrex_b200/ may reference it."""
    pkg = os.path.join(ROOT, "ethrex_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cpu_oracle" not in txt and "import pyref" not in txt and "libb200zk_oracle" not in txt, f


def test_c_demo_fails_loudly_without_a_gpu():
    """The plain-C demo links against the product library and reports `no CUDA device` (status 6) on this box: the
    product has no CPU fallback to fall into."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "examples", "build", "c_abi_demo")
    if not os.path.exists(exe):
        pytest.skip("examples/build/c_abi_demo not built")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "no CUDA device" in r.stderr

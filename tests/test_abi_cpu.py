"""CPU: the C-ABI library loads, exports every symbol include/whisper_b200.h declares, and fails
loudly (never falls back to the CPU) when no GPU is present."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

from whisper_mojo_b200 import _lib

HEADER = os.path.join(ROOT, "include", "whisper_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(w[bmt]_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_plain_c():
    subprocess.check_call(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-std=c99", "-fsyntax-only",
                           "-x", "c", HEADER])


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (w[bmt]_\w+)", out))
    assert exported == set(names), exported ^ set(names)
    assert lib.wb_abi_version() == 1


def test_no_oracle_or_cpu_fallback_in_product():
    """The product package must not import / link anything under oracle/."""
    pkg = os.path.join(ROOT, "whisper_mojo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "from oracle" not in txt and "import oracle" not in txt and "whisper_oracle" not in txt, f
    out = subprocess.check_output(["ldd", _lib.LIB_PATH], text=True)
    assert "oracle" not in out


def test_config_struct_matches_header():
    from whisper_mojo_b200 import WhisperConfig

    arr = WhisperConfig.tiny().as_c_array()
    assert list(arr)[:14] == [384, 6, 4, 51865, 1500, 448, 80, 50258, 50259, 50359, 50363, 50257, 195, 1]
    n = _lib.load().wm_weight_count(ctypes.cast(arr, ctypes.c_void_p))
    assert n == 37_760_640 == WhisperConfig.tiny().weight_count()
    small = WhisperConfig.small_shaped()
    assert _lib.load().wm_weight_count(ctypes.cast(small.as_c_array(), ctypes.c_void_p)) == small.weight_count()


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from whisper_mojo_b200 import Tensor, Whisper

    with pytest.raises(_lib.WhisperB200Error) as e:
        Tensor(4, 4)
    assert e.value.code == _lib.WB_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.WhisperB200Error) as e:
        Whisper()
    assert e.value.code == _lib.WB_ERR_CUDA
    out = np.zeros(4, np.float32)
    rc = _lib.load().wb_debug_gemm(1, out.ctypes.data, 1, 1, 64, 64, 1, 1, 0, 1, out.ctypes.data, 1, None, 3, out.ctypes.data)
    assert rc == _lib.WB_ERR_CUDA


def test_bad_handles_are_rejected():
    lib = _lib.load()
    assert lib.wt_gelu(123456) == _lib.WB_ERR_ARG
    assert "bad tensor handle" in _lib.last_error()
    assert lib.wm_destroy(987654) == _lib.WB_ERR_ARG
    assert lib.wt_tensor_free(0) == _lib.WB_OK


def test_cpp_driver_links_against_the_abi_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/main.cpp is main.mojo (main.mojo:11-45) over the C ABI in plain C++: it must build with g++ alone,
    need nothing but the library + libcudart at run time, and -- with no device -- stop at wm_create with the
    library's error text instead of producing ids some other way."""
    import torch

    from whisper_mojo_b200 import build as B

    exe = B.build_example()
    needed = subprocess.check_output(["readelf", "-d", exe], text=True)
    assert "libwhisper_b200.so" in needed and "torch" not in needed and "python" not in needed
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run is covered by the gpu test")
    r = subprocess.run([exe, str(tmp_path / "missing.bin")], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and "Token IDs" not in r.stdout


def test_every_option_key_is_documented_in_the_header():
    """wm_set_option is a string-keyed entry point: every key the library accepts (csrc/api.cu) must be described in
    include/whisper_b200.h, the only document an FFI caller reads."""
    api = open(os.path.join(ROOT, "whisper_mojo_b200", "csrc", "api.cu")).read()
    keys = sorted(set(re.findall(r'!strcmp\(key, "([a-z_0-9]+)"\)', api)))
    assert len(keys) >= 10, keys
    header = open(HEADER).read()
    missing = [k for k in keys if f'"{k}"' not in header]
    assert not missing, f"options without a description in the header: {missing}"

"""The C-ABI shared library builds for sm_100a without a GPU, loads, and exports every symbol that
include/pylda_b200.h declares -- and the ctypes binding covers exactly that set.  No compute calls
here (no GPU in the build container); the one call made, pylda_create, must FAIL loudly without a
device instead of falling back to anything."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pylda_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pylda_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from pylda_b200 import build
    return build.build_native()


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for required in ("pylda_create", "pylda_destroy", "pylda_last_error", "pylda_set_corpus", "pylda_estep",
                     "pylda_estep_resident", "pylda_get_results", "pylda_comm_init", "pylda_comm_unique_id"):
        assert required in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_symbols():
        assert hasattr(lib, name), "libpylda_b200.so does not export %s" % name


def test_binding_covers_exactly_the_header(lib_path):
    from pylda_b200 import native
    assert sorted(native._SIGNATURES) == declared_symbols()
    lib = native.load_library()
    assert lib.pylda_abi_version() == native.ABI_VERSION
    text = open(HEADER).read()
    assert int(re.search(r"#define\s+PYLDA_ABI_VERSION\s+(\d+)", text).group(1)) == native.ABI_VERSION
    assert int(re.search(r"#define\s+PYLDA_NCCL_ID_BYTES\s+(\d+)", text).group(1)) == native.NCCL_ID_BYTES


def test_stats_struct_matches_header(lib_path):
    from pylda_b200 import native
    text = open(HEADER).read()
    body = re.search(r"typedef struct pylda_stats \{(.*?)\} pylda_stats;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int64_t|int32_t|double)\s+([a-z_]+)\s*;", body)
    ctype = {"int64_t": ctypes.c_int64, "int32_t": ctypes.c_int32, "double": ctypes.c_double}
    assert [(n, ctype[t]) for t, n in fields] == list(native.Stats._fields_)


def test_no_cpu_fallback_without_a_device(lib_path):
    """Without a usable B200 the product path raises; it never routes to the oracle or numpy."""
    from pylda_b200 import native
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = os.path.exists("/dev/nvidia0")
    if has_gpu:
        pytest.skip("a GPU is present; the failure path is exercised on the CPU-only container")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|failed"):
        native.EStepContext(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pylda_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            code = re.sub(r'""".*?"""', "", src, flags=re.S)
            code = "\n".join(line.split("#")[0] for line in code.split("\n"))
            assert "/root/reference" not in code, fn    # nothing reads the reference at run time


def test_kernel_sass_uses_bulk_async_engine(lib_path):
    """The staging/scatter path must be the bulk-async (TMA) engine: UBLKCP for the row gathers,
    UBLKRED for the f64 reduce-add scatter (B200_PROFILING.md 'What proves a Blackwell-native kernel')."""
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "UBLKCP" in out.stdout
    assert "UBLKRED" in out.stdout or "UBLKRED".lower() in out.stdout.lower()
    assert "sm_100a" in out.stdout

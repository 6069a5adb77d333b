"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/wdgh_b200.h declares, and the ctypes binding covers exactly that set.  No compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wdgh_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wdgh_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    fns = header_functions()
    for must in ("wdgh_spmm_csr", "wdgh_structure_counts", "wdgh_gram", "wdgh_edge_cosine", "wdgh_pipeline_host",
                 "wdgh_coo_to_csr", "wdgh_degree_scale", "wdgh_las_score", "wdgh_gntk_transform"):
        assert must in fns


def test_library_exports_every_header_symbol():
    import wdgh_b200._lib as L
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.wdgh_version() == 100


def test_binding_matches_header():
    import wdgh_b200._lib as L
    assert sorted(L.SIGNATURES) == header_functions()


def test_argument_counts_match_header():
    import wdgh_b200._lib as L
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in L.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, binding has {len(args)}"


def test_no_cpu_fallback_without_device():
    import torch
    import wdgh_b200._lib as L
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L._device_checked = False
    with pytest.raises(L.WdghError):
        L.require_device()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "when-do-gnns-help_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header must compile as strict C99 and a C program must link against the library
    (a cgo / JNI / ctypes-free consumer sees exactly this).  No compute call: there is no GPU here."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "wdgh_b200.h"\n'
                   'int main(void) {\n'
                   '  int sm = 0, major = 0, minor = 0;\n'
                   '  (void)wdgh_device_info(&sm, &major, &minor);   /* fails without a device, must not crash */\n'
                   '  return (wdgh_version() == WDGH_VERSION && wdgh_last_error() != 0) ? 0 : 1;\n'
                   '}\n')
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "when-do-gnns-help_b200", "wdgh_b200")
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe),
                           f"-L{libdir}", "-lwdgh_b200", f"-Wl,-rpath,{libdir}"])
    assert subprocess.run([str(exe)], timeout=60).returncode == 0

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/rm_accel.h declares,
and fails LOUDLY without a GPU (there is no CPU fallback in the product path)."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(built):
    from runmat_b200 import _capi

    names = _capi.declared_symbols()
    assert len(names) > 80
    missing = [n for n in names if not hasattr(_capi.lib, n)]
    assert not missing, f"declared in include/rm_accel.h but not exported: {missing}"
    assert _capi.lib.rm_abi_version() == 1


def test_handle_struct_layout_matches_header(built):
    from runmat_b200 import _capi

    assert C.sizeof(_capi.Handle) == 8 + 4 + 4 + 8 * 16
    assert _capi.Handle.shape_arr.offset == 16
    assert C.sizeof(_capi.Telemetry) == 6 * 16 + 5 * 8


def test_product_path_does_not_touch_the_oracle(built):
    """The product may not import/link/call anything under oracle/ (nor any CPU fallback)."""
    for f in list((ROOT / "runmat_b200").rglob("*.py")) + list((ROOT / "runmat_b200" / "csrc").glob("*")):
        if f.name == "build.py":  # build.py compiles the checker; building it is not using it
            continue
        text = f.read_text(errors="ignore")
        assert "librm_oracle" not in text and "oracle_binding" not in text and "rm_oracle" not in text, f
    out = subprocess.run(["ldd", str(ROOT / "runmat_b200" / "librm_accel_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_no_gpu_fails_loudly(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from runmat_b200 import B200Provider, ProviderError

    with pytest.raises(ProviderError, match="RM_NO_DEVICE"):
        B200Provider(0)


def test_missing_extension_raises_at_import(built, tmp_path):
    code = "import os; os.environ['RUNMAT_B200_LIB']='/nonexistent/lib.so'; import runmat_b200"
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_header_is_plain_c_and_a_c_caller_links(built, tmp_path):
    """The boundary is a C ABI: include/rm_accel.h must compile as C11 (no C++ in the signatures), and a C program must link
    against the shared library and reach an entry point. Without a GPU the create call has to fail with RM_NO_DEVICE and a
    message; with one it has to succeed (the program only creates and destroys the provider: no compute)."""
    src = tmp_path / "caller.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "rm_accel.h"\n'
        "int main(void) {\n"
        "  rm_provider* p = NULL;\n"
        "  if (rm_abi_version() != 1) return 2;\n"
        "  rm_status st = rm_provider_create(0, 0, RM_F64, &p);\n"
        '  if (st == RM_OK) { rm_provider_destroy(p); printf("created\\n"); return 0; }\n'
        '  printf("status %d: %s\\n", (int)st, rm_last_error());\n'
        "  return st == RM_NO_DEVICE && strlen(rm_last_error()) > 0 ? 0 : 3;\n"
        "}\n"
    )
    libdir = ROOT / "runmat_b200"
    exe = tmp_path / "caller"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe),
                    f"-L{libdir}", "-lrm_accel_b200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr

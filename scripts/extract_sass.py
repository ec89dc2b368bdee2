#!/usr/bin/env python
"""Evidence extracts: the SASS mnemonics that prove what each hand-written kernel runs on.

  python scripts/extract_sass.py            # needs only cuobjdump + the built library / kernel cache (no GPU)

Writes profiles/sass_<kernel>.txt: an opcode histogram of the kernel plus the first lines that carry the instructions of
interest (UTCIMMA / UTMALDG / LDTM / UTCBAR for the tcgen05 engine, DMMA for the FP64 engine, LDG.E...256 / STG.E...256 for
the generated streaming kernels, FMUL/FADD counts for the bit-exact imfilter).
"""
from __future__ import annotations

import collections
import ctypes as C
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
LIB = ROOT / "runmat_b200" / "librm_accel_b200.so"
OUT = ROOT / "profiles"

INTEREST = {
    "ozaki_gemm_kernel": ["UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS"],
    "slice_kernel": ["F2I", "DMUL", "STS"],
    "dgemm_dmma_kernel": ["DMMA", "LDGSTS"],
    "evolve_kernel": ["DFMA", "DMUL", "DADD", "MUFU"],
    "imfilter_regblock_kernel": ["FMUL", "FADD", "FFMA"],
    "imfilter_tma_f32_kernelILi5ELi8ELi1": ["FMUL2", "FFMA2", "UTMALDG", "SYNCS", "LDS"],
    "imfilter_packed_f32_kernelILi5": ["FMUL2", "FFMA2", "LDS"],
    "normalize_fixed_clampgamma_f32_kernel": ["MUFU", "FMUL", "FADD", "FMNMX"],
    "trsm_block_kernelILi0": ["SHFL", "DFMA", "LDS", "BAR"],
    "moments_partial_fold_kernel": ["SHFL", "DADD", "DFMA", "ATOM"],
    "normalize_fixed_kernel": ["MUFU", "FMUL", "FADD"],
    "lu_panel_smem_kernel": ["DFMA", "DMUL", "BAR", "UCGABAR", "MEMBAR"],
    "lu_panel_push_kernel": ["STAS", "SYNCS", "CREDUX", "VOTE", "DFMA", "MUFU", "BAR", "UCGABAR", "MEMBAR"],
    "dgemm_dmma_kernelILb1ELb1E": ["DMMA", "LDGSTS"],
    "p2p_publish_kernel": ["ST.E", "STG", "MEMBAR", "FENCE"],
    "p2p_combine_kernel": ["LD.E", "LDG", "NANOSLEEP"],
    "rm_fused_ew": ["LDG", "STG"],
    "rm_fused_red": ["LDG", "DADD", "SHFL", "ST.E", "STG"],
}


def sass_functions(path: Path) -> dict[str, list[str]]:
    txt = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True).stdout
    funcs: dict[str, list[str]] = {}
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            funcs[cur].append(line.strip())
    return funcs


def opcode(line: str) -> str:
    body = re.sub(r"^/\*[0-9a-f]+\*/\s*", "", line)
    body = re.sub(r"^@!?U?P\d+\s+", "", body)
    return body.split()[0].rstrip(";") if body.split() else ""


def demangle(name: str) -> str:
    r = subprocess.run(["c++filt", name], capture_output=True, text=True)
    return r.stdout.strip() or name


def write_extract(kernel: str, mangled: str, lines: list[str], source: str):
    ops = collections.Counter(opcode(l) for l in lines)
    keys = INTEREST.get(kernel, [])
    with open(OUT / f"sass_{kernel}.txt", "a") as f:
        f.write(f"== {demangle(mangled)}\n   source: {source}; {len(lines)} SASS instructions\n")
        f.write("   opcode histogram (top 24): " + ", ".join(f"{k} x{v}" for k, v in ops.most_common(24)) + "\n")
        for key in keys:
            hits = [l for l in lines if key in opcode(l)]
            f.write(f"   {key}: {len(hits)} instructions" + (f"; e.g. {re.sub(r'/\\*[0-9a-f]+\\*/', '', hits[0]).strip()}" if hits else "") + "\n")
        f.write("\n")


def main():
    OUT.mkdir(exist_ok=True)
    for k in INTEREST:
        (OUT / f"sass_{k}.txt").unlink(missing_ok=True)
    funcs = sass_functions(LIB)
    for mangled, lines in funcs.items():
        for kernel in INTEREST:
            if kernel in mangled and not kernel.startswith("rm_fused"):
                write_extract(kernel, mangled, lines, "runmat_b200/librm_accel_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false)")
    # the generated (NVRTC) headline kernels: lower them through the library's debug entry points and disassemble the cubin
    from runmat_b200 import _capi
    import fusion_text as ft

    lib = _capi.lib
    for name, shader, fn, args in (("rm_fused_ew", ft.sin_mul_add_wgsl(), lib.rm_debug_lower_elementwise, (0, 4)),
                                   ("rm_fused_red", ft.sum_sin_mul_add_wgsl(), None, None)):
        need = C.c_size_t()
        if name == "rm_fused_ew":
            fn(shader.encode(), *args, None, 0, C.byref(need))
            buf = C.create_string_buffer(need.value)
            fn(shader.encode(), *args, buf, need.value, None)
        else:
            lib.rm_debug_lower_reduction(shader.encode(), 0, 0, None, 0, C.byref(need), None, None)
            buf = C.create_string_buffer(need.value)
            lib.rm_debug_lower_reduction(shader.encode(), 0, 0, buf, need.value, None, None, None)
        with tempfile.TemporaryDirectory() as td:
            import os

            os.environ["RUNMAT_B200_KCACHE"] = td
            if lib.rm_debug_compile(buf.value, name.encode(), None) != 0:
                raise RuntimeError(lib.rm_last_error().decode())
            cub = next(Path(td).glob("*.cubin"))
            for mangled, lines in sass_functions(cub).items():
                write_extract(name, mangled, lines, f"NVRTC cubin of the lowered headline program ({'C = sin(A).*B+1' if name == 'rm_fused_ew' else 'sum(sin(A).*B+1)'}, f64)")
    print("wrote", ", ".join(sorted(p.name for p in OUT.glob("sass_*.txt"))))


if __name__ == "__main__":
    main()

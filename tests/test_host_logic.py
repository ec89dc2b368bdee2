"""CPU tests of device-side index logic that can be compiled for the host (no GPU needed)."""
import pathlib
import subprocess

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_imfilter_tile_iterator_matches_divisions(tmp_path):
    """The persistent TMA imfilter kernel (image.cu) walks its tiles t = blockIdx.x + k * gridDim.x with an incremental iterator instead
    of two divisions per tile; producer lane and compute warps both rely on it. Cut the struct out of the source, compile it with
    g++ and compare every step with the division-based decomposition; all CTAs together must cover every tile exactly once."""
    src = (ROOT / "runmat_b200" / "csrc" / "image.cu").read_text()
    a, b = src.index("// ---- imf_tile_iter begin"), src.index("// ---- imf_tile_iter end")
    (tmp_path / "imf_tile_iter.inc").write_text(src[a:b])
    exe = tmp_path / "check_imf_tile_iter"
    subprocess.run(["g++", "-std=c++17", "-O2", f"-I{tmp_path}", "-o", str(exe), str(ROOT / "tests" / "golden" / "check_imf_tile_iter.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr

"""The batched Jacobi eigensolver's source (micmec_b200/csrc/mm_eigh.cu) compiled for the host and run serially
(tests/eigh_host_check.cpp): round-robin schedule, reconstruction, orthogonality, ordering, degenerate spectra."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_jacobi_eigh_source_on_the_host(tmp_path):
    exe = str(tmp_path / "eigh_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-o", exe, os.path.join(ROOT, "tests", "eigh_host_check.cpp")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout
    assert "FAIL" not in out.stdout and out.stdout.count(" ok") >= 15

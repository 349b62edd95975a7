"""The drop-in boundary used from plain C (tests/c/abi_smoke.c): compiled with gcc against
include/*.h and linked with libsema_b200.so — no Python or torch in the process."""
import os
import subprocess

import pytest

from sema_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    so_dir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["/usr/bin/gcc", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", so_dir, "-lsema_b200", "-lm", f"-Wl,-rpath,{so_dir}"])
    return exe


@pytest.mark.skipif(not os.path.exists(_lib.SO_PATH), reason="libsema_b200.so not built")
def test_c_program_links_and_refuses_to_run_without_a_gpu(tmp_path):
    exe = _build(tmp_path)
    if _lib.lib().sema_device_count() > 0:
        pytest.skip("a GPU is present (covered by the gpu-marked test)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_program_drives_both_seams_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi_smoke ok" in r.stdout

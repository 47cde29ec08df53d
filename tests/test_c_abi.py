"""The drop-in boundary from plain C (tests/c/abi_smoke.c): a C translation unit that includes
include/pfrx.h, fills a pfrx_config by hand and steps host-resident state -- no Python, no ctypes
in the loop -- and the route from that C configuration to a specialised cubin:
pfrx_config_write / pfrx_config_dump -> python -m pflotran_elm_interface_b200.specialize -> pfrx_load_specialized."""
import os
import shutil
import subprocess
import sys

import pytest

from pflotran_elm_interface_b200 import abi, specialize, workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
PKG = os.path.join(ROOT, "pflotran_elm_interface_b200")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "c"),
           os.path.join(HERE, "c", "abi_smoke.c"), "-L", PKG, "-lpfrx_b200", "-lm", "-Wl,-rpath," + PKG, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_c_translation_unit_builds_and_writes_the_configuration(tmp_path):
    """gcc compiles the header as C99 and links every symbol the program uses; the set-up half
    (pfrx_config_write, no device) produces a file whose signature is the Python configuration's"""
    exe = _build(tmp_path)
    dump = str(tmp_path / "calcite.cfg")
    out = subprocess.run([exe, dump], env=dict(os.environ, PFRX_SMOKE_SETUP_ONLY="1"), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    cfg = abi.ReactionConfig.from_dump(dump)
    ref = W.by_name("c2", ncell=2).cfg
    sig = specialize.signature(ref)
    assert f"signature {sig:016x}" in out.stdout
    assert cfg.dump_signature == sig == specialize.signature(cfg)
    # the generator writes the same kernel for the C-built configuration as for the deck-built one
    assert specialize.generate_source(cfg) == specialize.generate_source(ref)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc")
def test_cli_builds_cubins_from_a_dump(tmp_path):
    exe = _build(tmp_path)
    dump = str(tmp_path / "calcite.cfg")
    subprocess.check_call([exe, dump], env=dict(os.environ, PFRX_SMOKE_SETUP_ONLY="1"))
    out = subprocess.run([sys.executable, "-m", "pflotran_elm_interface_b200.specialize", dump, "--out", str(tmp_path / "spec"),
                          "--styles", "default,straight"], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0, out.stdout + out.stderr
    paths = out.stdout.split()
    assert len(paths) == 2 and all(os.path.exists(p) and p.endswith("f2.cubin") for p in paths)


@pytest.mark.gpu
def test_c_program_steps_through_the_generic_and_the_specialised_kernel(tmp_path):
    """the whole program on a GPU: pfrx_create, pfrx_rstep_host on 64 cells against the oracle's
    numbers in calcite_fixture.h, pfrx_config_dump; then the CLI builds the cubin from that dump
    and the program steps again through pfrx_load_specialized"""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    exe = _build(tmp_path)
    dump = str(tmp_path / "calcite.cfg")
    out = subprocess.run([exe, dump], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
    assert "lanes 0" in out.stdout or "lanes 1" in out.stdout or "lanes 4" in out.stdout   # a generic kernel
    r = subprocess.run([sys.executable, "-m", "pflotran_elm_interface_b200.specialize", dump, "--out", str(tmp_path / "spec")],
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    cubin = r.stdout.split()[-1]
    out2 = subprocess.run([exe, dump, cubin], capture_output=True, text=True)
    assert out2.returncode == 0 and "OK" in out2.stdout, out2.stdout + out2.stderr
    assert "lanes -1" in out2.stdout   # the specialised kernel

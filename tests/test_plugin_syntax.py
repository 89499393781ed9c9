"""The onika plugin shim (plugin/xsb_onika_plugin.cpp) is real source: it is type-checked here against minimal stand-ins
of the onika / exaNBody headers (plugin/mock), in both naming modes, and every reference operator name of INTEGRATION.md's
table must be registered by it."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "plugin", "xsb_onika_plugin.cpp")


@pytest.mark.parametrize("defs", [[], ["-DXSB_REPLACE_REFERENCE_OPERATORS=1"]])
def test_plugin_source_type_checks_against_mock_headers(defs):
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "plugin", "mock"), "-I" + os.path.join(ROOT, "include")] + defs + [SRC]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]


def test_plugin_registers_every_operator_of_the_hot_path():
    src = open(SRC).read()
    names = set(re.findall(r'XSB_OPNAME\("([a-z0-9_]+)"\)', src))
    want = {"chunk_neighbors", "lj_compute_force", "lj_compute_force_symetric", "lj_multi_force", "zbl_compute_force", "zbl_multi_force",
            "exp6_compute_force", "buckingham_compute_force", "johnson_force", "johnson_emb", "johnson_force_reuse_emb", "eam_alloy_init",
            "eam_alloy_force", "snap_force", "ghost_update_r", "ghost_update_all_no_fv", "ghost_update_opt", "update_force_energy_from_ghost",
            "update_virial_force_energy_from_ghost", "update_opt_from_ghost", "zero_force_energy",
            "yukawa_compute_force", "relax_compute_force", "zero_compute_force", "sutton_chen_force", "sutton_chen_emb",
            "sutton_chen_force_reuse_emb", "vniitf_force", "vniitf_emb", "vniitf_force_reuse_emb", "snap_force_fp32",
            "yukawa_compute_force_symetric", "buckingham_compute_force_symetric", "exp6_compute_force_symetric"}
    assert want <= names, sorted(want - names)
    # every C entry point the shim calls is declared by the public header
    hdr = open(os.path.join(ROOT, "include", "xsb200.h")).read()
    for call in set(re.findall(r"\b(xsb_[a-z0-9_]+)\(", src)):
        if call in ("xsb_bind_grid",):
            continue
        assert re.search(r"\b%s\(" % call, hdr), "%s is not declared in include/xsb200.h" % call

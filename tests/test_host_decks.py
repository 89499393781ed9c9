"""Host operator layer (host/, C++): YAML decks with the reference's operator names and slots -> C ABI.

CPU tests: the deck reader (YAML subset, units, includes, anchors), graph resolution (aliases, batches, rebind,
conditions), slot validation and failure behaviour -- through the `xsb200-run` binary, the same way a user drives it.
GPU tests: whole decks run on the device and the dumped forces / energies are checked against the CPU oracle on the
dumped positions (tolerance 1e-10 relative to the field maximum, FP64 mode), plus NVE energy conservation."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from helpers import EV, JOHNSON_CU, SC_CU, SC_XX, GridSystem, johnson_params, write_setfl  # noqa: E402

RUN = os.path.join(ROOT, "host", "xsb200-run")
DECKS = os.path.join(ROOT, "tests", "decks")
TOL = 1e-10


@pytest.fixture(scope="module", autouse=True)
def built():
    import exastamp_b200 as xsb
    xsb.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])


def run(*args, cwd=None, check=True):
    p = subprocess.run([RUN] + [str(a) for a in args], capture_output=True, text=True, cwd=cwd, timeout=600)
    if check and p.returncode != 0:
        raise AssertionError("xsb200-run %s failed (%d):\n%s\n%s" % (" ".join(map(str, args)), p.returncode, p.stdout, p.stderr))
    return p


def quantity(text):
    return float(run("--quantity", text).stdout)


# ------------------------------------------------------------------------------------------------ units
def test_units_match_the_internal_unit_system():
    # internal units: angstrom, Da, ps, e, K (include/exaStamp/unit_system.h:28-36)
    assert quantity("3.4 ang") == 3.4
    assert quantity("0.244E-09 m") == pytest.approx(2.44, rel=1e-15)
    assert quantity("1.0e-3 ps") == 1.0e-3
    assert quantity("2 fs") == pytest.approx(2e-3, rel=1e-15)
    assert quantity("39.948 Da") == 39.948
    assert quantity("1 eV") == pytest.approx(EV, rel=1e-14)
    assert quantity("0.0104 eV") == pytest.approx(0.0104 * EV, rel=1e-14)
    assert quantity("2.522E-20 J") == pytest.approx(2.522e-20 / 1.602176634e-19 * EV, rel=1e-13)
    assert quantity("50. m/s") == pytest.approx(50.0 * 1e10 / 1e12, rel=1e-14)            # ang / ps
    assert quantity("1 kcal/mol") == pytest.approx(4184.0 / 6.02214076e23 / 1.602176634e-19 * EV, rel=1e-13)
    assert quantity("1 1/ang") == 1.0
    assert quantity("2 ang^-1") == 2.0
    assert quantity("1 eV/ang") == pytest.approx(EV, rel=1e-14)
    assert quantity("300 K") == 300.0
    assert quantity("42") == 42.0
    assert quantity("0 e-") == 0.0
    assert run("--quantity", "1 parsec", check=False).returncode != 0


# ------------------------------------------------------------------------------------------------ yaml subset
def test_yaml_subset_anchors_flow_includes(tmp_path):
    (tmp_path / "base.yaml").write_text(
        "ghost_cfg: &g\n  gpu_buffer_pack: true\n  wait_all: false\n"
        "ghost_update_r: *g\n"
        "merged:\n  <<: *g\n  wait_all: true\n"
        "species:\n  - Al: { mass: 26.982 Da , z: 13 }\n  - Cu: { mass: 63.546 Da , z: 29 }\n"
        "global:\n  dt: 2.0e-3 ps   # comment\n  rcut_inc: 1.0 ang\n")
    (tmp_path / "deck.msp").write_text(
        "includes:\n  - base.yaml\n"
        "global:\n  dt: 1.0e-3 ps\n"
        "lj_compute_force:\n  parameters: { epsilon: 0.0104 eV , sigma: 3.4 ang }\n  rcut: 8.0 ang\n"
        "bounds: [[0 ang ,0 ang,0 ang],\n         [40. ang, 40. ang, 40. ang]]\n"
        "quoted: \"a: b # not a comment\"\n"
        "compute_force: lj_compute_force\n")
    out = run("--parse", tmp_path / "deck.msp").stdout
    assert 'ghost_update_r: {gpu_buffer_pack: "true", wait_all: "false"}' in out
    assert 'dt: "1.0e-3 ps", rcut_inc: "1.0 ang"' in out                  # including file overrides key by key
    assert 'bounds: [["0 ang", "0 ang", "0 ang"], ["40. ang", "40. ang", "40. ang"]]' in out
    assert 'quoted: "a: b # not a comment"' in out
    assert 'merged: {wait_all: "true", gpu_buffer_pack: "true"}' in out or 'merged: {gpu_buffer_pack: "true", wait_all: "true"}' in out
    assert 'species: [{Al: {mass: "26.982 Da", z: "13"}}, {Cu: {mass: "63.546 Da", z: "29"}}]' in out
    assert "includes" not in out


def test_yaml_errors_are_reported_with_line_numbers(tmp_path):
    (tmp_path / "bad.msp").write_text("a:\n  b: 1\n c: 2\n")
    p = run("--parse", tmp_path / "bad.msp", check=False)
    assert p.returncode != 0 and "line 3" in p.stderr
    (tmp_path / "bad2.msp").write_text("a: *nowhere\n")
    p = run("--parse", tmp_path / "bad2.msp", check=False)
    assert p.returncode != 0 and "alias" in p.stderr


# ------------------------------------------------------------------------------------------------ operators and graphs
def test_operator_factory_lists_the_reference_names():
    names = set(run("--list-operators").stdout.split())
    # the hot-path operators of SURVEY.md 8(b), exact spelling
    for n in ["lj_compute_force", "lj_multi_force", "lj_compute_force_symetric", "johnson_force", "johnson_emb", "johnson_force_reuse_emb", "johnson_init",
              "eam_alloy_force", "eam_alloy_init", "snap_force", "chunk_neighbors", "ghost_update_r", "ghost_update_opt", "ghost_update_all_no_fv",
              "update_force_energy_from_ghost", "update_opt_from_ghost", "ghost_comm_scheme", "zero_force_energy", "force_to_accel", "push_f_v_r", "push_f_v",
              "particle_displ_over", "backup_r", "move_particles", "simulation_thermodynamic_state", "domain", "lattice",
              "yukawa_compute_force", "yukawa_multi_force", "relax_compute_force", "zero_compute_force", "sutton_chen_force", "sutton_chen_emb",
              "sutton_chen_force_reuse_emb", "sutton_chen_init", "vniitf_force", "vniitf_emb", "vniitf_force_reuse_emb", "vniitf_init"]:
        assert n in names, n


def graph_of(deck, *extra):
    out = run(deck, "--dry-run", *extra).stdout
    g = re.search(r"^graph: (.*)$", out, re.M).group(1).split()
    vals = dict(zip(*[iter(re.search(r"^rcut_max .*$", out, re.M).group(0).split())] * 2))
    return g, {k: float(v) for k, v in vals.items()}


def test_lj_deck_resolves_to_the_reference_step_sequence():
    g, v = graph_of(os.path.join(DECKS, "lj_single_specy_nosym.msp"))
    assert v["rcut_max"] == 8.0 and v["nbh_dist"] == 9.0 and v["max_displ"] == 0.5 and v["dt"] == 1e-3
    s = " ".join(g)
    # verlet_nve body (config_numerical_schemes.msp:44-52) with compute_force bound to lj_compute_force
    assert "push_f_v_r push_f_v particle_displ_over" in s
    assert "zero_force_energy lj_compute_force force_to_accel push_f_v" in s
    # full update path (config_move_particles.msp:89-125)
    assert re.search(r"move_particles .*migrate_cell_particles .*backup_r ghost_comm_scheme .*ghost_update_all_no_fv .*chunk_neighbors", s)
    assert s.index("chunk_neighbors") < s.index("lj_compute_force", s.index("chunk_neighbors"))
    assert g[-2:] == ["dump_particles", "finalize_cuda"]


def test_eam_deck_rebind_and_three_phase_graph():
    g, v = graph_of(os.path.join(DECKS, "eam_alloy_nosym.msp"))
    s = " ".join(g)
    assert v["rcut_max"] == 6.0 and v["species"] == 2
    assert "eam_alloy_init" in g
    assert "zero_force_energy eam_alloy_force eam_alloy_force ghost_update_opt eam_alloy_force force_to_accel" in s


def test_set_overrides_and_slot_validation(tmp_path):
    deck = os.path.join(DECKS, "lj_single_specy_nosym.msp")
    _, v = graph_of(deck, "--set", "lj_compute_force.rcut", "6.5 ang", "--set", "global.rcut_inc", "0.5 ang")
    assert v["rcut_max"] == 6.5 and v["nbh_dist"] == 7.0 and v["max_displ"] == 0.25
    # REQUIRED slot missing (pair_potential_impl.hxx:107 rcut is REQUIRED) -> fatal, non-zero exit like fatal_error()
    (tmp_path / "d.msp").write_text("lj_compute_force:\n  parameters: { epsilon: 0.0104 eV , sigma: 3.4 ang }\ncompute_force: lj_compute_force\n")
    p = run(tmp_path / "d.msp", "--dry-run", check=False)
    assert p.returncode != 0 and "required slot 'rcut'" in p.stderr
    # unknown slot
    (tmp_path / "e.msp").write_text("lj_compute_force:\n  parameters: { epsilon: 0.0104 eV , sigma: 3.4 ang }\n  rcut: 8 ang\n  rcutt: 1\ncompute_force: lj_compute_force\n")
    p = run(tmp_path / "e.msp", "--dry-run", check=False)
    assert p.returncode != 0 and "no slot named 'rcutt'" in p.stderr
    # unknown operator
    (tmp_path / "f.msp").write_text("compute_force: does_not_exist\n")
    p = run(tmp_path / "f.msp", "--dry-run", check=False)
    assert p.returncode != 0 and "unknown operator 'does_not_exist'" in p.stderr
    # chunk_size must be a power of two (type_pair_rcut_neighbors.cpp:90-104): checked when the operator runs, so only the
    # deck-level type check is visible on CPU
    (tmp_path / "g.msp").write_text("compute_force: [ a, b ]\na: b\nb: a\n")
    p = run(tmp_path / "g.msp", "--dry-run", check=False)
    assert p.returncode != 0 and "recursion" in p.stderr


def test_no_cpu_fallback_in_the_host_layer():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = run(os.path.join(DECKS, "lj_single_specy_nosym.msp"), check=False)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/data/regression_new"), reason="reference tree not present")
@pytest.mark.parametrize("deck,rcut_max", [
    ("pair/lj/single_specy_nosym.msp", 8.0), ("pair/lj/multi_species_nosym.msp", 8.0), ("pair/lj/single_specy_sym.msp", 8.0),
    ("eam/eam_alloy/single_specy_nosym_cs1.msp", 5.3), ("eam/eam_alloy/multi_species_nosym_cs4.msp", 6.6825),
    ("eam/eam_alloy/multi_species_singlepass_cs1.msp", 6.6825), ("eam/eam_alloy/multi_species_sym_cs1.msp", 6.6825),
    ("eam/eam_alloy/benchmark_Al_Cu.msp", 6.6825), ("eam/eam_johnson/single_specy.msp", 6.1), ("snap/multi_WBe.msp", 4.8123),
    ("eam/eam_sutton_chen/single_specy.msp", 7.29), ("eam/eam_vniitf/single_specy.msp", 5.599), ("pair/yukawa/single_specy_nosym.msp", 10.0),
    ("pair/yukawa/multi_species_nosym.msp", 8.0), ("pair/zero/single_specy_nosym.msp", 8.0), ("pair/yukawa/single_specy_sym.msp", 8.0),
    ("pair/buckingham/single_specy_sym.msp", None), ("pair/exp6/single_specy_sym.msp", None), ("pair/relax/single_specy_sym.msp", None),
    ("snap/multi_WBe_fp32.msp", 4.8123)])
def test_unmodified_reference_decks_resolve(deck, rcut_max):
    """the reference's own regression decks build a graph here: same names, same slots, same layering"""
    g, v = graph_of(os.path.join("/root/reference/data/regression_new/potentials", deck), "--data-dir", "/root/reference/data/config")
    if rcut_max is not None:
        assert v["rcut_max"] == pytest.approx(rcut_max, rel=1e-15)
    assert "chunk_neighbors" in g and "force_to_accel" in g
    if "multi_WBe_fp32" in deck:   # the reference's SNAP_FP32_MATH variant: same graph with snap_force_fp32 (-> XSB_FLAG_MIXED)
        assert " ".join(g).count("zero_force_energy zbl_multi_force snap_force_fp32 update_force_energy_from_ghost force_to_accel") >= 2
    elif "multi_WBe" in deck:   # zbl_multi_force + snap_force with the reference's own WBe_Wood_PRB2019 files, symmetric-force epilog
        assert " ".join(g).count("zero_force_energy zbl_multi_force snap_force update_force_energy_from_ghost force_to_accel") >= 2


# ------------------------------------------------------------------------------------------------ GPU: decks against the oracle
def read_dump(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"XSBDUMP1"
    n, nf = np.frombuffer(raw, dtype=np.uint64, count=2, offset=8)
    n = int(n)
    meta = np.frombuffer(raw, dtype=np.float64, count=16, offset=24)
    o = 24 + 128
    d = {"bounds_min": meta[0:3], "bounds_max": meta[3:6], "cell_size": meta[6], "xform": meta[7:16].reshape(3, 3)}
    d["id"] = np.frombuffer(raw, dtype=np.uint64, count=n, offset=o); o += 8 * n
    d["type"] = np.frombuffer(raw, dtype=np.uint8, count=n, offset=o); o += n
    for name in ["rx", "ry", "rz", "vx", "vy", "vz", "ax", "ay", "az", "ep", "rho_dEmb"]:
        d[name] = np.frombuffer(raw, dtype=np.float64, count=n, offset=o); o += 8 * n
    assert o == len(raw)
    return d


def thermo_lines(stdout):
    rows = []
    for line in stdout.splitlines():
        c = line.split()
        if len(c) == 8 and re.fullmatch(r"\d+", c[0]):
            rows.append([float(x) for x in c])
    return np.array(rows)


def oracle_system(d, nbh_dist):
    from oracle import oracle as O
    box = d["bounds_max"] - d["bounds_min"]
    pos = np.stack([d["rx"], d["ry"], d["rz"]], axis=1) - d["bounds_min"]
    pos = np.mod(pos, box); pos = np.where(pos >= box, 0.0, pos)
    scale = np.linalg.norm(d["xform"], axis=0).min()        # narrowest physical extent of a grid-space cell
    gs = GridSystem(pos, d["type"], box, d["cell_size"], int(np.ceil(nbh_dist / (d["cell_size"] * scale) - 1e-12)), xform=d["xform"])
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nbh_dist, 1, True)
    return O, gs, g, nb


def compare(d, gs, masses, fx, fy, fz, ep):
    own = ~gs.is_ghost
    m = np.asarray(masses)[d["type"]]
    ref = {k: np.zeros(len(m)) for k in "xyze"}
    for k, a in zip("xyze", (fx, fy, fz, ep)):
        ref[k][gs.src_index[own]] = a[own]
    fmax = max(np.abs(ref[k]).max() for k in "xyz")
    for k, got in zip("xyz", (d["ax"] * m, d["ay"] * m, d["az"] * m)):
        assert np.abs(got - ref[k]).max() <= TOL * fmax
    assert np.abs(d["ep"] - ref["e"]).max() <= TOL * np.abs(ref["e"]).max()


@pytest.mark.gpu
def test_lj_deck_forces_match_oracle_and_nve_conserves_energy(tmp_path):
    p = run(os.path.join(DECKS, "lj_single_specy_nosym.msp"), "--set", "global.max_iteration", "0", cwd=tmp_path)
    d = read_dump(tmp_path / "lj_single.xsbdump")
    assert len(d["id"]) == 2048 and len(set(d["id"].tolist())) == 2048
    O, gs, g, nb = oracle_system(d, 9.0)
    fx, fy, fz, ep = [gs.zeros() for _ in range(4)]
    O.pair_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, [0.0104 * EV, 3.4], 8.0, 0, fx, fy, fz, ep, None)
    compare(d, gs, [39.948], fx, fy, fz, ep)
    t0 = thermo_lines(p.stdout)
    assert t0.shape[0] == 1 and t0[0, 0] == 0 and t0[0, 7] == 2048
    own = ~gs.is_ghost
    assert t0[0, 4] == pytest.approx(ep[own].sum() / EV, rel=1e-10)                      # Pot. E. column
    ke = 0.5 * 39.948 * (d["vx"] ** 2 + d["vy"] ** 2 + d["vz"] ** 2).sum()
    assert t0[0, 3] == pytest.approx(ke / EV, rel=1e-10)
    # 40 NVE steps at dt = 1 fs: total energy drift of velocity Verlet stays far below the kinetic energy
    p = run(os.path.join(DECKS, "lj_single_specy_nosym.msp"), "--set", "global.max_iteration", "40", "--trace", cwd=tmp_path)
    t = thermo_lines(p.stdout)
    assert list(t[:, 0]) == [0, 5, 10, 15, 20, 25, 30, 35, 40]
    # the noisy compressed lattice starts far from equilibrium: tens of eV move from potential to kinetic energy in the
    # first 50 fs.  Velocity Verlet keeps the total within O((w dt)^2) of that exchange, and halving dt divides the
    # error by ~4 (second order), which is what pins the integrator + force + thermo chain
    exch = np.abs(t[:, 3] - t[0, 3]).max()
    err1 = np.abs(t[:, 2] - t[0, 2]).max()
    assert exch > 1.0 and err1 < 5e-4 * exch
    p2 = run(os.path.join(DECKS, "lj_single_specy_nosym.msp"), "--set", "global.max_iteration", "80", "--set", "global.dt", "0.5e-3 ps",
             "--set", "global.simulation_thermostate_screen_frequency", "10", cwd=tmp_path)
    t2 = thermo_lines(p2.stdout)
    assert list(t2[:, 0]) == list(range(0, 81, 10))
    err2 = np.abs(t2[:, 2] - t2[0, 2]).max()
    assert 3.0 < err1 / err2 < 5.0, (err1, err2)
    assert abs(t[-1, 4] - t[0, 4]) > 1e-6                                                # the system did move
    tr = re.search(r"^trace: (.*)$", p.stdout, re.M).group(1).split()
    assert tr.count("lj_compute_force") == 42                                             # pre-pass + first iteration + 40 steps
    assert tr.count("ghost_update_r") + tr.count("chunk_neighbors") - 1 == 40            # every step: fast update or full rebuild


@pytest.mark.gpu
def test_lj_multi_deck_matches_oracle(tmp_path):
    run(os.path.join(DECKS, "lj_multi_species_nosym.msp"), cwd=tmp_path)
    d = read_dump(tmp_path / "lj_multi.xsbdump")
    O, gs, g, nb = oracle_system(d, 9.0)
    J = 1.0 / 1.602176634e-19 * EV
    rows = []                                        # unique_pair_id order (0,0), (0,1), (1,1); Cu = 0, Zn = 1
    for e, s, rc in ((9.340e-20 * J, 2.27, 5.68), (4.853e-20 * J, 2.36, 5.89), (2.522e-20 * J, 2.44, 6.10)):
        q6 = (s / rc) ** 6
        rows.append([e, s, rc, 4 * e * (q6 * q6 - q6)])
    fx, fy, fz, ep = [gs.zeros() for _ in range(4)]
    O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, np.array(rows), 8.0, 0, fx, fy, fz, ep, None)
    compare(d, gs, [63.546, 65.38], fx, fy, fz, ep)


@pytest.mark.gpu
def test_eam_alloy_deck_matches_oracle_and_conserves_energy(tmp_path):
    setfl = write_setfl(str(tmp_path / "synthetic_CuXx.eam.alloy"), [SC_CU, SC_XX], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    deck = os.path.join(DECKS, "eam_alloy_nosym.msp")
    run(deck, "--data-dir", tmp_path, "--set", "global.max_iteration", "0", cwd=tmp_path)
    d = read_dump(tmp_path / "eam_alloy.xsbdump")
    O, gs, g, nb = oracle_system(d, 7.0)
    eam = O.EamAlloy(setfl)
    fx, fy, fz, ep, emb = [gs.zeros() for _ in range(5)]
    own = ~gs.is_ghost
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, 6.0, 1 | 2 | 16, fx, fy, fz, ep, None, emb)
    owner_of = np.zeros(len(d["id"]), dtype=np.int64); owner_of[gs.src_index[own]] = np.nonzero(own)[0]
    emb[:] = emb[owner_of[gs.src_index]]
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, 6.0, 8 | 16, fx, fy, fz, ep, None, emb)
    compare(d, gs, [63.546, 26.982], fx, fy, fz, ep)
    ref_emb = np.zeros(len(d["id"])); ref_emb[gs.src_index[own]] = emb[own]
    assert np.abs(d["rho_dEmb"] - ref_emb).max() <= TOL * np.abs(ref_emb).max()
    p = run(deck, "--data-dir", tmp_path, cwd=tmp_path)
    t = thermo_lines(p.stdout)
    assert list(t[:, 0]) == [0, 5, 10, 15, 20]
    # the synthetic Sutton-Chen tables are truncated at rc without smoothing (rho and phi jump there), so pairs crossing
    # the cutoff leak a little energy: bound the error by 1 % of the potential <-> kinetic exchange
    assert np.abs(t[:, 2] - t[0, 2]).max() < 1e-2 * max(1.0, np.abs(t[:, 3] - t[0, 3]).max())


@pytest.mark.gpu
def test_eam_alloy_lj_chain_deck_is_fused_and_matches_oracle(tmp_path):
    """compute_force: [eam_rho, eam_rho2emb, ghost_update_opt, eam_force, lj_multi_force] (configs[4]): the deck runs the
    operators one by one, the library evaluates lj_multi_force inside the EAM force pass -- forces and energies are the
    oracle's sum of both operators, with and without XSB_NO_CHAIN_FUSION"""
    setfl = write_setfl(str(tmp_path / "synthetic_CuXx.eam.alloy"), [SC_CU, SC_XX], nrho=2000, drho=0.1, nr=2000, rc=6.0)
    deck = os.path.join(DECKS, "eam_alloy_lj_chain.msp")
    p = run(deck, "--data-dir", tmp_path, "--trace", cwd=tmp_path)
    assert int(re.search(r"^fused_pair_operators: (\d+)$", p.stdout, re.M).group(1)) >= 1
    d = read_dump(tmp_path / "eam_alloy_lj.xsbdump")
    O, gs, g, nb = oracle_system(d, 7.0)
    eam = O.EamAlloy(setfl)
    fx, fy, fz, ep, emb = [gs.zeros() for _ in range(5)]
    own = ~gs.is_ghost
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, 6.0, 1 | 2 | 16, fx, fy, fz, ep, None, emb)
    owner_of = np.zeros(len(d["id"]), dtype=np.int64); owner_of[gs.src_index[own]] = np.nonzero(own)[0]
    emb[:] = emb[owner_of[gs.src_index]]
    O.eam_alloy(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, eam, 6.0, 8 | 16, fx, fy, fz, ep, None, emb)
    J = 1.0 / 1.602176634e-19 * EV
    rows = []                                        # unique_pair_id order (0,0), (0,1), (1,1); Cu = 0, Xx = 1; (0,0) <- common_parameters
    for e, s, rc in ((0.0, 0.0, 5.9), (4.853e-20 * J, 2.36, 5.50), (2.522e-20 * J, 2.44, 5.90)):
        q6 = (s / rc) ** 6
        rows.append([e, s, rc, 4 * e * (q6 * q6 - q6)])
    O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, np.array(rows), 5.9, 0, fx, fy, fz, ep, None)
    compare(d, gs, [63.546, 26.982], fx, fy, fz, ep)
    env = dict(os.environ, XSB_NO_CHAIN_FUSION="1")
    q = subprocess.run([RUN, deck, "--data-dir", str(tmp_path), "--trace"], capture_output=True, text=True, cwd=tmp_path, timeout=600, env=env)
    assert q.returncode == 0 and re.search(r"^fused_pair_operators: 0$", q.stdout, re.M)
    compare(read_dump(tmp_path / "eam_alloy_lj.xsbdump"), gs, [63.546, 26.982], fx, fy, fz, ep)


@pytest.mark.gpu
def test_johnson_deck_matches_oracle(tmp_path):
    run(os.path.join(DECKS, "johnson_single_specy.msp"), cwd=tmp_path)
    d = read_dump(tmp_path / "johnson.xsbdump")
    from oracle import oracle as O
    # single-species EAM computes the embedding of ghost atoms itself (ComputeGhostEmb): needs ghosts out to 2 rcut + skin,
    # so the oracle runs on a grid with 4 ghost layers of 3.615 ang
    box = d["bounds_max"] - d["bounds_min"]
    pos = np.mod(np.stack([d["rx"], d["ry"], d["rz"]], axis=1) - d["bounds_min"], box)
    gs = GridSystem(pos, d["type"], box, d["cell_size"], 4)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 7.0, 1, True)
    fx, fy, fz, ep, emb = [gs.zeros() for _ in range(5)]
    O.eam_johnson(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, johnson_params(JOHNSON_CU), 6.0, 7, fx, fy, fz, ep, None, emb)
    compare(d, gs, [63.546], fx, fy, fz, ep)


@pytest.mark.gpu
def test_sutton_chen_deck_matches_oracle(tmp_path):
    """the parameters of the reference's own regression deck (eam_sutton_chen/single_specy.msp) through the deck layer:
    unit conversion (J, m), sutton_chen_force on the GPU, oracle restatement of the same three functions"""
    run(os.path.join(DECKS, "sutton_chen_single_specy.msp"), cwd=tmp_path)
    d = read_dump(tmp_path / "sutton_chen.xsbdump")
    from oracle import oracle as O
    box = d["bounds_max"] - d["bounds_min"]
    pos = np.mod(np.stack([d["rx"], d["ry"], d["rz"]], axis=1) - d["bounds_min"], box)
    gs = GridSystem(pos, d["type"], box, d["cell_size"], 5)          # ghosts out to 2 rcut + skin (ComputeGhostEmb)
    g = gs.oracle_grid()
    nb = O.Neighbors.build(g, gs.cell_off, gs.rx, gs.ry, gs.rz, 8.29, 1, True)
    fx, fy, fz, ep, emb = [gs.zeros() for _ in range(5)]
    prm = [3.317e1, 3.605e-21 / 1.602176634e-19 * EV, 3.27, 9.05, 5.005]
    O.eam_analytic(g, gs.cell_off, gs.rx, gs.ry, gs.rz, nb, O.EAM_SUTTON_CHEN, prm, 7.29, 7, fx, fy, fz, ep, None, emb)
    compare(d, gs, [63.546], fx, fy, fz, ep)


def write_snap_files(tmp_path, twoj, ncoef, seed=1):
    beta = np.random.default_rng(seed).normal(0, 1, ncoef + 1) * 1e-2
    (tmp_path / "synthetic_Ta.snapparam").write_text("# synthetic\nrcutfac 4.7\ntwojmax %d\nrfac0 0.99363\nrmin0 0\nbzeroflag 0\nquadraticflag 0\n" % twoj)
    (tmp_path / "synthetic_Ta.snapcoeff").write_text("# synthetic\n1 %d\nTa 0.5 1\n%s\n" % (ncoef + 1, "\n".join("%.17g" % b for b in beta)))
    return beta


@pytest.mark.gpu
def test_snap_deck_matches_oracle(tmp_path):
    import exastamp_b200 as xsb
    twoj = 6
    ncoef = xsb.load_library().xsb_snap_ncoeff(twoj)
    beta = write_snap_files(tmp_path, twoj, ncoef)
    run(os.path.join(DECKS, "snap_monomat.msp"), "--data-dir", tmp_path, cwd=tmp_path)
    d = read_dump(tmp_path / "snap.xsbdump")
    O, gs, g, nb = oracle_system(d, 5.7)
    S = O.Snap(twoj, 4.7, [0.5], [1.0], (beta * EV).reshape(1, -1), bzeroflag=0)
    fx, fy, fz, ep = [gs.zeros() for _ in range(4)]
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2, fx, fy, fz, ep, None)
    # Newton-on: forces written on ghost copies fold back onto their owners (update_force_energy_from_ghost)
    n = len(d["id"])
    tot = [np.zeros(n) for _ in range(3)]
    for t, a in zip(tot, (fx, fy, fz)):
        np.add.at(t, gs.src_index, a)
    own = ~gs.is_ghost
    e = np.zeros(n); e[gs.src_index[own]] = ep[own]
    fmax = max(np.abs(t).max() for t in tot)
    for got, ref in zip((d["ax"], d["ay"], d["az"]), tot):
        assert np.abs(got * 180.95 - ref).max() <= TOL * fmax
    assert np.abs(d["ep"] - e).max() <= TOL * np.abs(e).max()


@pytest.mark.gpu
def test_deck_failure_behaviour_on_gpu(tmp_path):
    (tmp_path / "bad.msp").write_text(open(os.path.join(DECKS, "lj_single_specy_nosym.msp")).read() + "\nchunk_neighbors:\n  config: { chunk_size: 3 }\n")
    p = run(tmp_path / "bad.msp", cwd=tmp_path, check=False)
    assert p.returncode != 0 and "power of two" in p.stderr


@pytest.mark.gpu
def test_snap_zbl_two_species_deck_matches_oracle(tmp_path):
    """the operator pairing of the reference's multi_WBe.msp: zbl_multi_force + snap_force, two elements, bzeroflag 1"""
    import exastamp_b200 as xsb
    twoj = 4
    ncoef = xsb.load_library().xsb_snap_ncoeff(twoj)
    rng = np.random.default_rng(7)
    beta = rng.normal(0, 1, (2, ncoef + 1)) * 1e-2
    (tmp_path / "synthetic_WBe.snapparam").write_text("rcutfac 4.8123\ntwojmax %d\nrfac0 0.99363\nrmin0 0\nbzeroflag 1\nquadraticflag 0\n" % twoj)
    (tmp_path / "synthetic_WBe.snapcoeff").write_text("# synthetic\n2 %d\nW 0.5 1\n%s\nBe 0.417932 0.959049\n%s\n" % (
        ncoef + 1, "\n".join("%.17g" % b for b in beta[0]), "\n".join("%.17g" % b for b in beta[1])))
    run(os.path.join(DECKS, "snap_zbl_multi.msp"), "--data-dir", tmp_path, cwd=tmp_path)
    d = read_dump(tmp_path / "snap_zbl.xsbdump")
    rad, wj = [0.5, 0.417932], [1.0, 0.959049]
    rc_snap = 2 * 0.5 * 4.8123
    O, gs, g, nb = oracle_system(d, rc_snap + 1.0)
    S = O.Snap(twoj, 4.8123, rad, wj, beta * EV, bzeroflag=1)
    fx, fy, fz, ep = [gs.zeros() for _ in range(4)]
    z = [74, 4]
    rows = []
    for hi in range(2):
        for lo in range(hi + 1):
            prm = [4.0, 4.8, z[lo], z[hi]]
            rows.append(prm + [4.8, O.pair_ecut(1, prm, 4.8)])
    O.pair_multi_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, np.array(rows), 4.8, 0, fx, fy, fz, ep, None, pot=1)
    O.snap_force(g, gs.cell_off, gs.rx, gs.ry, gs.rz, gs.type, nb, S, 2, fx, fy, fz, ep, None)
    n = len(d["id"])
    tot = [np.zeros(n) for _ in range(3)]
    for t, a in zip(tot, (fx, fy, fz)):
        np.add.at(t, gs.src_index, a)
    own = ~gs.is_ghost
    e = np.zeros(n); e[gs.src_index[own]] = ep[own]
    m = np.array([183.84, 9.012182])[d["type"]]
    fmax = max(np.abs(t).max() for t in tot)
    for got, ref in zip((d["ax"], d["ay"], d["az"]), tot):
        assert np.abs(got * m - ref).max() <= TOL * fmax
    assert np.abs(d["ep"] - e).max() <= TOL * np.abs(e).max()

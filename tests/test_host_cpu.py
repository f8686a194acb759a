"""Host plugin layer without a GPU: factory / registration / settings behaviour of the
MultiConfigurationCalculator API (reference: cpp/tests/test_mc.cpp factory + metadata cases,
cpp/tests/test_macis.cpp settings round trips, algorithm.hpp:262-372)."""
import os

import numpy as np
import pytest

from qdk_chemistry_b200 import algorithms as alg
from qdk_chemistry_b200 import data

MC = "multi_configuration_calculator"
PMC = "projected_multi_configuration_calculator"


def test_factory_contents_and_defaults():
    assert alg.available(MC) == ["b200_asci", "b200_cas"]
    assert alg.available(PMC) == ["b200_pmc"]
    assert alg.show_default(MC) == "b200_cas"
    F = alg.MultiConfigurationCalculatorFactory
    assert F.algorithm_type_name() == MC and F.has("b200_cas") and not F.has("macis_cas")
    c = F.create("")  # empty name -> default algorithm (algorithm.hpp:262-266)
    assert c.name() == "b200_cas" and c.type_name() == MC and c.aliases() == ["b200_cas"]
    with pytest.raises(RuntimeError, match="not found in registry, available options are"):
        F.create("nope")
    with pytest.raises(KeyError):
        alg.create("scf_solver")


def test_reference_names_resolve_to_drop_ins():
    assert alg.create(MC, "macis_cas").name() == "b200_cas"
    assert alg.create(MC, "macis_asci").name() == "b200_asci"
    assert alg.create(PMC, "macis_pmc").name() == "b200_pmc"


def test_settings_defaults_match_the_reference():
    s = alg.create(MC, "b200_cas").settings()
    assert s.get("ci_residual_tolerance") == 1e-6          # mc.hpp:47-49
    assert s.get("max_solver_iterations") == 200
    assert s.get("iterative_solver_dimension_cutoff") == 2000
    assert s.get("ci_matel_tol") == np.finfo(np.float64).eps  # mcscf.hpp:51
    assert s.get("calculate_one_rdm") is False
    a = alg.create(MC, "b200_asci").settings()
    want = dict(ntdets_max=100000, ntdets_min=100, ncdets_max=100, search_matel_tol=1e-8,
                rv_prune_tol=1e-8, pair_size_max=500000000, grow_factor=8.0, min_grow_factor=1.01,
                growth_backoff_rate=0.5, growth_recovery_rate=1.1, max_refine_iter=6,
                refine_energy_tol=1e-6, warm_start_davidson=True, min_warm_start_overlap=0.5,
                min_patch_overlap=0.3, grow_ci_residual_tolerance=0.0, taper_grow_factor=0.0,
                constraint_level=2, hamiltonian_build_algorithm="", just_singles=False,
                core_selection_strategy="percentage", core_selection_threshold=0.95)
    for k, v in want.items():                               # determinant_search.hpp:95-199
        assert a.get(k) == v, k
    assert a.has_description("ntdets_max") and a.get_type_name("grow_factor") == "double"


def test_settings_errors():
    s = alg.create(MC, "b200_asci").settings()
    with pytest.raises(data.SettingNotFound):
        s.set("no_such_key", 1)
    with pytest.raises(data.SettingNotFound):
        s.get("no_such_key")
    with pytest.raises(data.SettingTypeMismatch):
        s.set("ntdets_max", "many")
    with pytest.raises(ValueError):
        s.set("ntdets_max", 0)                      # BoundConstraint{1, max}
    with pytest.raises(ValueError):
        s.set("core_selection_strategy", "random")  # ListConstraint
    s.set("grow_factor", 2)                         # int accepted for a double setting
    assert s.get("grow_factor") == 2.0
    s.update({"ntdets_max": 2000, "core_selection_strategy": "fixed"})
    assert s["ntdets_max"] == 2000 and "ntdets_max" in s
    s.lock()
    with pytest.raises(data.SettingsAreLocked):
        s.set("ntdets_max", 10)


def test_create_forwards_kwargs_and_hash_changes_with_settings():
    h = data.Hamiltonian(np.eye(2), np.zeros(16), 0.25)
    c1 = alg.create(MC, "b200_cas")
    c2 = alg.create(MC, "b200_cas", ci_residual_tolerance=1e-9)
    assert c2.settings().get("ci_residual_tolerance") == 1e-9
    assert c1.hash(h, 1, 1) != c2.hash(h, 1, 1)
    assert c1.hash(h, 1, 1) == alg.create(MC, "b200_cas").hash(h, 1, 1)
    assert c1.hash(h, 1, 1) != c1.hash(h, 1, 0)


def test_python_subclass_registration_through_the_trampoline():
    class Fixed(alg.MultiConfigurationCalculator):
        def __init__(self):
            super().__init__()

        def name(self):
            return "fixed_energy"

        def _run_impl(self, hamiltonian, na, nb):
            wfn = alg.create(MC, "b200_cas")  # any object; the energy is what is checked
            return -1.25 + hamiltonian.get_core_energy(), None

    alg.register(lambda: Fixed())
    try:
        assert "fixed_energy" in alg.available(MC)
        with pytest.raises(alg.DuplicateRegistrationError):
            alg.register(lambda: Fixed())
        c = alg.create(MC, "fixed_energy")
        E, w = c.run(data.Hamiltonian(np.eye(2), np.zeros(16), 0.25), 1, 1)
        assert E == -1.0 and w is None
    finally:
        alg.unregister(MC, "fixed_energy")
    assert "fixed_energy" not in alg.available(MC)
    with pytest.raises(KeyError):
        alg.unregister(MC, "fixed_energy")


def test_data_stand_ins():
    c = data.Configuration("2ud0")
    assert (c.alpha_word(), c.beta_word()) == (0b0011, 0b0101) and c.to_string() == "2ud0"
    assert c.get_n_electrons() == (2, 2) and c == data.Configuration(3, 5, 4)
    with pytest.raises(ValueError):
        data.Configuration("2x")
    with pytest.raises(ValueError):
        data.Hamiltonian(np.eye(3), np.zeros(16), 0.0)
    h = data.Hamiltonian(np.arange(4.0).reshape(2, 2), np.arange(16.0), -2.0)
    assert h.num_active_orbitals() == 2 and h.get_core_energy() == -2.0
    assert np.array_equal(h.get_two_body_integrals(), np.arange(16.0))


def test_run_without_a_gpu_fails_loudly_and_locks_settings():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    c = alg.create(MC, "b200_cas")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.run(data.Hamiltonian(np.eye(2), np.zeros(16), 0.0), 1, 1)
    with pytest.raises(data.SettingsAreLocked):   # run() locks before _run_impl (algorithm.hpp:67-70)
        c.settings().set("ci_residual_tolerance", 1e-9)
    unres = data.Hamiltonian(np.eye(2), np.zeros(16), 0.0, True)
    with pytest.raises(RuntimeError, match="does not support unrestricted orbitals"):
        alg.create(MC, "b200_asci").run(unres, 1, 1)
    # every hamiltonian_build_algorithm of the reference is accepted (macis_asci.hpp:174-179)
    for algo in ("", "sorted_double_loop", "residue_arrays", "dynamic_bit_masking"):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            alg.create(MC, "b200_asci", hamiltonian_build_algorithm=algo, grow_with_rot=True).run(
                data.Hamiltonian(np.eye(2), np.zeros(16), 0.0), 1, 1)
    with pytest.raises(RuntimeError, match="grow_factor must be > 1.0"):
        alg.create(MC, "b200_asci", grow_factor=1.0).run(
            data.Hamiltonian(np.eye(2), np.zeros(16), 0.0), 1, 1)


# ---- FCIDUMP / RDM file formats in the C++ host layer (macis/util/fcidump.hpp) -------------------
def _core_io():
    from qdk_chemistry_b200 import _core
    return _core.io


def test_fcidump_cpp_round_trip_and_reference_known_answers(tmp_path):
    # external/macis/tests/fcidump.cxx:26-61: header fields, core energy and the integral sums of
    # the water / cc-pVDZ file; here the file is rewritten from the committed integrals
    import os
    from qdk_chemistry_b200 import workloads as W
    io = _core_io()
    water = W.load_sparse_npz(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h2o_ccpvdz.ints.npz"))
    n = water.norb
    f = str(tmp_path / "water.fcidump")
    io.write_fcidump(f, n, 10, 0, np.asarray(water.T).reshape(n, n), np.asarray(water.V).reshape(-1), water.core_energy)
    h = io.fcidump_read_header(f)
    assert (h["norb"], h["nelec"], h["ms2"], h["isym"]) == (24, 10, 0, 1) and h["orbsym"] == [1] * 24
    assert io.read_fcidump_norb(f) == 24
    T, V, core = io.read_fcidump_all(f)
    assert abs(core - 9.191200742618042) < 1e-10 and abs(io.read_fcidump_core(f) - core) == 0
    assert abs(T.sum() - (-1.095432762653e+02)) < 1e-9
    assert abs(V.sum() - 2.701609068389e+02) < 1e-9
    # %25.14e keeps 15 significant digits
    assert np.allclose(T, np.ravel(water.T), rtol=1e-14, atol=1e-15)
    assert np.allclose(V, np.ravel(water.V), rtol=1e-14, atol=1e-15)
    assert np.array_equal(io.read_fcidump_1body(f), T) and np.array_equal(io.read_fcidump_2body(f), V)
    # the Python reader of the workload module and the C++ reader agree
    sp = W.read_fcidump(f)
    assert sp.norb == 24 and sp.nalpha == 5 and sp.nbeta == 5
    assert np.array_equal(np.ravel(sp.T), T) and np.array_equal(np.ravel(sp.V), V)


def test_fcidump_cpp_both_layouts_and_errors(tmp_path):
    io = _core_io()
    body_first = ["0.5 1 1 1 1", "0.25 2 1 1 1", "-1.5 1 1 0 0", "0.125 2 1 0 0", "3.0 0 0 0 0"]
    idx_first = ["1 1 1 1 0.5", "2 1 1 1 0.25", "1 1 0 0 -1.5", "2 1 0 0 0.125", "0 0 0 0 3.0"]
    out = []
    for name, lines in (("a", body_first), ("b", idx_first)):
        f = str(tmp_path / name)
        with open(f, "w") as fh:
            fh.write("&FCI NORB = 2; NELEC = 2; MS2 = 0,\n  ORBSYM=1,1,\n  ISYM=1,\n&END\n\n" + "\n".join(lines) + "\n")
        out.append(io.read_fcidump_all(f))
    (T1, V1, c1), (T2, V2, c2) = out
    assert np.array_equal(T1, T2) and np.array_equal(V1, V2) and c1 == c2 == 3.0
    assert T1.tolist() == [-1.5, 0.125, 0.125, 0.0]
    V = V1.reshape(2, 2, 2, 2, order="F")
    assert V[0, 0, 0, 0] == 0.5
    # (21|11) under its eight permutations
    for idx in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1)):
        assert V[idx] == 0.25
    with pytest.raises(RuntimeError, match="No FCIDUMP header"):
        f = str(tmp_path / "c")
        open(f, "w").write("1 1 1 1 0.5\n")
        io.fcidump_read_header(f)
    with pytest.raises(RuntimeError, match="Could not open"):
        io.read_fcidump_core(str(tmp_path / "missing"))


def test_rdm_binary_round_trip(tmp_path):
    io = _core_io()
    rng = np.random.default_rng(0)
    o, t = rng.normal(size=9), rng.normal(size=81)
    f = str(tmp_path / "rdm.bin")
    io.write_rdms_binary(f, 3, o, t)
    o2, t2 = io.read_rdms_binary(f, 3)
    assert np.array_equal(o, o2) and np.array_equal(t, t2)
    with pytest.raises(RuntimeError, match="doesn't match"):
        io.read_rdms_binary(f, 4)


@pytest.mark.skipif(not os.path.exists("/root/reference/external/macis/tests/ref_data/h2o.ccpvdz.fci.dat"),
                    reason="reference tree not present (build container only)")
def test_fcidump_cpp_reads_the_reference_fixture():
    from qdk_chemistry_b200 import workloads as W
    io = _core_io()
    f = "/root/reference/external/macis/tests/ref_data/h2o.ccpvdz.fci.dat"
    h = io.fcidump_read_header(f)
    assert (h["norb"], h["nelec"], h["ms2"], h["isym"]) == (24, 10, 0, 1) and h["orbsym"] == [1] * 24
    T, V, core = io.read_fcidump_all(f)
    assert abs(core - 9.191200742618042) < 1e-10
    assert abs(T.sum() - (-1.095432762653e+02)) < 1e-9 and abs(V.sum() - 2.701609068389e+02) < 1e-9
    water = W.load_sparse_npz(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h2o_ccpvdz.ints.npz"))
    assert np.array_equal(T, np.ravel(water.T)) and np.array_equal(V, np.ravel(water.V))


def test_core_selection_equals_stable_full_sort():
    """asci_iter's core set (iteration.hpp:62-100) is found by selection instead of the reference's
    full sort; the prefix and its order must be those of a stable sort by |c| descending."""
    from qdk_chemistry_b200 import _core
    sel = _core.algorithms.select_core_indices
    rng = np.random.default_rng(11)
    for n in (1, 2, 37, 5000, 70000):
        X = rng.normal(size=n) * np.exp(-rng.random(n) * 12)
        X[rng.integers(0, n, size=n // 5)] = 0.125          # many exact ties
        X /= np.linalg.norm(X)
        order = np.argsort(-np.abs(X), kind="stable")
        for k in (1, 100, 4096, 10 ** 6):
            assert np.array_equal(sel(X, True, k, 0.0), order[:min(k, n)])
        for thr in (1e-3, 0.5, 0.95, 0.999999, 1.0):
            w = np.cumsum(X[order] ** 2)
            reach = np.nonzero(w >= thr)[0]
            nkeep = (reach[0] + 1) if len(reach) else n
            assert np.array_equal(sel(X, False, 0, thr), order[:nkeep])


def test_wavefunction_text_io_cpp_matches_python_and_fixture(tmp_path):
    """macis/wavefunction_io.hpp in the C++ host layer: the reference's o2.wfn.dat fixture read, written
    and re-read without loss, byte-identical to what the Python module writes."""
    from qdk_chemistry_b200 import wavefunction_io as wio
    io = _core_io()
    fixture = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "o2.wfn.dat")
    a, b, c, meta = io.read_wavefunction(fixture)
    pa, pb, pc, pmeta = wio.read_wavefunction(fixture)
    assert tuple(meta) == pmeta == (120, 6, 5, 3)
    assert np.array_equal(a, pa) and np.array_equal(b, pb) and np.array_equal(c, pc)
    assert io.to_canonical_string(0b000111, 0b001011, 6) == "22ud00"
    assert io.from_canonical_string("222uu0") == (0b011111, 0b000111)
    f1, f2 = str(tmp_path / "cpp.dat"), str(tmp_path / "py.dat")
    io.write_wavefunction(f1, 6, a.tolist(), b.tolist(), c.tolist())
    wio.write_wavefunction(f2, 6, pa, pb, pc)
    assert open(f1).read() == open(f2).read()
    assert open(f1).read().splitlines()[1] == "       -6.8728389771404168e-09 222uu0 "
    with pytest.raises(RuntimeError, match="Invalid Wave Function Dimensions"):
        io.write_wavefunction(f1, 6, [1, 2], [1], [0.5, 0.5])


def test_hamiltonian_from_fcidump(tmp_path):
    from qdk_chemistry_b200 import workloads as W
    io = _core_io()
    sp = W.config("tiny_cas6")
    f = str(tmp_path / "tiny.fcidump")
    W.write_fcidump(f, sp)                       # the Python writer (unique integrals only)
    ham, na, nb = io.hamiltonian_from_fcidump(f)
    assert (ham.num_active_orbitals(), na, nb) == (6, 3, 3) and ham.get_core_energy() == sp.core_energy
    assert np.allclose(np.ravel(ham.get_one_body_integrals()), np.ravel(sp.T), rtol=0, atol=0)
    assert np.allclose(np.ravel(ham.get_two_body_integrals()), np.ravel(sp.V), rtol=0, atol=0)


def test_cis_and_cisd_spaces():
    """generate_cis_hilbert_space / generate_cisd_hilbert_space (sd_operations.hpp:60-383): sizes of the
    reference's water test (csr_hamiltonian.cxx:44-47,76: 12636 determinants), generation order, and the
    same set as the independent enumeration the parity tests use."""
    from qdk_chemistry_b200 import _core
    from helpers import cisd_space
    hf = (1 << 5) - 1
    d = _core.algorithms.generate_cisd_hilbert_space(24, hf, hf)
    assert d.shape == (12636, 2) and len({tuple(x) for x in d.tolist()}) == 12636
    a, b = cisd_space(24, 5, 5)
    assert sorted(map(tuple, d.tolist())) == list(zip(a.tolist(), b.tolist()))
    s = _core.algorithms.generate_cis_hilbert_space(24, hf, hf)
    assert s.shape == (1 + 2 * 5 * 19, 2) and np.array_equal(s, d[: len(s)])
    # order: the reference first, alpha singles (virtual index outermost) ...
    assert tuple(d[0]) == (hf, hf)
    assert tuple(d[1]) == (hf ^ 1 ^ (1 << 5), hf) and tuple(d[2]) == (hf ^ 2 ^ (1 << 5), hf)
    assert tuple(d[1 + 95]) == (hf, hf ^ 1 ^ (1 << 5))          # ... then beta singles
    # open-shell reference, small space
    t = _core.algorithms.generate_cisd_hilbert_space(6, 0b0111, 0b0011)
    assert len({tuple(x) for x in t.tolist()}) == len(t) == 1 + 9 + 8 + 9 + 6 + 72

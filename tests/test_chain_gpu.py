"""GPU parity: chain DP, batched multiplier search and mask -> intervals vs the oracle (bit-exact gates)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rocco_b200
    return rocco_b200


@pytest.fixture(params=["scan", "default"])
def path_mode(request):
    """'scan' forces the parallel clamp-map scan even on short inputs; 'default' lets chromosomes of
    <= 4096 bins take the exact sequential kernel."""
    from rocco_b200 import _lib
    prev = _lib.load().rocco_b200_chain_set_seq_max(0 if request.param == "scan" else 4096)
    yield request.param
    _lib.load().rocco_b200_chain_set_seq_max(prev)


def _scores(n, seed, peak_frac=0.03):
    rng = np.random.default_rng(seed)
    return rng.normal(size=n) - 0.8 + 4.0 * (rng.random(n) < peak_frac) * rng.random(n)


# ------------------------------------------------------------------ fixed multiplier
def test_bruteforce_vectors(rb, golden, path_mode):
    """reference tests/test_rocco.py:397-415 (n = 9, random costs, four multipliers)"""
    s, c = golden["dp_bf_scores"], golden["dp_bf_costs"]
    for k in range(4):
        pen, val, cnt = golden[f"dp_bf_{k}_meta"]
        sol, v, n_sel = rb.solve_penalized_chain(s, c, pen)
        assert sol.dtype == np.uint8 and isinstance(v, float) and isinstance(n_sel, int)
        assert np.array_equal(sol, golden[f"dp_bf_{k}_mask"])
        assert n_sel == int(cnt)
        assert abs(v - val) <= 1e-6 * max(1.0, abs(val))          # objective gate: 1e-6 relative


@pytest.mark.parametrize("n", [1, 2, 3, 17, 4095, 4096, 4097, 8192, 12289, 70001])
def test_fixed_multiplier_matches_oracle_across_tile_edges(rb, oracle, n, path_mode):
    s = _scores(n, seed=n)
    rng = np.random.default_rng(n + 1)
    for costs in (np.full(max(n - 1, 0), 1.0), rng.uniform(0.0, 2.5, size=max(n - 1, 0)), np.full(max(n - 1, 0), 40.0)):
        for pen in (-0.7, 0.0, 0.31, 1.9):
            want = oracle.solve_penalized_chain(s, costs, pen)
            got = rb.solve_penalized_chain(s, costs, pen)
            assert np.array_equal(got[0], want[0]), (n, pen, int(np.sum(got[0] != want[0])))
            assert got[2] == want[2]
            assert abs(got[1] - want[1]) <= 1e-6 * max(1.0, abs(want[1]))


def test_input_validation_mirrors_reference(rb):
    with pytest.raises(ValueError):
        rb.solve_penalized_chain(np.zeros((2, 2)), np.zeros(1), 0.0)
    with pytest.raises(ValueError):
        rb.solve_penalized_chain(np.zeros(0), np.zeros(0), 0.0)
    with pytest.raises(ValueError):
        rb.solve_penalized_chain(np.zeros(5), np.zeros(3), 0.0)
    with pytest.raises(ValueError):
        rb.calibrate_selection_penalty(np.zeros(0), np.zeros(0), 1)


# ------------------------------------------------------------------ exact ties: (value, fewer-count) rule
@pytest.mark.parametrize("tag", ["const", "ints"])
def test_tie_break_cases_match_reference_golden(rb, golden, tag):
    arr = golden[f"dp_tie_{tag}_scores"]
    for k in range(3):
        b, gm, obj, pen, cnt, lam = golden[f"dp_tie_{tag}_{k}_meta"]
        sol, o, det = rb.solve_chrom_exact(arr, budget=None if b < 0 else b, gamma=gm, return_details=True)
        assert np.array_equal(sol, golden[f"dp_tie_{tag}_{k}_mask"]), (tag, k)
        assert det["selected_count"] == int(cnt) and det["selection_penalty"] == lam
        assert abs(o - obj) <= 1e-9 and abs(det["penalized_objective"] - pen) <= 1e-9


def test_integer_scores_exact_ties_random(rb, oracle, path_mode):
    rng = np.random.default_rng(5)
    for trial in range(12):
        n = int(rng.integers(2, 9000))
        s = rng.integers(-3, 4, size=n).astype(np.float64)
        c = rng.integers(0, 3, size=n - 1).astype(np.float64)
        for pen in (-1.0, 0.0, 0.5, 1.0, 2.0):
            want = oracle.solve_penalized_chain(s, c, pen)
            got = rb.solve_penalized_chain(s, c, pen)
            assert np.array_equal(got[0], want[0]), (trial, n, pen)
            assert got[2] == want[2] and got[1] == want[1]


def test_short_inputs_replay_reference_bits(rb, oracle):
    """n <= 4096: value, count, mask and the searched multiplier are the reference's bits, even where
    the 60-step bisection has converged to neighbouring doubles around a breakpoint."""
    rng = np.random.default_rng(12)
    for trial in range(25):
        n = int(rng.integers(1, 4097))
        s = np.round(rng.normal(size=n) * 2.0, int(rng.integers(0, 4)))
        gamma = float(rng.choice([0.0, 0.5, 1.0, 6.86]))
        budget = float(rng.choice([0.01, 0.1, 0.3, 0.9]))
        want_sol, want_obj, want = oracle.solve_chrom_exact(s, budget=budget, gamma=gamma, return_details=True)
        sol, obj, det = rb.solve_chrom_exact(s, budget=budget, gamma=gamma, return_details=True)
        assert np.array_equal(sol, want_sol), (trial, n, gamma, budget)
        assert det["selection_penalty"] == want["selection_penalty"]
        assert det["penalized_objective"] == want["penalized_objective"]
        assert det["selected_count"] == want["selected_count"]
        assert abs(obj - want_obj) <= 1e-9 * max(1.0, abs(want_obj))


# ------------------------------------------------------------------ budget search
def test_known_budget_case(rb, golden):
    """reference tests/test_rocco.py:418-437"""
    s = golden["dp_s8_scores"]
    sol, obj, det = rb.solve_chrom_exact(s, budget=0.375, gamma=1.0, return_details=True)
    assert sol.dtype == np.uint8 and sol.tolist() == [0, 0, 0, 0, 1, 1, 0, 0]
    assert np.sum(sol) <= 3 and det["selected_fraction"] <= 0.375
    assert np.isclose(obj, rb.objective_value(sol, s, rb.build_switch_costs(s, gamma=1.0)))
    o, pen, cnt, lam = golden["dp_s8_meta"]
    assert det["selection_penalty"] == lam and det["selected_count"] == int(cnt)
    assert np.isclose(obj, o, rtol=1e-12) and np.isclose(det["penalized_objective"], pen, rtol=1e-12)


@pytest.mark.parametrize("tag", ["g1", "g7", "g0"])
def test_solve_chrom_exact_matches_reference_golden(rb, golden, tag):
    budget, gamma, obj, pen, cnt, lam = golden[f"dp_{tag}_meta"]
    sol, o, det = rb.solve_chrom_exact(golden["score_a_scores"], budget=budget, gamma=gamma, return_details=True)
    assert np.array_equal(sol, golden[f"dp_{tag}_mask"])
    assert det["selected_count"] == int(cnt)
    assert det["selection_penalty"] == lam                     # same dyadic lattice point
    assert abs(o - obj) <= 1e-6 * abs(obj) and abs(det["penalized_objective"] - pen) <= 1e-6 * abs(pen)


@pytest.mark.parametrize("n,gamma,budget,seed", [(50_000, 1.0, 0.02, 0), (200_000, 6.86, 0.03, 1),
                                                 (934_200, 1.0, 0.02, 21), (33_333, 0.0, 0.1, 3), (4096, 2.0, 0.5, 4)])
def test_budget_search_matches_oracle(rb, oracle, n, gamma, budget, seed):
    s = _scores(n, seed)
    want_sol, want_obj, want = oracle.solve_chrom_exact(s, budget=budget, gamma=gamma, return_details=True)
    for levels in (1, 3, 4):
        from rocco_b200.pipeline import solve_chromosomes
        got = solve_chromosomes([s], [budget], [gamma], levels_per_round=levels)[0]
        assert np.array_equal(got["solution"], want_sol), (levels, int(np.sum(got["solution"] != want_sol)))
        assert got["selected_count"] == want["selected_count"] <= int(np.floor(n * budget))
        # Long inputs use the scan: its decisions differ from the reference's only inside the reference's own
        # rounding noise around a breakpoint, which can move the returned lattice point by a few steps
        # (bracket width / 2^60 each) without changing the mask.  Gate: <= 1e-9 absolute.
        assert abs(got["selection_penalty"] - want["selection_penalty"]) <= 1e-9, (got["selection_penalty"], want["selection_penalty"])
        assert abs(got["objective"] - want_obj) <= 1e-6 * abs(want_obj)
        assert got["dp_passes"] == 62


def test_calibrate_with_vector_costs_matches_oracle(rb, oracle):
    n = 30_000
    s = _scores(n, 9)
    c = np.random.default_rng(10).uniform(0.1, 3.0, size=n - 1)
    want = oracle.calibrate_selection_penalty(s, c, 900)
    got = rb.calibrate_selection_penalty(s, c, 900)
    assert abs(got[0] - want[0]) <= 1e-9 and got[3] == want[3]
    assert np.array_equal(got[1], want[1])
    assert abs(got[2] - want[2]) <= 1e-6 * abs(want[2])
    # target == n short-circuits to lambda = 0 (dp.py:102-108)
    lam, sol, _, cnt = rb.calibrate_selection_penalty(s, c, n)
    w = oracle.calibrate_selection_penalty(s, c, n)
    assert lam == 0.0 and np.array_equal(sol, w[1]) and cnt == w[3]


def test_many_chromosomes_one_launch_set(rb, oracle):
    from rocco_b200.pipeline import solve_chromosomes
    sizes = [9342, 11723, 31208, 1, 4096, 257]
    budgets = [0.02, 0.045, 0.015, 0.5, None, 0.1]
    gammas = [1.0, 1.0, 1.0, 1.0, 0.5, 3.0]
    pens = [None, None, None, None, None, 0.25]
    sc = [_scores(n, 100 + i) for i, n in enumerate(sizes)]
    got = solve_chromosomes(sc, budgets, gammas, pens)
    for i in range(len(sizes)):
        sol, obj, det = oracle.solve_chrom_exact(sc[i], budget=budgets[i], gamma=gammas[i],
                                                 selection_penalty=pens[i], return_details=True)
        assert np.array_equal(got[i]["solution"], sol), i
        assert abs(got[i]["selection_penalty"] - det["selection_penalty"]) <= 1e-9
        assert got[i]["selected_count"] == det["selected_count"]


def test_solve_chrom_exact_accepts_cuda_tensor(rb, oracle):
    import torch
    s = _scores(20_000, 77)
    sol, obj = rb.solve_chrom_exact(torch.from_numpy(s).cuda(), budget=0.05, gamma=1.0)
    want_sol, want_obj = oracle.solve_chrom_exact(s, budget=0.05, gamma=1.0)
    assert isinstance(sol, np.ndarray) and sol.dtype == np.uint8 and np.array_equal(sol, want_sol)
    assert abs(obj - want_obj) <= 1e-6 * abs(want_obj)


# ------------------------------------------------------------------ mask -> intervals / BED
def test_bed_text_matches_reference_golden(rb, golden, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    iv = golden["bed_iv"]
    for tag, ml in (("all", None), ("min150", 150)):
        f = rb.chrom_solution_to_bed("chr21", iv, golden["dp_g1_mask"], ID="gold", min_length_bp=ml)
        assert f == "rocco_gold_chr21.bed"
        assert open(f, "rb").read() == golden[f"bed_{tag}_text"].tobytes()
    toy = np.array([0, 1, 1, 0, 0, 1, 0, 1, 1, 1], dtype=np.uint8)
    f = rb.chrom_solution_to_bed("chrT", np.arange(0, 500, 50), toy)
    assert open(f, "rb").read() == golden["bed_toy_text"].tobytes()
    with pytest.raises(ValueError):
        rb.chrom_solution_to_bed("chrT", np.array([0, 50, 150]), toy[:3])
    with pytest.raises(ValueError):
        rb.chrom_solution_to_bed("chrT", np.arange(0, 500, 50), toy[:4])


@pytest.mark.parametrize("n,seed", [(2, 0), (4097, 1), (300_001, 2)])
def test_intervals_match_oracle(rb, oracle, n, seed, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(seed)
    mask = (rng.random(n) < 0.3).astype(np.uint8)
    mask[-3:] = 1
    iv = np.arange(5000, 5000 + 20 * n, 20)
    for ml in (None, 60):
        want = oracle.solution_to_records("chrZ", iv, mask, min_length_bp=ml)
        f = rb.chrom_solution_to_bed("chrZ", iv, mask, min_length_bp=ml)
        got = oracle.read_bed_records(f)
        assert got == want


def test_combine_reproduces_reference_combined_bed(rb, bed_fixtures, tmp_path):
    files = []
    for name in ("ref_chr19", "ref_chr21", "ref_chrX"):
        p = tmp_path / f"{name}.bed"
        p.write_bytes(bed_fixtures[name].tobytes())
        files.append(str(p))
    out = rb.combine_chrom_results(files, str(tmp_path / "combined.bed"))
    assert open(out, "rb").read() == bed_fixtures["combined_ref"].tobytes()


# ------------------------------------------------------------------ multiplier sweep (BASELINE.json config 5)
@pytest.mark.parametrize("n,gamma,seed", [(5000, 1.0, 0), (70_001, 1.0, 1), (200_000, 6.86, 2), (33_000, 0.0, 3)])
def test_multiplier_sweep_matches_oracle(rb, oracle, n, gamma, seed):
    from rocco_b200 import _lib
    from rocco_b200.pipeline import sweep_multipliers
    prev = _lib.load().rocco_b200_chain_set_seq_max(0)
    try:
        s = _scores(n, seed)
        lams = np.concatenate([np.linspace(np.quantile(s, 0.5), np.quantile(s, 0.999), 61), [-3.0, 0.0, 50.0]])
        counts, pen, obj = sweep_multipliers(s, gamma, lams)
        c = oracle.build_switch_costs(s, gamma)
        for k, lam in enumerate(lams):
            sol, val, cnt = oracle.solve_penalized_chain(s, c, lam)
            assert counts[k] == cnt, (k, lam)
            assert abs(pen[k] - val) <= 1e-6 * max(1.0, abs(val))
            want_obj = oracle.objective_value(sol, s, c)
            assert abs(obj[k] - want_obj) <= 1e-6 * max(1.0, abs(want_obj))
        assert np.all(np.diff(counts[:61]) <= 0)                   # count(lambda) is non-increasing
    finally:
        _lib.load().rocco_b200_chain_set_seq_max(prev)


def test_multiplier_sweep_256_in_one_launch_set(rb, oracle):
    from rocco_b200.pipeline import sweep_multipliers
    s = _scores(120_000, 42)
    lams = np.linspace(np.quantile(s, 0.5), np.quantile(s, 0.999), 256)
    counts, pen, obj = sweep_multipliers(s, 1.0, lams)
    c = oracle.build_switch_costs(s, 1.0)
    for k in (0, 17, 128, 255):
        assert counts[k] == oracle.solve_penalized_chain(s, c, lams[k])[2]


@pytest.mark.parametrize("gamma", [1.0, 6.86])
def test_budget_search_many_seeds_masks_identical(rb, oracle, gamma):
    """Documented near-tie policy on long inputs: over 8 seeds x 500k bins the mask and count are the reference's in every
    case; the returned multiplier may sit a few lattice steps (bracket / 2^60) away when the reference's own rounding
    noise (~1e-10 on |V| ~ 1e5) decides its last bisection steps."""
    from rocco_b200.pipeline import solve_chromosomes
    n, budget = 500_000, 0.03
    scores = [_scores(n, 1000 + k) for k in range(8)]
    got = solve_chromosomes(scores, [budget] * 8, [gamma] * 8)
    exact = 0
    for k in range(8):
        want_sol, want_obj, want = oracle.solve_chrom_exact(scores[k], budget=budget, gamma=gamma, return_details=True)
        assert np.array_equal(got[k]["solution"], want_sol), (k, int(np.sum(got[k]["solution"] != want_sol)))
        assert got[k]["selected_count"] == want["selected_count"]
        assert abs(got[k]["selection_penalty"] - want["selection_penalty"]) <= 1e-9
        assert abs(got[k]["objective"] - want_obj) <= 1e-6 * abs(want_obj)
        exact += got[k]["selection_penalty"] == want["selection_penalty"]
    print(f"gamma={gamma}: multiplier bit-identical in {exact}/8 cases")


def test_mask_to_runs_with_hundreds_of_tiny_contigs(rb, oracle):
    """more than 64 chromosomes / contigs per launch (alt and unplaced scaffolds of a few bins each, many per tile)"""
    import torch
    from rocco_b200 import pipeline
    rng = np.random.default_rng(4)
    lengths = [int(v) for v in rng.integers(1, 60, size=700)] + [5000, 1, 2, 9000]
    offsets, total = pipeline.layout_offsets(lengths)
    masks = [(rng.random(n) < 0.4).astype(np.uint8) for n in lengths]
    buf = np.zeros(total, dtype=np.uint8)
    for off, m in zip(offsets, masks):
        buf[off:off + len(m)] = m
    chrom, starts, ends = pipeline.masks_to_runs(torch.from_numpy(buf).cuda(), offsets, lengths)
    got = list(zip(chrom.tolist(), starts.tolist(), ends.tolist()))
    want = []
    for c, m in enumerate(masks):
        for _, a, b in oracle.solution_to_records("c", np.arange(len(m)), m):
            want.append((c, a, b))
    assert got == want


@pytest.fixture
def exact_search():
    from rocco_b200 import _lib
    prev = _lib.load().rocco_b200_chain_set_exact_search(1)
    yield
    _lib.load().rocco_b200_chain_set_exact_search(prev)


@pytest.mark.parametrize("n,gamma,budget,seed", [(50_000, 1.0, 0.02, 0), (200_000, 6.86, 0.03, 1),
                                                 (934_200, 1.0, 0.02, 21), (33_333, 0.0, 0.1, 3)])
def test_exact_search_returns_the_reference_multiplier_bit_for_bit(rb, oracle, exact_search, n, gamma, budget, seed):
    """SURVEY.md 8d gate "lambda from the search: identical double" on long inputs, with the exact replay switched on"""
    from rocco_b200.pipeline import solve_chromosomes
    s = _scores(n, seed)
    want_sol, want_obj, want = oracle.solve_chrom_exact(s, budget=budget, gamma=gamma, return_details=True)
    for levels in (2, 3):
        got = solve_chromosomes([s], [budget], [gamma], levels_per_round=levels)[0]
        assert got["selection_penalty"] == want["selection_penalty"], (levels, got["selection_penalty"], want["selection_penalty"])
        assert np.array_equal(got["solution"], want_sol)
        assert got["selected_count"] == want["selected_count"]
        assert abs(got["objective"] - want_obj) <= 1e-6 * abs(want_obj)


@pytest.mark.parametrize("gamma", [1.0, 6.86])
def test_exact_search_many_seeds_all_identical(rb, oracle, exact_search, gamma):
    from rocco_b200.pipeline import solve_chromosomes
    n, budget = 500_000, 0.03
    scores = [_scores(n, 1000 + k) for k in range(8)]
    got = solve_chromosomes(scores, [budget] * 8, [gamma] * 8)
    for k in range(8):
        want_sol, want_obj, want = oracle.solve_chrom_exact(scores[k], budget=budget, gamma=gamma, return_details=True)
        assert got[k]["selection_penalty"] == want["selection_penalty"], (k, got[k]["selection_penalty"], want["selection_penalty"])
        assert np.array_equal(got[k]["solution"], want_sol)
        assert got[k]["selected_count"] == want["selected_count"]


def test_exact_search_with_vector_costs(rb, oracle, exact_search):
    n = 30_000
    s = _scores(n, 9)
    c = np.random.default_rng(10).uniform(0.1, 3.0, size=n - 1)
    want = oracle.calibrate_selection_penalty(s, c, 900)
    got = rb.calibrate_selection_penalty(s, c, 900)
    assert got[0] == want[0] and got[3] == want[3] and np.array_equal(got[1], want[1])


def test_tile_freezing_changes_nothing_but_the_time(rb, oracle):
    """search rounds answer tiles whose decisions cannot change inside the remaining bracket from stored outputs: counts,
    multipliers and masks must be those of the run that evaluates every tile in every round (and the oracle's)"""
    from rocco_b200 import _lib
    from rocco_b200.pipeline import solve_chromosomes
    rng = np.random.default_rng(12)
    sc = [_scores(300_000, 41), _scores(70_001, 42), rng.normal(size=120_000) + 3.0 * (rng.random(120_000) < 0.2)]   # the last: many unsaturated tile boundaries
    budgets, gammas = [0.02, 0.05, 0.1], [1.0, 2.5, 0.7]
    runs = {}
    for on in (1, 0):
        prev = _lib.load().rocco_b200_chain_set_tile_freezing(on)
        try:
            runs[on] = solve_chromosomes(sc, budgets, gammas)
        finally:
            _lib.load().rocco_b200_chain_set_tile_freezing(prev)
    for a, b, s, bud, g in zip(runs[1], runs[0], sc, budgets, gammas):
        assert a["selection_penalty"] == b["selection_penalty"] and a["selected_count"] == b["selected_count"]
        assert np.array_equal(a["solution"], b["solution"])
        want_sol, _, want = oracle.solve_chrom_exact(s, budget=bud, gamma=g, return_details=True)
        assert np.array_equal(a["solution"], want_sol) and a["selected_count"] == want["selected_count"]

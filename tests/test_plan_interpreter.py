"""CPU tier: the static launch plans of the numeric phases (csrc/plan.cpp) checked without a GPU
(tools/plancheck/plancheck.py).

1. Hazards.  The tasks of one launch run concurrently on the GPU and the kernels use no atomics, so within every launch
   no two tasks may write the same arena entry and none may read what another one writes.  compute-sanitizer's racecheck
   is refused on the GPU pool; here the read and write sets of every task are derived from the same Task records the
   kernels consume and intersected on the host.  Negative controls: merging two extend-add launches of sibling children,
   or a POTRF launch with the apply-inverse launch that follows it, must be reported.
2. Interpretation.  Every launch kind is restated in NumPy from its kernel and the plans are executed on NaN-poisoned
   arenas: the factor must satisfy L L' = P A P', the kept inverses W L11 = I, the selected inverse must equal
   inv(P A P') on the pattern of L - which also pins the read / write model of (1) and shows that nothing is read before
   it is written."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pc():
    spec = importlib.util.spec_from_file_location("plancheck", os.path.join(ROOT, "tools", "plancheck", "plancheck.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return mod


def _check(res):
    assert res["violations"] == []
    assert res["factor_err"] < 1e-13 and res["selinv_err"] < 1e-11 and res["zdiag_err"] < 1e-11
    if "wide_err" in res:
        assert res["wide_err"] < 1e-12


@pytest.mark.parametrize("nx,ordering,use_wide", [(12, "nd", False), (40, "nd", False), (58, "nd", True), (45, "amd", False),
                                                  (45, "nd_amd", False)])
def test_plans_of_a_2d_posterior(pc, W, nx, ordering, use_wide):
    prob = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.2, seed=nx)
    res = pc.check_matrix(prob["Qpost"], ordering=ordering, coords=prob["nodes"] if ordering == "nd" else None,
                          use_wide=use_wide, wide_min=65)
    _check(res)
    if use_wide:
        assert res["n_wide"] > 0 and res["max_front"] > pc.SMALL_FRONT_MAX


@pytest.mark.parametrize("ordering,use_wide", [("nd", True), ("amd", False)])
def test_plans_of_a_spacetime_precision(pc, W, ordering, use_wide):
    """3-D separators: fronts of several hundred columns (recursive blocking over several 64-column blocks, multi-level
    recursive-doubling inverses, many children per parent)."""
    st = W.heat_spacetime_sparse(14, 6)
    res = pc.check_matrix(st["A"], ordering=ordering, coords=st["coords"] if ordering == "nd" else None, use_wide=use_wide,
                          wide_min=65)
    _check(res)
    assert res["max_front"] > 256 and res["factor"]["tasks"] > res["nsuper"]


def test_plans_of_a_banded_chain(pc):
    """Natural ordering of a banded matrix: a chain of supernodes (one child each), given permutation path."""
    n, bw = 700, 40
    rng = np.random.default_rng(0)
    B = sp.diags([rng.standard_normal(n - k) for k in range(1, bw)], list(range(1, bw)), format="csc")
    A = (B + B.T + sp.identity(n) * (2.0 * bw)).tocsc()
    res = pc.check_matrix(A, perm=np.arange(n))
    _check(res)


def test_hazard_checker_reports_injected_hazards(pc, W):
    prob = W.matern_posterior(58, obs_frac=0.2, q_eps=1e2, corr_range=0.2, seed=58)
    P = pc.Plans(prob["Qpost"], ordering="nd", coords=prob["nodes"], wide_min=65)
    plan = P.plan("factor")
    L = plan["launches"]
    kinds = [int(k) for k in L["kind"]]

    def merged(i, j):
        """a plan with the single launch made of the tasks of launches i and j"""
        ti = plan["tasks"][int(L[i]["task0"]):int(L[i]["task0"]) + int(L[i]["ntasks"])]
        tj = plan["tasks"][int(L[j]["task0"]):int(L[j]["task0"]) + int(L[j]["ntasks"])]
        one = np.zeros(1, dtype=pc.LAUNCH)
        one[0]["kind"] = L[i]["kind"]
        one[0]["ntasks"] = len(ti) + len(tj)
        return dict(tasks=np.concatenate([ti, tj]), launches=one)

    # (a) two child ranks of the same parents in ONE launch: both add into the same separator entries
    pairs = [i for i in range(len(kinds) - 1) if kinds[i] == pc.LK_EXTEND_ADD and kinds[i + 1] == pc.LK_EXTEND_ADD]
    assert pairs
    found = False
    for i in pairs:
        _, _, bad = pc.hazards(P, merged(i, i + 1))
        found = found or any(b[0] == "write/write" for b in bad)
    assert found
    # (b) the apply-inverse launch reads the inverse blocks the POTRF launch before it writes
    i = next(i for i in range(len(kinds) - 1) if kinds[i] == pc.LK_POTRF and kinds[i + 1] == pc.LK_TRSM_RLT)
    mp = merged(i, i + 1)
    # the merged launch is interpreted task by task with each task's own kind for the access model
    kind_of = [pc.LK_POTRF] * int(L[i]["ntasks"]) + [pc.LK_TRSM_RLT] * int(L[i + 1]["ntasks"])
    writes = {}
    hit = False
    for k, t in enumerate(mp["tasks"]):
        R, Wr = pc.task_access(P, kind_of[k], t)
        for a, ix in R:
            for (a2, k2), ix2 in writes.items():
                if a2 == a and k2 != k and np.intersect1d(ix, ix2).size:
                    hit = True
        for a, ix in Wr:
            writes[(a, k)] = ix
    assert hit
    # and the unmodified plans are clean
    for w in ("zero", "factor", "selinv"):
        assert pc.hazards(P, P.plan(w))[2] == []


def test_poison_shows_a_missing_clear(pc, W):
    """Without the zero plan the factorisation reads entries nobody wrote: the NaN poison reaches the factor."""
    prob = W.matern_posterior(40, obs_frac=0.2, q_eps=1e2, corr_range=0.2, seed=40)
    P = pc.Plans(prob["Qpost"], ordering="nd", coords=prob["nodes"])
    S = pc.State(P, poison=True)
    pc.scatter_values(P, S)
    try:
        pc.run(P, S, P.plan("factor"))
        Lm = pc.factor_matrix(P, S)
        assert not np.all(np.isfinite(Lm.data))
    except np.linalg.LinAlgError:
        pass  # a NaN front is not positive definite either


@pytest.mark.parametrize("case,nr", [("mesh12", 3), ("mesh58", 7), ("mesh58", 32), ("spacetime", 9)])
def test_panel_sweep_plans(pc, W, case, nr):
    """The multi-right-hand-side sweeps (build_solve_mr_plans): forward and backward plans are hazard-free and, run
    after the factor plan on poisoned arenas, solve (P A P') X' = B' for a node-major panel of nr right-hand sides."""
    if case == "spacetime":
        st = W.heat_spacetime_sparse(14, 6)
        A, coords = st["A"], st["coords"]
    else:
        nx = int(case[4:])
        prob = W.matern_posterior(nx, obs_frac=0.2, q_eps=1e2, corr_range=0.2, seed=nx)
        A, coords = prob["Qpost"], prob["nodes"]
    res = pc.check_panel_solves(A, nr=nr, ordering="nd", coords=coords)
    assert res["violations"] == []
    assert res["fwd_err"] < 1e-12 and res["solve_err"] < 1e-11


def test_plans_of_random_patterns(pc):
    """Seeded fuzz over patterns the meshes do not produce - random, arrow (one dense row), disconnected blocks, fully
    dense; n from 1 to 333; every ordering kind - through the hazard check and the interpreter (factor, selected
    inversion, kept inverses, panel sweeps)."""
    rng = np.random.default_rng(123)

    def spd(n, dens, kind):
        seed = int(rng.integers(1 << 30))
        if kind == "dense":
            B = sp.csc_matrix(rng.standard_normal((n, n)))
        elif kind == "arrow":
            B = sp.random(n, n, density=dens, random_state=seed, format="lil")
            B[0, :] = 1.0
            B[:, 0] = 1.0
            B = B.tocsc()
        elif kind == "blocks" and n >= 6:
            k = n // 3
            B = sp.block_diag([sp.random(k, k, density=min(1.0, 3 * dens), random_state=seed + i) for i in range(3)] +
                              [sp.identity(n - 3 * k)], format="csc")
        else:
            B = sp.random(n, n, density=dens, random_state=seed, format="csc")
        S = (abs(B) + abs(B).T).tocsc()
        A = (-S + sp.diags(2.0 * (np.asarray(S.sum(axis=1)).ravel() + 1.0))).tocsc()  # diagonally dominant
        A.sort_indices()
        return A

    for _ in range(45):
        n = int(rng.choice([1, 2, 3, 5, 17, 64, 65, 130, 200, 333]))
        kind = str(rng.choice(["random", "arrow", "blocks", "dense"]))
        n = min(n, 130) if kind == "dense" else n
        A = spd(n, float(rng.choice([0.005, 0.02, 0.1, 0.4])), kind)
        ordering = str(rng.choice(["nd", "amd", "nd_amd", "natural"]))
        res = pc.check_matrix(A, ordering=ordering, use_wide=bool(rng.integers(2)), wide_min=65)
        assert res["violations"] == [], (n, kind, ordering, res["violations"][:2])
        assert res["factor_err"] < 1e-12 and res["selinv_err"] < 1e-9, (n, kind, ordering, res)
        r2 = pc.check_panel_solves(A, nr=int(rng.choice([1, 5, 8, 33])), ordering=ordering)
        assert r2["violations"] == [] and r2["solve_err"] < 1e-9, (n, kind, ordering, r2)


@pytest.mark.parametrize("b,nblocks", [(1, 3), (64, 3), (65, 4), (100, 4), (193, 4), (320, 4), (700, 3)])
def test_btd_lookahead_schedule(pc, b, nblocks):
    """The three-stream look-ahead schedule of the block-tridiagonal factor (btd.cu: btd_run_factor): POTRF_i, TRSM_{i+1}
    and the rank-512 SYRK_{i+1} run on separate streams ordered by per-panel events only.  The host issue sequence is
    replayed into a happens-before relation; no two launches that may overlap touch a common entry unless both read it,
    everything is ordered before the end of the stream the host waits for, and the launches - in host order and in random
    topological orders of that relation - produce the factor of src/tridiagonal_cholesky.jl:70-80 from lower triangles
    only (the rest of the arena and the inverse slots are NaN)."""
    res = pc.check_btd_lookahead(b, nblocks=nblocks, orders=2)
    assert res["violations"] == []
    assert max(res["factor_err"]) < 1e-13
    if b > 64:
        assert res["concurrent_pairs"] > 0  # the schedule does overlap launches: the check is not vacuous


@pytest.mark.parametrize("b,nblocks", [(2, 3), (100, 4), (129, 4), (320, 3)])
def test_btd_time_sharded_lane_schedule(pc, b, nblocks):
    """The fourth stream of the time-sharded factor (gmrfb_btd_dist_factor's after_block): W_i = L_i^-1 by recursive
    doubling and the spike step of block i run behind POTRF_i, ordered by one event, while the chains of the next blocks
    go on.  Same check; in every order the interpreted launches give W_i L_i = I, the spike blocks
    S_1 = E_l' W_1', S_i = -S_{i-1} C_i' W_i' and Q = sum S_i S_i'."""
    res = pc.check_btd_lookahead(b, nblocks=nblocks, orders=2, lane=True)
    assert res["violations"] == []
    assert max(res["factor_err"]) < 1e-12


@pytest.mark.parametrize("mutation,b", [(dict(drop_trsm_waits=True), 320), (dict(drop_potrf_wait=True), 320),
                                        (dict(syrk_wait_shift=1), 700), (dict(lane=True, drop_lane_wait=True), 200)])
def test_btd_schedule_checker_reports_missing_synchronisation(pc, mutation, b):
    """Negative controls: without the per-panel waits of the TRSM, without the wait of POTRF_{i+1} for the last update,
    with a rank-512 update waiting one column block too early, or without the event between POTRF_i and the time-sharded
    lane, the checker names the unordered launches."""
    BP = pc.BtdPlans(b)
    nodes, issues = pc.btd_schedule(BP, 3, **mutation)
    races, _ = pc.btd_races(BP, nodes)
    assert races
    kinds = {r[0] for r in races}
    assert "read after write" in kinds


def test_btd_schedule_model_bounds(pc):
    """The fluid model of the look-ahead schedule (profiles/r02_btd_schedule_model.md): its makespan lies between the two
    lower bounds (critical path with unlimited SMs; SM-seconds of work over 148 SMs) and the serial schedule."""
    BP = pc.BtdPlans(700)
    nodes, issues = pc.btd_schedule(BP, 5)
    assert not issues
    t, cp, work = pc.btd_simulate(nodes)
    serial, _, _ = pc.btd_simulate(nodes, serial=True)
    assert max(cp, work) <= t * (1 + 1e-9) and t <= serial * (1 + 1e-9)
    assert t < 0.9 * serial  # the look-ahead does overlap the three chains

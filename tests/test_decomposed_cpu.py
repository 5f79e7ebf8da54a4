"""The clique-decomposed SDP built from the LIBRARY'S hand-off (tests/golden/handoff_*.npz, dumped on a B200 by
tests/golden/make_handoff.py from nnsdp_affine_get + nnsdp_cliques) -- /root/reference/src/Methods/chordal_sdp.jl:19-57
(setupZs!, SingleDecomp and DoubleDecomp with D_k1 / D_k2), :60-93 (setupZksum!), :125-153 (setupReach!) -- checked
WITHOUT a solver: the stored solutions of oracle/sdp_decomposed.py are re-verified as certificates.

  * primal: (gamma*, Z_k*) is feasible for the problem built from the hand-off, so its optimum is AT MOST c'x*;
  * dual:   the stored multipliers X_k >= 0 give, by weak duality, c'y >= dual objective + rd'y for every feasible
            y, rd = the dual residual (max |rd| <= 1e-6 here; rd'x* ~ 1e-8 is what the bound is corrected by);
  * the bracket contains the optimum of the DENSE LMI Z(gamma) <= 0 the oracle solved independently
    (tests/golden/scale_W10_D10_optimum.json) -- by Agler's theorem the two problems have the same optimum exactly
    when the cliques cover Z's pattern chordally and the hand-off places every coefficient on the right entry;
  * and it lies within the 3-digit band of the optima the reference recorded with MOSEK (dump/scale/*W10-D10*.csv).
"""
import json
import os

import numpy as np
import pytest

import sdp_decomposed as sd

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [(f"W10-D10_beta{b}", m) for b in range(8) for m in ("single", "double")] + [("W10-D20_beta2", "single"), ("W10-D20_beta2", "double")]
# dump/scale/{deepsdp,chordalsdp,chordalsdp2}-scale-I2-O2-W10-D20.nnet.csv, beta = 2 (recorded by the reference)
REF_W10_D20_BETA2 = (1.34675, 1.34711, 1.34714)


def _load(name, mode):
    h = np.load(os.path.join(GOLD, f"handoff_{name}.npz"))
    path = os.path.join(GOLD, f"decomposed_{name}_{mode}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not generated")
    return h, np.load(path)


@pytest.mark.parametrize("name,mode", CASES)
def test_stored_decomposed_solution_is_a_certificate(name, mode):
    h, sol = _load(name, mode)
    cliques = sd.cliques_from_npz(h)
    prob = sd.problem_from_handoff(h, cliques, mode)
    x, U = sol["x"], float(sol["U"])
    ng = prob["ng"]
    assert len(x) == ng + prob["ns"]
    # primal feasibility: multipliers in [0, U], every block negative semidefinite
    assert np.all(x[:ng] >= 0) and np.all(x[:ng] <= U)
    X, o = [], 0
    for b in prob["blocks"]:
        m = len(b["idx"])
        X.append(sol["X"][o:o + m * m].reshape(m, m))
        o += m * m
    assert o == sol["X"].size
    pobj, lam, dobj, rd, rdx = sd.certificate(prob, x, X, sol["xl"], sol["xu"], U)
    assert lam <= 0.0, lam
    assert abs(pobj - float(sol["obj"])) <= 1e-12 * abs(pobj)
    # dual side: X_k, xl, xu >= 0, small residual; weak duality c'y >= dobj + rd'y with y ~ x*
    for Xj in X:
        assert np.linalg.eigvalsh(0.5 * (Xj + Xj.T)).min() >= -1e-9 * max(1.0, np.abs(Xj).max())
    assert sol["xl"].min() >= 0 and sol["xu"].min() >= 0
    assert rd <= 2e-6
    lower = dobj - 10 * abs(rdx)
    width = (pobj - lower) / abs(pobj)
    assert 0 <= width <= (3e-5 if name.startswith("W10-D10") else 5e-3), width
    if name.startswith("W10-D10"):
        beta = name.split("beta")[1]
        res = json.load(open(os.path.join(GOLD, "scale_W10_D10_optimum.json")))
        dense = res["oracle_optimum"][beta]["obj"]
        # (i) the decomposed optimum equals the dense-Z optimum to solver tolerance
        assert abs(pobj - dense) <= 5e-6 * dense, (pobj, dense)
        assert lower <= dense * (1 + 1e-6)
        # (ii) inside the band of the reference's recorded optima (3 digits)
        ref = [res["reference_obj_val"][k][int(beta)] for k in res["reference_obj_val"]]
        assert abs(pobj - np.mean(ref)) <= 2e-3 * np.mean(ref)
    else:
        assert abs(pobj - np.mean(REF_W10_D20_BETA2)) <= 2e-3 * np.mean(REF_W10_D20_BETA2)


def test_double_decomposition_blocks_follow_Dk():
    """DoubleDecomp (chordal_sdp.jl:25-46): cliques 1 and p keep one block, every other clique becomes the blocks
    Z_k[D_k1, D_k1] and Z_k[D_k2, D_k2]; single and double give the same optimum."""
    h = np.load(os.path.join(GOLD, "handoff_W10-D10_beta2.npz"))
    cliques = sd.cliques_from_npz(h)
    single, double = sd.blocks_of(cliques, "single"), sd.blocks_of(cliques, "double")
    p = len(cliques)
    assert len(single) == p and len(double) == 2 * p - 2
    k = 3
    Ck, _, (D1, D2) = cliques[k]
    assert np.array_equal(double[1 + 2 * (k - 1)], Ck[D1 - 1] - 1) and np.array_equal(double[2 + 2 * (k - 1)], Ck[D2 - 1] - 1)
    a, b = (np.load(os.path.join(GOLD, f"decomposed_W10-D10_beta2_{m}.npz")) for m in ("single", "double"))
    assert abs(float(a["obj"]) - float(b["obj"])) <= 5e-6 * float(a["obj"])


def test_a_misplaced_entry_breaks_the_certificate():
    """Sensitivity: swap the coefficients of two cover entries in the hand-off -- the stored solution is no longer
    feasible, i.e. the check above really depends on the entry numbering."""
    h = dict(np.load(os.path.join(GOLD, "handoff_W10-D10_beta2.npz")))
    sol = np.load(os.path.join(GOLD, "decomposed_W10-D10_beta2_single.npz"))
    cliques = sd.cliques_from_npz(h)
    ent = h["coo_ent"].copy()
    vals, counts = np.unique(ent, return_counts=True)
    e1, e2 = vals[np.argsort(-counts)[:2]]                      # the two entries with most coefficients
    ent[h["coo_ent"] == e1], ent[h["coo_ent"] == e2] = e2, e1
    h["coo_ent"] = ent
    prob = sd.problem_from_handoff(h, cliques, "single")
    assert sd.lambda_max(prob, sol["x"]) > 1e-6


FROM_DENSE = [n for n in [f"W10-D10_beta{b}" for b in range(8)] + ["W10-D20_beta2"]
              if os.path.exists(os.path.join(GOLD, f"decomposed_{n}_single_from_dense.npz"))]


@pytest.mark.parametrize("name", FROM_DENSE)
def test_point_built_from_the_dense_optimum_is_feasible_for_the_decomposed_problem(name):
    """Agler's theorem made constructive (oracle/sdp_decomposed.certificate_from_dense, tests/golden/make_from_dense.py):
    the optimum of the DENSE LMI of the hand-off, factorised without fill along the cliques of makeCliques
    (chordal_cliques.jl:13-59), gives blocks Z_k < 0 with Z .== Zksum (chordal_sdp.jl:60-93) -- a strictly feasible
    point of the decomposed problem built from the library's hand-off at the dense optimum.  With the easy direction
    (a decomposed-feasible point has sum_k Z_k = Z(gamma) <= 0, so the decomposed optimum cannot be below the dense
    one) this pins  |decomposed optimum - dense optimum| <= 3e-6  without trusting a solver, also on W10-D20, where the
    block interior-point method stalls at a bracket of 2e-3."""
    h = np.load(os.path.join(GOLD, f"handoff_{name}.npz"))
    sol = np.load(os.path.join(GOLD, f"decomposed_{name}_single_from_dense.npz"))
    cliques = sd.cliques_from_npz(h)
    prob = sd.problem_from_handoff(h, cliques, "single")
    x, U = sol["x"], float(sol["U"])
    ng = prob["ng"]
    assert len(x) == ng + prob["ns"] and np.all(x[:ng] > 0) and np.all(x[:ng] < U)
    X, o = [], 0
    for b in prob["blocks"]:
        m = len(b["idx"])
        X.append(sol["X"][o:o + m * m].reshape(m, m))
        o += m * m
    pobj, lam, dobj, rd, rdx = sd.certificate(prob, x, X, sol["xl"], sol["xu"], U)
    assert lam < 0.0, lam                                        # every block strictly negative definite
    assert abs(pobj - float(sol["obj"])) <= 1e-12 * abs(pobj)
    # the blocks add up to the dense Z(gamma) of the same hand-off, which is therefore negative semidefinite as well
    Z0, A = sd.dense_from_handoff(h, prob["keep"])
    Zg = Z0 + np.tensordot(x[:ng], A, 1)
    Zsum = np.zeros_like(Zg)
    for b, Zk in zip(prob["blocks"], sd.block_matrices(prob, x)):
        Zsum[np.ix_(b["idx"], b["idx"])] += Zk
    assert np.abs(Zsum - Zg).max() <= 1e-11 * max(1.0, np.abs(Zg).max())
    assert np.linalg.eigvalsh(Zg).max() < 0.0
    # at the dense optimum (the barrier's last iterate, moved 1e-3 of the way back to a centred point for the margin)
    dense = float(sol["dense_obj"])
    assert 0.0 <= pobj - dense <= 3e-6 * dense, (pobj, dense)
    # the stored interior-point multipliers bound the optimum from below: a solver-free bracket around the dense optimum
    assert rd <= 2e-6
    lower = dobj - 10 * abs(rdx)
    assert lower <= dense and (pobj - lower) / pobj <= (3e-5 if name.startswith("W10-D10") else 1.5e-3)
    if name.startswith("W10-D10"):
        ipm = np.load(os.path.join(GOLD, f"decomposed_{name}_single.npz"))
        assert abs(pobj - float(ipm["obj"])) <= 5e-6 * pobj      # and the interior-point optimum is the same number
    else:
        assert abs(pobj - np.mean(REF_W10_D20_BETA2)) <= 2e-3 * np.mean(REF_W10_D20_BETA2)


def test_chordal_split_of_a_random_banded_matrix():
    """The factorisation behind it on its own: a random negative definite matrix with the pattern of overlapping
    cliques plus a common tail splits into negative definite blocks that add up to it."""
    rng = np.random.default_rng(3)
    n, tail = 40, [38, 39]
    blocks = [np.array(sorted(set(range(s, min(s + 14, 38))) | set(tail))) for s in range(0, 30, 6)]
    M = np.zeros((n, n))
    for B in blocks:
        G = rng.standard_normal((len(B), len(B)))
        M[np.ix_(B, B)] -= G @ G.T + 0.1 * np.eye(len(B))
    parts = sd.chordal_split(M, blocks, eps=1e-6)
    S = np.zeros_like(M)
    for B, P in zip(blocks, parts):
        assert np.linalg.eigvalsh(P).max() <= -0.5e-6
        S[np.ix_(B, B)] += P
    assert np.abs(S - M).max() <= 1e-10

"""CPU restatement of the decision shortcuts of the K6 batch kernel (ellp_b200/csrc/batch.cuh, round 2).

The kernel decides the entering position (primal_simplex_solver.rs:253-292) and the leaving row (:320-400) from ONE reduction pass
that carries (best, second best, position of the best); the reference's sequential tie folds only run when the best is not
"isolated".  Round 1 established isolation with a second pass that counted the candidates within EPS / 2 EPS of the best
(nF == 1 and nBand == 0).  This test pins the claim that replaced it: because the band tests `kmax - k < 2 EPS` and
`l < L + 2 EPS` are monotone in k (l), the count formulation is equivalent to applying the band test to the SECOND best alone --
in IEEE double arithmetic, for ties, near-ties at the EPS scale, duplicates, infinities, and with the entering variable's own
bound-flip ratio as an extra candidate.  No GPU needed: both formulations are evaluated with numpy float64 scalars using exactly
the expressions of the kernel."""
import numpy as np
import pytest

EPS = 1e-10  # kEps of the kernels (device_types.cuh) = EPS of the reference (src/util.rs:1)
INF = np.inf


def _two_pass_max(keys):
    """Round-1 pricing: keys > 0 are candidates, -1 marks none.  Returns (isolated, position) or (False, None)."""
    cand = [(k, j) for j, k in enumerate(keys) if k != -1.0]
    if not cand:
        return None
    kmax = max(k for k, _ in cand)
    nF = nBand = 0
    idxF = None
    for k, j in cand:
        if np.float64(kmax) - np.float64(k) < EPS:
            nF += 1
            idxF = j if idxF is None else min(idxF, j)
        elif np.float64(kmax) - np.float64(k) < 2.0 * EPS:
            nBand += 1
    return (nF == 1 and nBand == 0), idxF


def _top2_max(keys, lanes=7):
    """Round-2 pricing: per-"thread" (best, second, position) folds, merged the way the warp / CTA combines do."""
    parts = []
    for t in range(lanes):
        a1, a2, i1 = 0.0, 0.0, 0
        for j in range(t, len(keys), lanes):
            k = keys[j]
            if k > a1:
                a2, a1, i1 = a1, k, j
            elif k > a2:
                a2 = k
        parts.append((a1, a2, i1))
    g1 = max(p[0] for p in parts)
    if g1 == 0.0:
        return None
    holders = [p for p in parts if p[0] == g1]
    g2 = g1 if len(holders) > 1 else max((p[1] if p[0] == g1 else p[0]) for p in parts)
    isolated = not (np.float64(g1) - np.float64(g2) < 2.0 * EPS)
    return isolated, holders[0][2]


def _two_pass_min(lams, lam_q):
    """Round-1 ratio test with the entering variable's own ratio lam_q (f0 / band0)."""
    cand = [(l, i) for i, l in enumerate(lams) if l is not None]
    finite = [l for l, _ in cand if l < INF]
    lmin = min(finite) if finite else INF
    L = min(lmin, lam_q)
    if not (L < INF):
        return "none", None
    nF = nBand = 0
    idxF = None
    for l, i in cand:
        if l < L + EPS:
            nF += 1
            idxF = i if idxF is None else min(idxF, i)
        elif l < L + 2.0 * EPS:
            nBand += 1
    f0 = 1 if lam_q < L + EPS else 0
    band0 = 1 if (not f0 and lam_q < L + 2.0 * EPS) else 0
    if nF + f0 == 1 and nBand + band0 == 0:
        return ("flip", None) if f0 else ("row", idxF)
    return "fold", None


def _top2_min(lams, lam_q, lanes=5):
    parts = []
    for t in range(lanes):
        r1, r2, ri = INF, INF, 0
        for i in range(t, len(lams), lanes):
            l = lams[i]
            if l is None:
                continue
            if l < r1:
                r2, r1, ri = r1, l, i
            elif l < r2:
                r2 = l
        parts.append((r1, r2, ri))
    lmin = min(p[0] for p in parts)
    L = min(lmin, lam_q)
    if not (L < INF):
        return "none", None
    holders = [p for p in parts if p[0] == lmin]
    l2 = lmin if len(holders) > 1 else min((p[1] if p[0] == lmin else p[0]) for p in parts)
    q_is_min = lam_q < lmin
    second = lmin if q_is_min else (lmin if lam_q == lmin else min(l2, lam_q))
    if not (second < L + 2.0 * EPS):
        return ("flip", None) if q_is_min else ("row", holders[0][2])
    return "fold", None


def _adversarial(rng, n, scale):
    base = rng.uniform(0.5, 2.0) * scale
    v = base - rng.choice([0.0, 0.3 * EPS, 0.999 * EPS, EPS, 1.001 * EPS, 1.5 * EPS, 1.999 * EPS, 2 * EPS, 2.001 * EPS, 5 * EPS, 1e-3 * scale, 0.5 * scale], size=n) * rng.choice([0, 1], size=n, p=[0.2, 0.8])
    return np.abs(v) + 100 * EPS


@pytest.mark.parametrize("seed", range(40))
def test_top2_isolation_equals_the_count_formulation_for_pricing(seed):
    rng = np.random.default_rng(seed)
    for trial in range(200):
        n = int(rng.integers(1, 40))
        keys = _adversarial(rng, n, 10.0 ** rng.integers(-6, 4)) if trial % 2 else rng.uniform(1e-8, 10.0, size=n)
        keys = [float(k) if rng.random() > 0.25 else -1.0 for k in keys]
        if rng.random() < 0.3 and n > 2:  # exact duplicates of the maximum
            m = max(keys)
            if m > 0:
                keys[int(rng.integers(0, n))] = m
        ref, got = _two_pass_max(keys), _top2_max(keys, lanes=int(rng.integers(1, 9)))
        assert (ref is None) == (got is None)
        if ref is None:
            continue
        assert ref[0] == got[0], (keys, ref, got)
        if ref[0]:
            assert ref[1] == got[1], (keys, ref, got)


@pytest.mark.parametrize("seed", range(40))
def test_top2_isolation_equals_the_count_formulation_for_the_ratio_test(seed):
    rng = np.random.default_rng(1000 + seed)
    for trial in range(200):
        n = int(rng.integers(1, 30))
        vals = _adversarial(rng, n, 10.0 ** rng.integers(-3, 3)) if trial % 2 else rng.uniform(0.0, 5.0, size=n)
        lams = []
        for v in vals:
            u = rng.random()
            lams.append(None if u < 0.2 else (INF if u < 0.3 else (0.0 if u < 0.35 else float(v))))
        u = rng.random()
        finite = [l for l in lams if l is not None and l < INF]
        lam_q = INF if u < 0.4 else (0.0 if u < 0.45 else float(rng.uniform(0.0, 5.0)))
        if finite and rng.random() < 0.4:  # the entering variable's own ratio inside the bands of the minimum
            lam_q = max(0.0, min(finite) + float(rng.choice([-2.0, -1.0, -0.5, 0.0, 0.5, 1.0, 1.5, 2.0, 3.0])) * EPS)
        ref, got = _two_pass_min(lams, lam_q), _top2_min(lams, lam_q, lanes=int(rng.integers(1, 7)))
        assert ref == got, (lams, lam_q, ref, got)

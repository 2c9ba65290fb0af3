"""CPU model (numpy) of two pieces of selection logic the GPU kernels of round 2 rely on, checked for SOUNDNESS:

* the large-batch chain's fused tail (gf_gemm.cu collect_tail_kernel): candidates collected above a
  threshold, the best kp_sel by APPROXIMATE score selected, the best K + 32 re-scored exactly, L = the K-th best exact
  score so far, then every further candidate with approx >= L - delta re-scored; certificate: K-th exact > max(threshold,
  approx of the first candidate not re-scored, approx of the kp_sel-th candidate when more were collected) + delta.
  Claim: whenever |approx - exact| <= delta for every row, a CERTIFIED answer is the exact top-K of the whole set.
* the int8 COLLECT epilogue's group test: the per-score test sign(fma(-float(acc), rs, thr)) is implied by
  the group test !(fma(-float_ru(max acc), rs, min thr) > 0) for row scales >= 0.
"""
import numpy as np


def _tail(approx, exact, thr, K, kp_sel, delta):
    """approx / exact: scores of ALL rows; thr: COLLECT threshold.  Returns (top-K row ids, certified)."""
    n = len(approx)
    cand = np.nonzero(approx > thr)[0]
    order = cand[np.argsort(-approx[cand], kind="stable")]
    sel = order[:kp_sel]
    have = len(sel)
    k1 = min(have, K + 32)
    rescored = list(sel[:k1])
    need = have
    if K <= k1 < have:
        L = np.sort(exact[sel[:k1]])[::-1][K - 1]
        cut = L - delta * 1.001
        need = k1 + int(np.count_nonzero(approx[sel[k1:]] >= cut))   # sorted by approx: a prefix
        rescored += list(sel[k1:need])
    rescored = np.array(rescored, dtype=np.int64)
    best = rescored[np.argsort(-exact[rescored], kind="stable")][:K]
    ok = len(cand) >= min(K, n)
    if ok and len(cand) < n:
        ceil = thr
        if need < have:
            ceil = max(ceil, approx[sel[need]])
        elif len(cand) > kp_sel:
            ceil = max(ceil, approx[sel[kp_sel - 1]])
        ok = len(best) >= min(K, n) and exact[best[min(K, n) - 1]] > ceil + delta
    return best, ok


def test_certified_tail_answers_are_exact():
    rng = np.random.default_rng(11)
    certified = refused = 0
    for trial in range(300):
        n = int(rng.integers(2000, 20000))
        K = int(rng.choice([1, 10, 33, 100]))
        delta = float(rng.choice([1e-3, 8e-3, 3e-2]))
        if trial % 3 == 0:      # clustered: many rows within delta of each other at the top
            exact = np.concatenate([rng.normal(0, 0.05, n - 400), 0.9 + rng.uniform(-2 * delta, 2 * delta, 400)])
        else:
            exact = rng.normal(0, 0.05, n)
        exact = exact.astype(np.float64)
        approx = exact + rng.uniform(-delta, delta, n)              # the pass's bound holds for every row
        kp_sel = int(rng.choice([64, 512, 1024]))
        target = int(rng.choice([K + 40, 3 * K + 64, 800]))
        thr = np.sort(approx)[::-1][min(target, n - 1)]
        got, ok = _tail(approx, exact, thr, K, kp_sel, delta)
        if ok:
            certified += 1
            want = np.argsort(-exact, kind="stable")[:K]
            assert np.array_equal(np.sort(exact[got])[::-1], np.sort(exact[want])[::-1]), (trial, n, K, delta)
        else:
            refused += 1
    assert certified > 100 and refused > 0      # the test exercises both outcomes


def test_tail_without_the_second_stage_would_certify_wrong_answers():
    """The teeth of the model: drop the second stage AND its ceiling and a wrong answer gets through."""
    rng = np.random.default_rng(3)
    wrong = 0
    for _ in range(200):
        n, K, delta = 5000, 10, 2e-2
        exact = np.concatenate([rng.normal(0, 0.05, n - 300), 0.9 + rng.uniform(-delta, delta, 300)])
        approx = exact + rng.uniform(-delta, delta, n)
        thr = np.sort(approx)[::-1][600]
        sel = np.argsort(-approx, kind="stable")[:K + 32]
        best = sel[np.argsort(-exact[sel], kind="stable")][:K]
        want = np.argsort(-exact, kind="stable")[:K]
        wrong += not np.array_equal(np.sort(best), np.sort(want))
    assert wrong > 0


def _f32(x):
    return np.float32(x)


def _fma(a, b, c):
    """fma.rn.f32 through float64: a * b is exact in float64 for a 24-bit a and b; the sum is rounded once to float64 and
    once to float32 -- the double rounding cannot change the SIGN, which is all the tests below read."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def test_collect_group_test_is_a_superset_of_the_per_score_test():
    rng = np.random.default_rng(5)
    hits = 0
    for _ in range(20000):
        acc = rng.integers(-3_000_000, 3_000_000, 8).astype(np.int64)
        rs = _f32(10.0 ** rng.uniform(-7, 1)) if rng.random() > 0.05 else _f32(0.0)
        centre = float(acc.max()) * float(rs)
        thr = (centre * rng.uniform(0.7, 1.3, 8) + rng.normal(0, 1e-3, 8)).astype(np.float32)
        if rng.random() < 0.1:
            thr[rng.integers(0, 8)] = np.float32(-0.0)
        per_score = [np.signbit(_fma(-_f32(a), rs, t)) for a, t in zip(acc, thr)]
        m_ru = np.nextafter(_f32(acc.max()), _f32(np.inf)) if float(_f32(acc.max())) < float(acc.max()) else _f32(acc.max())
        tp = _fma(-m_ru, rs, thr.min())
        group = not (tp > 0)
        if any(per_score):
            hits += 1
            assert group, (acc, rs, thr)
    assert hits > 1000

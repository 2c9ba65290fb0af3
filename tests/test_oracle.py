"""CPU tests of the oracle itself: golden vectors of the reference's only test, the
restated util.go helpers, and the internal consistency of the three arithmetic modes."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import MODE_CANONICAL, MODE_F64, MODE_SEQ32
from oracle import lib as ol
from oracle import model

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hex32(h):
    return struct.unpack("<f", bytes.fromhex(h))[0]


@pytest.fixture(scope="module")
def basic():
    with open(os.path.join(GOLD, "basic_search.json")) as f:
        return json.load(f)


def _basic_set(basic):
    c = model.Cache(basic["cache"]["blockNum"], basic["cache"]["blockSize"])
    s = basic["set"]
    c.new_set(s["name"], s["dims"], s["precision"], s["batch"])
    st = c.get_set(s["name"])
    vec = {k: ol.normalize(v) for k, v in basic["raw"].items()}   # search_test.go:87-88
    st.add(np.stack([vec["f1"], vec["f2"]]), ["f1", "f2"])          # search_test.go:98
    return st, vec


@pytest.mark.parametrize("mode", [MODE_CANONICAL, MODE_SEQ32, MODE_F64])
def test_basic_search_golden(basic, mode):
    """K1-K5 of TestBasicSearch (search_test.go:102-144) hold for the oracle in every mode."""
    st, vec = _basic_set(basic)
    for k in basic["pinned"]:
        if "delete" in k:
            assert st.delete(k["delete"]) == k["deleted"]
        call, exp = k["call"], k["expect"]
        ret = st.search(call["threshold"], call["limit"], [vec[n] for n in call["queries"]], mode)
        assert len(ret) == exp["batch"], k["id"]
        assert [len(r) for r in ret] == exp["counts"], k["id"]
        assert [r[0][1] for r in ret] == exp["first_ids"], k["id"]
        for r in ret:
            if "first_score_min" in exp:
                assert np.float32(r[0][0]) >= np.float32(exp["first_score_min"]), k["id"]
            if "first_score_max" in exp:
                assert np.float32(r[0][0]) <= np.float32(exp["first_score_max"]), k["id"]


def test_normalize_matches_fixture(basic):
    for name, key in (("f1", "v1"), ("f2", "v2")):
        got = ol.normalize(basic["raw"][name])
        want = np.array([_hex32(h) for h in basic["oracle"][key]], dtype=np.float32)
        assert got.tobytes() == want.tobytes()
    assert ol.normalize(np.zeros(7)).tolist() == [0.0] * 7          # util.go:131-133


def test_survey_constants(basic):
    """<v1,v1>=1.0000001, <v2,v2>=0.99999994, <v1,v2>=-0.048969507 in sequential fp32."""
    v1 = np.array([_hex32(h) for h in basic["oracle"]["v1"]], dtype=np.float32)
    v2 = np.array([_hex32(h) for h in basic["oracle"]["v2"]], dtype=np.float32)
    assert np.float32(ol.dot(v1, v1, MODE_SEQ32)) == np.float32(1.0000001)
    assert np.float32(ol.dot(v2, v2, MODE_SEQ32)) == np.float32(0.99999994)
    assert np.float32(ol.dot(v1, v2, MODE_SEQ32)) == np.float32(-0.048969507)


def test_canonical_dot_fixture_and_avx2_twin():
    with open(os.path.join(GOLD, "canonical_dots.json")) as f:
        cases = json.load(f)["cases"]
    for c in cases:
        x = np.frombuffer(bytes.fromhex(c["x"]), dtype=np.float32)
        q = np.frombuffer(bytes.fromhex(c["q"]), dtype=np.float32)
        want = np.float32(_hex32(c["canonical"]))
        assert np.float32(ol.dot_canonical_scalar(x, q)) == want
        assert np.float32(ol.dot(x, q, MODE_CANONICAL)) == want       # AVX2 path when dims % 128 == 0
        assert np.float32(ol.dot(x, q, MODE_SEQ32)) == np.float32(_hex32(c["seq32"]))
        assert abs(ol.dot(x, q, MODE_F64) - c["f64"]) < 1e-12
        assert abs(float(want) - c["f64"]) < 1e-5


def test_canonical_order_definition():
    """The canonical order written out in numpy: 128 fma partial sums, fixed tree."""
    rng = np.random.default_rng(3)
    for dims in (5, 130, 512):
        x = rng.uniform(-1, 1, dims).astype(np.float32)
        q = rng.uniform(-1, 1, dims).astype(np.float32)
        P = [np.float32(0)] * 128
        for e in range(dims):  # fma = one rounding of the exact a*b+c; float64 holds a*b exactly
            P[e % 128] = np.float32(np.float64(x[e]) * np.float64(q[e]) + np.float64(P[e % 128]))
        T = [np.float32(np.float32(P[4 * l] + P[4 * l + 1]) + np.float32(P[4 * l + 2] + P[4 * l + 3]))
             for l in range(32)]
        off = 16
        while off >= 1:
            for l in range(off):
                T[l] = np.float32(T[l] + T[l + off])
            off //= 2
        # float64 a*b+c then one rounding to float32 is a correct fma except on double-rounding
        # ties, which these random inputs do not hit; allow 1 ulp on the whole dot
        got = np.float32(ol.dot_canonical_scalar(x, q))
        assert abs(float(got) - float(T[0])) <= abs(float(T[0])) * 2.0 ** -22


def test_exact_arithmetic_modes_agree():
    """Components k/64 with few terms: every partial sum is exact, all three modes agree bit for bit."""
    rng = np.random.default_rng(5)
    x = (rng.integers(-8, 9, (50, 512)) / 64.0).astype(np.float32)
    q = (rng.integers(-8, 9, 512) / 64.0).astype(np.float32)
    for r in x:
        a, b, c = ol.dot(r, q, MODE_CANONICAL), ol.dot(r, q, MODE_SEQ32), ol.dot(r, q, MODE_F64)
        assert a == b == c


def test_maxn_float32_ties_nan_zero():
    v = np.array([0.5, np.nan, 0.5, -0.0, 0.0, 1.0, -np.inf], dtype=np.float32)
    idx, val = ol.maxn_float32(v, 10)                                  # limit > len -> all (util.go:88)
    assert idx.tolist() == [5, 0, 2, 3, 4, 6, 1]                       # ties by index; -0.0 == 0.0; NaN last
    idx, _ = ol.maxn_float32(v, 2)
    assert idx.tolist() == [5, 0]
    assert ol.maxn_float32(np.zeros(0, dtype=np.float32), 3)[0].tolist() == []


def test_transpose1d():
    a = np.arange(6, dtype=np.float32).tobytes()
    b = np.arange(10, 16, dtype=np.float32).tobytes()
    t = np.frombuffer(ol.transpose1d(4, [a, b]), dtype=np.float32)
    assert t.tolist() == [0, 10, 1, 11, 2, 12, 3, 13, 4, 14, 5, 15]   # ret[i*row+j] = features[j][i]
    assert ol.transpose1d(4, []) == b""
    with pytest.raises(ValueError):
        ol.transpose1d(4, [b"12345"])                                  # util.go:26-28


def test_model_add_delete_refill_and_blocks():
    """set.go:77-116 block growth, block.go:93-129 refill order, set.go:163-193 delete walk."""
    c = model.Cache(4, 3 * 4 * 4)            # 3 rows of 4 floats per block
    c.new_set("s", 4, 4, 2)
    s = c.get_set("s")
    rows = np.eye(4, dtype=np.float32)
    s.add(rows, ["a", "b", "c", "d"])        # 2 blocks: [a b c] [d]
    assert [b.next_index for b in s.blocks] == [3, 1]
    assert s.delete(["b", "zz", "d"]) == ["b", "d"]
    assert s.blocks[0].empty == [1] and s.blocks[1].empty == [0]
    s.add(np.ones((2, 4), dtype=np.float32), ["e", "f"])   # refills slot 1 of block 0, slot 0 of block 1
    assert s.blocks[0].ids == ["a", "e", "c"] and s.blocks[1].ids[0] == "f"
    with pytest.raises(model.OracleError, match="ErrNotEnoughBlocks"):
        s.add(np.ones((20, 4), dtype=np.float32), [str(i) for i in range(20)])
    with pytest.raises(model.OracleError, match="ErrOutOfBatch"):
        s.search(0, 1, np.ones((3, 4), dtype=np.float32))
    with pytest.raises(model.OracleError, match="ErrFeatureSetExist"):
        c.new_set("s", 4, 4, 2)
    c.destroy_set("s")
    assert all(not b.is_owned() for b in c.all_blocks)
    with pytest.raises(model.OracleError, match="ErrFeatureSetNotFound"):
        c.get_set("s")


def test_tombstone_quirk_in_model():
    """block.go:236-243: zeroed tombstones compete for the block's top-limit before being dropped."""
    c = model.Cache(2, 4 * 2 * 4)            # 4 rows of 2 floats per block
    c.new_set("s", 2, 4, 1)
    s = c.get_set("s")
    rows = np.array([[1, 0], [-1, 0], [-0.5, 0], [0.5, 0]], dtype=np.float32)
    s.add(rows, ["p", "n1", "n2", "h"])
    s.delete(["h"])                          # tombstone scores exactly 0.0 > n2 > n1
    r = s.search(-2, 2, np.array([[1, 0]], dtype=np.float32))
    assert [i for _, i in r[0]] == ["p"]     # block top-2 = {p, tombstone}; n2 never surfaces
    r = s.search(-2, 3, np.array([[1, 0]], dtype=np.float32))
    assert [i for _, i in r[0]] == ["p", "n2"]


def test_scan_topk_matches_set_search_and_threads():
    X = ol.synth_rows(11, 0, 3000, 128)
    Q = ol.synth_rows(12, 0, 2, 128)
    one = ol.scan_topk(X, Q, 7, MODE_CANONICAL, 1)
    many = ol.scan_topk(X, Q, 7, MODE_CANONICAL, 5)
    for a, b in zip(one, many):
        assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist()
    ref = ol.set_search([X], [3000], [np.ones(3000, dtype=np.uint8)], Q, 7, -2.0, MODE_CANONICAL)
    for a, r in zip(one, ref):
        assert a[1].tolist() == [row for _, _, row in r]
        assert a[0].tolist() == [sc for sc, _, _ in r]


def test_synth_rows_are_deterministic_and_normalised():
    a = ol.synth_rows(20260101, 5, 4, 512)
    b = ol.synth_rows(20260101, 0, 9, 512)[5:]
    assert a.tobytes() == b.tobytes()
    assert np.allclose(np.linalg.norm(a.astype(np.float64), axis=1), 1.0, atol=1e-6)
    raw = ol.synth_rows(20260101, 0, 4, 16, normalize=False)
    assert raw.min() >= -1.0 and raw.max() < 1.0

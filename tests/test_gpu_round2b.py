"""GPU parity of the kernels reworked late in round 2, each against a FULL oracle scan (ids, order, float32 scores bit
for bit):

* the large-batch chain's fused tail (collect_tail_kernel: selection, two-stage exact re-score whose depth comes from the
  K-th EXACT score, top-K, certificate -- one launch for up to 1 024 queries) at the corners of its ranges: K = 1, K just
  above the one-pass limit, K = 256 and K = 512 (selection depth 1 024), batches that end in a padded pass, more than
  one tail group (> 1 024 queries);
* the int8 COLLECT epilogue's group test (maximum of 8 accumulators against the smallest of 8 thresholds) with rows whose
  scales differ by orders of magnitude, zero rows and negative-only scores;
* the exact scan's owner-computes candidate lists + mailboxes on rows SORTED by score, where every row is a candidate
  until the very end and every mailbox runs full -- the case that deadlocked the first two versions: an
  owner waiting for a tile whose stage a blocked sender still held, and two owners spinning on each other's unwritten
  slots -- for 1 .. 8 queries, once per shape and 40 times in a row for the shapes with mutual owners.
"""
import os

import numpy as np
import pytest

from oracle import lib as ol

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


def _set(gpu_ctx, name, dims, rows, batch, block_rows=16384, shadow="int8"):
    from gofeature_b200 import api
    cache = api.NewCache(gpu_ctx, (len(rows) + block_rows - 1) // block_rows + 1, block_rows * dims * 4, shadow=shadow)
    cache.NewSet(name, dims, 4, batch)
    st = cache.GetSet(name)
    st.AddArray(rows, ["r%07d" % i for i in range(len(rows))])
    return cache, st


def _check(got, want):
    for qi, (g, (sc, idx)) in enumerate(zip(got, want)):
        assert [r.ID for r in g] == ["r%07d" % i for i in idx], "query %d ids" % qi
        assert np.array([r.Score for r in g], dtype=np.float32).tobytes() == np.asarray(sc, dtype=np.float32).tobytes(), \
            "query %d scores" % qi


@pytest.mark.parametrize("q,limit", [(70, 1), (70, 33), (300, 100), (64, 256), (40, 512), (1100, 10)])
def test_fused_tail_ranges(gpu_ctx, q, limit):
    dims, n = 128, 150_000
    rows = ol.synth_rows(9090, 0, n, dims)
    cache, st = _set(gpu_ctx, "tail", dims, rows, 2048)
    queries = ol.synth_rows(17 + q, 0, q, dims)
    queries[0] = rows[n - 1]
    want = ol.scan_topk(rows, queries, limit, ol.MODE_CANONICAL, THREADS)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-2.0, limit, queries)
    s = gpu_ctx.Stats()
    assert s["tensor_passes"] >= (q + 255) // 256, s          # the tcgen05 chain answered, not the scan
    _check(got, want)
    # (K = 512 sits at the edge of the design -- the selection keeps 1 024 candidates -- and may refuse a few queries,
    # which the exact scan then answers)
    assert s["uncertified_queries"] <= (q // 4 if limit >= 512 else max(1, q // 50)), s
    cache.Close()


def test_collect_group_test_with_mixed_row_scales(gpu_ctx):
    """Row norms from 1/4 to 4 in one set (int8 row scales differ 16x), zero rows (scale 0) and queries whose best
    scores are negative: the group test (max of 8 int32 accumulators x row scale against the smallest threshold of the
    group) must stay a superset of the per-score test whatever the signs and magnitudes -- the answers stay exact and
    the tensor chain still certifies nearly all of them."""
    dims, n = 128, 140_000
    rng = np.random.default_rng(5)
    rows = ol.synth_rows(31, 0, n, dims)
    mag = (2.0 ** rng.uniform(-2, 2, n)).astype(np.float32)
    rows = (rows * mag[:, None]).astype(np.float32)
    rows[rng.integers(0, n, 500)] = 0.0
    q = 96
    queries = ol.synth_rows(41, 0, q, dims)
    queries[1] = -rows[7] / np.linalg.norm(rows[7])
    cache, st = _set(gpu_ctx, "mixed", dims, rows, 256)
    want = ol.scan_topk(rows, queries, 10, ol.MODE_CANONICAL, THREADS)
    gpu_ctx.ResetStats()
    got = st.SearchArray(-3.0e38, 10, queries)
    s = gpu_ctx.Stats()
    _check(got, want)
    assert s["tensor_passes"] >= 1 and s["uncertified_queries"] <= q // 4, s
    cache.Close()


@pytest.mark.parametrize("q", [1, 2, 4, 8])
@pytest.mark.parametrize("dims", [128, 512])
def test_scan_on_sorted_rows(gpu_ctx, q, dims, repeats=1):
    n, limit = 60_000, 10
    rows = ol.synth_rows(777, 0, n, dims)
    queries = ol.synth_rows(778, 0, q, dims)
    # ascending by the first query's score: every row beats everything before it, so every row is a candidate
    order = np.argsort(rows.astype(np.float64) @ queries[0].astype(np.float64), kind="stable")
    rows = np.ascontiguousarray(rows[order])
    for k in range(1, q):                            # the other queries: near the first one, same ordering pressure
        v = queries[0] + 0.05 * queries[k]
        queries[k] = (v / np.linalg.norm(v)).astype(np.float32)
    cache, st = _set(gpu_ctx, "sorted", dims, rows, 16, shadow=False)
    st.SetSearchPath(1)                              # the fused exact scan
    want = ol.scan_topk(rows, queries, limit, ol.MODE_CANONICAL, THREADS)
    gpu_ctx.ResetStats()
    for _ in range(repeats):
        got = st.SearchArray(-2.0, limit, queries)
        _check(got, want)
    assert gpu_ctx.Stats()["scan_launches"] >= repeats
    cache.Close()


@pytest.mark.parametrize("q,dims", [(2, 128), (4, 128), (8, 128), (4, 256), (2, 512)])
def test_scan_on_sorted_rows_many_times(gpu_ctx, q, dims):
    test_scan_on_sorted_rows(gpu_ctx, q, dims, repeats=40)

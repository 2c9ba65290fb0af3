#!/usr/bin/env python
"""bench.py -- search QPS of the brute-force cosine top-K hot path on B200.

Headline workload (BASELINE.json configs[1]): 1 000 000 x 512-d L2-normalised float32 features per GPU
(2.048 GB, 62 blocks of 32 MiB = demo geometry, demo/main.go:19-23), top-10, 10 concurrent
single-query searchers.  One "step" serves one query from each of the 10 searchers.

  value     queries/s with the 10 queries already resident in HBM (gf_set_search_device_ex: query
            preparation, the tcgen05 candidate pass over the int8 shadow slab, the tail kernel that re-scores
            the candidates from the float32 rows and certifies them; at N > 1 also the exchange of 10 keys
            per query per rank over NVLink peer memory and the cross-rank merge), CUDA-event timed, max
            over ranks.  The driver's K steps are repeated back to back until >= 1.2 s of device time has been
            measured (every repetition bracketed by its own events); value / ms_per_step are the MEDIAN
            repetition.  Weak scaling: every rank holds its own 1 M-row shard of an N-million-row set, and a
            query is counted once per 1 M-row shard it scans (N per query), so value is the 1 M-row-equivalent
            QPS of the whole job; `qps_user` is the plain end-user QPS.
  e2e       the same through the public host-buffer call: N = 1: 10 real host threads x gf_set_search(q=1),
            coalesced by the library (tools/gf_demo); N > 1: one caller, gf_set_search(q=10) on a set the library
            itself shards over the N devices of the process (gf_context_create_multi) when available, else the
            collective ShardedSet.search_keys.  Queries are read from host memory and winners written back to
            it inside the timed region.
  parity_check
            after the timed regions every batch of the query pool is answered again by the exact float32
            scan kernel (path 1) and compared key for key with what the timed path returned; stored rows owned
            by every rank are searched as needles.  mismatches must be 0 (bench exits non-zero otherwise).
  roofline  the dominant kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json).  With the
            int8 shadow slab (default) that kernel is the tcgen05 kind::i8 candidate pass and its algorithmic
            bytes are the SHADOW bytes it must stream (rows x dims x 1 + one float scale per row), not the
            float32 bytes: the fraction says how well the kernel uses HBM, the smaller traffic shows up in
            `value`.  `float32_paths` times the same step WITHOUT the shadow (kind::tf32 over the float32
            rows, and the fused exact scan), whose algorithmic bytes are SURVEY 8(d)'s rows x dims x 4.
  legs      config3 (1024 queries x top-100, tensor roofline), config4_shard (12.5 M rows per GPU, 64 queries:
            at --gpus 8 the 100 M-row set), strong_scaling (8 M rows split over the ranks), churn (10 M rows,
            Add / Delete 100 k rows under concurrent searches), baselines (the reference's own cuBLAS path
            replayed on this GPU, a 1-thread CPU scan, numpy/BLAS) -- extra keys of the same JSON line.
  cpu_baseline / --impl reference
            the oracle's plain multi-threaded CPU scan ("port": the reference has no CPU build and
            there is no Go toolchain) over the FULL 1 M rows, persistent thread pool, all host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = 512
BLOCK_SIZE = 1024 * 1024 * 32      # demo/main.go:19
CAP = BLOCK_SIZE // (DIMS * 4)
SEED = 20260101
QUERY_SEED = 20260102
UNIT = "queries/s"
METRIC = "search_qps_1Mx512_top10_parallel10"
MIN_TIMED_S = 1.2


def make_queries(n, dims, seed):
    """U(-1,1) rows (search_test.go:65) normalised like util.go:125-139, float32."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-1.0, 1.0, (n, dims)).astype(np.float32)
    q /= np.sqrt((q.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
    return np.ascontiguousarray(q)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7]
        inside = len(rows)
        if not rows:
            rows = [r for t, r in self.rows if t0 - 0.1 <= t <= t1 + 0.2 and len(r) >= 7] or \
                   [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "samples_inside_timed_region": inside,
                "power_w_max": max(float(r[2]) for r in rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm": float(d["hbm_gbs"]), "hbm_src": "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "bf16": float(d.get("bf16_tflops", 1642.4)),
                "bf16_src": "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst)"}
    return {"hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md 6.65 TB/s)", "bf16": 1650.0,
            "bf16_src": "fallback (B200_PROFILING.md)"}


def traffic_from_profile(eb):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture of the same
    kernel (eb = bytes per element it streams: 4 float32 rows, 2 bf16 shadow, 1 int8 shadow), if any."""
    p = os.path.join(ROOT, "profiles", {4: "scan_traffic.json", 2: "shadow_traffic.json", 1: "shadow_i8_traffic.json"}[eb])
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


# --------------------------------------------------------------------------- CPU arms (the oracle: checker / baseline only)
def cpu_port_qps(rows, nq, limit, min_seconds, threads, max_passes=400):
    """The oracle's plain scan over `rows` rows (persistent thread pool) -> (qps, ms per pass, passes)."""
    from oracle import lib as ol
    X = ol.synth_rows(SEED, 0, rows, DIMS)
    Q = make_queries(nq, DIMS, QUERY_SEED)
    ol.scan_topk(X[:4096], Q, limit, ol.MODE_CANONICAL, threads)  # warm: the pool's threads exist
    ol.scan_topk(X, Q, limit, ol.MODE_CANONICAL, threads)
    t0 = time.perf_counter()
    times = []
    while True:
        a = time.perf_counter()
        ol.scan_topk(X, Q, limit, ol.MODE_CANONICAL, threads)
        times.append(time.perf_counter() - a)
        if (time.perf_counter() - t0 >= min_seconds and len(times) >= 3) or len(times) >= max_passes:
            break
    per_pass = float(np.median(times))
    return nq / per_pass, per_pass * 1e3, len(times), X, Q


def cpu_blas_qps(X, Q, limit, min_seconds):
    """'Best library CPU': numpy (OpenBLAS sgemm, all cores) scores + argpartition + sort of the K best."""
    def one():
        S = X @ Q.T                                           # [rows, q]
        idx = np.argpartition(-S, limit - 1, axis=0)[:limit]  # [limit, q]
        top = np.take_along_axis(S, idx, axis=0)
        order = np.argsort(-top, axis=0, kind="stable")
        return np.take_along_axis(idx, order, axis=0)
    one()
    t0, times = time.perf_counter(), []
    while True:
        a = time.perf_counter()
        one()
        times.append(time.perf_counter() - a)
        if (time.perf_counter() - t0 >= min_seconds and len(times) >= 3) or len(times) >= 200:
            break
    per = float(np.median(times))
    return Q.shape[0] / per, per * 1e3, len(times)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path.  goFeature has no non-cublas build and this image
    has no Go toolchain, so this is the oracle port (plain brute-force scan, all host threads, persistent
    pool) over the full workload: every step scans all rows_per_gpu rows for the step's queries."""
    if rank != 0:
        return
    from oracle import lib as ol
    cores = os.cpu_count() or 1
    rows = args.rows_per_gpu
    X = ol.synth_rows(SEED, 0, rows, DIMS)
    Q = make_queries(args.queries, DIMS, QUERY_SEED)
    ol.scan_topk(X[:4096], Q, args.limit, ol.MODE_CANONICAL, cores)
    for _ in range(max(args.warmup, 1)):
        ol.scan_topk(X, Q, args.limit, ol.MODE_CANONICAL, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ol.scan_topk(X, Q, args.limit, ol.MODE_CANONICAL, cores)
    dt = time.perf_counter() - t0
    ms_step = dt / args.steps * 1e3
    # same unit as the GPU arm (1 M-row shard scans per second): the host cores are one pool, so it does not grow with N
    value = args.queries / (ms_step / 1e3)
    sample_txt = ("each step scans all %d rows x %d-d for %d queries top-%d with %d host threads (persistent pool); "
                  "%d steps, %.1f s" % (rows, DIMS, args.queries, args.limit, cores, args.steps, dt))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_txt},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def bench_config(args, world):
    fmt = "none" if getattr(args, "no_shadow", False) else getattr(args, "shadow", "int8")
    return {"shadow": {"int8": "int8 shadow slab + one float scale per row (+26% HBM per row), candidates re-scored "
                               "from the float32 rows and certified against the measured quantisation residuals",
                       "bf16": "bf16 shadow slab (+50% HBM per row), candidates re-scored from the float32 rows",
                       "none": "none (float32 rows only)"}[fmt],
            "workload": "BASELINE configs[1]: %d x %d-d normalised float32 rows per GPU (%d blocks of 32 MiB), "
                        "top-%d, %d concurrent single-query searchers per step"
                        % (args.rows_per_gpu, DIMS, (args.rows_per_gpu + CAP - 1) // CAP, args.limit, args.queries),
            "rows_per_gpu": args.rows_per_gpu, "rows_total": args.rows_per_gpu * world, "dims": DIMS,
            "limit": args.limit, "queries_per_step": args.queries, "sharding": "row-sharded x%d" % world,
            "l2": "inputs exceed L2 (0.51 GB int8 shadow / 2.05 GB float32 rows per GPU streamed per pass vs 126 MB L2)"}


# --------------------------------------------------------------------------- GPU arm
class Harness:
    """One rank's view of the job: torch.distributed plumbing + event timing helpers."""

    def __init__(self, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.host_group = None
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.host_group = dist.new_group(backend="gloo")

    def host_barrier(self):
        """ranks meet on the HOST (gloo): a NCCL barrier would park a spinning kernel on every GPU, and rank 0's
        in-library multi-device leg needs those GPUs."""
        if self.world > 1:
            self.dist.barrier(group=self.host_group)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, xs):
        """element-wise max over ranks of a list of floats."""
        if self.world == 1:
            return list(xs)
        t = self.torch.tensor(list(xs), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.cpu().tolist()

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed_reps(self, step, steps, min_seconds, max_reps=20000):
        """`steps` calls of step(i) per repetition, repetitions back to back until min_seconds of device time;
        every repetition has its own CUDA events.  Returns (per-repetition ms, max over ranks; wall marks)."""
        torch = self.torch
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            step(i)
        b.record()
        torch.cuda.synchronize()
        first = self.max_over_ranks([a.elapsed_time(b)])[0]
        reps = int(min(max_reps, max(1, np.ceil(min_seconds * 1e3 / max(first, 1e-3)))))
        self.barrier()
        t0 = time.time()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        evs[0].record()
        k = 0
        for r in range(reps):
            for i in range(steps):
                step(k)
                k += 1
            evs[r + 1].record()
        self.barrier()
        t1 = time.time()
        ms = self.max_over_ranks([evs[r].elapsed_time(evs[r + 1]) for r in range(reps)])
        return ms, t0, t1


def kernel_ms(stats):
    if stats["tensor_passes"] > 0:
        return stats["gemm_kernel_ns"] / 1e6 / max(stats["gemm_kernel_timed"], 1), stats["gemm_kernel_timed"], True
    return stats["scan_kernel_ns"] / 1e6 / max(stats["scan_kernel_timed"], 1), stats["scan_kernel_timed"], False


def profile_kernel(ctx, fn, n):
    """n calls of fn() with CUDA events around every launch of the dominant kernel (the library brackets them on
    the launching stream; profiling runs the pipeline un-graphed, kernel for kernel)."""
    import torch
    ctx.ResetStats()
    ctx.SetProfiling(True)
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    st = ctx.Stats()
    ctx.SetProfiling(False)
    ms, launches, tensor = kernel_ms(st)
    return ms, launches / float(n), tensor, st


def new_sharded(H, api, ShardedSet, ctx, rows_total, batch, shadow, name):
    gblocks = (rows_total + CAP - 1) // CAP
    cache = api.NewCache(ctx, (gblocks + H.world - 1) // H.world + 1, BLOCK_SIZE, shadow=shadow)
    sset = ShardedSet(ctx, cache, name, DIMS, batch, H.rank, H.world)
    sset.AddSynthetic(rows_total, SEED)
    return cache, sset


def parity_check(H, sset, pool_dev, K, Q, rows_total):
    """Every batch of the pool: the default path's keys against the exact float32 scan's (path 1), key for key;
    plus needles -- one stored row owned by each rank must come back first with score ~1."""
    torch = H.torch
    npool = pool_dev.shape[0]
    mism = flagged = 0
    for b in range(npool):
        k0 = sset.search_device(K, Q, pool_dev[b]).clone()
        f0 = sset.flags(Q, K).clone()
        k1 = sset.search_device(K, Q, pool_dev[b], path=1).clone()
        torch.cuda.synchronize()
        bad = (k0 != k1).any(dim=1)
        fl = f0 != 0
        flagged += int(fl.sum().item())
        mism += int((bad & ~fl).sum().item())       # a flagged query is re-run by the caller; an unflagged one must be exact
    # needles: rank r reads a row it owns back through the C ABI (gf_set_read) and everyone searches for it
    needles = torch.zeros((H.world, DIMS), dtype=torch.float32, device="cuda")
    ordinals = []
    for r in range(H.world):
        g = r + H.world * max(0, ((rows_total // CAP) // H.world) // 2)      # a global block owned by rank r
        o = min(g * CAP + 77, rows_total - 1)
        if (o // CAP) % H.world != r:
            o = r * CAP + 77 if r * CAP + 77 < rows_total else None
        ordinals.append(o)
        if o is not None and r == H.rank:
            feats = sset.local.Read("f%011d" % o)
            if feats:
                needles[r] = torch.from_numpy(np.frombuffer(feats[0].Value, dtype=np.float32).copy()).cuda()
    if H.world > 1:
        H.dist.all_reduce(needles)
    ok = True
    from gofeature_b200 import api
    for r0 in range(0, H.world, Q):
        nb = min(Q, H.world - r0)
        qb = torch.zeros((Q, DIMS), dtype=torch.float32, device="cuda")
        qb[:nb] = needles[r0:r0 + nb]
        qb[nb:] = pool_dev[0][:Q - nb]
        keys = sset.search_device(K, Q, qb).clone()
        torch.cuda.synchronize()
        hk = keys.cpu().numpy().view(np.uint64)
        for j in range(nb):
            o = ordinals[r0 + j]
            if o is None:
                continue
            ok = ok and api.key_slot(hk[j][0]) == o and abs(float(api.key_score(hk[j][0])) - 1.0) < 1e-5
    return {"batches": npool, "queries": npool * Q, "mismatches": mism, "flagged_queries": flagged,
            "needles": sum(1 for o in ordinals if o is not None), "needles_ok": bool(ok),
            "against": "exact float32 scan kernel (path 1) on the same resident rows; both paths are bit-exact "
                       "against the CPU oracle in tests/test_gpu_configs.py"}


def run_tool(cmd, timeout):
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        if out.returncode == 0 and out.stdout.strip():
            return json.loads(out.stdout.strip().splitlines()[-1])
        return {"error": "rc=%d %s" % (out.returncode, (out.stderr or "")[-300:])}
    except Exception as exc:  # the bench line must still come out
        return {"error": str(exc)}


def run_ours(args, rank, world, local_rank):
    import torch

    from gofeature_b200 import _capi, api
    from gofeature_b200.sharded import ShardedSet

    _capi.lib()  # raises if the CUDA library is missing: no fallback
    H = Harness(rank, world, local_rank)
    PK = peaks()
    Q, K = args.queries, args.limit
    rows_total = args.rows_per_gpu * world
    ctx = api.NewContext(local_rank)
    shadow = False if args.no_shadow else args.shadow
    cache, sset = new_sharded(H, api, ShardedSet, ctx, rows_total, max(Q, 16, 1024 if world == 1 else 16), shadow, "bench")
    shadow_eb = cache.ShadowBytesPerElement()
    rows_local = sset.local.ScannedRows()

    npool = 32
    pool = make_queries(npool * Q, DIMS, QUERY_SEED).reshape(npool, Q, DIMS)
    dpool = torch.from_numpy(pool).cuda()
    hpool = torch.from_numpy(pool).pin_memory()
    lean = api.lib()
    handle = ctypes.c_void_p()

    def step_device(i):
        return sset.search_device(K, Q, dpool[i % npool])

    def step_e2e(i):
        if world == 1:
            hq = hpool[i % npool]
            rc = lean.gf_set_search(sset.local._h, -2.0, K, Q, hq.data_ptr(), Q * DIMS * 4, ctypes.byref(handle))
            if rc != 0:
                raise api.GoFeatureError(rc)
            n = lean.gf_result_total(handle)
            lean.gf_result_free(handle)
            return n
        return sset.search_keys(K, pool[i % npool])

    # ---- warm-up (>= 3), then the device-resident timed region
    W = max(args.warmup, 3)
    for i in range(W):
        step_device(i)
        step_e2e(i)
    H.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25 if rank == 0 else 0.0)
    flagged0 = sset.local.DeviceCounters()[0]
    ctx.ResetStats()
    rep_ms, t_mark0, t_mark1 = H.timed_reps(step_device, args.steps, MIN_TIMED_S)
    flagged_timed = H.sum_over_ranks(sset.local.DeviceCounters()[0] - flagged0)
    launches_timed = ctx.Stats()
    launches = (launches_timed["scan_launches"] + launches_timed["merge_launches"] +
                launches_timed["other_launches"] + launches_timed["gemm_launches"])
    reps = len(rep_ms)
    med_ms = float(np.median(rep_ms))
    scan_ms, per_step, tensor, stats = profile_kernel(ctx, step_device, max(args.steps, 20))
    H.barrier()

    # ---- where the cross-rank step spends its time (device clock, per rank; the wait includes the skew between ranks)
    xtrace = None
    if world > 1:
        ctx.SetProfiling(True)
        rows_ns = []
        for i in range(40):
            for j in range(12):          # back to back: the ranks fall into lock step, the last step is the sample
                step_device(i * 12 + j)
            tr = sset.exchange_trace(Q)
            if tr is None:
                break
            if i >= 5:
                t = tr.astype(np.int64)
                rows_ns.append([float((t[:, 1] - t[:, 0]).mean()), float((t[:, 2] - t[:, 1]).mean()),
                                float((t[:, 3] - t[:, 2]).mean()), float(t[:, 3].max() - t[:, 0].min())])
        ctx.SetProfiling(False)
        if rows_ns:
            m = np.median(np.array(rows_ns), axis=0) / 1e3
            worst = H.max_over_ranks(list(m))
            xtrace = {"push_publish_us": m[0], "wait_for_peers_us": m[1], "fold_us": m[2], "first_enter_to_last_done_us": m[3],
                      "max_over_ranks": {"push_publish_us": worst[0], "wait_for_peers_us": worst[1], "fold_us": worst[2],
                                         "first_enter_to_last_done_us": worst[3]},
                      "how": "globaltimer stamps inside the tail kernel's exchange (rank 0's medians over 35 samples, each the "
                             "last of 12 back-to-back steps)"}
    H.barrier()

    # ---- correctness of what was just timed, at this N
    parity = parity_check(H, sset, dpool, K, Q, rows_total)

    # ---- e2e timed region (host buffers, H2D + D2H inside), same >= 1 s rule
    def e2e_run(n, first, lat_out):
        H.barrier()
        t0 = time.perf_counter()
        for i in range(n):
            a = time.perf_counter()
            step_e2e(first + i)
            lat_out.append(time.perf_counter() - a)
        torch.cuda.synchronize()
        H.barrier()
        return H.max_over_ranks([(time.perf_counter() - t0) * 1e3])[0]

    probe_ms = e2e_run(args.steps, 0, [])
    n_e2e = int(min(20000, max(args.steps, np.ceil(MIN_TIMED_S * 1e3 / max(probe_ms, 1e-3) * args.steps))))
    lat = []
    ctx.ResetStats()
    e2e_ms = e2e_run(n_e2e, args.steps, lat)
    hs = ctx.Stats()
    host_breakdown = None
    if hs["host_calls"]:
        c = float(hs["host_calls"])
        host_breakdown = {"stage_queries_us": hs["host_stage_ns"] / c / 1e3, "enqueue_us": hs["host_launch_ns"] / c / 1e3,
                          "wait_for_device_us": hs["host_wait_ns"] / c / 1e3, "resolve_ids_us": hs["host_resolve_ns"] / c / 1e3,
                          "calls": int(c), "note": "inside gf_set_search; the rest of a call is the caller's own (ctypes) overhead"}

    line_extra = {}
    # ---- N > 1: the same N x 1 M-row set INSIDE the library (one process, N devices, gf_context_create_multi):
    # rank 0 drives all N GPUs through the reference-facing host call; the other ranks wait on the host
    multi = None
    if world > 1:
        ref_keys = [sset.search_device(K, Q, dpool[b]).clone().cpu().numpy().view(np.uint64) for b in range(4)]
        torch.cuda.synchronize()
        H.host_barrier()
        if rank == 0:
            try:
                multi = leg_multi(api, world, rows_total, shadow, pool, ref_keys, Q, K, args.steps)
            except Exception as exc:   # the bench line must still come out
                multi = {"error": repr(exc)}
        H.host_barrier()
    lat1, q1 = [], None
    f32_paths = {}
    if world == 1:
        # ---- single-query latency through the host call (north_star: sub-0.5 ms p50)
        for i in range(300):
            hq = hpool[i % npool][:1]
            a = time.perf_counter()
            lean.gf_set_search(sset.local._h, -2.0, K, 1, hq.data_ptr(), DIMS * 4, ctypes.byref(handle))
            lat1.append(time.perf_counter() - a)
            lean.gf_result_free(handle)
        q1_ms, _, q1_tensor, _ = profile_kernel(ctx, lambda i: sset.search_device(K, 1, dpool[i % npool][:1]), 20)
        q1 = (q1_ms, q1_tensor)
        # ---- the same step on the float32 rows only (no shadow traffic saved): kind::tf32 candidate pass, and the
        # fused exact scan -- SURVEY 8(d)'s algorithmic bytes (rows x dims x 4) apply to these as they stand
        for name, path in (("tf32_candidate_pass", 3), ("exact_scan", 1)):
            fn = lambda i, p=path: sset.search_device(K, Q, dpool[i % npool], path=p)   # noqa: E731
            for i in range(3):
                fn(i)
            ms, _, _ = H.timed_reps(fn, 20, 0.3)
            fms = float(np.median(ms)) / 20
            kms, per, _, _ = profile_kernel(ctx, fn, 20)
            fbytes = rows_local * DIMS * 4 + (Q / max(per, 1)) * DIMS * 4
            f32_paths[name] = {"ms_per_step": fms, "value": Q / (fms * 1e-3), "unit": UNIT, "kernel_ms": kms,
                               "kernel_launches_per_step": per,
                               "achieved_gbs": fbytes / (kms * 1e-3) / 1e9 if kms > 0 else None,
                               "roofline_frac": (fbytes / (kms * 1e-3) / 1e9 / PK["hbm"]) if kms > 0 else None}
        # the exact scan kernel by batch size (one launch holds 1 / 2 / 4 / 8 queries; 10 queries = an 8 and a 2)
        by_q = {}
        for qq in (1, 2, 4, 8):
            fn = lambda i, n=qq: sset.search_device(K, n, dpool[i % npool][:n], path=1)   # noqa: E731
            for i in range(3):
                fn(i)
            kms, per, _, _ = profile_kernel(ctx, fn, 20)
            fbytes = rows_local * DIMS * 4 + qq * DIMS * 4
            by_q[str(qq)] = {"kernel_ms": kms, "achieved_gbs": fbytes / (kms * 1e-3) / 1e9 if kms > 0 else None,
                             "roofline_frac": (fbytes / (kms * 1e-3) / 1e9 / PK["hbm"]) if kms > 0 else None}
        f32_paths["exact_scan"]["kernel_by_queries_per_launch"] = by_q
        # ---- config 3: 1024 queries x top-100 on the same 1 M rows (tensor roofline)
        line_extra["config3"] = leg_config3(H, ctx, sset, rows_local, shadow_eb, PK)
    sset.close_exchange()
    cache.Close()
    cache = None

    if world == 1 and not args.quick:
        line_extra["clustered"] = leg_clustered(H, api, ctx, args.rows_per_gpu, Q, K, shadow)
    # ---- config 4: 12.5 M rows per GPU, 64 queries top-10 (8 GPUs: the 100 M-row set); strong scaling: 8 M rows
    if not args.quick:
        line_extra["config4_shard"] = leg_rows(H, api, ShardedSet, ctx, args.config4_rows_per_gpu * world, 64, 10,
                                               shadow, PK, "config4")
        if args.strong_rows >= world * 2 * CAP:
            line_extra["strong_scaling"] = leg_rows(H, api, ShardedSet, ctx, args.strong_rows, Q, K, shadow, PK, "strong")

    # ---- config[1] exactly as written: 10 concurrent host threads, one query each, through the C ABI
    # (tools/gf_demo, the counterpart of the reference's demo/main.go), request coalescing on
    demo = churn = None
    baselines = {}
    if world == 1 and rank == 0:
        tools = os.path.join(ROOT, "tools")
        if not (os.path.exists(os.path.join(tools, "gf_demo")) and os.path.exists(os.path.join(tools, "ref_replica"))):
            subprocess.call(["make", "-C", tools], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        exe = os.path.join(tools, "gf_demo")
        if os.path.exists(exe):
            # three runs of >= 1 s each, the MEDIAN run reported (all three listed): a batch shape the coalescer forms
            # for the first time on a workspace records its CUDA graph inside the timed rounds (tens of ms, once)
            demo_runs = [run_tool([exe, "--rows", str(args.rows_per_gpu), "--threads", str(Q), "--rounds", "10000",
                                   "--warm-rounds", str(max(args.warmup, 30)), "--limit", str(K), "--coalesce-us", "100",
                                   "--device", str(local_rank)], 300) for _ in range(3)]
            ok_runs = sorted((r for r in demo_runs if "qps" in r), key=lambda r: r["qps"])
            demo = ok_runs[len(ok_runs) // 2] if ok_runs else demo_runs[0]
            if ok_runs:
                demo["runs_qps"] = [r["qps"] for r in demo_runs if "qps" in r]
            if not args.quick:
                # config 5: 10 M rows, 4 searchers, Add / Delete 100 k rows in 10 k chunks meanwhile -- and the same
                # searchers on the same set without the churn
                common = [exe, "--rows", "10000000", "--threads", "4", "--limit", str(K), "--coalesce-us", "100",
                          "--device", str(local_rank), "--warm-rounds", "20"]
                churn = {"idle": run_tool(common + ["--rounds", "1500"], 300),
                         "churn": run_tool(common + ["--rounds", "1500", "--adds", "10", "--add-size", "10000",
                                                     "--deletes", "10", "--delete-size", "10000", "--add-gap-us",
                                                     "20000"], 300)}
        rep = os.path.join(tools, "ref_replica")
        if os.path.exists(rep) and not args.quick:
            # the reference's own path on this GPU: cublasSgemm per block + full D2H + CPU sort (block.go:196-246)
            baselines["ref_replica_gpu"] = run_tool([rep, "--rows", str(args.rows_per_gpu), "--threads", str(Q),
                                                     "--queries", "1", "--limit", str(K), "--seconds", "4",
                                                     "--device", str(local_rank)], 120)
            baselines["ref_replica_gpu"]["what"] = ("tools/ref_replica: the reference's Set.Search replayed with the "
                                                    "image's cuBLAS -- per block H2D of the transposed query, "
                                                    "cublasSgemm, D2H of the whole score matrix, full CPU sort "
                                                    "(block.go:196-246, util.go:76-93); %d concurrent callers" % Q)
    if sampler:
        sampler.stop()

    if rank != 0:
        ctx.Close()
        if world > 1:
            H.dist.destroy_process_group()
        return

    avg_q = Q / per_step if per_step else Q
    eb = shadow_eb if (tensor and shadow_eb) else 4   # bytes per element the dominant kernel streams
    alg_bytes = rows_local * DIMS * eb + (rows_local * 4 if eb == 1 else 0) + avg_q * DIMS * eb + avg_q * K * 8
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    qps_user = Q * args.steps / (med_ms * 1e-3)
    e2e_user = Q * n_e2e / (e2e_ms * 1e-3)
    lat_ms = sorted(x * 1e3 for x in lat)
    peak = PK["hbm"]
    line = {
        "metric": METRIC, "value": qps_user * world, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": med_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(args, world),
        "qps_user": qps_user, "rows_local_rank0": rows_local,
        "timing": {"repetitions": reps, "steps_per_repetition": args.steps, "timed_region_s": sum(rep_ms) / 1e3,
                   "ms_per_step_median": med_ms / args.steps, "ms_per_step_min": min(rep_ms) / args.steps,
                   "ms_per_step_p90": float(np.percentile(rep_ms, 90)) / args.steps,
                   "how": "the driver's K steps repeated back to back until >= %.1f s of device time; CUDA events "
                          "around every repetition, max over ranks per repetition, median reported" % MIN_TIMED_S},
        "e2e": {"value": e2e_user * world, "unit": UNIT, "h2d_bytes_per_step": Q * DIMS * 4,
                "d2h_bytes_per_step": Q * K * 8, "qps_user": e2e_user, "ms_per_step": e2e_ms / n_e2e, "calls": n_e2e,
                "p50_ms": lat_ms[len(lat_ms) // 2], "p99_ms": lat_ms[min(len(lat_ms) - 1, int(len(lat_ms) * 0.99))],
                "call": "one gf_set_search(q=%d) per step" % Q if world == 1 else "ShardedSet.search_keys(q=%d)" % Q,
                "host_breakdown": host_breakdown},
        "units": "a query counts once per %d-row shard it scans: value = qps_user x n_gpus (weak scaling)" % args.rows_per_gpu,
        "gpu_launches": int(launches), "gpu_launches_per_step": launches / float(max(reps * args.steps, 1)),
        "tensor_passes": stats["tensor_passes"],
        "uncertified_queries": int(flagged_timed),
        "uncertified_note": "queries of the timed region whose certificate refused (device counter, summed over ranks); "
                            "the host API re-runs such a query through the exact scan",
        "parity_check": parity, "exchange_trace": xtrace,
        "exactness": "ids, order and scores are the canonical float32 answer (bit-identical to the oracle in the GPU "
                     "tests): the int8 / bf16 / tf32 pass only selects candidates, every winner is re-scored from the "
                     "float32 rows and certified against the measured quantisation residual; a query that cannot be "
                     "certified (uncertified_queries) is re-run by the exact float32 scan",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic_from_profile(eb),
                     "kernel": ("gemm_kernel<16,1,%d> (tcgen05 %s candidate pass over the %s)"
                                % (eb, {1: "kind::i8 int8", 2: "kind::f16 bf16", 4: "kind::tf32"}[eb],
                                   {1: "int8 shadow slab", 2: "bf16 shadow slab", 4: "float32 rows"}[eb])) if tensor
                     else "scan_tma_kernel<512,QT>",
                     "bytes_per_element": eb,
                     "frac_if_counted_as_f32_bytes": (achieved * 4 / eb) / peak if peak else None,
                     "launches_per_step": per_step,
                     "avg_launch_ms": scan_ms, "alg_bytes_per_launch": alg_bytes, "peak_source": PK["hbm_src"],
                     "timed": "CUDA events around each launch, a second pass of %d steps" % max(args.steps, 20),
                     "frac_of_nominal_8000": achieved / 8000.0},
        "clocks": sampler.summary(t_mark0, t_mark1) if sampler else None,
    }
    line.update(line_extra)
    if multi is not None and "qps_user" in multi:
        # the headline end-to-end number at N > 1: ONE process, one caller, gf_set_search(q=10) with host buffers on
        # a set the library shards over the N devices; ids resolved
        line["e2e_torchrun_collective"] = line["e2e"]
        line["e2e"] = {"value": multi["qps_user"] * world, "unit": UNIT, "h2d_bytes_per_step": Q * DIMS * 4 * world,
                       "d2h_bytes_per_step": Q * K * 8 + Q * 4, "qps_user": multi["qps_user"],
                       "ms_per_step": multi["ms_per_step"], "calls": multi["calls"], "p50_ms": multi["p50_ms"],
                       "p99_ms": multi["p99_ms"],
                       "call": "one process, gf_context_create_multi(%d devices): gf_set_search(q=%d) per step, host "
                               "buffers in, ids out" % (world, Q),
                       "host_breakdown": multi["host_breakdown"], "parity_check": multi["parity_check"]}
    elif multi is not None:
        line["e2e_multi_error"] = multi
    if demo and "qps" in demo:
        # the headline end-to-end number: 10 real concurrent callers, host buffers, coalesced by the library
        line["e2e_batched_call"] = line["e2e"]
        line["e2e"] = {"value": demo["qps"], "unit": UNIT,
                       "h2d_bytes_per_step": int(demo["h2d_bytes"] / max(demo["searches"], 1) * Q),
                       "d2h_bytes_per_step": int(demo["d2h_bytes"] / max(demo["searches"], 1) * Q),
                       "qps_user": demo["qps"], "ms_per_step": demo["wall_ms"] / demo["rounds"],
                       "wall_s": demo["wall_ms"] / 1e3, "p50_ms": demo["p50_ms"], "p99_ms": demo["p99_ms"],
                       "call": "%d host threads x gf_set_search(q=1), coalescing 100 us (tools/gf_demo)" % Q,
                       "coalesced_batches": demo["coalesced_batches"], "tensor_passes": demo["tensor_passes"],
                       "self_check": demo["self_check"], "runs_qps": demo.get("runs_qps"),
                       "how": "median of three runs of 10 000 rounds (>= 1.2 s each)"}
    elif demo:
        line["e2e_parallel_error"] = demo
    if churn:
        c, i = churn.get("churn", {}), churn.get("idle", {})
        line["churn"] = {"workload": "BASELINE configs[4]: 10 M x 512 rows, 4 searcher threads x gf_set_search(q=1, top-%d), "
                                     "Add + Delete of 100 000 rows in chunks of 10 000 meanwhile (tools/gf_demo)" % K,
                         "idle": {k: i.get(k) for k in ("qps", "p50_ms", "p99_ms", "searches", "error") if k in i},
                         "under_churn": {k: c.get(k) for k in ("qps", "p50_ms", "p99_ms", "searches", "rows_added",
                                                               "add_rows_per_s", "deleted", "delete_ids_per_s",
                                                               "self_check", "error") if k in c}}
    if world == 1:
        line["float32_paths"] = f32_paths
        l1 = sorted(x * 1e3 for x in lat1)
        eb1 = shadow_eb if (q1[1] and shadow_eb) else 4
        b1 = rows_local * DIMS * eb1 + (rows_local * 4 if eb1 == 1 else 0) + DIMS * eb1 + K * 8
        line["single_query"] = {"p50_ms": l1[len(l1) // 2], "p99_ms": l1[min(len(l1) - 1, int(len(l1) * 0.99))],
                                "qps": 1e3 / (sum(l1) / len(l1)), "scan_kernel_ms": q1[0],
                                "kernel": "tcgen05 candidate pass" if q1[1] else "scan_tma_kernel<512,1>",
                                "bytes_per_element": eb1,
                                "roofline_frac": (b1 / (q1[0] * 1e-3) / 1e9 / peak) if q1[0] > 0 else None,
                                "achieved_gbs": (b1 / (q1[0] * 1e-3) / 1e9) if q1[0] > 0 else None}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            qps, ms_pass, passes, X, Qh = cpu_port_qps(args.rows_per_gpu, Q, K, args.cpu_seconds, cores)
            line["cpu_baseline"] = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "all %d rows x %d-d, %d queries top-%d per pass, %d passes, median %.1f ms/pass, "
                                              "oracle canonical-order AVX2 scan, persistent pool of %d threads"
                                              % (args.rows_per_gpu, DIMS, Q, K, passes, ms_pass, cores)}
            if not args.quick:
                sub = min(250_000, args.rows_per_gpu)
                from oracle import lib as ol
                t0 = time.perf_counter()
                n1 = 0
                while n1 < 2 or time.perf_counter() - t0 < 3.0:
                    ol.scan_topk(X[:sub], Qh, K, ol.MODE_CANONICAL, 1)
                    n1 += 1
                ms1 = (time.perf_counter() - t0) / n1 * 1e3 * args.rows_per_gpu / sub
                baselines["cpu_scan_1t"] = {"value": Q / (ms1 * 1e-3), "unit": UNIT, "cores": 1,
                                            "sample": "%d of %d rows, %d passes, scaled by the rows ratio" % (sub, args.rows_per_gpu, n1)}
                bq, bms, bn = cpu_blas_qps(X, Qh, K, 3.0)
                baselines["cpu_blas"] = {"value": bq, "unit": UNIT, "cores": cores, "ms_per_pass": bms, "passes": bn,
                                         "what": "numpy %s matmul (all rows x %d queries) + argpartition + sort of the "
                                                 "%d best per query" % (np.__version__, Q, K)}
            del X
    if baselines:
        for b in baselines.values():
            if "qps" in b and "value" not in b:
                b["value"], b["unit"] = b["qps"], UNIT
        line["baselines"] = baselines
    emit(line)
    ctx.Close()
    if world > 1:
        H.dist.destroy_process_group()
    if parity["mismatches"] != 0 or not parity["needles_ok"]:
        sys.stderr.write("bench.py: parity_check FAILED: %r\n" % (parity,))
        sys.exit(3)


def leg_multi(api, world, rows_total, shadow, pool, ref_keys, Q, K, steps):
    """rank 0 only: the N-device set behind the C ABI.  Host queries in, [][]FeatureSearchResult out."""
    lean = api.lib()
    mctx = api.NewContextMulti(list(range(world)))
    gblocks = (rows_total + CAP - 1) // CAP
    cache = api.NewCache(mctx, gblocks + world, BLOCK_SIZE, shadow=shadow)
    cache.NewSet("multi", DIMS, 4, max(Q, 16))
    st = cache.GetSet("multi")
    st.AddSynthetic(rows_total, SEED)
    npool = pool.shape[0]
    handle = ctypes.c_void_p()
    hq = [np.ascontiguousarray(pool[b]) for b in range(npool)]

    def call(i):
        a = hq[i % npool]
        rc = lean.gf_set_search(st._h, -2.0, K, Q, a.ctypes.data, a.nbytes, ctypes.byref(handle))
        if rc != 0:
            raise api.GoFeatureError(rc)
        n = lean.gf_result_total(handle)
        lean.gf_result_free(handle)
        return n

    # what the library returns (ids, slots, scores) against the keys the torchrun path merged for the same batches
    mism = 0
    for b, rk in enumerate(ref_keys):
        got = st.SearchArray(-2.0, K, pool[b])
        for i in range(Q):
            want = [(api.key_slot(k), np.float32(api.key_score(k))) for k in rk[i] if k != 0]
            have = [(r.Slot, np.float32(r.Score)) for r in got[i]]
            ids_ok = all(r.ID == "f%011d" % r.Slot for r in got[i])
            mism += 0 if (want == have and ids_ok) else 1
    for i in range(10):
        call(i)
    t0 = time.perf_counter()
    for i in range(steps):
        call(i)
    probe = time.perf_counter() - t0
    n = int(min(20000, max(steps, np.ceil(MIN_TIMED_S / max(probe, 1e-6) * steps))))
    lat = []
    t0 = time.perf_counter()
    for i in range(n):
        a = time.perf_counter()
        call(i)
        lat.append(time.perf_counter() - a)
    wall = time.perf_counter() - t0
    stats = mctx.Stats()
    c = float(max(stats["host_calls"], 1))
    hb = {"stage_queries_us": stats["host_stage_ns"] / c / 1e3, "enqueue_us": stats["host_launch_ns"] / c / 1e3,
          "wait_for_device_us": stats["host_wait_ns"] / c / 1e3, "resolve_ids_us": stats["host_resolve_ns"] / c / 1e3,
          "calls": int(c)}
    cache.Close()
    mctx.Close()
    l = sorted(x * 1e3 for x in lat)
    return {"qps_user": Q * n / wall, "ms_per_step": wall / n * 1e3, "calls": n, "p50_ms": l[len(l) // 2],
            "p99_ms": l[min(len(l) - 1, int(len(l) * 0.99))], "uncertified_queries": stats["uncertified_queries"],
            "host_breakdown": hb,
            "parity_check": {"queries": len(ref_keys) * Q, "mismatches": mism,
                             "against": "the keys the one-process-per-GPU path (NCCL harness, in-kernel exchange) merged "
                                        "for the same batches; ids must be f<slot>"}}


def _mix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def cluster_centre(seed, centre, dims):
    """centre vector of a clustered synthetic set (the hash gf_set_add_synthetic_ex uses, restated in numpy)."""
    with np.errstate(over="ignore"):
        cnt = np.uint64(centre) * np.uint64(dims) + np.arange(dims, dtype=np.uint64)
        h = _mix64(np.uint64(seed ^ 0xC3A7E25) ^ _mix64(cnt))
    return (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)


def leg_clustered(H, api, ctx, rows, Q, K, shadow):
    """The headline step on CLUSTERED data: 10 000 centres x ~100 members at cosine ~0.975 to their centre, 1 % exact
    duplicates, queries near centres (about 100 rows score within 0.01 of the winner).  How many queries does the
    certificate refuse, and what does the step cost?"""
    torch = H.torch
    seed, centres, spread, dup = 777, max(1, rows // 100), 0.228, 10000
    cache = api.NewCache(ctx, (rows + CAP - 1) // CAP + 1, BLOCK_SIZE, shadow=shadow)
    cache.NewSet("clustered", DIMS, 4, max(Q, 16))
    st = cache.GetSet("clustered")
    st.AddSynthetic(rows, seed, 0, True, "f", centres, spread, dup)
    rng = np.random.default_rng(5)
    npool = 32
    pool = np.empty((npool, Q, DIMS), dtype=np.float32)
    for b in range(npool):
        for i in range(Q):
            v = cluster_centre(seed, int(rng.integers(0, centres)), DIMS) + 0.1 * rng.uniform(-1, 1, DIMS).astype(np.float32)
            pool[b, i] = v / np.linalg.norm(v)
    dpool = torch.from_numpy(pool).cuda()
    keys = torch.zeros((Q, K), dtype=torch.int64, device="cuda")
    flags = torch.zeros((Q,), dtype=torch.int32, device="cuda")

    def step(i, path=0):
        st.SearchDeviceEx(K, Q, dpool[i % npool].data_ptr(), keys.data_ptr(), flags.data_ptr(), path, None)

    for i in range(5):
        step(i)
    f0 = st.DeviceCounters()[0]
    ms, _, _ = H.timed_reps(step, 20, 0.5)
    n_steps = 20 * (len(ms) + 1)
    refused = st.DeviceCounters()[0] - f0
    step_ms = float(np.median(ms)) / 20
    # correctness + spread of the winners, batch by batch, against the exact scan on the same rows
    mism = flagged = 0
    gaps = []
    for b in range(npool):
        step(b)
        k0, fl = keys.clone(), flags.clone()
        step(b, 1)
        torch.cuda.synchronize()
        bad = (k0 != keys).any(dim=1)
        flagged += int((fl != 0).sum().item())
        mism += int((bad & (fl == 0)).sum().item())
        hk = keys.cpu().numpy().view(np.uint64)
        gaps += [float(api.key_score(r[0]) - api.key_score(r[K - 1])) for r in hk]
    # through the host call: refused queries are re-run by the exact scan inside
    ctx.ResetStats()
    lean = api.lib()
    handle = ctypes.c_void_p()
    t0 = time.perf_counter()
    n = 400
    for i in range(n):
        a = pool[i % npool]
        lean.gf_set_search(st._h, -2.0, K, Q, a.ctypes.data, a.nbytes, ctypes.byref(handle))
        lean.gf_result_free(handle)
    e2e_ms = (time.perf_counter() - t0) / n * 1e3
    hs = ctx.Stats()
    cache.Close()
    return {"workload": "%d x %d-d rows in %d clusters (members at cosine ~0.975 to their centre), 1%% exact duplicates; "
                        "%d queries near centres, top-%d per step" % (rows, DIMS, centres, Q, K),
            "median_score_gap_best_to_kth": float(np.median(gaps)),
            "ms_per_step": step_ms, "value": Q / (step_ms * 1e-3), "unit": UNIT,
            "uncertified_rate": refused / float(max(n_steps * Q, 1)), "queries_timed": n_steps * Q,
            "e2e_ms_per_step": e2e_ms, "e2e_value": Q / (e2e_ms * 1e-3),
            "e2e_uncertified_rate": hs["uncertified_queries"] / float(n * Q),
            "parity_check": {"queries": npool * Q, "mismatches": mism, "flagged_queries": flagged,
                             "against": "exact float32 scan kernel (path 1)"}}


def leg_config3(H, ctx, sset, rows_local, shadow_eb, PK):
    """BASELINE configs[2]: 1024 queries x top-100 on 1 M rows -- four 256-query tcgen05 passes + ONE fused tail
    launch (selection, two-stage exact re-score, top-K, certificate) for all 1024 queries."""
    torch = H.torch
    Q3, K3 = 1024, 100
    q3 = torch.from_numpy(make_queries(Q3 * 2, DIMS, QUERY_SEED + 7).reshape(2, Q3, DIMS)).cuda()
    fn = lambda i: sset.search_device(K3, Q3, q3[i % 2])   # noqa: E731
    for i in range(3):
        fn(i)
    ms, _, _ = H.timed_reps(fn, 10, 0.5)
    step_ms = float(np.median(ms)) / 10
    kms, per, tensor, st = profile_kernel(ctx, fn, 6)
    flops = 2.0 * Q3 * rows_local * DIMS
    mult = {1: 2.0, 2: 1.0}.get(shadow_eb, 0.5)        # int8 runs at twice, tf32 at half the bf16 rate
    peak = PK["bf16"] * mult
    pass_flops = 2.0 * 256 * rows_local * DIMS
    k0 = sset.search_device(K3, Q3, q3[0]).clone()
    f0 = sset.flags(Q3, K3).clone()
    k1 = sset.search_device(K3, Q3, q3[0], path=1).clone()
    torch.cuda.synchronize()
    bad = ((k0 != k1).any(dim=1) & (f0 == 0)).sum().item()
    return {"workload": "BASELINE configs[2]: %d x %d-d rows, %d queries top-%d per step" % (rows_local, DIMS, Q3, K3),
            "ms_per_step": step_ms, "value": Q3 / (step_ms * 1e-3), "unit": UNIT,
            "roofline": {"bound": "tensor", "achieved": flops / (step_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (step_ms * 1e-3) / 1e12 / peak,
                         "what": "useful 2*Q*N*D of the WHOLE step (4 x (sample + threshold + pass) + one fused tail) against "
                                 "%.1fx the %s" % (mult, PK["bf16_src"]),
                         "pass_kernel_ms": kms, "passes_per_step": per,
                         "pass_kernel_tflops": pass_flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                         "pass_kernel_frac": (pass_flops / (kms * 1e-3) / 1e12 / peak) if kms > 0 else None},
            "parity_check": {"queries": Q3, "mismatches": int(bad), "flagged_queries": int((f0 != 0).sum().item())}}


def leg_rows(H, api, ShardedSet, ctx, rows_total, Qx, Kx, shadow, PK, name):
    """A sharded set of rows_total rows (block-striped over the ranks), Qx queries top-Kx per step."""
    torch = H.torch
    cache, sset = new_sharded(H, api, ShardedSet, ctx, rows_total, max(Qx, 16), shadow, name)
    rows_local = sset.local.ScannedRows()
    eb = cache.ShadowBytesPerElement() or 4
    qh = make_queries(Qx * 4, DIMS, QUERY_SEED + 11).reshape(4, Qx, DIMS)
    qd = torch.from_numpy(qh).cuda()
    fn = lambda i: sset.search_device(Kx, Qx, qd[i % 4])   # noqa: E731
    for i in range(3):
        fn(i)
        sset.search_keys(Kx, qh[i % 4])
    ms, _, _ = H.timed_reps(fn, 10, 0.5)
    step_ms = float(np.median(ms)) / 10
    kms, per, tensor, _ = profile_kernel(ctx, fn, 10)
    kms = H.max_over_ranks([kms])[0]
    H.barrier()
    t0 = time.perf_counter()
    n = 20
    for i in range(n):
        sset.search_keys(Kx, qh[i % 4])
    H.barrier()
    e2e_ms = H.max_over_ranks([(time.perf_counter() - t0) * 1e3])[0] / n
    k0 = sset.search_device(Kx, Qx, qd[0]).clone()
    f0 = sset.flags(Qx, Kx).clone()
    k1 = sset.search_device(Kx, Qx, qd[0], path=1).clone()
    torch.cuda.synchronize()
    bad = ((k0 != k1).any(dim=1) & (f0 == 0)).sum().item()
    alg = rows_local * DIMS * eb + (rows_local * 4 if eb == 1 else 0)
    out = {"workload": "%d x %d-d rows row-sharded over %d GPU(s) (%d on rank 0), %d queries top-%d per step"
                       % (rows_total, DIMS, H.world, rows_local, Qx, Kx),
           "rows_total": rows_total, "ms_per_step": step_ms, "qps_user": Qx / (step_ms * 1e-3), "unit": UNIT,
           "e2e_ms_per_step": e2e_ms, "e2e_qps_user": Qx / (e2e_ms * 1e-3),
           "pass_kernel_ms_max_over_ranks": kms, "pass_launches_per_step": per,
           "hbm_frac_per_rank": (alg / (kms * 1e-3) / 1e9 / PK["hbm"]) if (kms > 0 and tensor) else None,
           "parity_check": {"queries": Qx, "mismatches": int(bad), "flagged_queries": int((f0 != 0).sum().item())}}
    sset.close_exchange()
    cache.Close()
    return out


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL's version banner, torchrun notices) write to fd 1: keep stdout for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=10)
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shadow", action="store_true", help="float32 rows only (no shadow slab)")
    ap.add_argument("--shadow", default="int8", choices=["int8", "bf16"], help="format of the shadow slab")
    ap.add_argument("--quick", action="store_true", help="headline + parity only (no config 3/4/5, strong, baselines)")
    ap.add_argument("--config4-rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--strong-rows", type=int, default=8_000_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

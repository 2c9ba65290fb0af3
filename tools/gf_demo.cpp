// gf_demo -- load generator over the C ABI, the counterpart of the reference's demo/main.go:
// SearchParallel goroutines x Round searches of Batch random queries (demo/main.go:116-143), one
// thread adding 1-10 features at a time (demo/main.go:147-180), one thread deleting random ids
// (demo/main.go:184-201), average search delay printed at the end (demo/main.go:205-206).
// Host buffers in, results out: every search pays its H2D / D2H inside the timed call.
// Prints one JSON object.  Build: g++ -O2 -std=c++17 tools/gf_demo.cpp -Iinclude -Lgofeature_b200
//   -lgofeature_b200 -Wl,-rpath,$PWD/gofeature_b200 -lpthread -o tools/gf_demo
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "gofeature_b200.h"

static void die(const char *what, int rc)
{
    fprintf(stderr, "%s: %s (%d) %s\n", what, gf_strerror(rc), rc, gf_last_error_detail());
    exit(1);
}
#define CK(x)                    \
    do {                         \
        int _rc = (x);           \
        if (_rc) die(#x, _rc);   \
    } while (0)

static long arg(int argc, char **argv, const char *name, long def)
{
    for (int i = 1; i + 1 < argc; ++i)
        if (!strcmp(argv[i], name)) return atol(argv[i + 1]);
    return def;
}

int main(int argc, char **argv)
{
    const long rows = arg(argc, argv, "--rows", 1000000);       // SetSize   demo/main.go:25
    const int dims = (int)arg(argc, argv, "--dims", 512);       // Dimension demo/main.go:21
    const int threads = (int)arg(argc, argv, "--threads", 5);   // SearchParallel demo/main.go:27
    const int rounds = (int)arg(argc, argv, "--rounds", 200);   // Round     demo/main.go:24
    const int warm_rounds = (int)arg(argc, argv, "--warm-rounds", 30);  // untimed rounds of all searcher threads
    const int batch = (int)arg(argc, argv, "--batch", 1);       // Batch     demo/main.go:22
    const int limit = (int)arg(argc, argv, "--limit", 10);
    const int coalesce = (int)arg(argc, argv, "--coalesce-us", 30);
    const int adds = (int)arg(argc, argv, "--adds", 0);         // 50 in the demo
    const int dels = (int)arg(argc, argv, "--deletes", 0);      // 200 in the demo
    const int add_gap_us = (int)arg(argc, argv, "--add-gap-us", 2000);
    const int add_size = (int)arg(argc, argv, "--add-size", 0);      // rows per Add call (0 = 1..10 like the demo)
    const int del_size = (int)arg(argc, argv, "--delete-size", 1);   // ids per Delete call
    const int device = (int)arg(argc, argv, "--device", 0);
    const int ndev = (int)arg(argc, argv, "--devices", 1);   // > 1: one set striped over that many devices of this process
    const long block_size = 1024L * 1024 * 32;                  // BlockSize demo/main.go:19
    const long cap = block_size / (dims * 4L);
    const int block_num = (int)((rows + cap - 1) / cap) + 4;

    gf_context *ctx;
    gf_cache *cache;
    gf_set *set;
    if (ndev > 1) {
        int have = 0;
        CK(gf_device_count(&have));
        std::vector<int> devs;
        for (int i = 0; i < ndev; ++i) devs.push_back((device + i) % std::max(have, 1));
        CK(gf_context_create_multi(devs.data(), ndev, &ctx));
    } else {
        CK(gf_context_create(device, &ctx));
    }
    CK(gf_cache_create(ctx, block_num, block_size, &cache));
    CK(gf_cache_new_set(cache, "test0", dims, 4, std::max(batch, 16)));
    CK(gf_cache_get_set(cache, "test0", &set));
    auto t_fill = std::chrono::steady_clock::now();
    CK(gf_set_add_synthetic(set, rows, 20260101, 0, 1, "f"));
    double fill_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_fill).count();
    CK(gf_set_set_coalescing(set, coalesce));

    // query pool: U(-1,1) like demo/main.go:131, L2-normalised (util.go:125-139)
    const int pool = 256;
    std::vector<float> qpool((size_t)pool * batch * dims);
    {
        std::mt19937 rng(7);
        std::uniform_real_distribution<float> u(-1.f, 1.f);
        std::vector<float> raw(dims);
        for (int i = 0; i < pool * batch; ++i) {
            for (int j = 0; j < dims; ++j) raw[j] = u(rng);
            CK(gf_util_normalize_float32(raw.data(), qpool.data() + (size_t)i * dims, dims));
        }
    }
    // warm-up
    for (int i = 0; i < 5; ++i) {
        gf_result *r;
        CK(gf_set_search(set, 0.f, limit, batch, qpool.data(), (uint64_t)batch * dims * 4, &r));
        gf_result_free(r);
    }
    // ... and with every searcher thread at once: the batch shapes the coalescer forms (2 .. threads queries) are each
    // recorded as a CUDA graph on first use, which must not fall into the timed rounds
    if (warm_rounds > 0 && threads > 1) {
        std::vector<std::thread> wt;
        for (int p = 0; p < threads; ++p) {
            wt.emplace_back([&, p]() {
                for (int b = 0; b < warm_rounds; ++b) {
                    const float *q = qpool.data() + (size_t)((b * threads + p) % pool) * batch * dims;
                    gf_result *r;
                    int rc = gf_set_search(set, 0.f, limit, batch, q, (uint64_t)batch * dims * 4, &r);
                    if (rc) die("gf_set_search (warm-up)", rc);
                    gf_result_free(r);
                }
            });
        }
        for (auto &t : wt) t.join();
    }
    gf_context_reset_stats(ctx);

    std::vector<std::vector<double>> lat(threads);
    std::atomic<long> hits{0};
    std::atomic<bool> stop{false};
    std::atomic<int> n_adds{0}, n_dels{0}, n_deleted{0};
    std::atomic<long> rows_added{0}, add_us{0}, del_us{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int p = 0; p < threads; ++p) {
        th.emplace_back([&, p]() {
            lat[p].reserve(rounds);
            for (int b = 0; b < rounds; ++b) {
                const float *q = qpool.data() + (size_t)((b * threads + p) % pool) * batch * dims;
                auto a = std::chrono::steady_clock::now();
                gf_result *r;
                int rc = gf_set_search(set, 0.f, limit, batch, q, (uint64_t)batch * dims * 4, &r);
                if (rc) die("gf_set_search", rc);
                lat[p].push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count());
                hits += gf_result_total(r);
                gf_result_free(r);
            }
        });
    }
    std::thread adder([&]() {  // demo/main.go:147-180
        std::mt19937 rng(11);
        std::uniform_real_distribution<float> u(-1.f, 1.f);
        long next = 0;
        for (int i = 0; i < adds && !stop; ++i) {
            int size = add_size > 0 ? add_size : 1 + (int)(rng() % 10);
            std::vector<float> v((size_t)size * dims), raw(dims);
            std::string ids;
            std::vector<uint64_t> off(size + 1, 0);
            for (int k = 0; k < size; ++k) {
                for (int j = 0; j < dims; ++j) raw[j] = u(rng);
                gf_util_normalize_float32(raw.data(), v.data() + (size_t)k * dims, dims);
                char buf[32];
                snprintf(buf, sizeof buf, "a%011ld", next++);
                ids += buf;
                off[k + 1] = ids.size();
            }
            auto a0 = std::chrono::steady_clock::now();
            int rc = gf_set_add(set, size, v.data(), (uint64_t)size * dims * 4, ids.data(), off.data());
            if (rc) die("gf_set_add", rc);
            add_us += (long)std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a0).count();
            rows_added += size;
            ++n_adds;
            std::this_thread::sleep_for(std::chrono::microseconds(add_gap_us));
        }
    });
    std::thread deleter([&]() {  // demo/main.go:184-201
        std::mt19937_64 rng(13);
        for (int i = 0; i < dels && !stop; ++i) {
            std::string ids;
            std::vector<uint64_t> off(del_size + 1, 0);
            for (int k = 0; k < del_size; ++k) {
                char buf[32];
                snprintf(buf, sizeof buf, "f%011ld", (long)(rng() % (uint64_t)rows));
                ids += buf;
                off[k + 1] = ids.size();
            }
            std::vector<uint8_t> flags(del_size);
            int64_t n = 0;
            auto a0 = std::chrono::steady_clock::now();
            int rc = gf_set_delete(set, del_size, ids.data(), off.data(), flags.data(), &n);
            if (rc) die("gf_set_delete", rc);
            del_us += (long)std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a0).count();
            ++n_dels;
            n_deleted += (int)n;
            std::this_thread::sleep_for(std::chrono::microseconds(add_gap_us / 4 + 1));
        }
    });
    for (auto &t : th) t.join();
    double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    stop = true;
    adder.join();
    deleter.join();

    std::vector<double> all;
    for (auto &v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    double sum = 0;
    for (double x : all) sum += x;
    gf_stats st;
    gf_context_stats(ctx, &st);

    // self-check through the ABI only: a stored row must find itself first with score ~1
    std::vector<float> row(dims);
    uint8_t found = 0;
    int64_t nf = 0;
    const char *probe = "f00000000042";
    uint64_t poff[2] = {0, 12};
    CK(gf_set_read(set, 1, probe, poff, row.data(), (uint64_t)dims * 4, &found, &nf));
    bool ok = true;
    if (found) {
        gf_result *r;
        CK(gf_set_search(set, 0.f, 1, 1, row.data(), (uint64_t)dims * 4, &r));
        uint32_t len = 0;
        const char *id = gf_result_id(r, 0, 0, &len);
        ok = gf_result_count(r, 0) == 1 && len == 12 && !memcmp(id, probe, 12) && std::fabs(gf_result_score(r, 0, 0) - 1.f) < 1e-5f;
        gf_result_free(r);
    }
    const double nsearch = (double)all.size();
    printf("{\"devices\": %d, ", gf_context_device_count(ctx));
    printf("\"rows\": %ld, \"dims\": %d, \"threads\": %d, \"rounds\": %d, \"batch\": %d, \"limit\": %d, \"coalesce_us\": %d, "
           "\"fill_ms\": %.1f, \"wall_ms\": %.3f, \"searches\": %.0f, \"queries\": %.0f, \"qps\": %.1f, \"avg_ms\": %.4f, "
           "\"p50_ms\": %.4f, \"p99_ms\": %.4f, \"hits\": %ld, ",
           rows, dims, threads, rounds, batch, limit, coalesce, fill_ms, wall_ms, nsearch, nsearch * batch,
           nsearch * batch / (wall_ms / 1e3), sum / nsearch, all[all.size() / 2],
           all[std::min(all.size() - 1, (size_t)(all.size() * 0.99))], hits.load());
    printf("\"adds\": %d, \"rows_added\": %ld, \"add_rows_per_s\": %.0f, \"deletes\": %d, \"deleted\": %d, "
           "\"delete_ids_per_s\": %.0f, ",
           n_adds.load(), rows_added.load(), add_us.load() ? rows_added.load() / (add_us.load() / 1e6) : 0.0,
           n_dels.load(), n_deleted.load(),
           del_us.load() ? (double)n_dels.load() * del_size / (del_us.load() / 1e6) : 0.0);
    printf("\"add_stage_us_per_krow\": %.1f, \"add_place_us_per_krow\": %.1f, \"add_index_us_per_krow\": %.1f, ",
           st.add_rows ? st.add_stage_ns / 1e3 / (st.add_rows / 1e3) : 0.0, st.add_rows ? st.add_place_ns / 1e3 / (st.add_rows / 1e3) : 0.0,
           st.add_rows ? st.add_index_ns / 1e3 / (st.add_rows / 1e3) : 0.0);
    printf("\"scan_launches\": %llu, \"tensor_passes\": %llu, \"coalesced_batches\": %llu, \"quirk_fallbacks\": %llu, "
           "\"uncertified_queries\": %llu, \"h2d_bytes\": %llu, \"d2h_bytes\": %llu, \"self_check\": \"%s\"}\n",
           (unsigned long long)st.scan_launches, (unsigned long long)st.tensor_passes,
           (unsigned long long)st.coalesced_batches, (unsigned long long)st.quirk_fallbacks,
           (unsigned long long)st.uncertified_queries, (unsigned long long)st.h2d_bytes,
           (unsigned long long)st.d2h_bytes, ok ? "ok" : "FAILED");
    gf_cache_destroy(cache);
    gf_context_destroy(ctx);
    return ok ? 0 : 2;
}

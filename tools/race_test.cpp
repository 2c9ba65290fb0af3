// race_test -- concurrent Add / Delete / Update / Read / Search / DestroySet on one set through the C ABI, meant to be
// built with -fsanitize=thread (tools/Makefile: `make race_test_tsan`, which also builds the library's host code with
// TSAN) and run on a GPU box.  The reference's own races on these paths (set.go:77-116 Add vs Search on s.Blocks,
// cache.go:104-115 GetEmptyBlock without the mutex, block.go IDs[] read by Search while Insert writes) are what the
// library's locks replace; SURVEY section 5 asks for the tooling that shows it.
//   exit 0 + "race_test ok": invariants held (ThreadSanitizer reports go to stderr and make the exit code 66).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "gofeature_b200.h"

#define CK(x)                                                                              \
    do {                                                                                   \
        int rc_ = (x);                                                                     \
        if (rc_) {                                                                         \
            fprintf(stderr, "%s: %s (%d) %s\n", #x, gf_strerror(rc_), rc_, gf_last_error_detail()); \
            exit(2);                                                                       \
        }                                                                                  \
    } while (0)

int main(int argc, char **argv)
{
    const int dims = 64, cap = 256, ops = argc > 1 ? atoi(argv[1]) : 300;
    gf_context *ctx;
    gf_cache *cache;
    gf_set *set;
    CK(gf_context_create(0, &ctx));
    CK(gf_cache_create(ctx, 64, (int64_t)cap * dims * 4, &cache));
    CK(gf_cache_new_set(cache, "race", dims, 4, 8));
    CK(gf_cache_get_set(cache, "race", &set));
    CK(gf_set_set_coalescing(set, 50));
    std::atomic<long> added{0}, deleted{0}, searches{0}, errors{0};
    std::atomic<bool> stop{false};

    auto make_rows = [&](std::mt19937 &rng, int n, std::vector<float> &v) {
        std::uniform_real_distribution<float> u(-1.f, 1.f);
        v.resize((size_t)n * dims);
        for (auto &x : v) x = u(rng);
    };
    auto pack = [](const std::vector<std::string> &ids, std::string &blob, std::vector<uint64_t> &off) {
        blob.clear();
        off.assign(1, 0);
        for (auto &s : ids) {
            blob += s;
            off.push_back(blob.size());
        }
    };
    std::vector<std::thread> th;
    for (int a = 0; a < 2; ++a)  // adders: disjoint id families
        th.emplace_back([&, a]() {
            std::mt19937 rng(100 + a);
            std::vector<float> v;
            std::string blob;
            std::vector<uint64_t> off;
            for (int i = 0; i < ops; ++i) {
                const int n = 1 + (int)(rng() % 40);
                std::vector<std::string> ids;
                for (int k = 0; k < n; ++k) ids.push_back("t" + std::to_string(a) + "_" + std::to_string(i) + "_" + std::to_string(k));
                make_rows(rng, n, v);
                pack(ids, blob, off);
                int rc = gf_set_add(set, n, v.data(), (uint64_t)n * dims * 4, blob.data(), off.data());
                if (rc == 0) added += n;
                else if (rc != GF_ERR_NOT_ENOUGH_BLOCKS) ++errors;
            }
        });
    th.emplace_back([&]() {  // deleter: ids the adders may or may not have written yet
        std::mt19937 rng(7);
        std::string blob;
        std::vector<uint64_t> off;
        for (int i = 0; i < ops; ++i) {
            std::vector<std::string> ids;
            for (int k = 0; k < 5; ++k)
                ids.push_back("t" + std::to_string(rng() % 2) + "_" + std::to_string(rng() % (unsigned)ops) + "_" + std::to_string(rng() % 10));
            pack(ids, blob, off);
            uint8_t found[5];
            int64_t n = 0;
            if (gf_set_delete(set, 5, blob.data(), off.data(), found, &n) == 0) deleted += n; else ++errors;
        }
    });
    th.emplace_back([&]() {  // update + read
        std::mt19937 rng(9);
        std::vector<float> v, out(dims);
        std::string blob;
        std::vector<uint64_t> off;
        for (int i = 0; i < ops; ++i) {
            std::vector<std::string> ids{"t0_" + std::to_string(rng() % (unsigned)ops) + "_0"};
            pack(ids, blob, off);
            make_rows(rng, 1, v);
            uint8_t f = 0;
            int64_t n = 0;
            if (gf_set_update(set, 1, v.data(), (uint64_t)dims * 4, blob.data(), off.data(), &f, &n)) ++errors;
            if (gf_set_read(set, 1, blob.data(), off.data(), out.data(), (uint64_t)dims * 4, &f, &n)) ++errors;
        }
    });
    for (int s = 0; s < 4; ++s)  // searchers
        th.emplace_back([&, s]() {
            std::mt19937 rng(50 + s);
            std::vector<float> q;
            while (!stop) {
                make_rows(rng, 2, q);
                gf_result *r = nullptr;
                int rc = gf_set_search(set, -2.f, 5, 2, q.data(), (uint64_t)2 * dims * 4, &r);
                if (rc) {
                    if (rc != GF_ERR_FEATURE_SET_NOT_FOUND) ++errors;
                    continue;
                }
                for (int qi = 0; qi < 2; ++qi)
                    for (int i = 1; i < gf_result_count(r, qi); ++i)
                        if (gf_result_score(r, qi, i) > gf_result_score(r, qi, i - 1)) ++errors;
                gf_result_free(r);
                ++searches;
            }
        });
    for (int i = 0; i < 4; ++i) th[i].join();
    const long live = (long)gf_set_live_rows(set);
    const bool counts_ok = live == added.load() - deleted.load();
    // DestroySet while the searchers are still calling: they must get ErrFeatureSetNotFound or a complete result
    CK(gf_cache_destroy_set(cache, "race"));
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
    stop = true;
    for (size_t i = 4; i < th.size(); ++i) th[i].join();
    gf_cache_destroy(cache);
    gf_context_destroy(ctx);
    printf("{\"added\": %ld, \"deleted\": %ld, \"live\": %ld, \"searches\": %ld, \"errors\": %ld, \"counts_ok\": %s}\n", added.load(),
           deleted.load(), live, searches.load(), errors.load(), counts_ok ? "true" : "false");
    if (errors.load() || !counts_ok) return 3;
    printf("race_test ok\n");
    return 0;
}

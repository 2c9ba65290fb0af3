// ref_replica -- the reference's OWN GPU search path, replayed on this box and timed (a BASELINE, not product code).
//
// goFeature's Set.Search (set.go:118-159) posts one job per block to the set's SearchQueue; every block owns a worker
// goroutine (block.go:76, `go worker(...)`) that takes jobs off that queue and, per job (set.go:40-75):
//   1. transposes the queries to dims-major (util.go:21-39) and uploads them into the block's inputBuffer (set.go:57-63),
//   2. inside ONE closure on the single context thread (block.go:192-229): cublasSgemm(N, N, batch, height, dims, 1,
//      input, batch, block, dims, 0, output, batch) and a blocking read-back of the WHOLE score matrix (block.go:214),
//   3. on the worker itself: strided gather of each query's column, full sort of (value, index), first `limit`
//      (block.go:231-246, util.go:76-93);
// the caller then filters by threshold, concatenates and sorts again (set.go:138-157).
//
// This tool issues exactly those operations with the image's cuBLAS (12.9 instead of the reference's 8.0): W worker
// threads (one per block, capped at the host's cores) consume a queue of (search, block) jobs, device work is
// serialised by a mutex (the context thread), sorting happens on the workers.  It links nothing of gofeature_b200.
//
//   ref_replica --rows 1000000 --threads 10 --queries 1 --limit 10 --seconds 5
// prints one JSON line: qps, p50/p99 latency per Search call, what was run.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s failed: %s\n", #x, cudaGetErrorString(e_));                    \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ unsigned long long mix64(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// U(-1,1) rows, L2-normalised (search_test.go:65 + util.go:125-139): one warp per row
__global__ void fill_rows(float *rows, long long n, int dims, unsigned long long seed)
{
    const long long r = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (r >= n) return;
    const int lane = threadIdx.x & 31;
    float ss = 0.f;
    for (int c = lane; c < dims; c += 32) {
        const unsigned long long h = mix64(seed ^ mix64((unsigned long long)r * dims + c));
        const float v = (float)(unsigned)(h >> 40) * (1.0f / 8388608.0f) - 1.0f;
        rows[r * dims + c] = v;
        ss += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = ss > 0.f ? 1.0f / sqrtf(ss) : 0.f;
    for (int c = lane; c < dims; c += 32) rows[r * dims + c] *= inv;
}

struct Hit {
    float score;
    long long row;
};

struct Block {
    float *dev = nullptr;     // block.go:23 Buffer
    int height = 0;           // NextIndex
    long long first = 0;
};

struct Search {
    const float *queries;  // batch x dims, row-major (FeatureValue bytes)
    int batch, limit;
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
    std::vector<std::vector<Hit>> results;  // per query, concatenated per-block winners (set.go:150)
};

struct Job {
    Search *s;
    int block;
};

static std::mutex g_qmu;
static std::condition_variable g_qcv;
static std::deque<Job> g_queue;
static bool g_stop = false;
static std::mutex g_ctx;  // the single context thread (cuda.Context.Run)

struct Worker {
    float *d_in = nullptr, *d_out = nullptr;  // block.go:68-73 inputBuffer / outputBuffer
    cublasHandle_t handle;
    cudaStream_t stream;
};

static void worker_loop(Worker *w, const std::vector<Block> *blocks, int dims)
{
    std::vector<float> target, vec3;
    std::vector<std::pair<float, int>> scores;
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(g_qmu);
            g_qcv.wait(lk, [] { return g_stop || !g_queue.empty(); });
            if (g_queue.empty()) return;
            job = g_queue.front();
            g_queue.pop_front();
        }
        Search *s = job.s;
        const Block &b = (*blocks)[job.block];
        const int batch = s->batch, height = b.height;
        // FeatureValueTranspose1D (util.go:33-36): ret[i*row + j] = features[j][i]
        target.resize((size_t)batch * dims);
        for (int i = 0; i < dims; ++i)
            for (int j = 0; j < batch; ++j) target[(size_t)i * batch + j] = s->queries[(size_t)j * dims + i];
        vec3.resize((size_t)height * batch);
        {
            std::lock_guard<std::mutex> ctx(g_ctx);
            // inputBuffer.Write (buffer.go:34-44): synchronous copy from pageable memory
            CK(cudaMemcpy(w->d_in, target.data(), target.size() * 4, cudaMemcpyHostToDevice));
            const float alpha = 1.f, beta = 0.f;
            cublasStatus_t st = cublasSgemm(w->handle, CUBLAS_OP_N, CUBLAS_OP_N, batch, height, dims, &alpha, w->d_in, batch,
                                            b.dev, dims, &beta, w->d_out, batch);
            if (st != CUBLAS_STATUS_SUCCESS) {
                fprintf(stderr, "cublasSgemm failed: %d\n", (int)st);
                exit(2);
            }
            CK(cudaMemcpy(vec3.data(), w->d_out, vec3.size() * 4, cudaMemcpyDeviceToHost));  // block.go:214
        }
        std::vector<std::vector<Hit>> mine((size_t)batch);
        for (int i = 0; i < batch; ++i) {
            scores.resize((size_t)height);
            for (int j = 0; j < height; ++j) scores[j] = {vec3[(size_t)j * batch + i], j};  // block.go:231-235
            std::sort(scores.begin(), scores.end(),
                      [](const std::pair<float, int> &a, const std::pair<float, int> &c) { return a.first > c.first; });
            const int m = std::min(s->limit, height);
            for (int j = 0; j < m; ++j) mine[i].push_back({scores[j].first, b.first + scores[j].second});
        }
        {
            std::lock_guard<std::mutex> lk(s->mu);
            for (int i = 0; i < batch; ++i) s->results[i].insert(s->results[i].end(), mine[i].begin(), mine[i].end());
            if (--s->pending == 0) s->cv.notify_all();
        }
    }
}

// FeatureSet.Search (set.go:118-159)
static void set_search(const std::vector<Block> &blocks, const float *queries, int batch, int limit, float threshold,
                       std::vector<std::vector<Hit>> *out)
{
    Search s;
    s.queries = queries;
    s.batch = batch;
    s.limit = limit;
    s.results.resize((size_t)batch);
    s.pending = (int)blocks.size();
    {
        std::lock_guard<std::mutex> lk(g_qmu);
        for (int b = 0; b < (int)blocks.size(); ++b) g_queue.push_back({&s, b});
    }
    g_qcv.notify_all();
    {
        std::unique_lock<std::mutex> lk(s.mu);
        s.cv.wait(lk, [&] { return s.pending == 0; });
    }
    out->clear();
    for (int i = 0; i < batch; ++i) {
        std::vector<Hit> r;
        for (const Hit &h : s.results[i])
            if (h.score >= threshold) r.push_back(h);
        std::sort(r.begin(), r.end(), [](const Hit &a, const Hit &c) { return a.score > c.score; });  // set.go:155
        if ((int)r.size() > limit) r.resize((size_t)limit);
        out->push_back(std::move(r));
    }
}

int main(int argc, char **argv)
{
    long long rows = 1000000;
    int dims = 512, threads = 10, batch = 1, limit = 10, device = 0, workers = 0;
    double seconds = 5.0;
    long long block_bytes = 32ll << 20;  // demo/main.go:19
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i];
        const char *v = argv[i + 1];
        if (k == "--rows") rows = atoll(v);
        else if (k == "--dims") dims = atoi(v);
        else if (k == "--threads") threads = atoi(v);
        else if (k == "--queries") batch = atoi(v);
        else if (k == "--limit") limit = atoi(v);
        else if (k == "--seconds") seconds = atof(v);
        else if (k == "--device") device = atoi(v);
        else if (k == "--workers") workers = atoi(v);
    }
    CK(cudaSetDevice(device));
    const int cap = (int)(block_bytes / ((long long)dims * 4));
    const int nblocks = (int)((rows + cap - 1) / cap);
    const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
    if (workers <= 0) workers = std::min(nblocks, cores);  // one goroutine per block, GOMAXPROCS = cores run at once
    std::vector<Block> blocks((size_t)nblocks);
    for (int b = 0; b < nblocks; ++b) {
        blocks[b].first = (long long)b * cap;
        blocks[b].height = (int)std::min<long long>(cap, rows - blocks[b].first);
        CK(cudaMalloc((void **)&blocks[b].dev, (size_t)block_bytes));
        fill_rows<<<(blocks[b].height + 7) / 8, 256>>>(blocks[b].dev, blocks[b].height, dims,
                                                      20260101ull + (unsigned long long)b);
    }
    CK(cudaDeviceSynchronize());
    std::vector<Worker> ws((size_t)workers);
    for (auto &w : ws) {
        CK(cudaMalloc((void **)&w.d_in, (size_t)batch * dims * 4));
        CK(cudaMalloc((void **)&w.d_out, (size_t)batch * cap * 4));
        if (cublasCreate(&w.handle) != CUBLAS_STATUS_SUCCESS) {
            fprintf(stderr, "cublasCreate failed\n");
            return 2;
        }
    }
    std::vector<std::thread> pool;
    for (auto &w : ws) pool.emplace_back(worker_loop, &w, &blocks, dims);

    // queries: stored-looking unit vectors (values do not matter for the timing)
    std::vector<float> hq((size_t)threads * batch * dims);
    for (size_t i = 0; i < hq.size(); ++i) hq[i] = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
    for (int t = 0; t < threads * batch; ++t) {
        double ss = 0;
        for (int c = 0; c < dims; ++c) ss += (double)hq[(size_t)t * dims + c] * hq[(size_t)t * dims + c];
        const float inv = (float)(1.0 / std::sqrt(ss));
        for (int c = 0; c < dims; ++c) hq[(size_t)t * dims + c] *= inv;
    }
    {  // warm-up: one search (cuBLAS picks and loads its kernel)
        std::vector<std::vector<Hit>> r;
        set_search(blocks, hq.data(), batch, limit, -2.f, &r);
    }
    std::atomic<long long> searches{0};
    std::vector<std::vector<double>> lat((size_t)threads);
    const auto t0 = std::chrono::steady_clock::now();
    const auto deadline = t0 + std::chrono::duration<double>(seconds);
    std::vector<std::thread> callers;
    for (int t = 0; t < threads; ++t)
        callers.emplace_back([&, t] {
            std::vector<std::vector<Hit>> r;
            while (std::chrono::steady_clock::now() < deadline) {
                const auto a = std::chrono::steady_clock::now();
                set_search(blocks, hq.data() + (size_t)t * batch * dims, batch, limit, -2.f, &r);
                lat[t].push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count());
                ++searches;
            }
        });
    for (auto &c : callers) c.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    {
        std::lock_guard<std::mutex> lk(g_qmu);
        g_stop = true;
    }
    g_qcv.notify_all();
    for (auto &p : pool) p.join();
    std::vector<double> all;
    for (auto &l : lat) all.insert(all.end(), l.begin(), l.end());
    std::sort(all.begin(), all.end());
    const double p50 = all.empty() ? 0 : all[all.size() / 2], p99 = all.empty() ? 0 : all[(size_t)(all.size() * 0.99)];
    int cuv = 0;
    cublasGetVersion(ws[0].handle, &cuv);
    printf("{\"qps\": %.3f, \"searches\": %lld, \"wall_s\": %.3f, \"p50_ms\": %.3f, \"p99_ms\": %.3f, \"rows\": %lld, "
           "\"dims\": %d, \"blocks\": %d, \"caller_threads\": %d, \"queries_per_call\": %d, \"limit\": %d, "
           "\"block_workers\": %d, \"host_cores\": %d, \"cublas_version\": %d, \"d2h_bytes_per_search\": %lld}\n",
           (double)searches.load() * batch / wall, searches.load(), wall, p50, p99, rows, dims, nblocks, threads, batch,
           limit, workers, cores, cuv, (long long)rows * batch * 4);
    return 0;
}

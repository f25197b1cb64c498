// host_rows.cpp -- host-side companions of nbb200_own_slab_to_host (several ranks, callers that keep coordinates and gradients in host
// arrays): row gather / scatter-add over the atoms a rank owns.  Plain CPU loops, no CUDA.
#include "../../include/nbabfs_b200.h"
#include <algorithm>
#include <cstdint>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

static int host_threads()
{
    static const int t = []() { const char *e = std::getenv("NBB200_HOST_THREADS"); const int v = e ? std::atoi(e) : 4; return std::max(1, std::min(v, 64)); }();
    return t;
}

// a small persistent pool (creating threads costs ~20 us each: more than the work of a piece): the workers sleep on a condition variable,
// a call hands out [k0, k1) ranges and waits for the last one
namespace {
struct Pool {
    std::mutex mu;
    std::condition_variable wake, done;
    std::vector<std::thread> workers;
    std::function<void(long, long)> body;
    long count = 0, chunk = 0;
    int parts = 0, next = 0, pending = 0;
    unsigned long generation = 0;
    bool stop = false;

    explicit Pool(int nworkers)
    {
        for (int w = 0; w < nworkers; w++) workers.emplace_back([this]() { run(); });
    }
    ~Pool()
    {
        { std::lock_guard<std::mutex> g(mu); stop = true; }
        wake.notify_all();
        for (auto &t : workers) t.join();
    }
    void run()
    {
        unsigned long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            wake.wait(lk, [&]() { return stop || (generation != seen && next < parts); });
            if (stop) return;
            seen = generation;
            while (next < parts) {
                const int p = next++;
                lk.unlock();
                body(std::min(count, p * chunk), std::min(count, (p + 1) * chunk));
                lk.lock();
                if (--pending == 0) done.notify_all();
            }
        }
    }
    void parallel(long n, int T, const std::function<void(long, long)> &f)
    {
        std::unique_lock<std::mutex> lk(mu);
        body = f; count = n; parts = T; chunk = (n + T - 1) / T; next = 1; pending = T - 1; generation++;
        lk.unlock();
        wake.notify_all();
        f(0L, std::min(n, chunk));                            // the caller takes the first part
        lk.lock();
        done.wait(lk, [&]() { return pending == 0; });
    }
};
}  // namespace

template <typename F> static void host_parallel(long count, F body)
{
    const int T = (count < 32768) ? 1 : host_threads();
    if (T == 1) { body(0L, count); return; }
    static Pool pool(host_threads() - 1);
    static std::mutex callers;                                // one parallel region at a time
    std::lock_guard<std::mutex> g(callers);
    pool.parallel(count, T, body);
}

extern "C" {

/* out[k] = x[atoms[k]] over rows of three doubles: the rows are scattered over a 24 n byte array, hence the software prefetch */
void nbb200_host_gather_rows(const double *x, const int *atoms, long count, double *out)
{
    if (x == nullptr || atoms == nullptr || out == nullptr || count <= 0) return;
    host_parallel(count, [=](long k0, long k1) {
        for (long k = k0; k < k1; k++) {
            if (k + 16 < k1) __builtin_prefetch(x + 3 * (long) atoms[k + 16]);
            const double *r = x + 3 * (long) atoms[k];
            out[3 * k] = r[0]; out[3 * k + 1] = r[1]; out[3 * k + 2] = r[2];
        }
    });
}

/* g[atoms[k]] += in[k]; the atoms of a slab are distinct, so the threads never meet */
void nbb200_host_scatter_add_rows(double *g, const int *atoms, long count, const double *in)
{
    if (g == nullptr || atoms == nullptr || in == nullptr || count <= 0) return;
    host_parallel(count, [=](long k0, long k1) {
        for (long k = k0; k < k1; k++) {
            if (k + 16 < k1) __builtin_prefetch(g + 3 * (long) atoms[k + 16], 1);
            double *r = g + 3 * (long) atoms[k];
            r[0] += in[3 * k]; r[1] += in[3 * k + 1]; r[2] += in[3 * k + 2];
        }
    });
}

/* dst[0 .. m) = src[0 .. m), threaded: a caller's pageable array into page-locked staging memory that a DMA reads next.  Streaming stores:
 * measured on the GPU box, a DMA out of lines that eight cores have just written through their caches runs at a fifth of its speed (13.4 MB:
 * 1.2 ms instead of 0.25 ms) -- the copy must not leave the data dirty in the caches. */
void nbb200_host_copy(double *dst, const double *src, long m)
{
    if (dst == nullptr || src == nullptr || m <= 0) return;
    host_parallel(m, [=](long k0, long k1) {
#if defined(__SSE2__)
        long k = k0;
        while (k < k1 && (reinterpret_cast<uintptr_t>(dst + k) & 15u) != 0) { dst[k] = src[k]; k++; }
        for (; k + 2 <= k1; k += 2) _mm_stream_pd(dst + k, _mm_loadu_pd(src + k));
        for (; k < k1; k++) dst[k] = src[k];
        _mm_sfence();
#else
        std::copy(src + k0, src + k1, dst + k0);
#endif
    });
}

/* dst[0 .. m) += src[0 .. m), threaded */
void nbb200_host_add(double *dst, const double *src, long m)
{
    if (dst == nullptr || src == nullptr || m <= 0) return;
    host_parallel(m, [=](long k0, long k1) { for (long k = k0; k < k1; k++) dst[k] += src[k]; });
}

}  // extern "C"

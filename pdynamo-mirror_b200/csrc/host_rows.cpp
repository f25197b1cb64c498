// host_rows.cpp -- host-side companions of nbb200_own_slab_to_host (several ranks, callers that keep coordinates and gradients in host
// arrays): row gather / scatter-add over the atoms a rank owns.  Plain CPU loops, no CUDA.
#include "../../include/nbabfs_b200.h"
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

static int host_threads()
{
    static const int t = []() { const char *e = std::getenv("NBB200_HOST_THREADS"); const int v = e ? std::atoi(e) : 4; return std::max(1, std::min(v, 64)); }();
    return t;
}

template <typename F> static void host_parallel(long count, F body)
{
    const int T = (count < 32768) ? 1 : host_threads();
    if (T == 1) { body(0L, count); return; }
    std::vector<std::thread> pool;
    const long chunk = (count + T - 1) / T;
    for (int t = 1; t < T; t++) pool.emplace_back([=]() { body(std::min(count, t * chunk), std::min(count, (t + 1) * chunk)); });
    body(0L, std::min(count, chunk));
    for (auto &th : pool) th.join();
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

}  // extern "C"

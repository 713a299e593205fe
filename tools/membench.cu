// membench.cu -- access-pattern floor of the relaxation sweep on one B200 (measurement aid, not product code).
// Same traffic shape as relax_kernel at C3 (n = 1M vertices, m = 10M destination-sorted edges): a 24 B/edge
// stream (src, dst int32; cost, time f64) plus ONE random gather per edge of 8 / 16 / 32 bytes from an
// L2-resident vertex array, plus the (nearly coalesced) dist[dst] read.  No winner resolution, no atomics: what
// is left is the cost of the memory pattern itself, i.e. the practical roofline of an edge-parallel sweep.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef long long ll;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void init_edges(int *src, int *dst, double *c, double *t, ll m, ll n) {
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)e * 0x9E3779B97F4A7C15ull + 12345;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        src[e] = (int)(x % (unsigned long long)n);
        dst[e] = (int)(e * n / m);
        c[e] = (double)(x & 1023) * 0.01 - 5.0;
        t[e] = 0.1 + (double)((x >> 10) & 1023) * 0.004;
    }
}

// MODE: 0 stream only; 1 + gather 8 B; 2 + gather 16 B; 4 + gather 32 B (one 256-bit load); +8: also read dist[dst]
template <int G, bool DST, int ROWS>
__global__ void __launch_bounds__(256) sweep(const int *__restrict__ src, const int *__restrict__ dst,
                                             const double *__restrict__ cost, const double *__restrict__ time,
                                             const double *dist, int ks, ll m, double lam, double *sink) {
    const int lane = threadIdx.x & 31;
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((ll)gridDim.x * blockDim.x) >> 5;
    const ll CH = 32 * ROWS, nch = m / CH;
    double acc = 0.0;
    for (ll c = warp; c < nch; c += nw) {
        int s[ROWS], d[ROWS]; double cs[ROWS], tm[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; j++) {
            const ll e = c * CH + 32 * j + lane;
            s[j] = __ldcs(src + e); d[j] = __ldcs(dst + e); cs[j] = __ldcs(cost + e); tm[j] = __ldcs(time + e);
        }
#pragma unroll
        for (int j = 0; j < ROWS; j++) {
            double w = cs[j] - lam * tm[j];
            if (G == 1) w += __ldcg(dist + (size_t)s[j] * ks);
            if (G == 2) { double2 v = __ldcg(reinterpret_cast<const double2 *>(dist + (size_t)s[j] * ks)); w += v.x + v.y; }
            if (G == 4) {
                double a, b, cc, dd;
                asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(cc), "=d"(dd) : "l"(dist + (size_t)s[j] * ks));
                w += a + b + cc + dd;
            }
            if (DST) {
                if (G == 4) {
                    double a, b, cc, dd;
                    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(cc), "=d"(dd) : "l"(dist + (size_t)d[j] * ks));
                    w -= a + b + cc + dd;
                } else w -= __ldcg(dist + (size_t)d[j] * ks);
            } else w += (double)d[j];
            acc += w;
        }
    }
    if (acc == 1.2345e300) *sink = acc;
}

template <int G, bool DST, int ROWS>
static void run(const char *name, int bps, const int *src, const int *dst, const double *c, const double *t, const double *dist,
                int ks, ll m, double *sink) {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = sms * bps;
    for (int i = 0; i < 5; i++) sweep<G, DST, ROWS><<<grid, 256>>>(src, dst, c, t, dist, ks, m, -3.0, sink);
    float best = 1e9f, sum = 0;
    for (int i = 0; i < 20; i++) {
        CK(cudaEventRecord(e0));
        sweep<G, DST, ROWS><<<grid, 256>>>(src, dst, c, t, dist, ks, m, -3.0, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best; sum += ms;
    }
    printf("%-34s rows=%d ctas/SM=%d  best %7.1f us  mean %7.1f us  -> %6.0f GB/s of the 24 B/edge stream\n", name, ROWS, bps,
           best * 1e3, sum / 20 * 1e3, 24.0 * m / (best * 1e-3) / 1e9);
}

int main(int argc, char **argv) {
    const ll n = argc > 1 ? atoll(argv[1]) : 1000000, m = argc > 2 ? atoll(argv[2]) : 10000000;
    int *src, *dst; double *c, *t, *dist, *sink;
    CK(cudaMalloc(&src, m * 4)); CK(cudaMalloc(&dst, m * 4)); CK(cudaMalloc(&c, m * 8)); CK(cudaMalloc(&t, m * 8));
    CK(cudaMalloc(&dist, n * 8 * 8)); CK(cudaMalloc(&sink, 8));
    CK(cudaMemset(dist, 0, n * 64));
    init_edges<<<1184, 256>>>(src, dst, c, t, m, n);
    CK(cudaDeviceSynchronize());
    printf("n=%lld m=%lld  (edge stream %.0f MB)\n", n, m, 24.0 * m / 1e6);
    for (int bps : {4, 6, 8}) {
        run<0, false, 2>("stream only", bps, src, dst, c, t, dist, 1, m, sink);
        run<1, false, 2>("stream + gather 8 B [n][1]", bps, src, dst, c, t, dist, 1, m, sink);
        run<1, true, 2>("stream + gather 8 B + dst [n][1]", bps, src, dst, c, t, dist, 1, m, sink);
        run<1, true, 2>("stream + gather 8 B + dst [n][4]", bps, src, dst, c, t, dist, 4, m, sink);
        run<2, true, 2>("stream + gather 16 B + dst [n][2]", bps, src, dst, c, t, dist, 2, m, sink);
        run<4, false, 2>("stream + gather 32 B [n][4]", bps, src, dst, c, t, dist, 4, m, sink);
        run<4, true, 2>("stream + gather 32 B + dst [n][4]", bps, src, dst, c, t, dist, 4, m, sink);
    }
    run<1, true, 1>("stream + gather 8 B + dst [n][1]", 8, src, dst, c, t, dist, 1, m, sink);
    run<1, true, 4>("stream + gather 8 B + dst [n][1]", 4, src, dst, c, t, dist, 1, m, sink);
    run<4, true, 1>("stream + gather 32 B + dst [n][4]", 8, src, dst, c, t, dist, 4, m, sink);
    run<4, true, 4>("stream + gather 32 B + dst [n][4]", 4, src, dst, c, t, dist, 4, m, sink);
    return 0;
}

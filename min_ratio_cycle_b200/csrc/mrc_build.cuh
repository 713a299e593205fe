// mrc_build.cuh -- device-side array build: the stable sort by destination of
// _build_arrays_if_needed (np.argsort(V, kind="stable") + four fancy-index gathers, solver.py:476-478), and the
// duplicate-(u,v) census that tells the caller whether _remove_dominated_edges (:531-575) has anything to look at.
// At m = 10^7 the NumPy versions cost ~2 s and ~1 s on the host; here: one LSD radix sort of (dst, index) pairs
// (cub::DeviceRadixSort, stable) + one gather sweep, and one 64-bit key sort + adjacent-compare.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

typedef long long ll;
typedef unsigned long long ull;

__global__ void build_iota_kernel(int *p, ll m) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (ll)gridDim.x * blockDim.x) p[i] = (int)i;
}
__global__ void build_gather_kernel(const int *order, const int *src, const double *cost, const double *time, int *osrc,
                                    double *oc, double *ot, ll m) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (ll)gridDim.x * blockDim.x) {
        const int o = order[i];
        osrc[i] = src[o]; oc[i] = cost[o]; ot[i] = time[o];
    }
}
__global__ void build_pair_keys_kernel(const int *src, const int *dst, ull *keys, ll m) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (ll)gridDim.x * blockDim.x)
        keys[i] = ((ull)(unsigned)src[i] << 32) | (unsigned)dst[i];
}
__global__ void build_count_dups_kernel(const ull *keys, ll m, ull *count) {
    unsigned c = 0;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x + 1; i < m; i += (ll)gridDim.x * blockDim.x)
        c += keys[i] == keys[i - 1];
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (ull)c);
}

// out_off[v] = first position in the src-sorted edge list whose source is >= v (lower bound), v in [0, n]
__global__ void build_out_offsets_kernel(const int *sorted_src, ll m, ll n, int *out_off) {
    for (ll v = (ll)blockIdx.x * blockDim.x + threadIdx.x; v <= n; v += (ll)gridDim.x * blockDim.x) {
        ll lo = 0, hi = m;
        while (lo < hi) {
            const ll mid = (lo + hi) >> 1;
            if (sorted_src[mid] < (int)v) lo = mid + 1; else hi = mid;
        }
        out_off[v] = (int)lo;
    }
}

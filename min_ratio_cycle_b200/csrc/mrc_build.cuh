// mrc_build.cuh -- device-side array build: the stable sort by destination of
// _build_arrays_if_needed (np.argsort(V, kind="stable") + four fancy-index gathers, solver.py:476-478), and the
// duplicate-(u,v) census that tells the caller whether _remove_dominated_edges (:531-575) has anything to look at.
// At m = 10^7 the NumPy versions cost ~2 s and ~1 s on the host; here: one LSD radix sort of (dst, index) pairs
// (cub::DeviceRadixSort, stable) + one gather sweep, and one 64-bit key sort + adjacent-compare.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
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

// ---- _remove_dominated_edges (solver.py:531-575) on the device ------------------------------------------------------
// pos runs over the edges sorted by (u,v) pair; idx[pos] is the insertion index.  An edge is dropped when another edge
// of its pair has cost <= and time <= with one of them strict -- the reference's double loop over each group of
// parallel edges, here a scan left and right while the pair key stays the same (groups of parallel edges are tiny).
__global__ void build_dominated_kernel(const ull *keys, const int *idx, const double *cost, const double *time, ll m,
                                       unsigned char *keep /* by insertion index */, ull *counts /* [0] repeated pairs, [1] dominated */) {
    unsigned dup = 0, dom = 0;
    for (ll p = (ll)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (ll)gridDim.x * blockDim.x) {
        const ull k = keys[p];
        const int me = idx[p];
        bool dominated = false;
        if (p > 0 && keys[p - 1] == k) dup++;
        if ((p > 0 && keys[p - 1] == k) || (p + 1 < m && keys[p + 1] == k)) {
            const double c = cost[me], t = time[me];
            for (ll q = p - 1; q >= 0 && keys[q] == k && !dominated; q--) {
                const double cq = cost[idx[q]], tq = time[idx[q]];
                dominated = cq <= c && tq <= t && (cq < c || tq < t);
            }
            for (ll q = p + 1; q < m && keys[q] == k && !dominated; q++) {
                const double cq = cost[idx[q]], tq = time[idx[q]];
                dominated = cq <= c && tq <= t && (cq < c || tq < t);
            }
        }
        keep[me] = dominated ? 0 : 1;
        dom += dominated;
    }
    dup = __reduce_add_sync(0xffffffffu, dup);
    dom = __reduce_add_sync(0xffffffffu, dom);
    if ((threadIdx.x & 31) == 0) { if (dup) atomicAdd(counts, (ull)dup); if (dom) atomicAdd(counts + 1, (ull)dom); }
}
__global__ void build_gather_dst_kernel(const int *kept, const int *dst, int *out, ll m) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (ll)gridDim.x * blockDim.x) out[i] = dst[kept[i]];
}
// sorted edge i came from insertion index order[i]; with removals, renumber to the index among the survivors
__global__ void build_renumber_kernel(int *order, const int *rank_of, ll m) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (ll)gridDim.x * blockDim.x) order[i] = rank_of[order[i]];
}
__global__ void build_widen_flags_kernel(const unsigned char *keep, int *wide, ll m) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (ll)gridDim.x * blockDim.x) wide[i] = keep[i];
}
// starts[v] = first sorted position whose destination is >= v, v in [0, n] (CSR by destination, solver.py:480-489)
__global__ void build_in_offsets_kernel(const int *sorted_dst, ll m, ll n, int *starts) {
    for (ll v = (ll)blockIdx.x * blockDim.x + threadIdx.x; v <= n; v += (ll)gridDim.x * blockDim.x) {
        ll lo = 0, hi = m;
        while (lo < hi) {
            const ll mid = (lo + hi) >> 1;
            if (sorted_dst[mid] < (int)v) lo = mid + 1; else hi = mid;
        }
        starts[v] = (int)lo;
    }
}

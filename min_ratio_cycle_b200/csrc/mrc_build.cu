// mrc_build.cu -- device-side array build entry point (own translation unit: CUB is slow to compile).
#include <algorithm>

#include "mrc_build.cuh"
#include "mrc_host.h"

// ---------------------------------------------------------------------------------------- device-side array build
/*
 * Stable sort of insertion-order edges by destination on the device (replaces np.argsort(V, kind="stable") and
 * the gathers of _build_arrays_if_needed, solver.py:476-478).  Host arrays in, host arrays out; `order_out[i]` is
 * the insertion index of sorted edge i.  dup_pairs_out (may be NULL): number of edges whose (u,v) pair also occurs
 * earlier -- 0 means _remove_dominated_edges (:531-575) has nothing to do.
 */
extern "C" int mrc_sort_edges_by_dst(int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const double *cost,
                                     const double *time, int device, int32_t *src_out, int32_t *dst_out,
                                     double *cost_out, double *time_out, int32_t *order_out, int64_t *dup_pairs_out) {
    if (n <= 0 || m <= 0 || n > 0x7ffffff0ll || m > 0x7ffffff0ll) return set_err(MRC_ERR_ARG, "n, m must be positive and fit int32");
    if (!src || !dst || !cost || !time || !src_out || !dst_out || !cost_out || !time_out)
        return set_err(MRC_ERR_ARG, "NULL array");
    int ndev = 0;
    MRC_TRY(mrc_device_count(&ndev));
    if (device < 0 || device >= ndev) return set_err(MRC_ERR_ARG, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int *d_src = nullptr, *d_dst = nullptr, *d_dst2 = nullptr, *d_idx = nullptr, *d_ord = nullptr, *d_src2 = nullptr;
    double *d_c = nullptr, *d_t = nullptr, *d_c2 = nullptr, *d_t2 = nullptr;
    ull *d_k = nullptr, *d_k2 = nullptr, *d_cnt = nullptr;
    void *d_tmp = nullptr;
    int rc = [&]() -> int {
        const size_t M = (size_t)m;
        CUDA_TRY(cudaMallocAsync((void **)&d_src, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_dst, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_dst2, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_idx, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_ord, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_src2, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_c, M * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_t, M * 8, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_c2, M * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_t2, M * 8, st));
        CUDA_TRY(cudaMemcpyAsync(d_src, src, M * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_dst, dst, M * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_c, cost, M * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_t, time, M * 8, cudaMemcpyHostToDevice, st));
        const int blocks = (int)std::min<ll>((m + 255) / 256, 148 * 16);
        build_iota_kernel<<<blocks, 256, 0, st>>>(d_idx, m);
        int bits = 1;
        while (((ll)1 << bits) < n) bits++;
        size_t tmp_bytes = 0, tmp2 = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_dst, d_dst2, d_idx, d_ord, (int)m, 0, bits, st));
        if (dup_pairs_out) {
            CUDA_TRY(cudaMallocAsync((void **)&d_k, M * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_k2, M * 8, st));
            CUDA_TRY(cudaMallocAsync((void **)&d_cnt, 8, st));
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp2, d_k, d_k2, (int)m, 0, 32 + bits, st));
        }
        tmp_bytes = std::max(tmp_bytes, tmp2);
        CUDA_TRY(cudaMallocAsync(&d_tmp, tmp_bytes, st));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_dst, d_dst2, d_idx, d_ord, (int)m, 0, bits, st));
        build_gather_kernel<<<blocks, 256, 0, st>>>(d_ord, d_src, d_c, d_t, d_src2, d_c2, d_t2, m);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(src_out, d_src2, M * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(dst_out, d_dst2, M * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cost_out, d_c2, M * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(time_out, d_t2, M * 8, cudaMemcpyDeviceToHost, st));
        if (order_out) CUDA_TRY(cudaMemcpyAsync(order_out, d_ord, M * 4, cudaMemcpyDeviceToHost, st));
        ull dups = 0;
        if (dup_pairs_out) {
            CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 8, st));
            build_pair_keys_kernel<<<blocks, 256, 0, st>>>(d_src, d_dst, d_k, m);
            CUDA_TRY(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_k, d_k2, (int)m, 0, 32 + bits, st));
            build_count_dups_kernel<<<blocks, 256, 0, st>>>(d_k2, m, d_cnt);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaMemcpyAsync(&dups, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        if (dup_pairs_out) *dup_pairs_out = (int64_t)dups;
        return MRC_OK;
    }();
    void *ptrs[] = {d_src, d_dst, d_dst2, d_idx, d_ord, d_src2, d_c, d_t, d_c2, d_t2, d_k, d_k2, d_cnt, d_tmp};
    for (void *q : ptrs) if (q) cudaFreeAsync(q, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

/*
 * Out-edge index of a device-resident graph (for the sparse relaxation passes): out_edge[out_off[u] .. out_off[u+1])
 * are the indices (into the destination-sorted arrays) of the edges leaving u, in increasing edge index.  One stable
 * radix sort of (src, edge index) pairs + n+1 binary searches.  The two result arrays are allocated here
 * (stream-ordered) and owned by the caller.
 */
int mrc_build_out_csr(cudaStream_t st, int64_t n, int64_t m, const int32_t *d_src, int32_t **d_off_out, int32_t **d_edge_out) {
    int *d_idx = nullptr, *d_keys = nullptr, *d_off = nullptr, *d_edge = nullptr;
    void *d_tmp = nullptr;
    int rc = [&]() -> int {
        const size_t M = (size_t)m;
        CUDA_TRY(cudaMallocAsync((void **)&d_idx, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_keys, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_edge, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_off, ((size_t)n + 1) * 4, st));
        const int blocks = (int)std::min<ll>((m + 255) / 256, 148 * 16);
        build_iota_kernel<<<blocks, 256, 0, st>>>(d_idx, m);
        int bits = 1;
        while (((ll)1 << bits) < n) bits++;
        size_t tmp_bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_src, d_keys, d_idx, d_edge, (int)m, 0, bits, st));
        CUDA_TRY(cudaMallocAsync(&d_tmp, tmp_bytes, st));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_src, d_keys, d_idx, d_edge, (int)m, 0, bits, st));
        build_out_offsets_kernel<<<(int)std::min<ll>((n + 256) / 256, 148 * 16), 256, 0, st>>>(d_keys, m, n, d_off);
        CUDA_TRY(cudaGetLastError());
        return MRC_OK;
    }();
    void *ptrs[] = {d_idx, d_keys, d_tmp};
    for (void *q : ptrs) if (q) cudaFreeAsync(q, st);
    if (rc) { if (d_off) cudaFreeAsync(d_off, st); if (d_edge) cudaFreeAsync(d_edge, st); return rc; }
    *d_off_out = d_off; *d_edge_out = d_edge;
    return MRC_OK;
}

/*
 * The whole of _preprocess_graph + _build_arrays_if_needed (solver.py:450-575) on device-resident insertion-order arrays:
 * dominated parallel edges removed (optional), survivors stable-sorted by destination, CSR offsets.  Outputs go to
 * caller-owned device buffers of m elements (o_starts: n + 1).  o_keep (may be NULL): one byte per INPUT edge, 1 = kept.
 * o_order[i] = index among the SURVIVORS (insertion order) of sorted edge i.  *m_out = number of survivors.
 */
int mrc_build_sorted_device(cudaStream_t st, int64_t n, int64_t m, const int32_t *d_src, const int32_t *d_dst, const double *d_c,
                            const double *d_t, int remove_dominated, int32_t *o_src, int32_t *o_dst, double *o_c, double *o_t,
                            int32_t *o_order, int32_t *o_starts, unsigned char *o_keep, int64_t *m_out, int64_t *dups_out,
                            int64_t *removed_out) {
    const size_t M = (size_t)m;
    int *d_idx = nullptr, *d_idx2 = nullptr, *d_kept = nullptr, *d_key32 = nullptr, *d_wide = nullptr, *d_rank = nullptr, *d_nsel = nullptr;
    ull *d_k = nullptr, *d_k2 = nullptr, *d_cnt = nullptr;
    unsigned char *d_keep = nullptr;
    void *d_tmp = nullptr;
    int rc = [&]() -> int {
        const int blocks = (int)std::min<ll>((m + 255) / 256, 148 * 16);
        int bits = 1;
        while (((ll)1 << bits) < n) bits++;
        CUDA_TRY(cudaMallocAsync((void **)&d_idx, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_idx2, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_kept, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_key32, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_k, M * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_k2, M * 8, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_cnt, 16, st)); CUDA_TRY(cudaMallocAsync((void **)&d_nsel, 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_keep, M, st));
        size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t1, d_k, d_k2, d_idx, d_idx2, (int)m, 0, 32 + bits, st));
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, t2, d_key32, o_dst, d_kept, o_order, (int)m, 0, bits, st));
        CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, t3, d_idx, d_keep, d_kept, d_nsel, (int)m, st));
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, t4, d_idx, d_idx2, (int)m, st));
        const size_t tmp_bytes = std::max(std::max(t1, t2), std::max(t3, t4));
        CUDA_TRY(cudaMallocAsync(&d_tmp, tmp_bytes, st));
        build_iota_kernel<<<blocks, 256, 0, st>>>(d_idx, m);
        // ---- repeated (u,v) pairs and, among them, dominated edges
        ull counts[2] = {0, 0};
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 16, st));
        build_pair_keys_kernel<<<blocks, 256, 0, st>>>(d_src, d_dst, d_k, m);
        size_t tb = tmp_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_k, d_k2, d_idx, d_idx2, (int)m, 0, 32 + bits, st));
        build_dominated_kernel<<<blocks, 256, 0, st>>>(d_k2, d_idx2, d_c, d_t, m, d_keep, d_cnt);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(counts, d_cnt, 16, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const ll removed = remove_dominated ? (ll)counts[1] : 0;
        ll mk = m;
        const int *kept = d_idx;                    // identity: nothing removed
        if (removed > 0) {
            tb = tmp_bytes;
            CUDA_TRY(cub::DeviceSelect::Flagged(d_tmp, tb, d_idx, d_keep, d_kept, d_nsel, (int)m, st));   // stable: insertion order kept
            kept = d_kept;
            mk = m - removed;
            // rank among the survivors of every input edge
            CUDA_TRY(cudaMallocAsync((void **)&d_wide, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_rank, M * 4, st));
            build_widen_flags_kernel<<<blocks, 256, 0, st>>>(d_keep, d_wide, m);
            tb = tmp_bytes;
            CUDA_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_wide, d_rank, (int)m, st));
        } else if (!remove_dominated) {
            CUDA_TRY(cudaMemsetAsync(d_keep, 1, M, st));
        }
        // ---- stable sort of the survivors by destination + gathers
        build_gather_dst_kernel<<<blocks, 256, 0, st>>>(kept, d_dst, d_key32, mk);
        tb = tmp_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tb, d_key32, o_dst, kept, o_order, (int)mk, 0, bits, st));
        build_gather_kernel<<<blocks, 256, 0, st>>>(o_order, d_src, d_c, d_t, o_src, o_c, o_t, mk);
        if (removed > 0) build_renumber_kernel<<<blocks, 256, 0, st>>>(o_order, d_rank, mk);
        build_in_offsets_kernel<<<(int)std::min<ll>((n + 256) / 256, 148 * 16), 256, 0, st>>>(o_dst, mk, n, o_starts);
        CUDA_TRY(cudaGetLastError());
        if (o_keep) CUDA_TRY(cudaMemcpyAsync(o_keep, d_keep, M, cudaMemcpyDeviceToDevice, st));
        *m_out = mk;
        if (dups_out) *dups_out = (int64_t)counts[0];
        if (removed_out) *removed_out = removed;
        return MRC_OK;
    }();
    void *ptrs[] = {d_idx, d_idx2, d_kept, d_key32, d_wide, d_rank, d_nsel, d_k, d_k2, d_cnt, d_keep, d_tmp};
    for (void *q : ptrs) if (q) cudaFreeAsync(q, st);
    return rc;
}

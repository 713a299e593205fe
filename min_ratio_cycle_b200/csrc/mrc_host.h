// mrc_host.h -- host-side error plumbing shared by the translation units of libmrc_b200.so
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "mrc_b200.h"

extern thread_local std::string g_mrc_err;   // defined in mrc_b200.cu; read by mrc_last_error()

static inline int set_err(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_mrc_err = buf;
    return code;
}
#define CUDA_TRY(x)                                                                                     \
    do {                                                                                                \
        cudaError_t e_ = (x);                                                                           \
        if (e_ != cudaSuccess)                                                                          \
            return set_err(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? MRC_ERR_NO_DEVICE \
                                                                                        : MRC_ERR_CUDA, \
                           "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);    \
    } while (0)
#define DALLOC(ptr, bytes) CUDA_TRY(cudaMallocAsync((void **)&(ptr), (bytes), g->stream))
#define MRC_TRY(x)            \
    do {                      \
        int rc_ = (x);        \
        if (rc_) return rc_;  \
    } while (0)

// mrc_build.cu: out-edge (by source) index of a device-resident destination-sorted graph
int mrc_build_out_csr(cudaStream_t st, int64_t n, int64_t m, const int32_t *d_src, int32_t **d_off_out, int32_t **d_edge_out);
// mrc_build.cu: _preprocess_graph + _build_arrays_if_needed on device-resident insertion-order arrays
int mrc_build_sorted_device(cudaStream_t st, int64_t n, int64_t m, const int32_t *d_src, const int32_t *d_dst, const double *d_c,
                            const double *d_t, int remove_dominated, int32_t *o_src, int32_t *o_dst, double *o_c, double *o_t,
                            int32_t *o_order, int32_t *o_starts, unsigned char *o_keep, int64_t *m_out, int64_t *dups_out,
                            int64_t *removed_out);

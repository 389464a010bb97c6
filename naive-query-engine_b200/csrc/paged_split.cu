// paged_split.cu -- host side of the paged partition streams (see paged_split.cuh)
#include <cstdlib>
#include <cstring>

#include "paged_split.cuh"

int32_t nqe_ps_create(nqe_ctx *ctx, int64_t max_rows, int P, PagedStreams *ps) {
    memset(ps, 0, sizeof *ps);
    if (P < 1 || P > PS_MAX_PARTS) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "internal: %d partitions", P);
    const uint64_t pages = (uint64_t)((max_rows + PS_PAGE_ROWS - 1) / PS_PAGE_ROWS);
    ps->P = P;
    ps->max_pages = (uint32_t)(pages + (uint64_t)P + 1);
    ps->pt_stride = (uint32_t)(pages + 2); // one partition may receive every row
    void *pool = nullptr, *meta = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &pool, (size_t)ps->max_pages * PS_PAGE_ROWS * sizeof(ulonglong2)));
    // meta: [P] cursors | pool_next (+pad) | page table
    const size_t meta_bytes = (size_t)P * 8 + 16 + (size_t)P * ps->pt_stride * 4;
    int32_t rc = nqe_dev_alloc(ctx, &meta, meta_bytes);
    if (rc != NQE_OK) {
        nqe_dev_free(ctx, pool);
        return rc;
    }
    ps->pool = (ulonglong2 *)pool;
    ps->cursor = (unsigned long long *)meta;
    ps->pool_next = (unsigned int *)(ps->cursor + P);
    ps->pt = (unsigned int *)(ps->cursor + P + 2);
    ps->status = (uint32_t *)(ctx->d_scratch + 1);
    NQE_CUDA(ctx, cudaMemsetAsync(meta, 0, meta_bytes, ctx->stream));
    return NQE_OK;
}

void nqe_ps_destroy(nqe_ctx *ctx, PagedStreams *ps) {
    if (ps->pool) nqe_dev_free(ctx, ps->pool);
    if (ps->cursor) nqe_dev_free(ctx, ps->cursor);
    memset(ps, 0, sizeof *ps);
}

int nqe_ps_split_shape() {
    static int shape = -1;
    if (shape < 0) {
        const char *e = getenv("NQE_PS_SPLIT_SHAPE");
        shape = e ? atoi(e) : 0;
        if (shape < 0 || shape > 1) shape = 0;
    }
    return shape;
}


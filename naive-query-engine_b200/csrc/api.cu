// api.cu -- context, device tables, host<->HBM transfer (the ScanPlan side of the boundary)
#include <cstring>

#include <cstdlib>

#include "nqe_internal.cuh"

int32_t nqe_fail(nqe_ctx *ctx, int32_t code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}

uint64_t nqe_next_pow2(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

extern "C" int32_t nqe_abi_version(void) { return NQE_ABI_VERSION; }

extern "C" int32_t nqe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static const size_t kStageBytes = 4 * (size_t)(16 << 20); // 4 x 16 MiB pinned ring

extern "C" int32_t nqe_ctx_create(int32_t device, nqe_ctx **out) {
    if (!out) return NQE_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return NQE_ERR_CUDA; // no CPU fallback: the CUDA path is the only path
    }
    if (device < 0 || device >= n) return NQE_ERR_INVALID_ARG;
    nqe_ctx *ctx = new nqe_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return NQE_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return NQE_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return NQE_ERR_CUDA; }
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaMallocHost(&ctx->h_scratch, 64 * sizeof(uint64_t));
    cudaMalloc(&ctx->d_scratch, 64 * sizeof(uint64_t));
    // keep freed blocks cached in the default pool (outputs are sized worst-case)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thresh = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
    }
    if (const char *e = getenv("NQE_L2_FETCH")) { // tuning knob: DRAM->L2 fetch granularity hint (32/64/128 bytes)
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
        cudaGetLastError();
    }
    if (cudaGetLastError() != cudaSuccess || !ctx->h_scratch || !ctx->d_scratch) {
        nqe_ctx_destroy(ctx);
        return NQE_ERR_CUDA;
    }
    *out = ctx;
    return NQE_OK;
}

extern "C" void nqe_ctx_destroy(nqe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
}

extern "C" const char *nqe_last_error(const nqe_ctx *ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

extern "C" int32_t nqe_ctx_set_stream(nqe_ctx *ctx, void *s) {
    if (!ctx) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (s) {
        ctx->stream = (cudaStream_t)s;
        ctx->own_stream = false;
    } else {
        NQE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return NQE_OK;
}

extern "C" int32_t nqe_ctx_sync(nqe_ctx *ctx) {
    if (!ctx) return NQE_ERR_INVALID_ARG;
    NQE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQE_OK;
}

extern "C" int64_t nqe_ctx_kernel_launches(const nqe_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" double nqe_ctx_last_op_ms(const nqe_ctx *ctx) { return ctx ? ctx->last_op_ms : 0.0; }

int32_t nqe_dev_alloc(nqe_ctx *ctx, void **p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 256;
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return nqe_fail(ctx, e == cudaErrorMemoryAllocation ? NQE_ERR_OOM : NQE_ERR_CUDA,
                        "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    return NQE_OK;
}

void nqe_dev_free(nqe_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

int32_t nqe_table_new(nqe_ctx *ctx, int64_t nrows, nqe_table **out) {
    nqe_table *t = new nqe_table();
    t->ctx = ctx;
    t->nrows = nrows;
    *out = t;
    return NQE_OK;
}

int32_t nqe_column_alloc(nqe_ctx *ctx, int32_t dtype, int64_t n, bool with_validity, DevColumn *c) {
    c->dtype = dtype;
    c->length = n;
    c->null_count = 0;
    c->owned = true;
    NQE_TRY(nqe_dev_alloc(ctx, &c->values, nqe_values_bytes(dtype, n)));
    if (with_validity) NQE_TRY(nqe_dev_alloc(ctx, (void **)&c->validity, nqe_bitmap_bytes(n)));
    return NQE_OK;
}

void nqe_column_release(nqe_ctx *ctx, DevColumn *c) {
    if (c->owned) {
        nqe_dev_free(ctx, c->values);
        nqe_dev_free(ctx, c->validity);
        nqe_dev_free(ctx, c->data);
    }
    c->values = nullptr;
    c->validity = nullptr;
    c->data = nullptr;
}

extern "C" void nqe_table_free(nqe_table *t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    for (auto &c : t->cols) nqe_column_release(t->ctx, &c);
    delete t;
}

extern "C" int64_t nqe_table_num_rows(const nqe_table *t) { return t ? t->nrows : -1; }
extern "C" int32_t nqe_table_num_columns(const nqe_table *t) { return t ? (int32_t)t->cols.size() : -1; }

extern "C" int32_t nqe_table_column(const nqe_table *t, int32_t i, nqe_column_desc *out) {
    if (!t || !out || i < 0 || i >= (int32_t)t->cols.size()) return NQE_ERR_INVALID_ARG;
    const DevColumn &c = t->cols[i];
    out->dtype = c.dtype;
    out->reserved = 0;
    out->length = t->nrows;
    out->null_count = c.null_count;
    out->values = c.values;
    out->validity = c.validity;
    out->data = c.data;
    out->data_bytes = c.data_bytes;
    return NQE_OK;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Host -> device copy.  Pinned sources are DMA'd directly; pageable sources go
// through a 4-deep pinned ring so the CPU memcpy of chunk i+1 overlaps the DMA
// of chunk i.
static int32_t h2d(nqe_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return NQE_OK;
    if (is_pinned(src)) {
        NQE_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
        return NQE_OK;
    }
    if (!ctx->h_stage) {
        NQE_CUDA(ctx, cudaMallocHost(&ctx->h_stage, kStageBytes));
        ctx->stage_bytes = kStageBytes;
    }
    const size_t chunk = ctx->stage_bytes / 4;
    cudaEvent_t ev[4];
    for (auto &e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    size_t off = 0;
    int k = 0;
    int32_t rc = NQE_OK;
    while (off < bytes) {
        size_t n = bytes - off < chunk ? bytes - off : chunk;
        int slot = k & 3;
        if (k >= 4) cudaEventSynchronize(ev[slot]);
        memcpy(ctx->h_stage + slot * chunk, (const uint8_t *)src + off, n);
        cudaError_t e = cudaMemcpyAsync((uint8_t *)dst + off, ctx->h_stage + slot * chunk, n,
                                        cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { rc = nqe_fail(ctx, NQE_ERR_CUDA, "h2d: %s", cudaGetErrorString(e)); break; }
        cudaEventRecord(ev[slot], ctx->stream);
        off += n;
        k++;
    }
    cudaStreamSynchronize(ctx->stream);
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

// Device -> host copy.  Pinned destinations are DMA'd directly; pageable ones go through the same 4-deep pinned ring
// as the uploads: the DMA of chunk i+1 overlaps the CPU memcpy of chunk i out of the ring (a cudaMemcpy into pageable
// memory runs at ~4.5 GB/s here, the ring at the speed of one core's memcpy).
static int32_t d2h(nqe_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return NQE_OK;
    if (is_pinned(dst) || bytes < ((size_t)1 << 20)) {
        NQE_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return NQE_OK;
    }
    if (!ctx->h_stage) {
        NQE_CUDA(ctx, cudaMallocHost(&ctx->h_stage, kStageBytes));
        ctx->stage_bytes = kStageBytes;
    }
    const size_t chunk = ctx->stage_bytes / 4;
    cudaEvent_t ev[4];
    for (auto &e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    const size_t n_chunks = (bytes + chunk - 1) / chunk;
    int32_t rc = NQE_OK;
    for (size_t k = 0; k < n_chunks + 3 && rc == NQE_OK; k++) {
        if (k < n_chunks) { // issue the DMA of chunk k into ring slot k % 4 (drained three iterations ago)
            const size_t off = k * chunk, n = bytes - off < chunk ? bytes - off : chunk;
            if (cudaMemcpyAsync(ctx->h_stage + (k & 3) * chunk, (const uint8_t *)src + off, n, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                rc = nqe_fail(ctx, NQE_ERR_CUDA, "d2h: %s", cudaGetErrorString(cudaGetLastError()));
            cudaEventRecord(ev[k & 3], ctx->stream);
        }
        if (k >= 3) { // drain chunk k - 3
            const size_t j = k - 3, off = j * chunk, n = bytes - off < chunk ? bytes - off : chunk;
            cudaEventSynchronize(ev[j & 3]);
            memcpy((uint8_t *)dst + off, ctx->h_stage + (j & 3) * chunk, n);
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    return rc;
}

static int32_t check_desc(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, int64_t *nrows) {
    if (!cols && n_cols > 0) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "null column array");
    if (n_cols < 0) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "negative column count");
    *nrows = n_cols ? cols[0].length : 0;
    for (int i = 0; i < n_cols; i++) {
        const nqe_column_desc &c = cols[i];
        if (c.dtype < NQE_BOOL || c.dtype > NQE_UTF8) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "column %d: bad dtype %d", i, c.dtype);
        if (c.length < 0 || c.data_bytes < 0) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "column %d: negative length", i);
        if (c.length != *nrows) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "column %d: length %lld != %lld", i, (long long)c.length, (long long)*nrows);
        if (c.length > 0 && !c.values) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "column %d: null values buffer", i);
        if (c.null_count > 0 && !c.validity) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "column %d: null_count > 0 without validity", i);
    }
    return NQE_OK;
}

extern "C" int32_t nqe_table_upload(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, nqe_table **out) {
    if (!ctx || !out) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    int64_t nrows = 0;
    NQE_TRY(check_desc(ctx, cols, n_cols, &nrows));
    nqe_table *t;
    nqe_table_new(ctx, nrows, &t);
    t->cols.resize(n_cols);
    for (int i = 0; i < n_cols; i++) {
        const nqe_column_desc &s = cols[i];
        DevColumn &d = t->cols[i];
        bool with_valid = s.validity != nullptr && s.null_count != 0;
        int32_t rc = nqe_column_alloc(ctx, s.dtype, nrows, with_valid, &d);
        size_t vbytes = s.dtype == NQE_BOOL ? (size_t)((nrows + 7) / 8)
                        : s.dtype == NQE_UTF8 ? (size_t)(nrows + 1) * 4 : (size_t)nrows * 8;
        if (rc == NQE_OK && s.dtype == NQE_BOOL)
            if (cudaMemsetAsync(d.values, 0, nqe_bitmap_bytes(nrows), ctx->stream) != cudaSuccess) rc = NQE_ERR_CUDA;
        if (rc == NQE_OK) rc = h2d(ctx, d.values, s.values, vbytes);
        if (rc == NQE_OK && with_valid) {
            if (cudaMemsetAsync(d.validity, 0, nqe_bitmap_bytes(nrows), ctx->stream) != cudaSuccess) rc = NQE_ERR_CUDA;
            if (rc == NQE_OK) rc = h2d(ctx, d.validity, s.validity, (size_t)((nrows + 7) / 8));
            d.null_count = s.null_count;
        }
        if (rc == NQE_OK && s.dtype == NQE_UTF8) {
            d.data_bytes = s.data_bytes;
            rc = nqe_dev_alloc(ctx, (void **)&d.data, (size_t)s.data_bytes + 64);
            if (rc == NQE_OK) rc = h2d(ctx, d.data, s.data, (size_t)s.data_bytes);
        }
        if (rc != NQE_OK) {
            nqe_table_free(t);
            return rc;
        }
    }
    // pinned sources are DMA'd asynchronously: the caller's buffers are free to go when this returns
    NQE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = t;
    return NQE_OK;
}

extern "C" int32_t nqe_table_from_device(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, nqe_table **out) {
    if (!ctx || !out) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    int64_t nrows = 0;
    NQE_TRY(check_desc(ctx, cols, n_cols, &nrows));
    nqe_table *t;
    nqe_table_new(ctx, nrows, &t);
    t->cols.resize(n_cols);
    for (int i = 0; i < n_cols; i++) {
        DevColumn &d = t->cols[i];
        d.dtype = cols[i].dtype;
        d.length = nrows;
        d.null_count = cols[i].validity ? cols[i].null_count : 0;
        d.values = const_cast<void *>(cols[i].values);
        d.validity = d.null_count ? const_cast<uint8_t *>(cols[i].validity) : nullptr;
        d.data = const_cast<uint8_t *>(cols[i].data);
        d.data_bytes = cols[i].data_bytes;
        d.owned = false;
    }
    *out = t;
    return NQE_OK;
}

extern "C" int32_t nqe_table_download_column(nqe_ctx *ctx, const nqe_table *t, int32_t i, void *values,
                                             int64_t values_bytes, uint8_t *validity, int64_t validity_bytes,
                                             uint8_t *data, int64_t data_bytes) {
    if (!ctx || !t || i < 0 || i >= (int32_t)t->cols.size()) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    const DevColumn &c = t->cols[i];
    int64_t n = t->nrows;
    size_t vbytes = c.dtype == NQE_BOOL ? (size_t)((n + 7) / 8) : c.dtype == NQE_UTF8 ? (size_t)(n + 1) * 4 : (size_t)n * 8;
    if (values) {
        if ((size_t)values_bytes < vbytes) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "values buffer too small");
        NQE_TRY(d2h(ctx, values, c.values, vbytes));
    }
    if (validity) {
        size_t mb = (size_t)((n + 7) / 8);
        if ((size_t)validity_bytes < mb) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "validity buffer too small");
        if (c.validity) {
            if (mb) NQE_CUDA(ctx, cudaMemcpyAsync(validity, c.validity, mb, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            memset(validity, 0xff, mb);
        }
    }
    if (data && c.data) {
        if (data_bytes < c.data_bytes) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "data buffer too small");
        NQE_TRY(d2h(ctx, data, c.data, (size_t)c.data_bytes));
    }
    NQE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NQE_OK;
}

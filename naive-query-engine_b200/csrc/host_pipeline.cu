// host_pipeline.cu -- ProjectionPlan(SelectionPlan(ScanPlan)) over HOST Arrow columns with the
// result delivered into HOST buffers (what `PhysicalPlan::execute()` returns to the reference's
// caller, plan.rs:14-23), as a chunked three-stream pipeline:
//
//     H2D(chunk c+1)   |   filter/project kernel(chunk c)   |   D2H(result of chunk c-1)
//
// Only the columns the predicate / projections read are uploaded (the reference's scan hands all
// columns to SelectionPlan, selection.rs:65-101, but a fused plan never looks at the others), both
// PCIe directions are busy at once, and the filter kernel hides under the copies.  Row order is
// preserved: chunks are contiguous row ranges and their results are appended in order.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nqe_internal.cuh"

namespace {

// rows per pipeline chunk; default 8 Mi = 64 MiB per 8-byte column: ~1.2 ms of PCIe gen5 per column
// (knob NQE_HOST_CHUNK_ROWS, rounded down to a multiple of 2048)
int64_t chunk_rows() {
    static int64_t v = 0;
    if (!v) {
        const char *e = getenv("NQE_HOST_CHUNK_ROWS");
        v = e ? atoll(e) : (int64_t)8 << 20;
        v = std::max<int64_t>(1 << 16, v & ~(int64_t)2047);
    }
    return v;
}

void referenced_columns(const nqe_expr *e, std::vector<char> &used) {
    if (!e) return;
    for (int i = 0; i < e->n_nodes; i++)
        if (e->nodes[i].kind == NQE_NODE_COLUMN && e->nodes[i].column >= 0 && e->nodes[i].column < (int)used.size())
            used[e->nodes[i].column] = 1;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

} // namespace

extern "C" int32_t nqe_filter_project_host(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, const nqe_expr *predicate,
                                           const nqe_expr *projs, int32_t n_projs, nqe_column_desc *out_cols,
                                           int64_t *out_rows) {
    if (!ctx || !cols || n_cols <= 0 || !out_cols || !out_rows || n_projs <= 0 || !projs) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    const auto t_enter = std::chrono::steady_clock::now();
    const int64_t n = cols[0].length;
    const int64_t CHUNK_ROWS = chunk_rows();
    std::vector<char> used(n_cols, 0);
    referenced_columns(predicate, used);
    for (int i = 0; i < n_projs; i++) referenced_columns(&projs[i], used);
    // the pipelined path handles NULL-free 8-byte inputs and outputs from/to pinned memory;
    // everything else goes through upload -> operator -> download
    bool pipelined = n >= 2 * CHUNK_ROWS;
    for (int c = 0; c < n_cols && pipelined; c++) {
        if (cols[c].length != n) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "columns of different lengths");
        if (!used[c]) continue;
        const int dt = cols[c].dtype;
        if ((dt != NQE_INT64 && dt != NQE_UINT64 && dt != NQE_FLOAT64) || (cols[c].validity && cols[c].null_count) ||
            !is_pinned(cols[c].values))
            pipelined = false;
    }
    for (int i = 0; i < n_projs && pipelined; i++)
        if (!out_cols[i].values || out_cols[i].length < n || !is_pinned(out_cols[i].values)) pipelined = false;
    if (pipelined) {
        // type-check the expressions before the first chunk moves (reference error behaviour: raised before any work)
        // and keep plans whose outputs are Boolean or may be NULL (a NULL literal, a NULL-able predicate keeping NULL
        // rows) off the pipeline: it only drains NULL-free 8-byte result columns
        std::vector<nqe_column_desc> d0(cols, cols + n_cols);
        for (auto &c : d0) { c.length = 0; c.null_count = 0; c.validity = nullptr; } // a zero-row table of the same schema
        nqe_table *probe = nullptr;
        NQE_TRY(nqe_table_from_device(ctx, d0.data(), n_cols, &probe));
        std::vector<const nqe_expr *> list;
        if (predicate) list.push_back(predicate);
        for (int i = 0; i < n_projs; i++) list.push_back(&projs[i]);
        DevProgramSet ps;
        memset(&ps, 0, sizeof ps);
        std::vector<ExprInfo> info(list.size());
        const int32_t trc = nqe_compile_exprs(ctx, probe, list.data(), (int32_t)list.size(), &ps, info.data());
        nqe_table_free(probe);
        if (trc != NQE_OK) return trc;
        for (size_t i = 0; i < info.size(); i++)
            if (info[i].nullable || (info[i].result_dtype == NQE_BOOL && !(predicate && i == 0))) pipelined = false;
    }

    if (!pipelined) {
        nqe_table *in = nullptr, *out = nullptr;
        NQE_TRY(nqe_table_upload(ctx, cols, n_cols, &in));
        int32_t rc = nqe_filter_project(ctx, in, predicate, projs, n_projs, &out);
        nqe_table_free(in);
        if (rc != NQE_OK) return rc;
        *out_rows = nqe_table_num_rows(out);
        for (int i = 0; i < n_projs && rc == NQE_OK; i++) {
            nqe_column_desc d;
            nqe_table_column(out, i, &d);
            out_cols[i].dtype = d.dtype;
            out_cols[i].null_count = d.null_count;
            if (out_cols[i].length < *out_rows) rc = nqe_fail(ctx, NQE_ERR_INVALID_ARG, "output buffer %d too small", i);
            else
                rc = nqe_table_download_column(ctx, out, i, const_cast<void *>(out_cols[i].values),
                                               d.dtype == NQE_BOOL ? (*out_rows + 7) / 8 : *out_rows * 8,
                                               const_cast<uint8_t *>(out_cols[i].validity),
                                               out_cols[i].validity ? (*out_rows + 7) / 8 : 0, nullptr, 0);
        }
        if (rc == NQE_OK) rc = nqe_ctx_sync(ctx);
        nqe_table_free(out);
        return rc;
    }

    // ---- three-stream pipeline
    if (!ctx->s_h2d) {
        NQE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        NQE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    }
    // chunk boundaries: full chunks, then a tapered tail (quarter chunks) -- what follows the last upload (its kernel and
    // the download of its result) is not hidden under anything, so the last chunks are small
    std::vector<int64_t> start;
    {
        const int64_t small = std::max<int64_t>(1 << 16, (CHUNK_ROWS / 4) & ~(int64_t)2047);
        int64_t r = 0;
        while (n - r > CHUNK_ROWS + CHUNK_ROWS / 2) { start.push_back(r); r += CHUNK_ROWS; }
        while (r < n) { start.push_back(r); r += std::min(small, n - r); }
        start.push_back(n);
    }
    const int n_chunks = (int)start.size() - 1;
    constexpr int NBUF = 3; // chunk c's inputs live until its kernel is done; two more can be in flight
    std::vector<void *> dev_in((size_t)NBUF * n_cols, nullptr);
    int32_t rc = NQE_OK;
    for (int b = 0; b < NBUF && rc == NQE_OK; b++)
        for (int c = 0; c < n_cols && rc == NQE_OK; c++)
            if (used[c]) rc = nqe_dev_alloc(ctx, &dev_in[(size_t)b * n_cols + c], (size_t)CHUNK_ROWS * 8);
    std::vector<cudaEvent_t> up(NBUF), done(NBUF), drained(NBUF);
    for (int b = 0; b < NBUF; b++) {
        cudaEventCreateWithFlags(&up[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&drained[b], cudaEventDisableTiming);
    }
    std::vector<nqe_table *> pending_out(NBUF, nullptr);
    auto upload = [&](int c) {
        const int b = c % NBUF;
        const int64_t r0 = start[c], rows = start[c + 1] - r0;
        if (c >= NBUF) cudaStreamWaitEvent(ctx->s_h2d, done[b], 0); // the kernel of chunk c-NBUF has read this buffer
        for (int k = 0; k < n_cols; k++)
            if (used[k])
                cudaMemcpyAsync(dev_in[(size_t)b * n_cols + k], (const uint8_t *)cols[k].values + r0 * 8, (size_t)rows * 8,
                                cudaMemcpyHostToDevice, ctx->s_h2d);
        cudaEventRecord(up[b], ctx->s_h2d);
    };
    int64_t total = 0;
    double kernel_ms = 0.0;
    // NQE_HOST_PROF=1: where the wall time of one call goes (setup, H2D stream busy time, D2H tail)
    static const bool prof = getenv("NQE_HOST_PROF") && atoi(getenv("NQE_HOST_PROF"));
    cudaEvent_t pe[3] = {nullptr, nullptr, nullptr}; // first upload issued, last upload done, last download done
    const auto t_setup = std::chrono::steady_clock::now();
    if (prof) {
        for (auto &e : pe) cudaEventCreate(&e);
        cudaEventRecord(pe[0], ctx->s_h2d);
    }
    if (rc == NQE_OK) {
        upload(0);
        if (n_chunks > 1) upload(1);
    }
    for (int c = 0; c < n_chunks && rc == NQE_OK; c++) {
        const int b = c % NBUF;
        const int64_t rows = start[c + 1] - start[c];
        if (c + 2 < n_chunks) upload(c + 2);
        if (prof && c + 3 == n_chunks) cudaEventRecord(pe[1], ctx->s_h2d);
        // chunk table over the device buffers (unreferenced columns are never dereferenced)
        std::vector<nqe_column_desc> d(n_cols);
        for (int k = 0; k < n_cols; k++) {
            d[k] = cols[k];
            d[k].length = rows;
            d[k].null_count = 0;
            d[k].validity = nullptr;
            d[k].values = used[k] ? dev_in[(size_t)b * n_cols + k] : (const void *)ctx->d_scratch;
            if (!used[k]) d[k].dtype = NQE_INT64; // placeholder: a fused plan never reads this column
        }
        nqe_table *in = nullptr, *out = nullptr;
        rc = nqe_table_from_device(ctx, d.data(), n_cols, &in);
        if (rc != NQE_OK) break;
        cudaStreamWaitEvent(ctx->stream, up[b], 0);
        rc = nqe_filter_project(ctx, in, predicate, projs, n_projs, &out); // returns with the row count known
        kernel_ms += ctx->last_op_ms;
        cudaEventRecord(done[b], ctx->stream);
        nqe_table_free(in);
        if (rc != NQE_OK) break;
        const int64_t r = nqe_table_num_rows(out);
        // the previous result that used this slot has been copied out?
        if (pending_out[b]) {
            cudaEventSynchronize(drained[b]);
            nqe_table_free(pending_out[b]);
            pending_out[b] = nullptr;
        }
        cudaStreamWaitEvent(ctx->s_d2h, done[b], 0);
        for (int i = 0; i < n_projs; i++) {
            nqe_column_desc od;
            nqe_table_column(out, i, &od);
            out_cols[i].dtype = od.dtype;
            out_cols[i].null_count = 0;
            if (od.dtype == NQE_BOOL || od.null_count) { rc = nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "pipelined host path: unexpected output type"); break; }
            if (r) cudaMemcpyAsync((uint8_t *)const_cast<void *>(out_cols[i].values) + total * 8, od.values, (size_t)r * 8, cudaMemcpyDeviceToHost, ctx->s_d2h);
        }
        cudaEventRecord(drained[b], ctx->s_d2h);
        pending_out[b] = out;
        total += r;
    }
    if (prof) cudaEventRecord(pe[2], ctx->s_d2h);
    cudaStreamSynchronize(ctx->s_d2h);
    cudaStreamSynchronize(ctx->s_h2d);
    cudaStreamSynchronize(ctx->stream);
    if (prof) {
        float h2d = 0, all = 0;
        if (n_chunks >= 3) cudaEventElapsedTime(&h2d, pe[0], pe[1]);
        cudaEventElapsedTime(&all, pe[0], pe[2]);
        const double setup = std::chrono::duration<double, std::milli>(t_setup - t_enter).count();
        const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enter).count();
        fprintf(stderr, "[nqe host pipeline] chunks %d setup %.3f ms, uploads %.3f ms, first upload -> last download %.3f ms, call %.3f ms\n",
                n_chunks, setup, h2d, all, wall);
        for (auto &e : pe) cudaEventDestroy(e);
    }
    for (auto *t : pending_out)
        if (t) nqe_table_free(t);
    for (void *p : dev_in) nqe_dev_free(ctx, p);
    for (int b = 0; b < NBUF; b++) {
        cudaEventDestroy(up[b]);
        cudaEventDestroy(done[b]);
        cudaEventDestroy(drained[b]);
    }
    if (rc != NQE_OK) return rc;
    if (cudaGetLastError() != cudaSuccess) return nqe_fail(ctx, NQE_ERR_CUDA, "host pipeline failed");
    ctx->last_op_ms = kernel_ms;
    *out_rows = total;
    return NQE_OK;
}

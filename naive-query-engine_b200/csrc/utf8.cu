// utf8.cu -- Utf8 (StringArray) columns riding along through selection and join.
//
// The reference compacts every input column in SelectionPlan, strings included
// (selection.rs:82-97), and `take`s string payload columns in HashJoin (hash_join.rs:236-246).
// Here the numeric kernels additionally emit the surviving input row numbers (a hidden Int64
// row-id column travels through the same compaction / join), and the string columns are then
// gathered with those row ids:
//   lengths[i] = offsets[id+1] - offsets[id]   (0 for NULL)  ->  exclusive scan  ->  byte copy
#include <cstring>
#include <vector>

#include "nqe_internal.cuh"

int32_t nqe_pack_bytes(nqe_ctx *ctx, const uint8_t *bytes, int64_t n, uint32_t *words, unsigned long long *zeros);

namespace {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8; // 2048 elements per block

__global__ void iota_kernel(unsigned long long *out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (unsigned long long)i;
}

// lengths of the gathered strings; valid[i] = the source row exists and is non-NULL
__global__ void utf8_len_kernel(const int32_t *__restrict__ src_off, const uint32_t *__restrict__ src_valid,
                                const long long *__restrict__ idx, const uint32_t *__restrict__ idx_valid, int64_t n,
                                int32_t *__restrict__ len, uint8_t *__restrict__ valid) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool ok = !idx_valid || ((idx_valid[i >> 5] >> (i & 31)) & 1u);
    int32_t l = 0;
    if (ok) {
        const long long r = idx[i];
        if (src_valid && !((src_valid[r >> 5] >> (r & 31)) & 1u)) ok = false;
        else l = src_off[r + 1] - src_off[r];
    }
    len[i] = l;
    if (valid) valid[i] = (uint8_t)ok;
}

// two-level exclusive scan of int32 (sums kept in 64 bits to detect > 2 GiB of string data)
__global__ void scan_block_sums(const int32_t *__restrict__ in, int64_t n, unsigned long long *__restrict__ block_sums) {
    __shared__ unsigned long long s[SC_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SC_THREADS * SC_ITEMS;
    unsigned long long v = 0;
    for (int k = 0; k < SC_ITEMS; k++) {
        const int64_t i = base + (int64_t)k * SC_THREADS + threadIdx.x;
        if (i < n) v += (unsigned long long)in[i];
    }
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < SC_THREADS / 32; w++) t += s[w];
        block_sums[blockIdx.x] = t;
    }
}
__global__ void scan_sums_serial(unsigned long long *block_sums, int64_t n_blocks, unsigned long long *total) {
    // one thread: n_blocks = n / 2048 (48k for 1e8 rows)
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long run = 0;
    for (int64_t b = 0; b < n_blocks; b++) {
        const unsigned long long v = block_sums[b];
        block_sums[b] = run;
        run += v;
    }
    *total = run;
}
__global__ void scan_finish(const int32_t *__restrict__ in, int64_t n, const unsigned long long *__restrict__ block_sums,
                            int32_t *__restrict__ out_offsets) {
    // each block rescans its 2048 elements sequentially per thread-chunk: thread t owns SC_ITEMS consecutive items
    __shared__ unsigned long long s[SC_THREADS];
    const int64_t base = (int64_t)blockIdx.x * SC_THREADS * SC_ITEMS + (int64_t)threadIdx.x * SC_ITEMS;
    int32_t v[SC_ITEMS];
    unsigned long long sum = 0;
    for (int k = 0; k < SC_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        sum += (unsigned long long)v[k];
    }
    s[threadIdx.x] = sum;
    __syncthreads();
    // exclusive scan over the 256 thread sums (Hillis-Steele)
    for (int o = 1; o < SC_THREADS; o <<= 1) {
        unsigned long long t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    unsigned long long run = block_sums[blockIdx.x] + s[threadIdx.x] - sum;
    for (int k = 0; k < SC_ITEMS; k++) {
        if (base + k < n) out_offsets[base + k] = (int32_t)run;
        run += (unsigned long long)v[k];
    }
}

__global__ void utf8_copy_kernel(const int32_t *__restrict__ src_off, const uint8_t *__restrict__ src_data,
                                 const long long *__restrict__ idx, const int32_t *__restrict__ out_off, int64_t n,
                                 uint8_t *__restrict__ out_data) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t o0 = out_off[i], l = out_off[i + 1] - o0;
    if (l <= 0) return;
    const uint8_t *s = src_data + src_off[idx[i]];
    uint8_t *d = out_data + o0;
    for (int32_t b = 0; b < l; b++) d[b] = s[b];
}

// ---------------------------------------------------------------------------
// Utf8 keys (hash_join.rs:146-160,205-225; aggregate/mod.rs:170-216): strings are turned into
// 64-bit ids -- id = a row of the dictionary column that holds the same string -- and the integer
// join / group-by kernels run on the ids.  The dictionary is an open-addressing table of row
// numbers (claimed by CAS; the strings themselves are the immutable input, so an occupied slot can
// always be compared), hashed with FNV-1a over the bytes.
// ---------------------------------------------------------------------------
constexpr unsigned long long DICT_EMPTY = ~0ull;
constexpr unsigned long long DICT_NOMATCH = 1ull << 63; // | probe row: never equal to a dictionary row

__device__ __forceinline__ uint64_t utf8_hash(const uint8_t *p, int32_t len) {
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int32_t i = 0; i < len; i++) h = (h ^ p[i]) * 0x100000001b3ULL;
    h ^= h >> 32;
    h *= 0xd6e8feb86659fd93ULL;
    h ^= h >> 32;
    return h;
}
__device__ __forceinline__ bool utf8_eq(const uint8_t *a, int32_t la, const uint8_t *b, int32_t lb) {
    if (la != lb) return false;
    for (int32_t i = 0; i < la; i++)
        if (a[i] != b[i]) return false;
    return true;
}

struct DictParams {
    unsigned long long *slots; // row numbers of the dictionary column
    uint64_t cap;
    const int32_t *d_off;      // dictionary column
    const uint8_t *d_data;
};

__global__ void dict_clear_kernel(unsigned long long *slots, uint64_t cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) slots[i] = DICT_EMPTY;
}

// INSERT: rows of the dictionary column itself (ids[i] = the row that first claimed the string);
// !INSERT: rows of another column looked up in the finished dictionary (absent -> DICT_NOMATCH | i)
template <bool INSERT>
__global__ void dict_ids_kernel(DictParams dp, const int32_t *__restrict__ off, const uint8_t *__restrict__ data, int64_t n,
                                unsigned long long *__restrict__ ids, uint32_t *status) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *s = data + off[i];
    const int32_t len = off[i + 1] - off[i];
    uint64_t slot = __umul64hi(utf8_hash(s, len), dp.cap);
    for (uint64_t probe = 0; probe < dp.cap; probe++) {
        unsigned long long r = INSERT ? atomicCAS(dp.slots + slot, DICT_EMPTY, (unsigned long long)i)
                                      : *(volatile unsigned long long *)(dp.slots + slot);
        if (r == DICT_EMPTY) {
            ids[i] = INSERT ? (unsigned long long)i : (DICT_NOMATCH | (unsigned long long)i);
            return;
        }
        if (utf8_eq(dp.d_data + dp.d_off[r], dp.d_off[r + 1] - dp.d_off[r], s, len)) {
            ids[i] = r;
            return;
        }
        slot = slot + 1 == dp.cap ? 0 : slot + 1;
    }
    atomicOr(status, DEV_ERR_TABLE_FULL);
}

} // namespace

// hidden row-id column 0..n-1 (owned by the caller)
// ids of a Utf8 key column.  probe == nullptr: ids of `dict` itself (group keys, the build side of a join);
// otherwise also the ids of `probe` looked up in dict's dictionary (the probe side of a join).
// The id columns are UInt64 and borrow the validity bitmaps of their sources.
int32_t nqe_utf8_key_ids(nqe_ctx *ctx, const DevColumn &dict, const DevColumn *probe, DevColumn *dict_ids, DevColumn *probe_ids) {
    const int64_t n = dict.length;
    DictParams dp;
    dp.cap = (uint64_t)n * 2 + 16;
    dp.d_off = (const int32_t *)dict.values;
    dp.d_data = dict.data;
    void *slots = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &slots, dp.cap * 8));
    dp.slots = (unsigned long long *)slots;
    uint32_t *status = (uint32_t *)(ctx->d_scratch + 1);
    cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
    dict_clear_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(dp.slots, dp.cap);
    ctx->launches++;
    int32_t rc = NQE_OK;
    auto make_ids = [&](const DevColumn &src, DevColumn *out, bool insert) {
        *out = DevColumn();
        out->dtype = NQE_UINT64;
        out->length = src.length;
        if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, &out->values, (size_t)(src.length > 0 ? src.length : 1) * 8);
        if (rc != NQE_OK || src.length == 0) return;
        const unsigned grid = (unsigned)((src.length + 255) / 256);
        if (insert)
            dict_ids_kernel<true><<<grid, 256, 0, ctx->stream>>>(dp, (const int32_t *)src.values, src.data, src.length,
                                                               (unsigned long long *)out->values, status);
        else
            dict_ids_kernel<false><<<grid, 256, 0, ctx->stream>>>(dp, (const int32_t *)src.values, src.data, src.length,
                                                                (unsigned long long *)out->values, status);
        ctx->launches++;
    };
    make_ids(dict, dict_ids, true);
    if (probe) make_ids(*probe, probe_ids, false);
    if (rc == NQE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        rc = nqe_fail(ctx, NQE_ERR_CUDA, "utf8 dictionary failed: %s", cudaGetErrorString(cudaGetLastError()));
    nqe_dev_free(ctx, slots);
    if (rc != NQE_OK) {
        nqe_dev_free(ctx, dict_ids->values);
        dict_ids->values = nullptr;
        if (probe) { nqe_dev_free(ctx, probe_ids->values); probe_ids->values = nullptr; }
        return rc;
    }
    // validity travels with the ids (group-by drops NULL keys; the join ignores it, like the reference)
    dict_ids->validity = dict.validity;
    dict_ids->null_count = dict.null_count;
    if (probe) {
        probe_ids->validity = probe->validity;
        probe_ids->null_count = probe->null_count;
    }
    return NQE_OK;
}

int32_t nqe_make_rowid_column(nqe_ctx *ctx, int64_t n, DevColumn *c) {
    NQE_TRY(nqe_column_alloc(ctx, NQE_INT64, n, false, c));
    if (n > 0) {
        iota_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>((unsigned long long *)c->values, n);
        ctx->launches++;
        NQE_CUDA(ctx, cudaGetLastError());
    }
    return NQE_OK;
}

// arrow `take` for a Utf8 column: out[i] = src[idx[i]] (NULL when idx[i] is NULL or src[idx[i]] is NULL)
int32_t nqe_take_utf8(nqe_ctx *ctx, const DevColumn &src, const DevColumn &idx, int64_t n, DevColumn *out) {
    memset((void *)out, 0, sizeof *out);
    *out = DevColumn();
    out->dtype = NQE_UTF8;
    out->length = n;
    out->owned = true;
    NQE_TRY(nqe_dev_alloc(ctx, &out->values, (size_t)(n + 1) * 4 + 64));
    const bool nullable = src.validity || idx.validity;
    void *len = nullptr, *vbytes = nullptr, *sums = nullptr;
    int32_t rc = nqe_dev_alloc(ctx, &len, (size_t)(n + 1) * 4 + 64);
    if (rc == NQE_OK && nullable) rc = nqe_dev_alloc(ctx, &vbytes, (size_t)n + 64);
    const int64_t n_blocks = (n + SC_THREADS * SC_ITEMS - 1) / (SC_THREADS * SC_ITEMS) + 1;
    if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, &sums, (size_t)n_blocks * 8);
    if (rc == NQE_OK) {
        cudaMemsetAsync(ctx->d_scratch + 40, 0, 2 * sizeof(uint64_t), ctx->stream);
        if (n > 0) {
            utf8_len_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
                (const int32_t *)src.values, (const uint32_t *)src.validity, (const long long *)idx.values,
                (const uint32_t *)idx.validity, n, (int32_t *)len, (uint8_t *)vbytes);
            scan_block_sums<<<(unsigned)n_blocks, SC_THREADS, 0, ctx->stream>>>((const int32_t *)len, n, (unsigned long long *)sums);
            scan_sums_serial<<<1, 32, 0, ctx->stream>>>((unsigned long long *)sums, n_blocks, (unsigned long long *)(ctx->d_scratch + 40));
            scan_finish<<<(unsigned)n_blocks, SC_THREADS, 0, ctx->stream>>>((const int32_t *)len, n, (const unsigned long long *)sums,
                                                                           (int32_t *)out->values);
            // offsets[n] = total bytes (low 32 bits of the 64-bit total; > 2 GiB is rejected below)
            cudaMemcpyAsync((int32_t *)out->values + n, ctx->d_scratch + 40, 4, cudaMemcpyDeviceToDevice, ctx->stream);
            ctx->launches += 4;
        } else {
            cudaMemsetAsync(out->values, 0, 4, ctx->stream);
        }
        cudaMemcpyAsync(ctx->h_scratch + 40, ctx->d_scratch + 40, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = nqe_fail(ctx, NQE_ERR_CUDA, "utf8 take failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == NQE_OK) {
        const uint64_t total = ctx->h_scratch[40];
        if (total > 0x7fffffffull) rc = nqe_fail(ctx, NQE_ERR_PANIC, "Utf8 column exceeds 2 GiB of string data (i32 offsets overflow)");
        out->data_bytes = (int64_t)total;
    }
    if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, (void **)&out->data, (size_t)out->data_bytes + 64);
    if (rc == NQE_OK && n > 0) {
        utf8_copy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            (const int32_t *)src.values, src.data, (const long long *)idx.values, (const int32_t *)out->values, n, out->data);
        ctx->launches++;
        if (nullable) {
            rc = nqe_dev_alloc(ctx, (void **)&out->validity, nqe_bitmap_bytes(n));
            if (rc == NQE_OK) {
                cudaMemsetAsync(ctx->d_scratch + 41, 0, sizeof(uint64_t), ctx->stream);
                rc = nqe_pack_bytes(ctx, (const uint8_t *)vbytes, n, (uint32_t *)out->validity, (unsigned long long *)(ctx->d_scratch + 41));
                cudaMemcpyAsync(ctx->h_scratch + 41, ctx->d_scratch + 41, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
                if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "utf8 take failed");
                out->null_count = (int64_t)ctx->h_scratch[41];
                if (rc == NQE_OK && out->null_count == 0) {
                    nqe_dev_free(ctx, out->validity);
                    out->validity = nullptr;
                }
            }
        }
    }
    nqe_dev_free(ctx, len);
    nqe_dev_free(ctx, vbytes);
    nqe_dev_free(ctx, sums);
    if (rc != NQE_OK) nqe_column_release(ctx, out);
    return rc;
}

// ---------------------------------------------------------------------------
// operator wrappers: run the numeric operator with a hidden row-id column, then gather
// the Utf8 columns with the surviving row ids
// ---------------------------------------------------------------------------
static DevColumn borrow(const DevColumn &c) {
    DevColumn b = c;
    b.owned = false;
    return b;
}

// projs[i] with utf8_src[i] >= 0 is a bare reference to Utf8 column utf8_src[i] of `in`
int32_t nqe_filter_project_strings(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                                   int32_t n_projs, const int *utf8_src, nqe_table **out) {
    const int ncols = (int)in->cols.size();
    nqe_table aug;
    aug.ctx = ctx;
    aug.nrows = in->nrows;
    for (auto &c : in->cols) aug.cols.push_back(borrow(c));
    DevColumn rowid;
    NQE_TRY(nqe_make_rowid_column(ctx, in->nrows, &rowid));
    aug.cols.push_back(borrow(rowid));
    std::vector<nqe_expr> p2;
    for (int i = 0; i < n_projs; i++)
        if (utf8_src[i] < 0) p2.push_back(projs[i]);
    nqe_expr_node rid{NQE_NODE_COLUMN, 0, ncols, 0, 0, 0, {0}};
    p2.push_back(nqe_expr{&rid, 1, 0});
    nqe_table *tmp = nullptr;
    int32_t rc = (int)p2.size() > 16 ? nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "more than 15 output columns next to Utf8 columns")
                                     : nqe_filter_project(ctx, &aug, predicate, p2.data(), (int32_t)p2.size(), &tmp);
    nqe_table *res = nullptr;
    if (rc == NQE_OK) {
        nqe_table_new(ctx, tmp->nrows, &res);
        res->cols.resize(n_projs);
        const DevColumn &ids = tmp->cols.back();
        int k = 0;
        for (int i = 0; i < n_projs && rc == NQE_OK; i++) {
            if (utf8_src[i] < 0) {
                res->cols[i] = tmp->cols[k];
                tmp->cols[k].owned = false; // moved
                k++;
            } else {
                rc = nqe_take_utf8(ctx, in->cols[utf8_src[i]], ids, tmp->nrows, &res->cols[i]);
            }
        }
    }
    if (tmp) nqe_table_free(tmp);
    nqe_column_release(ctx, &rowid);
    if (rc != NQE_OK) {
        if (res) nqe_table_free(res);
        return rc;
    }
    *out = res;
    return NQE_OK;
}

// ---- Utf8 comparisons inside expressions (binary.rs:127-132: eq_dyn / neq_dyn / lt_dyn / lt_eq_dyn / gt_dyn / gt_eq_dyn work
// on any arrow-comparable dtype, Utf8 included: bytewise lexicographic order, NULL if either side is NULL) --------------------
// A Utf8 value only ever comes from a leaf (a column or a literal: the reference has no string-valued function that is
// not todo!()), so every comparison of two Utf8 leaves is evaluated by one kernel into a Boolean column appended to a view
// of the input, and the expression is rewritten to reference that column; the numeric machinery then runs unchanged.
struct Utf8Operand {
    const int32_t *off;    // column: offsets
    const uint8_t *data;   // column: bytes / literal: bytes (device)
    const uint32_t *valid; // column: validity words or nullptr
    int32_t lit_len;       // literal: byte length
    int32_t is_lit, lit_null, pad;
};

__global__ void utf8_compare_kernel(const Utf8Operand a, const Utf8Operand b, int op, int64_t n, uint8_t *val, uint8_t *ok) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto fetch = [&](const Utf8Operand &x, const uint8_t **p, int32_t *len) -> bool {
        if (x.is_lit) {
            *p = x.data;
            *len = x.lit_len;
            return !x.lit_null;
        }
        const int32_t o0 = x.off[i];
        *p = x.data + o0;
        *len = x.off[i + 1] - o0;
        return !x.valid || ((x.valid[i >> 5] >> (i & 31)) & 1u);
    };
    const uint8_t *pa, *pb;
    int32_t la, lb;
    const bool va = fetch(a, &pa, &la), vb = fetch(b, &pb, &lb);
    int cmp = 0;
    const int32_t m = la < lb ? la : lb;
    for (int32_t k = 0; k < m && cmp == 0; k++) cmp = (int)pa[k] - (int)pb[k];
    if (cmp == 0) cmp = la < lb ? -1 : la > lb ? 1 : 0;
    bool r;
    switch (op) {
    case NQE_OP_EQ: r = cmp == 0; break;
    case NQE_OP_NOT_EQ: r = cmp != 0; break;
    case NQE_OP_LT: r = cmp < 0; break;
    case NQE_OP_LT_EQ: r = cmp <= 0; break;
    case NQE_OP_GT: r = cmp > 0; break;
    default: r = cmp >= 0; break;
    }
    const bool v = va && vb;
    val[i] = v && r;
    ok[i] = v;
}

static bool utf8_leaf(const nqe_table *in, const nqe_expr_node &nd) {
    if (nd.kind == NQE_NODE_COLUMN) return nd.column >= 0 && nd.column < (int)in->cols.size() && in->cols[nd.column].dtype == NQE_UTF8;
    return nd.kind == NQE_NODE_LITERAL && nd.dtype == NQE_UTF8;
}

// Rewrites the comparisons of Utf8 leaves in (predicate, projs) and runs the operator on the augmented view.
// *handled == false: nothing to rewrite, the caller carries on.
int32_t nqe_filter_project_utf8_compares(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                                         int32_t n_projs, nqe_table **out, bool *handled) {
    *handled = false;
    std::vector<const nqe_expr *> list;
    if (predicate) list.push_back(predicate);
    for (int i = 0; i < n_projs; i++) list.push_back(&projs[i]);
    bool any = false;
    for (const nqe_expr *e : list)
        for (int i = 2; e->nodes && i < e->n_nodes; i++)
            if (e->nodes[i].kind == NQE_NODE_BINARY && e->nodes[i].op >= NQE_OP_EQ && e->nodes[i].op <= NQE_OP_GT_EQ &&
                utf8_leaf(in, e->nodes[i - 1]) && utf8_leaf(in, e->nodes[i - 2]))
                any = true;
    if (!any) return NQE_OK;
    *handled = true;
    const int64_t n = in->nrows;
    nqe_table aug;
    aug.ctx = ctx;
    aug.nrows = n;
    for (auto &c : in->cols) aug.cols.push_back(borrow(c));
    std::vector<DevColumn> made;   // Boolean result columns (owned here)
    std::vector<void *> scratch;   // literal bytes, byte-per-row buffers
    std::vector<std::vector<nqe_expr_node>> nodes(list.size());
    int32_t rc = NQE_OK;
    auto operand = [&](const nqe_expr_node &nd, Utf8Operand *o) -> int32_t {
        memset(o, 0, sizeof *o);
        if (nd.kind == NQE_NODE_COLUMN) {
            const DevColumn &c = in->cols[nd.column];
            o->off = (const int32_t *)c.values;
            o->data = c.data;
            o->valid = (const uint32_t *)c.validity;
            return NQE_OK;
        }
        o->is_lit = 1;
        o->lit_null = nd.is_null != 0;
        o->lit_len = nd.is_null ? 0 : nd.reserved;
        if (o->lit_len < 0) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "Utf8 literal with a negative length");
        void *d = nullptr;
        NQE_TRY(nqe_dev_alloc(ctx, &d, (size_t)o->lit_len + 16));
        scratch.push_back(d);
        if (o->lit_len) NQE_CUDA(ctx, cudaMemcpyAsync(d, (const void *)(uintptr_t)nd.value.u64, (size_t)o->lit_len, cudaMemcpyHostToDevice, ctx->stream));
        o->data = (const uint8_t *)d;
        return NQE_OK;
    };
    for (size_t x = 0; x < list.size() && rc == NQE_OK; x++) {
        const nqe_expr *e = list[x];
        std::vector<nqe_expr_node> &o = nodes[x];
        for (int i = 0; i < e->n_nodes && rc == NQE_OK; i++) {
            const nqe_expr_node &nd = e->nodes[i];
            const size_t m = o.size();
            if (nd.kind == NQE_NODE_BINARY && nd.op >= NQE_OP_EQ && nd.op <= NQE_OP_GT_EQ && m >= 2 && i >= 2 &&
                utf8_leaf(in, e->nodes[i - 1]) && utf8_leaf(in, e->nodes[i - 2]) && utf8_leaf(in, o[m - 1]) && utf8_leaf(in, o[m - 2])) {
                Utf8Operand a, b;
                rc = operand(o[m - 2], &a);
                if (rc == NQE_OK) rc = operand(o[m - 1], &b);
                const bool nullable = (!a.is_lit && a.valid) || (!b.is_lit && b.valid) || a.lit_null || b.lit_null;
                DevColumn col;
                uint8_t *val = nullptr, *ok = nullptr;
                if (rc == NQE_OK) rc = nqe_column_alloc(ctx, NQE_BOOL, n, nullable, &col);
                if (rc == NQE_OK) made.push_back(col);
                if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, (void **)&val, (size_t)n + 64);
                if (rc == NQE_OK) scratch.push_back(val);
                if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, (void **)&ok, (size_t)n + 64);
                if (rc == NQE_OK) scratch.push_back(ok);
                if (rc == NQE_OK && n > 0) {
                    utf8_compare_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a, b, nd.op, n, val, ok);
                    ctx->launches++;
                    cudaMemsetAsync(ctx->d_scratch + 8, 0, sizeof(uint64_t), ctx->stream);
                    rc = nqe_pack_bytes(ctx, val, n, (uint32_t *)made.back().values, nullptr);
                    if (rc == NQE_OK && nullable) rc = nqe_pack_bytes(ctx, ok, n, (uint32_t *)made.back().validity, (unsigned long long *)(ctx->d_scratch + 8));
                    if (rc == NQE_OK && nullable) {
                        cudaMemcpyAsync(ctx->h_scratch + 8, ctx->d_scratch + 8, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
                        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "Utf8 comparison failed");
                        made.back().null_count = (int64_t)ctx->h_scratch[8];
                    }
                }
                if (rc == NQE_OK) {
                    DevColumn view = borrow(made.back());
                    if (view.null_count == 0) view.validity = nullptr; // arrow: no nulls => no bitmap
                    aug.cols.push_back(view);
                    o.resize(m - 2);
                    o.push_back(nqe_expr_node{NQE_NODE_COLUMN, 0, (int32_t)aug.cols.size() - 1, 0, 0, 0, {0}});
                }
            } else {
                o.push_back(nd);
            }
        }
    }
    if (rc == NQE_OK) {
        std::vector<nqe_expr> ex(list.size());
        for (size_t x = 0; x < list.size(); x++) ex[x] = nqe_expr{nodes[x].data(), (int32_t)nodes[x].size(), 0};
        const nqe_expr *pred2 = predicate ? &ex[0] : nullptr;
        rc = nqe_filter_project(ctx, &aug, pred2, ex.data() + (predicate ? 1 : 0), n_projs, out);
    }
    for (auto &c : made) nqe_column_release(ctx, &c);
    for (void *p : scratch) nqe_dev_free(ctx, p);
    return rc;
}

int32_t nqe_hash_join_strings(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right, int32_t left_key,
                              int32_t right_key, nqe_table **out) {
    const nqe_table *side[2] = {left, right};
    const int key[2] = {left_key, right_key};
    nqe_table aug[2];
    DevColumn rowid[2], keyid[2];
    std::vector<int> map[2]; // original column -> column in aug (numeric) or -1 (Utf8)
    int akey[2] = {0, 0};
    int32_t rc = NQE_OK;
    // Utf8 keys (hash_join.rs:146-160): both sides get an id column from the build side's dictionary
    const bool utf8_key = left->cols[left_key].dtype == NQE_UTF8;
    if (utf8_key) rc = nqe_utf8_key_ids(ctx, left->cols[left_key], &right->cols[right_key], &keyid[0], &keyid[1]);
    for (int s = 0; s < 2 && rc == NQE_OK; s++) {
        aug[s].ctx = ctx;
        aug[s].nrows = side[s]->nrows;
        for (size_t c = 0; c < side[s]->cols.size(); c++) {
            if (side[s]->cols[c].dtype == NQE_UTF8) map[s].push_back(-1);
            else {
                map[s].push_back((int)aug[s].cols.size());
                aug[s].cols.push_back(borrow(side[s]->cols[c]));
            }
        }
        if (utf8_key) {
            akey[s] = (int)aug[s].cols.size();
            DevColumn k = borrow(keyid[s]);
            k.validity = nullptr; // key validity is ignored (hash_join.rs:67,86)
            k.null_count = 0;
            aug[s].cols.push_back(k);
        } else {
            akey[s] = map[s][key[s]];
        }
        rc = nqe_make_rowid_column(ctx, side[s]->nrows, &rowid[s]);
        aug[s].cols.push_back(borrow(rowid[s]));
    }
    nqe_table *tmp = nullptr;
    if (rc == NQE_OK) rc = nqe_hash_join(ctx, &aug[0], &aug[1], akey[0], akey[1], &tmp);
    nqe_table *res = nullptr;
    if (rc == NQE_OK) {
        nqe_table_new(ctx, tmp->nrows, &res);
        const int nl_aug = (int)aug[0].cols.size();
        const DevColumn *ids[2] = {&tmp->cols[nl_aug - 1], &tmp->cols.back()};
        for (int s = 0; s < 2 && rc == NQE_OK; s++) {
            for (size_t c = 0; c < side[s]->cols.size() && rc == NQE_OK; c++) {
                DevColumn col;
                if (map[s][c] >= 0) {
                    const int t = (s ? nl_aug : 0) + map[s][c];
                    col = tmp->cols[t];
                    tmp->cols[t].owned = false; // moved
                } else {
                    rc = nqe_take_utf8(ctx, side[s]->cols[c], *ids[s], tmp->nrows, &col);
                }
                res->cols.push_back(col);
            }
        }
    }
    if (tmp) nqe_table_free(tmp);
    nqe_column_release(ctx, &rowid[0]);
    nqe_column_release(ctx, &rowid[1]);
    for (int s = 0; s < 2; s++)
        if (keyid[s].values) nqe_dev_free(ctx, keyid[s].values);
    if (rc != NQE_OK) {
        if (res) nqe_table_free(res);
        return rc;
    }
    *out = res;
    return NQE_OK;
}

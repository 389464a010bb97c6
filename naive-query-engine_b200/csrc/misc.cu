// misc.cu -- row slicing (PhysicalLimitPlan / PhysicalOffsetPlan), key-radix
// partitioning for the multi-GPU shuffle, and on-device synthetic columns.
#include <cstring>

#include "hash_common.cuh"
#include "nqe_internal.cuh"

namespace {

// dst bit i = src bit (i + shift), i < n
__global__ void bitmap_slice_kernel(const uint32_t *__restrict__ src, int64_t shift, int64_t n, uint32_t *__restrict__ dst,
                                    unsigned long long *zeros) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nwords = (n + 31) / 32;
    unsigned int z = 0;
    if (w < nwords) {
        const int64_t b = w * 32 + shift;
        const int64_t rem = n - w * 32;
        const int64_t last = b + (rem < 32 ? rem : 32) - 1;
        const uint32_t lo = src[b >> 5], hi = (last >> 5) != (b >> 5) ? src[(b >> 5) + 1] : 0u;
        uint32_t v = (b & 31) ? (lo >> (b & 31)) | (hi << (32 - (b & 31))) : lo;
        if (rem < 32) v &= (1u << rem) - 1u;
        dst[w] = v;
        z = (unsigned)(rem < 32 ? rem : 32) - __popc(v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (zeros && (threadIdx.x & 31) == 0 && z) atomicAdd(zeros, (unsigned long long)z);
}

__global__ void utf8_rebase_kernel(const int32_t *__restrict__ src, int64_t n_plus_1, int32_t base, int32_t *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_plus_1) dst[i] = src[i] - base;
}

// concat_batches: OR the first n bits of `src` (nullptr = all ones) into `dst` starting at bit `at`.  `dst` was
// zeroed; words shared by two pieces are why this is an atomicOr.
__global__ void bitmap_place_kernel(const uint32_t *__restrict__ src, int64_t n, uint32_t *__restrict__ dst, int64_t at) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; // source word
    if (w * 32 >= n) return;
    const int64_t rem = n - w * 32;
    uint32_t v = src ? src[w] : 0xffffffffu;
    if (rem < 32) v &= (1u << rem) - 1u;
    const int64_t b = at + w * 32;
    const int sh = (int)(b & 31);
    if (v << sh) atomicOr(dst + (b >> 5), v << sh);
    if (sh && (v >> (32 - sh))) atomicOr(dst + (b >> 5) + 1, v >> (32 - sh));
}
__global__ void utf8_offsets_place_kernel(const int32_t *__restrict__ src, int64_t n, int32_t add, int32_t *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}

constexpr int MAX_PARTS = 64;

struct PartParams {
    const unsigned long long *keys;
    int64_t n;
    int32_t n_parts;
    int32_t n_cols;
    const unsigned long long *in[16];
    unsigned long long *out[16];
    unsigned long long *counts;  // [n_parts]
    unsigned long long *cursors; // [n_parts] running write positions (start at exclusive prefix)
};

__device__ __forceinline__ int part_of(unsigned long long key, int n_parts) { return (int)(nqe_mix64(key) % (unsigned)n_parts); }

__global__ void part_hist_kernel(PartParams pp) {
    __shared__ unsigned int h[MAX_PARTS];
    if (threadIdx.x < MAX_PARTS) h[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pp.n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&h[part_of(pp.keys[i], pp.n_parts)], 1u);
    __syncthreads();
    if (threadIdx.x < pp.n_parts && h[threadIdx.x]) atomicAdd(pp.counts + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}

// Block-local counting sort of a tile by destination, then each destination's
// run is copied out with consecutive threads writing consecutive rows.
constexpr int PT_THREADS = 256;
constexpr int PT_K = 8;
__global__ void __launch_bounds__(PT_THREADS) part_scatter_kernel(PartParams pp) {
    constexpr int TILE = PT_THREADS * PT_K;
    __shared__ unsigned int s_cnt[MAX_PARTS], s_start[MAX_PARTS];
    __shared__ unsigned long long s_base[MAX_PARTS];
    __shared__ unsigned int s_src[TILE];   // tile-local source row, ordered by destination
    const int tid = threadIdx.x;
    const int64_t num_tiles = (pp.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        if (tid < MAX_PARTS) s_cnt[tid] = 0;
        __syncthreads();
        int dest[PT_K];
        unsigned int rank[PT_K];
#pragma unroll
        for (int j = 0; j < PT_K; j++) {
            const int64_t e = base + j * PT_THREADS + tid;
            dest[j] = e < pp.n ? part_of(pp.keys[e], pp.n_parts) : -1;
            if (dest[j] >= 0) rank[j] = atomicAdd(&s_cnt[dest[j]], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int run = 0;
            for (int p = 0; p < pp.n_parts; p++) {
                s_start[p] = run;
                run += s_cnt[p];
            }
        }
        if (tid < pp.n_parts && s_cnt[tid]) s_base[tid] = atomicAdd(pp.cursors + tid, (unsigned long long)s_cnt[tid]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PT_K; j++)
            if (dest[j] >= 0) s_src[s_start[dest[j]] + rank[j]] = (unsigned)(j * PT_THREADS + tid) | ((unsigned)dest[j] << 24);
        __syncthreads();
        const int64_t n_here = pp.n - base < TILE ? pp.n - base : TILE;
        for (int i = tid; i < n_here; i += PT_THREADS) {
            const unsigned int v = s_src[i];
            const int p = (int)(v >> 24);
            const unsigned int local = v & 0xffffffu;
            const unsigned long long dst = s_base[p] + (unsigned)(i - s_start[p]);
            for (int c = 0; c < pp.n_cols; c++) pp.out[c][dst] = pp.in[c][base + local];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Radix partition fused with the exchange: the scatter kernel stores every row straight into the
// receive buffer of its destination GPU (peer-mapped memory, NVLink 5 / NVSwitch), so the shuffle
// is one pass over the local rows -- no partitioned staging copy and no separate all-to-all.
// Destination runs are formed block-locally (counting sort of the tile in shared memory), so the
// remote stores of one warp are consecutive rows of one destination.
// ---------------------------------------------------------------------------
struct ShuffleParams {
    const unsigned long long *keys;
    int64_t n;
    int32_t n_parts;
    int32_t n_cols;
    const unsigned long long *in[8];
    unsigned long long *dst[8 * 8];      // [col][part]: column receive buffer on rank `part`, as mapped here
    unsigned long long *cursors;         // [n_parts] next row to write in the destination (starts at this rank's offset)
};

__global__ void __launch_bounds__(PT_THREADS) shuffle_scatter_kernel(const __grid_constant__ ShuffleParams sp) {
    constexpr int TILE = PT_THREADS * PT_K;
    __shared__ unsigned int s_cnt[MAX_PARTS], s_start[MAX_PARTS];
    __shared__ unsigned long long s_base[MAX_PARTS];
    __shared__ unsigned int s_src[TILE]; // tile-local source row, ordered by destination
    const int tid = threadIdx.x;
    const int64_t num_tiles = (sp.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE;
        if (tid < MAX_PARTS) s_cnt[tid] = 0;
        __syncthreads();
        int dest[PT_K];
        unsigned int rank[PT_K];
#pragma unroll
        for (int j = 0; j < PT_K; j++) {
            const int64_t e = base + j * PT_THREADS + tid;
            dest[j] = e < sp.n ? part_of(sp.keys[e], sp.n_parts) : -1;
            if (dest[j] >= 0) rank[j] = atomicAdd(&s_cnt[dest[j]], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int run = 0;
            for (int p = 0; p < sp.n_parts; p++) {
                s_start[p] = run;
                run += s_cnt[p];
            }
        }
        if (tid < sp.n_parts && s_cnt[tid]) s_base[tid] = atomicAdd(sp.cursors + tid, (unsigned long long)s_cnt[tid]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PT_K; j++)
            if (dest[j] >= 0) s_src[s_start[dest[j]] + rank[j]] = (unsigned)(j * PT_THREADS + tid) | ((unsigned)dest[j] << 24);
        __syncthreads();
        const int64_t n_here = sp.n - base < TILE ? sp.n - base : TILE;
        for (int i = tid; i < n_here; i += PT_THREADS) {
            const unsigned int v = s_src[i];
            const int p = (int)(v >> 24);
            const unsigned int local = v & 0xffffffu;
            const unsigned long long d = s_base[p] + (unsigned)(i - s_start[p]);
            for (int c = 0; c < sp.n_cols; c++) sp.dst[c * 8 + p][d] = sp.in[c][base + local]; // remote (or local) store
        }
        __syncthreads();
    }
}

__global__ void synth_kernel(int kind, uint64_t seed, int64_t start, int64_t n, uint64_t a, uint64_t b, double scale,
                             unsigned long long *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)(start + i);
        unsigned long long v;
        if (kind == 0) v = nqe_splitmix(seed + g) % a;
        else if (kind == 1) v = (unsigned long long)__double_as_longlong(scale * ((double)(nqe_splitmix(seed + g) >> 11) * 0x1.0p-53));
        else v = (unsigned long long)(((unsigned __int128)g * a) % b);
        out[i] = v;
    }
}

} // namespace

extern "C" int32_t nqe_synth_column(nqe_ctx *ctx, int32_t kind, uint64_t seed, int64_t start, int64_t n,
                                    uint64_t mod_or_mul, uint64_t mod2, double scale, void *device_out) {
    if (!ctx || !device_out || n < 0 || kind < 0 || kind > 2) return NQE_ERR_INVALID_ARG;
    if ((kind == 0 && mod_or_mul == 0) || (kind == 2 && mod2 == 0)) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "zero modulus");
    cudaSetDevice(ctx->device);
    if (n == 0) return NQE_OK;
    int grid = ctx->sm_count * 8;
    synth_kernel<<<grid, 256, 0, ctx->stream>>>(kind, seed, start, n, mod_or_mul, mod2, scale, (unsigned long long *)device_out);
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}

extern "C" int32_t nqe_table_slice(nqe_ctx *ctx, const nqe_table *t, int64_t offset, int64_t len, nqe_table **out) {
    if (!ctx || !t || !out) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    // RecordBatch::slice semantics used by offset.rs:30-51 / limit.rs:32-49 after clamping
    if (offset < 0) offset = 0;
    if (offset > t->nrows) offset = t->nrows;
    if (len < 0) len = 0;
    if (offset + len > t->nrows) len = t->nrows - offset;
    nqe_table *r;
    nqe_table_new(ctx, len, &r);
    r->cols.resize(t->cols.size());
    int32_t rc = NQE_OK;
    cudaMemsetAsync(ctx->d_scratch + 8, 0, 32 * sizeof(uint64_t), ctx->stream);
    bool any_valid = false;
    for (size_t i = 0; i < t->cols.size() && rc == NQE_OK; i++) {
        const DevColumn &s = t->cols[i];
        DevColumn &d = r->cols[i];
        rc = nqe_column_alloc(ctx, s.dtype, len, s.validity != nullptr, &d);
        if (rc != NQE_OK) break;
        const unsigned grid = (unsigned)(((len + 31) / 32 + 255) / 256);
        if (s.dtype == NQE_BOOL) {
            if (len) bitmap_slice_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t *)s.values, offset, len, (uint32_t *)d.values, nullptr);
        } else if (s.dtype == NQE_UTF8) {
            int32_t ends[2] = {0, 0};
            cudaMemcpyAsync(&ends[0], (const int32_t *)s.values + offset, 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaMemcpyAsync(&ends[1], (const int32_t *)s.values + offset + len, 4, cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            d.data_bytes = ends[1] - ends[0];
            rc = nqe_dev_alloc(ctx, (void **)&d.data, (size_t)d.data_bytes + 64);
            if (rc != NQE_OK) break;
            utf8_rebase_kernel<<<(unsigned)((len + 1 + 255) / 256), 256, 0, ctx->stream>>>((const int32_t *)s.values + offset, len + 1, ends[0], (int32_t *)d.values);
            if (d.data_bytes) cudaMemcpyAsync(d.data, s.data + ends[0], (size_t)d.data_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        } else if (len) {
            cudaMemcpyAsync(d.values, (const uint64_t *)s.values + offset, (size_t)len * 8, cudaMemcpyDeviceToDevice, ctx->stream);
        }
        if (s.validity && len) {
            bitmap_slice_kernel<<<grid, 256, 0, ctx->stream>>>((const uint32_t *)s.validity, offset, len, (uint32_t *)d.validity,
                                                               (unsigned long long *)(ctx->d_scratch + 8 + (i & 31)));
            any_valid = true;
        }
        ctx->launches++;
    }
    if (rc == NQE_OK && cudaGetLastError() != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "slice failed");
    if (rc == NQE_OK && any_valid) {
        cudaMemcpyAsync(ctx->h_scratch + 8, ctx->d_scratch + 8, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "slice failed");
    }
    if (rc != NQE_OK) {
        nqe_table_free(r);
        return rc;
    }
    for (size_t i = 0; i < r->cols.size(); i++) {
        DevColumn &d = r->cols[i];
        if (d.validity) {
            d.null_count = len ? (int64_t)ctx->h_scratch[8 + (i & 31)] : 0;
            if (d.null_count == 0) {
                nqe_dev_free(ctx, d.validity);
                d.validity = nullptr;
            }
        }
    }
    *out = r;
    return NQE_OK;
}

static int32_t check_partition_input(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts, int max_parts,
                                     size_t max_cols) {
    if (n_parts < 1 || n_parts > max_parts) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "n_parts must be in [1,%d]", max_parts);
    if (key_column < 0 || key_column >= (int)in->cols.size()) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "bad key column");
    if (in->cols.size() > max_cols) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "more than %d columns", (int)max_cols);
    const int kd = in->cols[key_column].dtype;
    if (kd != NQE_INT64 && kd != NQE_UINT64) return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "partition key must be Int64/UInt64");
    for (auto &c : in->cols)
        if (c.dtype == NQE_BOOL || c.dtype == NQE_UTF8 || c.validity)
            return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "radix partition supports NULL-free 8-byte columns only");
    return NQE_OK;
}

extern "C" int32_t nqe_partition_counts(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts, int64_t *counts) {
    if (!ctx || !in || !counts) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    NQE_TRY(check_partition_input(ctx, in, key_column, n_parts, MAX_PARTS, 16));
    PartParams pp;
    memset(&pp, 0, sizeof pp);
    pp.keys = (const unsigned long long *)in->cols[key_column].values;
    pp.n = in->nrows;
    pp.n_parts = n_parts;
    pp.counts = (unsigned long long *)ctx->d_scratch;
    cudaMemsetAsync(pp.counts, 0, 64 * sizeof(uint64_t), ctx->stream);
    if (pp.n > 0) {
        part_hist_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(pp);
        ctx->launches++;
    }
    cudaMemcpyAsync(ctx->h_scratch, pp.counts, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return nqe_fail(ctx, NQE_ERR_CUDA, "partition histogram failed");
    for (int p = 0; p < n_parts; p++) counts[p] = (int64_t)ctx->h_scratch[p];
    return NQE_OK;
}

extern "C" int32_t nqe_shuffle_scatter(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts,
                                       void *const *dst_columns, const int64_t *dst_offsets) {
    if (!ctx || !in || !dst_columns || !dst_offsets) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    NQE_TRY(check_partition_input(ctx, in, key_column, n_parts, 8, 8));
    ShuffleParams sp;
    memset(&sp, 0, sizeof sp);
    sp.keys = (const unsigned long long *)in->cols[key_column].values;
    sp.n = in->nrows;
    sp.n_parts = n_parts;
    sp.n_cols = (int)in->cols.size();
    for (int c = 0; c < sp.n_cols; c++) {
        sp.in[c] = (const unsigned long long *)in->cols[c].values;
        for (int p = 0; p < n_parts; p++) sp.dst[c * 8 + p] = (unsigned long long *)dst_columns[c * n_parts + p];
    }
    OpTimer timer(ctx);
    void *cur = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &cur, 64 * sizeof(uint64_t)));
    for (int p = 0; p < n_parts; p++) ctx->h_scratch[p] = (uint64_t)dst_offsets[p];
    cudaMemcpyAsync(cur, ctx->h_scratch, 64 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    sp.cursors = (unsigned long long *)cur;
    int32_t rc = NQE_OK;
    if (sp.n > 0) {
        const int64_t tiles = (sp.n + PT_THREADS * PT_K - 1) / (PT_THREADS * PT_K);
        int grid = ctx->sm_count * 4;
        if (grid > tiles) grid = (int)tiles;
        shuffle_scatter_kernel<<<grid, PT_THREADS, 0, ctx->stream>>>(sp);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "shuffle scatter launch failed");
    }
    // h_scratch is reused by later calls: the staging copy must have been consumed before returning
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "shuffle scatter failed: %s", cudaGetErrorString(cudaGetLastError()));
    timer.stop();
    nqe_dev_free(ctx, cur);
    return rc;
}

extern "C" int32_t nqe_radix_partition(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts,
                                       nqe_table **out, int64_t *counts) {
    if (!ctx || !in || !out || !counts) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    if (n_parts < 1 || n_parts > MAX_PARTS) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "n_parts must be in [1,%d]", MAX_PARTS);
    if (key_column < 0 || key_column >= (int)in->cols.size()) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "bad key column");
    if (in->cols.size() > 16) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "more than 16 columns");
    const int kd = in->cols[key_column].dtype;
    if (kd != NQE_INT64 && kd != NQE_UINT64) return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "partition key must be Int64/UInt64");
    for (auto &c : in->cols)
        if (c.dtype == NQE_BOOL || c.dtype == NQE_UTF8 || c.validity)
            return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "radix partition supports NULL-free 8-byte columns only");
    const int64_t n = in->nrows;
    nqe_table *t;
    nqe_table_new(ctx, n, &t);
    t->cols.resize(in->cols.size());
    PartParams pp;
    memset(&pp, 0, sizeof pp);
    pp.keys = (const unsigned long long *)in->cols[key_column].values;
    pp.n = n;
    pp.n_parts = n_parts;
    pp.n_cols = (int)in->cols.size();
    int32_t rc = NQE_OK;
    for (int c = 0; c < pp.n_cols && rc == NQE_OK; c++) {
        rc = nqe_column_alloc(ctx, in->cols[c].dtype, n, false, &t->cols[c]);
        pp.in[c] = (const unsigned long long *)in->cols[c].values;
        pp.out[c] = (unsigned long long *)t->cols[c].values;
    }
    OpTimer timer(ctx);
    unsigned long long *d_counts = (unsigned long long *)ctx->d_scratch; // 64 words
    pp.counts = d_counts;
    if (rc == NQE_OK) {
        cudaMemsetAsync(d_counts, 0, 64 * sizeof(uint64_t), ctx->stream);
        if (n > 0) {
            int grid = ctx->sm_count * 8;
            part_hist_kernel<<<grid, 256, 0, ctx->stream>>>(pp);
            ctx->launches++;
        }
        cudaMemcpyAsync(ctx->h_scratch, d_counts, 64 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "partition histogram failed");
    }
    if (rc == NQE_OK) {
        uint64_t run = 0;
        for (int p = 0; p < n_parts; p++) {
            counts[p] = (int64_t)ctx->h_scratch[p];
            ctx->h_scratch[p] = run; // exclusive prefix = initial cursor
            run += (uint64_t)counts[p];
        }
        void *cur = nullptr;
        rc = nqe_dev_alloc(ctx, &cur, 64 * sizeof(uint64_t));
        if (rc == NQE_OK) {
            // h_scratch is pinned; the copy is ordered before the kernel on the stream
            cudaMemcpyAsync(cur, ctx->h_scratch, 64 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
            pp.cursors = (unsigned long long *)cur;
            if (n > 0) {
                const int64_t tiles = (n + PT_THREADS * PT_K - 1) / (PT_THREADS * PT_K);
                int grid = ctx->sm_count * 4;
                if (grid > tiles) grid = (int)tiles;
                part_scatter_kernel<<<grid, PT_THREADS, 0, ctx->stream>>>(pp);
                ctx->launches++;
            }
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "partition scatter failed");
            nqe_dev_free(ctx, cur);
        }
    }
    timer.stop();
    if (rc != NQE_OK) {
        nqe_table_free(t);
        return rc;
    }
    *out = t;
    return NQE_OK;
}

// concat_batches (hash_join.rs:258-273: column-wise arrow `concat`; used on the join's build side :131-132 and on the
// aggregate's input, aggregate/mod.rs:143-144) for device tables: one device-to-device copy per value buffer, bitmaps
// (validity, Boolean values) placed at their bit offsets, Utf8 offsets rebased.  Schemas must agree column by column.
extern "C" int32_t nqe_table_concat(nqe_ctx *ctx, const nqe_table *const *tables, int32_t n_tables, nqe_table **out) {
    if (!ctx || !out || n_tables < 0 || (n_tables > 0 && !tables)) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    if (n_tables == 0) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "concat of zero tables needs a schema: pass an empty table");
    const size_t ncols = tables[0]->cols.size();
    int64_t total = 0;
    for (int t = 0; t < n_tables; t++) {
        if (!tables[t] || tables[t]->cols.size() != ncols) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "concat: batch %d has a different column count", t);
        for (size_t c = 0; c < ncols; c++)
            if (tables[t]->cols[c].dtype != tables[0]->cols[c].dtype)
                return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "concat: column %d of batch %d has a different type", (int)c, t); // arrow: ArrowError::InvalidArgumentError
        total += tables[t]->nrows;
    }
    if (total >= ((int64_t)1 << 31)) {
        for (size_t c = 0; c < ncols; c++)
            if (tables[0]->cols[c].dtype == NQE_UTF8) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "concat: Utf8 column beyond 2^31 rows");
    }
    nqe_table *r;
    nqe_table_new(ctx, total, &r);
    r->cols.resize(ncols);
    int32_t rc = NQE_OK;
    for (size_t c = 0; c < ncols && rc == NQE_OK; c++) {
        const int32_t dt = tables[0]->cols[c].dtype;
        bool any_valid = false;
        int64_t data_bytes = 0, nulls = 0;
        for (int t = 0; t < n_tables; t++) {
            const DevColumn &s = tables[t]->cols[c];
            if (s.validity) any_valid = true;
            data_bytes += s.data_bytes;
            nulls += s.null_count;
        }
        DevColumn &d = r->cols[c];
        rc = nqe_column_alloc(ctx, dt, total, any_valid, &d);
        if (rc != NQE_OK) break;
        if (dt == NQE_UTF8) {
            if (data_bytes >= ((int64_t)1 << 31)) { rc = nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "concat: Utf8 data beyond 2^31 bytes"); break; }
            d.data_bytes = data_bytes;
            rc = nqe_dev_alloc(ctx, (void **)&d.data, (size_t)data_bytes + 64);
            if (rc != NQE_OK) break;
        }
        if (dt == NQE_BOOL) cudaMemsetAsync(d.values, 0, nqe_bitmap_bytes(total), ctx->stream);
        if (any_valid) cudaMemsetAsync(d.validity, 0, nqe_bitmap_bytes(total), ctx->stream);
        int64_t at = 0, byte_at = 0;
        for (int t = 0; t < n_tables; t++) {
            const DevColumn &s = tables[t]->cols[c];
            const int64_t n = tables[t]->nrows;
            const unsigned wgrid = (unsigned)(((n + 31) / 32 + 255) / 256);
            if (n == 0) continue;
            if (dt == NQE_BOOL) {
                bitmap_place_kernel<<<wgrid, 256, 0, ctx->stream>>>((const uint32_t *)s.values, n, (uint32_t *)d.values, at);
            } else if (dt == NQE_UTF8) {
                // offsets[0 .. n) of the piece, rebased; the closing offset is written after the last piece
                utf8_offsets_place_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const int32_t *)s.values, n, (int32_t)byte_at,
                                                                                                (int32_t *)d.values + at);
                if (s.data_bytes) cudaMemcpyAsync(d.data + byte_at, s.data, (size_t)s.data_bytes, cudaMemcpyDeviceToDevice, ctx->stream);
            } else {
                cudaMemcpyAsync((uint64_t *)d.values + at, s.values, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream);
            }
            if (any_valid) bitmap_place_kernel<<<wgrid, 256, 0, ctx->stream>>>((const uint32_t *)s.validity, n, (uint32_t *)d.validity, at);
            ctx->launches++;
            at += n;
            byte_at += s.data_bytes;
        }
        if (dt == NQE_UTF8) {
            const int32_t end = (int32_t)data_bytes;
            cudaMemcpyAsync((int32_t *)d.values + total, &end, 4, cudaMemcpyHostToDevice, ctx->stream);
            cudaStreamSynchronize(ctx->stream); // `end` lives on this stack frame
        }
        d.null_count = nulls;
        if (any_valid && nulls == 0) { // arrow `concat`: no nulls => no bitmap
            nqe_dev_free(ctx, d.validity);
            d.validity = nullptr;
        }
    }
    if (rc == NQE_OK && cudaGetLastError() != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "concat failed");
    if (rc != NQE_OK) {
        nqe_table_free(r);
        return rc;
    }
    *out = r;
    return NQE_OK;
}

// filter_project.cu -- SelectionPlan (selection.rs:58-107) fused with
// ProjectionPlan (projection.rs:43-70): predicate evaluation, stable
// selection-vector compaction and projection arithmetic in one pass over HBM.
//
// Layout: a tile is K*THREADS consecutive rows; thread t owns rows
// tile_base + j*THREADS + t (j < K), so every warp-level load/store touches one
// contiguous 256-byte span.  Kept rows are ranked with warp ballots, one
// warp-scan over the K*WARPS ballot counts and a decoupled look-back across
// tiles (tile ids come from an atomic ticket, so a predecessor is always
// resident).  Because ranks inside one ballot are consecutive, each warp store
// writes one contiguous run of the compacted output.
#include <climits>

#include "expr_eval.cuh"
#include "nqe_internal.cuh"

namespace {

constexpr int FP_THREADS = 256;
constexpr int FP_WARPS = FP_THREADS / 32;

struct FilterParams {
    int64_t n_rows;
    int32_t n_out;
    int32_t first_out_prog; // 1 when program 0 is the predicate
    void *out_values[16];   // 8-byte values, or one byte per row for Boolean outputs
    uint8_t *out_valid[16]; // one byte per row (1 = valid) or nullptr
    unsigned long long *tile_state; // look-back words: [63:62] status, [61:0] count
    unsigned int *ticket;
    unsigned long long *out_count;
    uint32_t *status;
    int32_t num_tiles;
};

constexpr unsigned long long LB_AGG = 1ull << 62, LB_PREFIX = 2ull << 62, LB_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v));
}

// exclusive prefix of kept rows over all tiles before `tile` (called by warp 0)
__device__ __forceinline__ unsigned long long lookback(unsigned long long *state, int tile, unsigned long long my_total,
                                                       int lane) {
    if (lane == 0) st_volatile_u64(state + tile, (tile == 0 ? LB_PREFIX : LB_AGG) | my_total);
    if (tile == 0) return 0;
    unsigned long long excl = 0;
    int idx = tile - 1;
    while (true) {
        const int my = idx - lane;
        unsigned long long s;
        do {
            s = my >= 0 ? ld_volatile_u64(state + my) : LB_PREFIX;
        } while (__any_sync(0xffffffffu, (s >> 62) == 0));
        const unsigned m = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        unsigned long long v = s & LB_MASK;
        if (m) {
            const int first = __ffs(m) - 1;
            if (lane > first) v = 0;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        excl += v;
        if (m) break;
        idx -= 32;
    }
    if (lane == 0) st_volatile_u64(state + tile, LB_PREFIX | (excl + my_total));
    return excl;
}

template <int K, bool HAS_PRED>
__global__ void __launch_bounds__(FP_THREADS)
filter_project_kernel(const __grid_constant__ DevProgramSet ps, const __grid_constant__ FilterParams fp) {
    constexpr int TILE = K * FP_THREADS;
    __shared__ unsigned int s_cnt[K * FP_WARPS];
    __shared__ unsigned long long s_tile_excl;
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    while (true) {
        int tile;
        if (HAS_PRED) {
            if (tid == 0) s_tile = (int)atomicAdd(fp.ticket, 1u);
            __syncthreads();
            tile = s_tile;
        } else {
            tile = blockIdx.x;
        }
        if (tile >= fp.num_tiles) break;
        const int64_t base = (int64_t)tile * TILE;
        const int64_t e0 = base + tid;
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (e0 + (int64_t)j * FP_THREADS < fp.n_rows) inrange |= 1u << j;

        uint32_t keep = inrange, rownull = 0;
        int64_t pos[K];
        if (HAS_PRED) {
            RowRegs<K> m;
            run_program<K>(ps, 0, e0, FP_THREADS, inrange, inrange, 0u, m, fp.status);
            uint32_t mt = 0;
#pragma unroll
            for (int j = 0; j < K; j++) mt |= (uint32_t)(m.v[j] & 1) << j;
            rownull = inrange & ~m.valid;            // predicate NULL: keep as an all-NULL row
            keep = inrange & ((mt & m.valid) | rownull);
            unsigned int rank[K];
#pragma unroll
            for (int j = 0; j < K; j++) {
                const unsigned b = __ballot_sync(0xffffffffu, (keep >> j) & 1u);
                rank[j] = __popc(b & ((1u << lane) - 1u));
                if (lane == 0) s_cnt[j * FP_WARPS + warp] = __popc(b);
            }
            __syncthreads();
            if (warp == 0) {
                // exclusive scan over the K*WARPS ballot counts (row order: j major, warp minor)
                constexpr int N = K * FP_WARPS, PER = (N + 31) / 32;
                unsigned int c[PER], sum = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    const int i = lane * PER + q;
                    c[q] = i < N ? s_cnt[i] : 0;
                    sum += c[q];
                }
                unsigned int incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                unsigned int run = incl - sum;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    const int i = lane * PER + q;
                    if (i < N) s_cnt[i] = run;
                    run += c[q];
                }
                const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
                const unsigned long long excl = lookback(fp.tile_state, tile, total, lane);
                if (lane == 0) {
                    s_tile_excl = excl;
                    if (tile == fp.num_tiles - 1) *fp.out_count = excl + total;
                }
            }
            __syncthreads();
            const unsigned long long tile_excl = s_tile_excl;
#pragma unroll
            for (int j = 0; j < K; j++) pos[j] = (int64_t)(tile_excl + s_cnt[j * FP_WARPS + warp] + rank[j]);
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) pos[j] = e0 + (int64_t)j * FP_THREADS;
        }

        const uint32_t active = keep & ~rownull;
        for (int o = 0; o < fp.n_out; o++) {
            RowRegs<K> r;
            run_program<K>(ps, fp.first_out_prog + o, e0, FP_THREADS, inrange, active, rownull, r, fp.status);
            if (ps.prog_type[fp.first_out_prog + o] == T_BOOL) {
                uint8_t *out = (uint8_t *)fp.out_values[o];
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((keep >> j) & 1u) out[pos[j]] = (uint8_t)(r.v[j] & 1);
            } else {
                uint64_t *out = (uint64_t *)fp.out_values[o];
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((keep >> j) & 1u) out[pos[j]] = ((r.valid >> j) & 1u) ? r.v[j] : 0ull;
            }
            if (fp.out_valid[o]) {
                uint8_t *ov = fp.out_valid[o];
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((keep >> j) & 1u) ov[pos[j]] = (uint8_t)((r.valid >> j) & 1u);
            }
        }
        if (!HAS_PRED) break;
        __syncthreads(); // s_cnt / s_tile reuse
    }
}

// bytes (0/1 per row) -> LSB-first bitmap; counts zero bytes (nulls) into *zeros
__global__ void pack_bytes_kernel(const uint8_t *__restrict__ bytes, int64_t n, uint32_t *__restrict__ words,
                                  unsigned long long *zeros) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nwords = (n + 31) / 32;
    unsigned int z = 0;
    if (w < nwords) {
        uint32_t bits = 0;
        const int64_t b0 = w * 32;
        if (b0 + 32 <= n) {
            const uint4 *p = (const uint4 *)(bytes + b0);
            const uint4 a = p[0], b = p[1];
            const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 8; q++)
#pragma unroll
                for (int k = 0; k < 4; k++) bits |= ((u[q] >> (8 * k)) & 1u) << (q * 4 + k);
            z = 32 - __popc(bits);
        } else {
            for (int k = 0; b0 + k < n; k++) {
                const uint32_t v = bytes[b0 + k] & 1u;
                bits |= v << k;
                z += 1 - v;
            }
        }
        words[w] = bits;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (zeros && (threadIdx.x & 31) == 0 && z) atomicAdd(zeros, (unsigned long long)z);
}

template <int K>
int32_t launch_fp(nqe_ctx *ctx, bool has_pred, const DevProgramSet &ps, FilterParams &fp) {
    constexpr int TILE = K * FP_THREADS;
    fp.num_tiles = (int32_t)((fp.n_rows + TILE - 1) / TILE);
    if (fp.num_tiles == 0) return NQE_OK;
    if (has_pred) {
        int occ = 0;
        NQE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_project_kernel<K, true>, FP_THREADS, 0));
        int grid = ctx->sm_count * (occ > 0 ? occ : 1);
        if (grid > fp.num_tiles) grid = fp.num_tiles;
        filter_project_kernel<K, true><<<grid, FP_THREADS, 0, ctx->stream>>>(ps, fp);
    } else {
        filter_project_kernel<K, false><<<fp.num_tiles, FP_THREADS, 0, ctx->stream>>>(ps, fp);
    }
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}

} // namespace

int32_t nqe_pack_bytes(nqe_ctx *ctx, const uint8_t *bytes, int64_t n, uint32_t *words, unsigned long long *zeros) {
    if (n <= 0) return NQE_OK;
    const int64_t nwords = (n + 31) / 32;
    pack_bytes_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, ctx->stream>>>(bytes, n, words, zeros);
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}

static int32_t status_to_error(nqe_ctx *ctx, uint32_t st) {
    if (st & DEV_ERR_DIV0) return nqe_fail(ctx, NQE_ERR_DIVIDE_BY_ZERO, "Divide by zero error");
    if (st & DEV_ERR_OVERFLOW) return nqe_fail(ctx, NQE_ERR_PANIC, "attempt to divide with overflow");
    return NQE_OK;
}

extern "C" int32_t nqe_filter_project(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate,
                                      const nqe_expr *projs, int32_t n_projs, nqe_table **out) {
    if (!ctx || !in || !out) return NQE_ERR_INVALID_ARG;
    if (n_projs < 0 || n_projs > 16) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "n_projs must be in [0,16]");
    cudaSetDevice(ctx->device);
    *out = nullptr;
    const int64_t n = in->nrows;

    // a bare SelectionPlan passes every input column through (selection.rs:65-101)
    std::vector<nqe_expr_node> pass_nodes;
    std::vector<nqe_expr> pass_exprs;
    if (n_projs == 0) {
        if (!predicate) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "neither predicate nor projections");
        n_projs = (int32_t)in->cols.size();
        if (n_projs > 16) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "selection over more than 16 columns");
        pass_nodes.resize(n_projs);
        pass_exprs.resize(n_projs);
        for (int i = 0; i < n_projs; i++) {
            pass_nodes[i] = nqe_expr_node{NQE_NODE_COLUMN, 0, i, 0, 0, 0, {0}};
            pass_exprs[i] = nqe_expr{&pass_nodes[i], 1, 0};
        }
        projs = pass_exprs.data();
    }

    DevProgramSet ps;
    memset(&ps, 0, sizeof ps);
    ExprInfo info[NQE_MAX_PROGS];
    std::vector<const nqe_expr *> list;
    if (predicate) list.push_back(predicate);
    for (int i = 0; i < n_projs; i++) list.push_back(&projs[i]);
    NQE_TRY(nqe_compile_exprs(ctx, in, list.data(), (int32_t)list.size(), &ps, info));
    const int first = predicate ? 1 : 0;
    if (predicate && info[0].result_dtype != NQE_BOOL)
        return nqe_fail(ctx, NQE_ERR_PANIC, "selection predicate is not Boolean (downcast_ref::<BooleanArray>().unwrap())");
    for (int i = 0; i < n_projs; i++)
        if (info[first + i].result_dtype == NQE_UTF8)
            return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "Utf8 columns are not implemented on the CUDA filter/project path yet");

    // a predicate that can be NULL makes every output nullable (NULL rows are kept)
    const bool pred_nullable = predicate && info[0].nullable;

    nqe_table *t;
    nqe_table_new(ctx, 0, &t);
    t->cols.resize(n_projs);
    FilterParams fp;
    memset(&fp, 0, sizeof fp);
    fp.n_rows = n;
    fp.n_out = n_projs;
    fp.first_out_prog = first;
    std::vector<uint8_t *> bool_bytes(n_projs, nullptr), valid_bytes(n_projs, nullptr);
    int32_t rc = NQE_OK;
    for (int i = 0; i < n_projs && rc == NQE_OK; i++) {
        const ExprInfo &ei = info[first + i];
        const bool nullable = ei.nullable || pred_nullable;
        rc = nqe_column_alloc(ctx, ei.result_dtype, n, nullable, &t->cols[i]);
        if (rc != NQE_OK) break;
        if (ei.result_dtype == NQE_BOOL) {
            rc = nqe_dev_alloc(ctx, (void **)&bool_bytes[i], (size_t)n + 64);
            fp.out_values[i] = bool_bytes[i];
        } else {
            fp.out_values[i] = t->cols[i].values;
        }
        if (rc == NQE_OK && nullable) {
            rc = nqe_dev_alloc(ctx, (void **)&valid_bytes[i], (size_t)n + 64);
            fp.out_valid[i] = valid_bytes[i];
        }
    }
    void *lb = nullptr;
    // d_scratch words: [0] out_count, [1] status, [2] ticket, [8..8+n_projs) null counts
    if (rc == NQE_OK && cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream) != cudaSuccess) rc = NQE_ERR_CUDA;
    fp.out_count = (unsigned long long *)ctx->d_scratch;
    fp.status = (uint32_t *)(ctx->d_scratch + 1);
    fp.ticket = (unsigned int *)(ctx->d_scratch + 2);
    constexpr int K = 8;
    const int64_t num_tiles = (n + K * FP_THREADS - 1) / (K * FP_THREADS);
    if (rc == NQE_OK && predicate) {
        rc = nqe_dev_alloc(ctx, &lb, (size_t)(num_tiles + 1) * 8);
        if (rc == NQE_OK && cudaMemsetAsync(lb, 0, (size_t)(num_tiles + 1) * 8, ctx->stream) != cudaSuccess) rc = NQE_ERR_CUDA;
        fp.tile_state = (unsigned long long *)lb;
    }
    OpTimer timer(ctx);
    if (rc == NQE_OK) rc = launch_fp<K>(ctx, predicate != nullptr, ps, fp);
    int64_t out_rows = n;
    if (rc == NQE_OK) {
        cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = nqe_fail(ctx, NQE_ERR_CUDA, "filter_project kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == NQE_OK) {
        if (predicate) out_rows = n ? (int64_t)ctx->h_scratch[0] : 0;
        rc = status_to_error(ctx, (uint32_t)ctx->h_scratch[1]);
    }
    // bitmaps for Boolean outputs / validity
    if (rc == NQE_OK) {
        bool any = false;
        for (int i = 0; i < n_projs && rc == NQE_OK; i++) {
            if (bool_bytes[i]) { rc = nqe_pack_bytes(ctx, bool_bytes[i], out_rows, (uint32_t *)t->cols[i].values, nullptr); any = true; }
            if (rc == NQE_OK && valid_bytes[i]) {
                rc = nqe_pack_bytes(ctx, valid_bytes[i], out_rows, (uint32_t *)t->cols[i].validity,
                                    (unsigned long long *)(ctx->d_scratch + 8 + i));
                any = true;
            }
        }
        if (rc == NQE_OK && any) {
            cudaMemcpyAsync(ctx->h_scratch + 8, ctx->d_scratch + 8, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "pack kernel failed");
        }
    }
    timer.stop();
    for (int i = 0; i < n_projs; i++) {
        nqe_dev_free(ctx, bool_bytes[i]);
        nqe_dev_free(ctx, valid_bytes[i]);
    }
    nqe_dev_free(ctx, lb);
    if (rc != NQE_OK) {
        nqe_table_free(t);
        return rc;
    }
    t->nrows = out_rows;
    for (int i = 0; i < n_projs; i++) {
        DevColumn &c = t->cols[i];
        c.length = out_rows;
        if (c.validity) {
            c.null_count = (int64_t)ctx->h_scratch[8 + i];
            if (c.null_count == 0) { // Arrow: no nulls => no bitmap (builder.finish())
                nqe_dev_free(ctx, c.validity);
                c.validity = nullptr;
            }
        }
    }
    *out = t;
    return NQE_OK;
}

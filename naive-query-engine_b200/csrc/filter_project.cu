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
#include <cstdlib>
#include <cstring>

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
__global__ void __launch_bounds__(FP_THREADS, (K >= 8 ? 2 : (K >= 4 ? 4 : 6)))
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

// ---------------------------------------------------------------------------
// TMA-staged variant.  A persistent CTA keeps STAGES tiles of every referenced
// column in flight with cp.async.bulk (1-D bulk copies completing on an
// mbarrier); the expression programs then read their operands from shared
// memory, so HBM latency is decoupled from the interpreter and a column used by
// several programs (`id` in the predicate and in the projection) is fetched once.
// ---------------------------------------------------------------------------
constexpr int TMA_STAGES_MAX = 4;

struct TmaParams {
    uint16_t voff16[NQE_MAX_COLS]; // stage-relative offset of the column's values / 16
    uint16_t boff16[NQE_MAX_COLS]; // ... of its validity bitmap / 16
    uint32_t stage_bytes;          // multiple of 128
    uint32_t tx_bytes;             // bytes landed per full tile
    int32_t stages;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void lb_publish(unsigned long long *state, int tile, unsigned long long total) {
    st_volatile_u64(state + tile, (tile == 0 ? LB_PREFIX : LB_AGG) | total);
}
// walk back over predecessors' aggregates until an inclusive prefix is found (warp 0)
__device__ __forceinline__ unsigned long long lb_walk(unsigned long long *state, int tile, unsigned long long my_total, int lane) {
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

// Software-pipelined persistent kernel: the COUNT phase of tile i+1 (predicate,
// ballots, publish the tile aggregate) runs before the WRITE phase of tile i
// (look-back, projection, stores), so aggregates reach the look-back chain as soon
// as a tile's data has landed and the look-back of tile i normally finds its
// predecessors already resolved.
template <int K>
__global__ void __launch_bounds__(FP_THREADS)
filter_project_tma_kernel(const __grid_constant__ DevProgramSet ps, const __grid_constant__ FilterParams fp,
                          const __grid_constant__ TmaParams tp) {
    constexpr int TILE = K * FP_THREADS;
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t s_bar[TMA_STAGES_MAX];
    __shared__ int s_tile_id[TMA_STAGES_MAX];
    __shared__ unsigned int s_cnt[2][K * FP_WARPS];
    __shared__ unsigned int s_total[2];
    __shared__ unsigned int s_flags[2][FP_THREADS]; // per thread: keep | rownull << 16
    __shared__ unsigned long long s_tile_excl;
    __shared__ uint16_t s_voff[NQE_MAX_COLS], s_boff[NQE_MAX_COLS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stages = tp.stages;

    // producer side (thread 0): claim the next tile and start its bulk copies into `stage`
    auto issue = [&](int stage) {
        const int tile = (int)atomicAdd(fp.ticket, 1u);
        s_tile_id[stage] = tile;
        uint64_t *bar = &s_bar[stage];
        if (tile < fp.num_tiles && (int64_t)(tile + 1) * TILE <= fp.n_rows) {
            mbar_arrive_expect_tx(bar, tp.tx_bytes);
            uint8_t *dst = smem_dyn + (size_t)stage * tp.stage_bytes;
            for (int c = 0; c < ps.n_cols; c++) {
                const DevColRef &col = ps.cols[c];
                if (col.dtype == NQE_BOOL)
                    bulk_g2s(dst + (size_t)tp.voff16[c] * 16, (const uint8_t *)col.values + (size_t)tile * (TILE / 8), TILE / 8, bar);
                else
                    bulk_g2s(dst + (size_t)tp.voff16[c] * 16, (const uint8_t *)col.values + (size_t)tile * TILE * 8, TILE * 8, bar);
                if (col.validity)
                    bulk_g2s(dst + (size_t)tp.boff16[c] * 16, (const uint8_t *)col.validity + (size_t)tile * (TILE / 8), TILE / 8, bar);
            }
        } else {
            mbar_arrive(bar); // past the end, or the ragged last tile (filled cooperatively below)
        }
    };

    // COUNT phase of pipeline slot `it`; returns the tile id (>= num_tiles: nothing left)
    auto count_phase = [&](int it) -> int {
        const int stage = it % stages, buf = it & 1;
        mbar_wait(&s_bar[stage], (uint32_t)((it / stages) & 1));
        const int tile = s_tile_id[stage];
        if (tile >= fp.num_tiles) return tile;
        uint8_t *stg = smem_dyn + (size_t)stage * tp.stage_bytes;
        const int64_t e0 = (int64_t)tile * TILE + tid;
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (e0 + (int64_t)j * FP_THREADS < fp.n_rows) inrange |= 1u << j;
        if ((int64_t)(tile + 1) * TILE > fp.n_rows) {
            // ragged last tile: every thread stages its own rows (and its warp's bitmap words)
            for (int c = 0; c < ps.n_cols; c++) {
                const DevColRef &col = ps.cols[c];
#pragma unroll
                for (int j = 0; j < K; j++) {
                    if (!((inrange >> j) & 1u)) continue;
                    const int r = j * FP_THREADS + tid;
                    const int64_t e = e0 + (int64_t)j * FP_THREADS;
                    if (col.dtype == NQE_BOOL)
                        *(uint32_t *)(stg + (size_t)s_voff[c] * 16 + (r >> 5) * 4) = ((const uint32_t *)col.values)[e >> 5];
                    else
                        *(uint64_t *)(stg + (size_t)s_voff[c] * 16 + (size_t)r * 8) = ((const uint64_t *)col.values)[e];
                    if (col.validity)
                        *(uint32_t *)(stg + (size_t)s_boff[c] * 16 + (r >> 5) * 4) = col.validity[e >> 5];
                }
            }
            __syncwarp();
        }
        const SmemRows rows{stg, s_voff, s_boff, tid, FP_THREADS};
        RowRegs<K> m;
        run_program_on<K>(ps, 0, rows, inrange, inrange, 0u, m, fp.status);
        uint32_t mt = 0;
#pragma unroll
        for (int j = 0; j < K; j++) mt |= (uint32_t)(m.v[j] & 1) << j;
        const uint32_t rownull = inrange & ~m.valid;            // predicate NULL: keep as an all-NULL row
        const uint32_t keep = inrange & ((mt & m.valid) | rownull);
        s_flags[buf][tid] = keep | (rownull << 16);
#pragma unroll
        for (int j = 0; j < K; j++) {
            const unsigned b = __ballot_sync(0xffffffffu, (keep >> j) & 1u);
            if (lane == 0) s_cnt[buf][j * FP_WARPS + warp] = __popc(b);
        }
        __syncthreads();
        if (warp == 0) {
            constexpr int N = K * FP_WARPS, PER = (N + 31) / 32;
            unsigned int c[PER], sum = 0;
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int i = lane * PER + q;
                c[q] = i < N ? s_cnt[buf][i] : 0;
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
                if (i < N) s_cnt[buf][i] = run;
                run += c[q];
            }
            if (lane == 31) {
                s_total[buf] = incl;
                lb_publish(fp.tile_state, tile, incl);
            }
        }
        return tile;
    };

    // WRITE phase of pipeline slot `it` (its COUNT phase already ran)
    auto write_phase = [&](int it, int tile) {
        const int stage = it % stages, buf = it & 1;
        if (warp == 0) {
            const unsigned int total = s_total[buf];
            const unsigned long long excl = lb_walk(fp.tile_state, tile, total, lane);
            if (lane == 0) {
                s_tile_excl = excl;
                if (tile == fp.num_tiles - 1) *fp.out_count = excl + total;
            }
        }
        __syncthreads();
        const unsigned long long tile_excl = s_tile_excl;
        const uint32_t fl = s_flags[buf][tid];
        const uint32_t keep = fl & 0xffffu, rownull = fl >> 16;
        const int64_t e0 = (int64_t)tile * TILE + tid;
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (e0 + (int64_t)j * FP_THREADS < fp.n_rows) inrange |= 1u << j;
        int64_t pos[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const unsigned b = __ballot_sync(0xffffffffu, (keep >> j) & 1u);
            pos[j] = (int64_t)(tile_excl + s_cnt[buf][j * FP_WARPS + warp] + __popc(b & ((1u << lane) - 1u)));
        }
        const SmemRows rows{smem_dyn + (size_t)stage * tp.stage_bytes, s_voff, s_boff, tid, FP_THREADS};
        const uint32_t active = keep & ~rownull;
        for (int o = 0; o < fp.n_out; o++) {
            RowRegs<K> r;
            run_program_on<K>(ps, 1 + o, rows, inrange, active, rownull, r, fp.status);
            if (ps.prog_type[1 + o] == T_BOOL) {
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
        __syncthreads(); // stage buffer, s_flags[buf], s_cnt[buf], s_tile_excl are free again
        if (tid == 0) issue(stage);
    };

    if (tid < NQE_MAX_COLS) {
        s_voff[tid] = tp.voff16[tid];
        s_boff[tid] = tp.boff16[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < stages; s++) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < stages; s++) issue(s);
    }
    __syncthreads();

    int tile = count_phase(0);
    for (int it = 0; tile < fp.num_tiles; it++) {
        const int next = count_phase(it + 1); // publishes tile it+1's aggregate before tile it is written
        write_phase(it, tile);
        tile = next;
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
    if (!has_pred) {
        filter_project_kernel<K, false><<<fp.num_tiles, FP_THREADS, 0, ctx->stream>>>(ps, fp);
        ctx->launches++;
        NQE_CUDA(ctx, cudaGetLastError());
        return NQE_OK;
    }
    // TMA-staged path: every referenced buffer 16-byte aligned and the stage ring fits in shared memory
    static int impl = -1; // tuning knob NQE_FP_IMPL=direct|tma
    if (impl < 0) {
        const char *e = getenv("NQE_FP_IMPL");
        impl = (e && !strcmp(e, "direct")) ? 0 : 1;
    }
    TmaParams tp;
    memset(&tp, 0, sizeof tp);
    uint32_t off = 0, tx = 0;
    bool ok = impl == 1 && ps.n_cols > 0;
    for (int c = 0; c < ps.n_cols && ok; c++) {
        const DevColRef &col = ps.cols[c];
        if (((uintptr_t)col.values & 15) || ((uintptr_t)col.validity & 15)) ok = false;
        const uint32_t vb = col.dtype == NQE_BOOL ? TILE / 8 : TILE * 8;
        tp.voff16[c] = (uint16_t)(off / 16);
        off += vb;
        tx += vb;
        if (col.validity) {
            tp.boff16[c] = (uint16_t)(off / 16);
            off += TILE / 8;
            tx += TILE / 8;
        }
    }
    tp.stage_bytes = (off + 127) & ~127u;
    tp.tx_bytes = tx;
    if (ok && tp.stage_bytes * 2 > 200 * 1024) ok = false;
    if (ok) {
        static int want_stages = 0;
        if (!want_stages) {
            const char *e = getenv("NQE_FP_STAGES");
            want_stages = e ? atoi(e) : 3;
            if (want_stages < 2 || want_stages > TMA_STAGES_MAX) want_stages = 3;
        }
        int stages = want_stages;
        while (stages > 2 && (size_t)stages * tp.stage_bytes > 100 * 1024) stages--; // leave room for >= 2 CTAs per SM
        tp.stages = stages;
        const size_t dyn = (size_t)stages * tp.stage_bytes;
        auto kern = filter_project_tma_kernel<K>;
        NQE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        int occ = 0;
        NQE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FP_THREADS, dyn));
        if (occ < 1) ok = false;
        if (ok) {
            int grid = ctx->sm_count * occ;
            if (grid > fp.num_tiles) grid = fp.num_tiles;
            kern<<<grid, FP_THREADS, dyn, ctx->stream>>>(ps, fp, tp);
            ctx->launches++;
            NQE_CUDA(ctx, cudaGetLastError());
            return NQE_OK;
        }
    }
    int occ = 0;
    NQE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_project_kernel<K, true>, FP_THREADS, 0));
    int grid = ctx->sm_count * (occ > 0 ? occ : 1);
    if (grid > fp.num_tiles) grid = fp.num_tiles;
    filter_project_kernel<K, true><<<grid, FP_THREADS, 0, ctx->stream>>>(ps, fp);
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
    static int K = 0;
    if (!K) {
        const char *e = getenv("NQE_FP_K"); // tuning knob: rows per thread per tile
        K = e ? atoi(e) : 8;
        if (K != 2 && K != 4 && K != 8) K = 8;
    }
    const int64_t num_tiles = (n + K * FP_THREADS - 1) / (K * FP_THREADS);
    if (rc == NQE_OK && predicate) {
        rc = nqe_dev_alloc(ctx, &lb, (size_t)(num_tiles + 1) * 8);
        if (rc == NQE_OK && cudaMemsetAsync(lb, 0, (size_t)(num_tiles + 1) * 8, ctx->stream) != cudaSuccess) rc = NQE_ERR_CUDA;
        fp.tile_state = (unsigned long long *)lb;
    }
    OpTimer timer(ctx);
    if (rc == NQE_OK)
        rc = K == 2 ? launch_fp<2>(ctx, predicate != nullptr, ps, fp)
           : K == 4 ? launch_fp<4>(ctx, predicate != nullptr, ps, fp)
                    : launch_fp<8>(ctx, predicate != nullptr, ps, fp);
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

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
#include <chrono>
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
    int32_t num_tiles;  // tiles of the whole input (the last one writes out_count)
    int32_t tile_base;  // direct kernel: first tile handled by this launch
    int32_t tile_end;   // TMA kernel: tiles [0, tile_end) are handled by this launch (all full)
};

#include "lookback_body.inc"

// publish + walk (warp 0 of a tile)
__device__ __forceinline__ unsigned long long lookback(unsigned long long *state, int tile, unsigned long long my_total,
                                                       int lane) {
    if (lane == 0) nqe_lb_publish(state, tile, my_total);
    return nqe_lb_walk(state, tile, my_total, lane);
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
            if (tid == 0) s_tile = fp.tile_base + (int)atomicAdd(fp.ticket, 1u);
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
constexpr int TMA_STAGES_MAX = 3;

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

// ---------------------------------------------------------------------------
// Warp-specialised, software-pipelined persistent kernel (the hot path).
//
//   warps 0..7  WORKERS  count phase of tile i   : predicate -> keep ballots -> smem
//                        write phase of tile i-1 : ranks + tile prefix -> projection -> stores
//   warp  8     CONTROL  claims tiles (atomic ticket) and issues their cp.async.bulk copies,
//                        scans the 8*K ballot counts of tile i, publishes the tile aggregate,
//                        runs the decoupled look-back and hands the offsets to the workers.
//
// All hand-offs are mbarriers (no __syncthreads in the steady state), so the look-back
// latency of tile i overlaps the workers' write of tile i-1 and count of tile i+1.
// ---------------------------------------------------------------------------
constexpr int WS_WORKERS = 8;                       // worker warps
constexpr int WS_THREADS = (WS_WORKERS + 1) * 32;   // + control warp

template <int K, int STAGES>
struct WsSmem {
    uint64_t full[STAGES];       // TMA bytes landed (tx) / "no tile" arrive
    uint64_t stage_free[STAGES]; // 8 worker arrives: stage buffer may be refilled
    uint64_t cnt_ready[2];       // 8 worker arrives: ballots of the tile are in smem
    uint64_t off_ready[2];       // 1 control arrive: offsets + tile prefix are in smem
    unsigned long long tile_excl[2];
    int tile_id[STAGES];
    unsigned int keep[2][K][WS_WORKERS];    // keep ballots
    unsigned int rownull[2][K][WS_WORKERS]; // predicate-was-NULL ballots (NULLS only)
    unsigned int offs[2][K][WS_WORKERS];    // exclusive offsets inside the tile
    uint16_t voff[NQE_MAX_COLS], boff[NQE_MAX_COLS];
};

template <int K, int STAGES, bool NULLS>
__global__ void __launch_bounds__(WS_THREADS)
filter_project_ws_kernel(const __grid_constant__ DevProgramSet ps, const __grid_constant__ FilterParams fp,
                         const __grid_constant__ TmaParams tp) {
    constexpr int TILE = K * FP_THREADS;
    constexpr uint32_t ALL = (1u << K) - 1u;
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    __shared__ __align__(16) WsSmem<K, STAGES> sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid < NQE_MAX_COLS) {
        sm.voff[tid] = tp.voff16[tid];
        sm.boff[tid] = tp.boff16[tid];
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.stage_free[s], WS_WORKERS);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(&sm.cnt_ready[b], WS_WORKERS);
            mbar_init(&sm.off_ready[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == WS_WORKERS) {
        // ============================ CONTROL WARP ============================
        auto issue = [&](int stage) { // lane 0 only
            const int tile = (int)atomicAdd(fp.ticket, 1u);
            sm.tile_id[stage] = tile;
            uint64_t *bar = &sm.full[stage];
            if (tile < fp.tile_end) {
                mbar_arrive_expect_tx(bar, tp.tx_bytes);
                uint8_t *dst = smem_dyn + (size_t)stage * tp.stage_bytes;
                for (int c = 0; c < ps.n_cols; c++) {
                    const DevColRef &col = ps.cols[c];
                    if (col.dtype == NQE_BOOL)
                        bulk_g2s(dst + (size_t)tp.voff16[c] * 16, (const uint8_t *)col.values + (size_t)tile * (TILE / 8), TILE / 8, bar);
                    else
                        bulk_g2s(dst + (size_t)tp.voff16[c] * 16, (const uint8_t *)col.values + (size_t)tile * TILE * 8, TILE * 8, bar);
                    if (NULLS && col.validity)
                        bulk_g2s(dst + (size_t)tp.boff16[c] * 16, (const uint8_t *)col.validity + (size_t)tile * (TILE / 8), TILE / 8, bar);
                }
            } else {
                mbar_arrive(bar); // nothing left for this stage
            }
        };
        if (lane == 0)
            for (int s = 0; s < STAGES; s++) issue(s);
        __syncwarp();
        int stage = 0;
        for (int it = 0;; it++) {
            const int buf = it & 1;
            const int tile = *(volatile int *)&sm.tile_id[stage];
            if (tile >= fp.tile_end) break;
            mbar_wait(&sm.cnt_ready[buf], (uint32_t)((it >> 1) & 1));
            // exclusive scan of the K*8 ballot popcounts (row order: j major, worker warp minor)
            constexpr int N = K * WS_WORKERS, PER = (N + 31) / 32;
            unsigned int c[PER], sum = 0;
            const unsigned int *kb = &sm.keep[buf][0][0];
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int i = lane * PER + q;
                c[q] = i < N ? __popc(kb[i]) : 0;
                sum += c[q];
            }
            unsigned int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            unsigned int run = incl - sum;
            unsigned int *ob = &sm.offs[buf][0][0];
#pragma unroll
            for (int q = 0; q < PER; q++) {
                const int i = lane * PER + q;
                if (i < N) ob[i] = run;
                run += c[q];
            }
            const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) nqe_lb_publish(fp.tile_state, tile, total);
            const unsigned long long excl = nqe_lb_walk(fp.tile_state, tile, total, lane);
            if (lane == 0) {
                sm.tile_excl[buf] = excl;
                if (tile == fp.num_tiles - 1) *fp.out_count = excl + total;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.off_ready[buf]);
                // refill the stage the workers are writing out right now (tile it-1) as soon as they are done
                if (it >= 1) {
                    const int ps_ = (stage + STAGES - 1) % STAGES;
                    mbar_wait(&sm.stage_free[ps_], (uint32_t)(((it - 1) / STAGES) & 1));
                    issue(ps_);
                }
            }
            __syncwarp();
            stage = stage + 1 == STAGES ? 0 : stage + 1;
        }
        return;
    }

    // ================================ WORKERS ================================
    const unsigned int ltmask = (1u << lane) - 1u;

    auto write_tile = [&](int it, int stage, int tile) {
        const int buf = it & 1;
        mbar_wait(&sm.off_ready[buf], (uint32_t)((it >> 1) & 1));
        const unsigned long long tile_excl = sm.tile_excl[buf];
        uint32_t keep = 0, rownull = 0;
        unsigned int idx[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const unsigned int b = sm.keep[buf][j][warp];
            keep |= ((b >> lane) & 1u) << j;
            idx[j] = sm.offs[buf][j][warp] + __popc(b & ltmask);
            if (NULLS) rownull |= ((sm.rownull[buf][j][warp] >> lane) & 1u) << j;
        }
        const SmemRows rows{smem_dyn + (size_t)stage * tp.stage_bytes, sm.voff, sm.boff, tid, FP_THREADS};
        const uint32_t active = keep & ~rownull;
        for (int o = 0; o < fp.n_out; o++) {
            RowRegs<K> r;
            run_program_on<K, NULLS>(ps, 1 + o, rows, ALL, active, rownull, r, fp.status);
            if (ps.prog_type[1 + o] == T_BOOL) {
                uint8_t *out = (uint8_t *)fp.out_values[o] + tile_excl;
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((keep >> j) & 1u) out[idx[j]] = (uint8_t)(r.v[j] & 1);
            } else {
                uint64_t *out = (uint64_t *)fp.out_values[o] + tile_excl;
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((keep >> j) & 1u) out[idx[j]] = (!NULLS || ((r.valid >> j) & 1u)) ? r.v[j] : 0ull;
            }
            if (NULLS && fp.out_valid[o]) {
                uint8_t *ov = fp.out_valid[o] + tile_excl;
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((keep >> j) & 1u) ov[idx[j]] = (uint8_t)((r.valid >> j) & 1u);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.stage_free[stage]);
    };

    int stage = 0, prev_stage = 0, prev_tile = -1, it = 0;
    uint32_t full_phase = 0;
    for (;; it++) {
        const int buf = it & 1;
        mbar_wait(&sm.full[stage], full_phase);
        const int tile = sm.tile_id[stage];
        if (tile >= fp.tile_end) break;
        {
            // ---- count phase of tile `it`
            const SmemRows rows{smem_dyn + (size_t)stage * tp.stage_bytes, sm.voff, sm.boff, tid, FP_THREADS};
            RowRegs<K> m;
            run_program_on<K, NULLS>(ps, 0, rows, ALL, ALL, 0u, m, fp.status);
#pragma unroll
            for (int j = 0; j < K; j++) {
                const bool isnull = NULLS && !((m.valid >> j) & 1u); // predicate NULL: keep as an all-NULL row
                const unsigned kb = __ballot_sync(0xffffffffu, isnull || (m.v[j] & 1));
                if (lane == 0) sm.keep[buf][j][warp] = kb;
                if (NULLS) {
                    const unsigned nb = __ballot_sync(0xffffffffu, isnull);
                    if (lane == 0) sm.rownull[buf][j][warp] = nb;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.cnt_ready[buf]);
        }
        if (prev_tile >= 0) write_tile(it - 1, prev_stage, prev_tile);
        prev_tile = tile;
        prev_stage = stage;
        if (++stage == STAGES) {
            stage = 0;
            full_phase ^= 1u;
        }
    }
    if (prev_tile >= 0) write_tile(it - 1, prev_stage, prev_tile);
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

// the same for up to 32 byte maps of one operator call in ONE launch (blockIdx.y = map); the row count is read from
// device memory when `n_dev` is set, so the launch can be queued right behind the kernel that produces the count
struct PackParams {
    const uint8_t *bytes[32];
    uint32_t *words[32];
    unsigned long long *zeros[32];
    const unsigned long long *n_dev;
    int64_t n;
};
__global__ void pack_bytes_multi_kernel(const __grid_constant__ PackParams pk) {
    const int64_t n = pk.n_dev ? (int64_t)*pk.n_dev : pk.n;
    const int64_t nwords = (n + 31) / 32;
    if ((int64_t)blockIdx.x * blockDim.x >= nwords) return;
    const uint8_t *__restrict__ bytes = pk.bytes[blockIdx.y];
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
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
        pk.words[blockIdx.y][w] = bits;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (pk.zeros[blockIdx.y] && (threadIdx.x & 31) == 0 && z) atomicAdd(pk.zeros[blockIdx.y], (unsigned long long)z);
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
    const int32_t n_full = (int32_t)(fp.n_rows / TILE);
    if (ok && (tp.stage_bytes * 2 > 200 * 1024 || n_full == 0)) ok = false;
    int32_t done = 0; // tiles [0, done) handled by the TMA kernel
    if (ok) {
        static int want_stages = 0;
        if (!want_stages) {
            const char *e = getenv("NQE_FP_STAGES");
            want_stages = e ? atoi(e) : 3;
            if (want_stages < 2 || want_stages > TMA_STAGES_MAX) want_stages = 3;
        }
        int stages = want_stages;
        while (stages > 2 && (size_t)stages * tp.stage_bytes > 100 * 1024) stages--; // leave room for >= 2 CTAs per SM
        stages = stages >= 3 ? 3 : 2;
        tp.stages = stages;
        const size_t dyn2 = (size_t)stages * tp.stage_bytes;
        void (*kern)(DevProgramSet, FilterParams, TmaParams) =
            stages == 3 ? (ps.any_nulls ? filter_project_ws_kernel<K, 3, true> : filter_project_ws_kernel<K, 3, false>)
                        : (ps.any_nulls ? filter_project_ws_kernel<K, 2, true> : filter_project_ws_kernel<K, 2, false>);
        NQE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn2));
        int occ = 0;
        NQE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WS_THREADS, dyn2));
        if (occ >= 1) {
            int grid = ctx->sm_count * occ;
            if (grid > n_full) grid = n_full;
            fp.tile_end = n_full;
            kern<<<grid, WS_THREADS, dyn2, ctx->stream>>>(ps, fp, tp);
            ctx->launches++;
            NQE_CUDA(ctx, cudaGetLastError());
            done = n_full;
        }
    }
    if (done < fp.num_tiles) {
        // the ragged last tile (or everything, when the TMA path is not applicable): direct-load kernel,
        // continuing the same look-back chain; it draws tickets from its own counter word
        fp.tile_base = done;
        fp.ticket += 1;
        int occ = 0;
        NQE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_project_kernel<K, true>, FP_THREADS, 0));
        int grid = ctx->sm_count * (occ > 0 ? occ : 1);
        if (grid > fp.num_tiles - done) grid = fp.num_tiles - done;
        filter_project_kernel<K, true><<<grid, FP_THREADS, 0, ctx->stream>>>(ps, fp);
        ctx->launches++;
        NQE_CUDA(ctx, cudaGetLastError());
    }
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

int32_t nqe_jit_filter_project(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                               int32_t n_projs, void *const *out_values, uint8_t *const *out_valid,
                               unsigned long long *tile_state, unsigned int *ticket, unsigned long long *out_count,
                               uint32_t *status, bool *used, std::string *source_out);

int32_t nqe_filter_project_strings(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                                   int32_t n_projs, const int *utf8_src, nqe_table **out);
int32_t nqe_filter_project_utf8_compares(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                                         int32_t n_projs, nqe_table **out, bool *handled);

static int32_t status_to_error(nqe_ctx *ctx, uint32_t st) {
    if (st & DEV_ERR_DIV0) return nqe_fail(ctx, NQE_ERR_DIVIDE_BY_ZERO, "Divide by zero error");
    if (st & DEV_ERR_OVERFLOW) return nqe_fail(ctx, NQE_ERR_PANIC, "attempt to divide with overflow");
    return NQE_OK;
}

extern "C" int32_t nqe_filter_project(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate,
                                      const nqe_expr *projs, int32_t n_projs, nqe_table **out) {
    if (!ctx || !in || !out) return NQE_ERR_INVALID_ARG;
    if (n_projs < 0 || n_projs > 16) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "n_projs must be in [0,16]");
    // NQE_FP_PROF=1: where the host side of one call spends its time (averages every 25 calls, stderr)
    static int prof = -1;
    if (prof < 0) prof = getenv("NQE_FP_PROF") ? 1 : 0;
    static double acc[6];
    static int acc_n = 0;
    auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tp[7] = {0, 0, 0, 0, 0, 0, 0};
    if (prof) tp[0] = now();
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

    // comparisons of Utf8 columns / literals (binary.rs:127-132) become Boolean columns of a view of the input (utf8.cu)
    {
        bool handled = false;
        const int32_t urc = nqe_filter_project_utf8_compares(ctx, in, predicate, projs, n_projs, out, &handled);
        if (handled || urc != NQE_OK) return urc;
    }
    // bare references to Utf8 columns ride along through a hidden row-id column (utf8.cu)
    {
        int utf8_src[16];
        bool any_utf8 = false;
        for (int i = 0; i < n_projs; i++) {
            utf8_src[i] = -1;
            const nqe_expr &e = projs[i];
            if (e.n_nodes == 1 && e.nodes && e.nodes[0].kind == NQE_NODE_COLUMN && e.nodes[0].column >= 0 &&
                e.nodes[0].column < (int)in->cols.size() && in->cols[e.nodes[0].column].dtype == NQE_UTF8) {
                utf8_src[i] = e.nodes[0].column;
                any_utf8 = true;
            }
        }
        if (any_utf8) return nqe_filter_project_strings(ctx, in, predicate, projs, n_projs, utf8_src, out);
    }

    DevProgramSet ps;
    memset(&ps, 0, sizeof ps);
    ExprInfo info[NQE_MAX_PROGS];
    std::vector<const nqe_expr *> list;
    if (predicate) list.push_back(predicate);
    for (int i = 0; i < n_projs; i++) list.push_back(&projs[i]);
    NQE_TRY(nqe_compile_exprs(ctx, in, list.data(), (int32_t)list.size(), &ps, info));
    if (prof) tp[1] = now();
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
    const int64_t num_tiles = (n + 255) / 256; // look-back words for the smallest tile any kernel variant uses
    if (rc == NQE_OK && predicate) {
        rc = nqe_dev_alloc(ctx, &lb, (size_t)(num_tiles + 1) * 8);
        if (rc == NQE_OK && cudaMemsetAsync(lb, 0, (size_t)(num_tiles + 1) * 8, ctx->stream) != cudaSuccess) rc = NQE_ERR_CUDA;
        fp.tile_state = (unsigned long long *)lb;
    }
    if (prof) tp[2] = now();
    OpTimer timer(ctx);
    bool jit_used = false;
    static int jit_nulls = -1; // knob NQE_JIT_NULLS=0: nullable inputs go to the interpreter kernels
    if (jit_nulls < 0) {
        const char *e = getenv("NQE_JIT_NULLS");
        jit_nulls = e ? atoi(e) : 1;
    }
    if (rc == NQE_OK && (!ps.any_nulls || jit_nulls)) // query-shape specialised kernel (jit.cu); falls through when not applicable
        rc = nqe_jit_filter_project(ctx, in, predicate, projs, n_projs, fp.out_values, fp.out_valid, fp.tile_state, fp.ticket,
                                    fp.out_count, fp.status, &jit_used, nullptr);
    if (rc == NQE_OK && !jit_used)
        for (const DevColumn &c : in->cols)
            if (c.via >= 0) { // gathered columns (library-internal) exist only in the specialised kernel: the caller falls back
                rc = nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "gathered columns need the shape-specialised kernel");
                break;
            }
    if (rc == NQE_OK && !jit_used)
        rc = K == 2 ? launch_fp<2>(ctx, predicate != nullptr, ps, fp)
           : K == 4 ? launch_fp<4>(ctx, predicate != nullptr, ps, fp)
                    : launch_fp<8>(ctx, predicate != nullptr, ps, fp);
    int64_t out_rows = n;
    // bitmaps for Boolean outputs / validity: one more launch, queued behind the kernel (it reads the row count on the
    // device), so the operator still needs a single host synchronisation
    PackParams pk;
    memset(&pk, 0, sizeof pk);
    int n_maps = 0;
    for (int i = 0; i < n_projs; i++) {
        if (bool_bytes[i]) { pk.bytes[n_maps] = bool_bytes[i]; pk.words[n_maps] = (uint32_t *)t->cols[i].values; n_maps++; }
        if (valid_bytes[i]) {
            pk.bytes[n_maps] = valid_bytes[i];
            pk.words[n_maps] = (uint32_t *)t->cols[i].validity;
            pk.zeros[n_maps] = (unsigned long long *)(ctx->d_scratch + 8 + i);
            n_maps++;
        }
    }
    if (rc == NQE_OK && n_maps && n > 0) {
        pk.n = n;
        pk.n_dev = predicate ? fp.out_count : nullptr;
        const int64_t nwords = (n + 31) / 32;
        pack_bytes_multi_kernel<<<dim3((unsigned)((nwords + 255) / 256), (unsigned)n_maps), 256, 0, ctx->stream>>>(pk);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "pack kernel launch failed");
    }
    if (rc == NQE_OK) {
        timer.mark_end();
        cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (prof) tp[3] = now();
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            rc = nqe_fail(ctx, NQE_ERR_CUDA, "filter_project kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc == NQE_OK) {
        if (predicate) out_rows = n ? (int64_t)ctx->h_scratch[0] : 0;
        rc = status_to_error(ctx, (uint32_t)ctx->h_scratch[1]);
    }
    if (prof) tp[4] = now();
    timer.stop();
    if (prof) tp[5] = now();
    for (int i = 0; i < n_projs; i++) {
        nqe_dev_free(ctx, bool_bytes[i]);
        nqe_dev_free(ctx, valid_bytes[i]);
    }
    nqe_dev_free(ctx, lb);
    if (prof) {
        tp[6] = now();
        for (int i = 0; i < 6; i++) acc[i] += tp[i + 1] - tp[i];
        if (++acc_n == 25) {
            fprintf(stderr, "nqe_filter_project host us: compile %.1f | alloc+memset %.1f | launch+memcpy enqueue %.1f | sync wait %.1f | timer.stop %.1f | frees %.1f\n",
                    acc[0] / 25, acc[1] / 25, acc[2] / 25, acc[3] / 25, acc[4] / 25, acc[5] / 25);
            for (double &a : acc) a = 0;
            acc_n = 0;
        }
    }
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

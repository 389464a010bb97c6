// hash_join_direct.cuh -- part of hash_join.cu (included inside its anonymous namespace, after the join table, the probe
// helpers and the single-pass probe kernel): nqe_hash_join over a DIRECT table.
// ---- direct table, plain 8-byte columns: probe pass + staged emit pass -------------------------------------------
// join_probe_kernel above is one pass with a decoupled look-back, and it is slow on 1e8 rows (2.3 ms where the traffic is
// worth 0.8 ms) for two reasons that were measured one by one (profiles/README_r02.md):
//  * the look-back: all ~600 resident tiles are in the same phase, so a tile's walk crosses hundreds of unresolved
//    predecessors (37 % of the stall samples sit at the barrier behind the walking warp);
//  * random table reads and heavy store traffic in ONE kernel: a kernel with the reads alone takes 0.88 ms, with the
//    stores alone 0.81 ms (5.9 TB/s), with both 2.31 ms -- the table reads queue behind the SM's outstanding stores in the
//    memory pipeline and every tile waits for them at its barrier.  L2 policies (evict_first streams, evict_last table,
//    a persisting set-aside) change nothing.
// So the work is cut where the dependency is: every CTA owns a CONTIGUOUS chunk of tiles;
//   pass 1 (join_direct_probe_kernel) reads the keys, probes the table, writes the row words as a stream and counts
//          the chunk's matches -- random reads, almost no stores, no barriers;
//   pass 2 (join_direct_emit_kernel) starts at the sum of the earlier chunks' counts and walks its chunk with a running
//          base: the probe-side columns and the row-word stream of its next tiles are in flight as cp.async.bulk
//          copies into a shared-memory ring, ranks come from ballots (a row has 0 or 1 match), one barrier per tile,
//          no global loads in the loop at all.
constexpr int DE_MAX = 4; // columns per side
struct DirectEmitParams {
    JoinTable jt;
    int64_t n_probe;
    int32_t nl, nr, left_key, right_key;
    const unsigned long long *right[DE_MAX]; // probe-side columns
    unsigned long long *out[2 * DE_MAX];
    unsigned long long *rowwords;            // per probe row: the table's row word (EMPTY_ROW: no match), written by pass 1;
                                             // with a narrow table the stream is narrow too: the 4-byte slot as it is
    unsigned long long *chunk_count, *out_count;
    int32_t num_tiles, tiles_per_chunk, stages, pad;
};

template <int K>
__global__ void __launch_bounds__(HJ_THREADS) join_direct_probe_kernel(const __grid_constant__ DirectEmitParams p) {
    constexpr int T = HJ_THREADS, TILE = T * HJ_K, STEP = T * K;
    __shared__ unsigned int s_warp[HJ_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * p.tiles_per_chunk;
    const int64_t t1 = t0 + p.tiles_per_chunk < p.num_tiles ? t0 + p.tiles_per_chunk : p.num_tiles;
    const int64_t r0 = t0 * TILE, r1 = t1 * TILE < p.n_probe ? t1 * TILE : p.n_probe;
    const unsigned long long *keys = p.right[p.right_key];
    const unsigned long long ef = pj_policy();
    unsigned int cnt = 0;
    for (int64_t base = r0 + tid; base < r1; base += STEP) {
        unsigned long long key[K], brow[K];
        uint64_t slot[K];
        uint32_t live = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = base + (int64_t)j * T;
            key[j] = e < r1 ? ld_ef(keys + e, ef) : 0ull;
            if (e < r1) live |= 1u << j;
        }
        probe_direct<K, false>(p.jt, key, live, brow, slot);
#pragma unroll
        for (int j = 0; j < K; j++) {
            cnt += brow[j] != EMPTY_ROW;
            if (!((live >> j) & 1u)) continue;
            if (p.jt.narrow) // the slot again (row words of a narrow table are pay_lo + a 32-bit value)
                ((unsigned int *)p.rowwords)[base + (int64_t)j * T] = brow[j] == EMPTY_ROW ? 0xffffffffu : (unsigned int)(brow[j] - (unsigned long long)p.jt.pay_lo);
            else
                st_ef(p.rowwords + base + (int64_t)j * T, brow[j], ef);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_warp[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        unsigned long long total = 0;
        for (int w = 0; w < HJ_WARPS; w++) total += s_warp[w];
        p.chunk_count[blockIdx.x] = total;
    }
}

// NL build-side columns (the key, and for NL == 2 the one column whose values ride in the table), NR probe-side columns:
// compile-time, so that every column loop unrolls and the stores need one address computation each.
template <int K, int NL, int NR, bool N32>
__global__ void __launch_bounds__(HJ_THREADS, 4) join_direct_emit_kernel(const __grid_constant__ DirectEmitParams p) {
    constexpr int T = HJ_THREADS, TILE = T * K;
    constexpr int STAGE_WORDS = NR * TILE + (N32 ? TILE / 2 : TILE); // 8-byte words of one stage: NR columns + the row words
    static_assert(K * HJ_WARPS == 32, "one (row group, warp) count per lane");
    extern __shared__ __align__(128) unsigned char de_smem[];
    unsigned long long *full = (unsigned long long *)de_smem, *empty = full + 8; // [stages] each
    unsigned long long *ring = (unsigned long long *)(de_smem + 128);            // [stages][STAGE_WORDS]: probe columns, row words
    __shared__ unsigned int s_cnt[2][K * HJ_WARPS];
    __shared__ unsigned long long s_part[HJ_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.stages;
    const int64_t full_tiles = p.n_probe / TILE;
    const int64_t t0 = (int64_t)blockIdx.x * p.tiles_per_chunk;
    const int64_t t1 = t0 + p.tiles_per_chunk < p.num_tiles ? t0 + p.tiles_per_chunk : p.num_tiles;
    if (tid == 0) {
        for (int s = 0; s < S; s++) {
            nqe_mbar_init(full + s, 1);
            nqe_mbar_init(empty + s, HJ_WARPS);
        }
        nqe_mbar_init_fence();
    }
    // first output row of this chunk: the matches of all earlier chunks
    unsigned long long base = 0;
    for (int i = tid; i < (int)blockIdx.x; i += T) base += p.chunk_count[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    if (lane == 0) s_part[warp] = base;
    __syncthreads();
    base = 0;
#pragma unroll
    for (int w = 0; w < HJ_WARPS; w++) base += s_part[w];
    const unsigned long long ef = pj_policy();
    auto issue = [&](int64_t tile, int stage) { // thread 0
        nqe_mbar_arrive_expect_tx(full + stage, (uint32_t)STAGE_WORDS * 8u);
#pragma unroll
        for (int c = 0; c < NR; c++)
            nqe_bulk_g2s(ring + (size_t)stage * STAGE_WORDS + c * TILE, p.right[c] + tile * TILE, TILE * 8u, full + stage, ef);
        if (N32) nqe_bulk_g2s(ring + (size_t)stage * STAGE_WORDS + NR * TILE, (const unsigned int *)p.rowwords + tile * TILE, TILE * 4u, full + stage, ef);
        else nqe_bulk_g2s(ring + (size_t)stage * STAGE_WORDS + NR * TILE, p.rowwords + tile * TILE, TILE * 8u, full + stage, ef);
    };
    if (tid == 0)
        for (int s = 0; s < S; s++)
            if (t0 + s < t1 && t0 + s < full_tiles) issue(t0 + s, s);
    uint32_t it = 0;
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t tile = t0; tile < t1; tile++, it++) {
        const bool staged = tile < full_tiles;
        const unsigned long long *st = ring + (size_t)stage * STAGE_WORDS + tid;
        const int64_t e0 = tile * TILE + tid;
        unsigned long long key[K], brow[K];
        auto widen = [&](unsigned int w) { return w == 0xffffffffu ? EMPTY_ROW : (unsigned long long)w + (unsigned long long)p.jt.pay_lo; };
        if (staged) {
            nqe_mbar_wait(full + stage, parity);
#pragma unroll
            for (int j = 0; j < K; j++) {
                key[j] = st[p.right_key * TILE + j * T];
                if (N32) brow[j] = widen(((const unsigned int *)(ring + (size_t)stage * STAGE_WORDS + NR * TILE))[j * T + tid]);
                else brow[j] = st[NR * TILE + j * T];
            }
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int64_t e = e0 + (int64_t)j * T;
                key[j] = e < p.n_probe ? ld_stream_u64(p.right[p.right_key] + e) : 0ull;
                if (e >= p.n_probe) brow[j] = EMPTY_ROW;
                else if (N32) brow[j] = widen(((const unsigned int *)p.rowwords)[e]);
                else brow[j] = ld_stream_u64(p.rowwords + e);
            }
        }
        unsigned off[K];
        uint32_t emit = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const bool m = brow[j] != EMPTY_ROW;
            const unsigned b = __ballot_sync(0xffffffffu, m);
            off[j] = __popc(b & ((1u << lane) - 1u));
            if (lane == 0) s_cnt[it & 1u][j * HJ_WARPS + warp] = __popc(b);
            if (m) emit |= 1u << j;
        }
        __syncthreads();
        // every warp scans the 32 (row group, warp) counts for itself: no second barrier
        const unsigned mine = s_cnt[it & 1u][lane];
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const unsigned excl = incl - mine, total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int j = 0; j < K; j++) off[j] += __shfl_sync(0xffffffffu, excl, j * HJ_WARPS + warp);
#pragma unroll
        for (int c = 0; c < NL; c++) {
            unsigned long long *out = p.out[c] + base;
            const bool is_key = NL == 1 || c == p.left_key; // the build key matched: its value is the probe key
#pragma unroll
            for (int j = 0; j < K; j++)
                if ((emit >> j) & 1u) st_ef(out + off[j], is_key ? key[j] : brow[j], ef);
        }
#pragma unroll
        for (int c = 0; c < NR; c++) {
            unsigned long long *out = p.out[NL + c] + base;
            unsigned long long v[K];
            if (staged) {
#pragma unroll
                for (int j = 0; j < K; j++) v[j] = st[c * TILE + j * T];
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) v[j] = ((emit >> j) & 1u) ? (unsigned long long)ld_stream_u64(p.right[c] + e0 + (int64_t)j * T) : 0ull;
            }
#pragma unroll
            for (int j = 0; j < K; j++)
                if ((emit >> j) & 1u) st_ef(out + off[j], v[j], ef);
        }
        base += total;
        if (staged) { // hand the stage back; thread 0 refills it with the tile S steps ahead
            __syncwarp();
            if (lane == 0) nqe_mbar_arrive(empty + stage);
            if (tid == 0 && tile + S < t1 && tile + S < full_tiles) {
                nqe_mbar_wait(empty + stage, parity);
                issue(tile + S, stage);
            }
        }
        if (++stage == S) {
            stage = 0;
            parity ^= 1u;
        }
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) *p.out_count = base;
}

template <int NL, int NR, bool N32>
int32_t join_direct_emit_launch_n(nqe_ctx *ctx, DirectEmitParams &de) {
    de.stages = NR <= 1 ? 3 : 2;
    const size_t smem = 128 + (size_t)de.stages * (NR * 8 + (N32 ? 4 : 8)) * HJ_K * HJ_THREADS;
    auto dk = join_direct_emit_kernel<HJ_K, NL, NR, N32>;
    NQE_CUDA(ctx, cudaFuncSetAttribute(dk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int docc = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&docc, dk, HJ_THREADS, smem);
    int chunks = ctx->sm_count * (docc > 0 ? docc : 1); // one wave of CTAs, a contiguous chunk of tiles each
    if (chunks > de.num_tiles) chunks = de.num_tiles;
    de.tiles_per_chunk = (de.num_tiles + chunks - 1) / chunks;
    chunks = (de.num_tiles + de.tiles_per_chunk - 1) / de.tiles_per_chunk;
    join_direct_probe_kernel<8><<<chunks, HJ_THREADS, 0, ctx->stream>>>(de);
    dk<<<chunks, HJ_THREADS, smem, ctx->stream>>>(de);
    ctx->launches += 2;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}
template <int NL, int NR>
int32_t join_direct_emit_launch(nqe_ctx *ctx, DirectEmitParams &de) {
    return de.jt.narrow ? join_direct_emit_launch_n<NL, NR, true>(ctx, de) : join_direct_emit_launch_n<NL, NR, false>(ctx, de);
}


// hash_join_paged.cuh -- part of hash_join.cu (included inside its anonymous namespace, after the fused single-pass
// kernel): the fused join -> group-by as streaming passes over paged 16-byte rows (paged_split.cuh), over the hashed
// table (first split by slot range + probe and re-split) and over a direct table (probe and split in one pass).
// ---- Fused join -> group-by over paged streams ------------------------------------------------------------------
// The direct kernel above makes one DRAM-random slot access per probe row into a table far larger than the L2
// (11.6 GB of DRAM traffic per 1e8 rows where the inputs are 1.76 GB) plus the L2-bound group-table updates.  When the
// build keys are unique and the group key rides in the slots (JoinTable::rowpay) the work is reorganised into three
// streaming passes over 16-byte rows (paged_split.cuh):
//   1. ps_split_kernel<PartBySlotRange>: (fk, value) rows split by the top bits of the key's hash = ranges of table
//      slots of <= 48 MiB;
//   2. ja_probe_scatter_kernel: pages are swept partition after partition (every CTA takes every gridDim-th page of
//      the concatenated page list, so the whole grid works inside one slot range at a time and the range stays in L2);
//      a probe row becomes (group key, value) and is split again, by group-key hash, into the group-by's partitions;
//   3. gp2_aggregate_kernel (hash_aggregate.cu): shared-memory aggregation of every partition.
// Unmatched probe rows simply vanish in pass 2 (inner join).
struct PartBySlotRange { // partition = top of the hash = contiguous range of table slots (join_slot_of is monotone in the hash)
    uint64_t P;
    __device__ __forceinline__ int operator()(unsigned long long key) const { return (int)__umul64hi(nqe_mix64(key), P); }
};

// A CTA works on PIECES of T * K rows (a quarter, a half or a whole page) through one cp.async.bulk-filled buffer:
// the rows go into registers, the buffer is handed back to thread 0 -- the elected producer -- which issues the copy
// of the CTA's next piece at once, so it flies while the CTA probes and scatters.  Several small CTAs per SM overlap
// each other's phases (bulk-copy wait, L2 probe latency, the scatter's barriers and its global cursor round trip).
template <int T, int K>
struct JpSmem {
    ulonglong2 piece[T * K];
    PsScatterSmem<T, K> sc;
    PsPageBuf buf;
    unsigned int pstart[PS_MAX_PARTS + 8]; // first page of every input partition in the concatenated page list
};

// piece i of the concatenated page list -> (partition, page, first row inside the page, rows); *p is a running cursor
template <int PIECE>
__device__ __forceinline__ unsigned jp_locate(const PagedStreams &in, const unsigned int *pstart, unsigned i, int *p, unsigned *q,
                                              unsigned *row0) {
    constexpr unsigned PER_PAGE = PS_PAGE_ROWS / PIECE;
    const unsigned w = i / PER_PAGE;
    int pp = *p;
    while (pp + 1 < in.P && w >= pstart[pp + 1]) pp++;
    *p = pp;
    *q = w - pstart[pp];
    *row0 = (i % PER_PAGE) * PIECE;
    const unsigned fill = ps_page_rows(in, pp, *q);
    return fill <= *row0 ? 0u : (fill - *row0 < (unsigned)PIECE ? fill - *row0 : (unsigned)PIECE);
}

template <int T, int K, int MINB>
__global__ void __launch_bounds__(T, MINB)
ja_probe_scatter_kernel(const __grid_constant__ PagedStreams in, const __grid_constant__ PagedStreams out, const JoinTable jt,
                        uint32_t P2, long long dense_lo, uint32_t dense_width) {
    constexpr int PIECE = T * K;
    extern __shared__ __align__(128) unsigned char jp_smem_raw[];
    JpSmem<T, K> &sm = *reinterpret_cast<JpSmem<T, K> *>(jp_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t dense_magic = dense_width ? ps_div_magic(dense_width) : 0u;
    if (tid == 0) {
        unsigned acc = 0;
        for (int p = 0; p < in.P; p++) {
            sm.pstart[p] = acc;
            acc += (unsigned)((in.cursor[p] + PS_PAGE_ROWS - 1) >> PS_PAGE_SHIFT);
        }
        sm.pstart[in.P] = acc;
    }
    ps_pagebuf_init(sm.buf, T / 32);
    ps_scatter_init(sm.sc); // __syncthreads inside
    const unsigned total = sm.pstart[in.P] * (PS_PAGE_ROWS / PIECE);
    // thread 0 walks the piece list one step ahead of the CTA (pn, in, next piece with rows)
    int pn = 0;
    unsigned nxt = blockIdx.x; // next piece to issue
    auto issue_next = [&]() {  // thread 0: issue the copy of the next non-empty piece at or after `nxt`
        while (nxt < total) {
            unsigned q, row0;
            const unsigned rows = jp_locate<PIECE>(in, sm.pstart, nxt, &pn, &q, &row0);
            nxt += gridDim.x;
            if (!rows) continue;
            const unsigned pte = in.pt[(size_t)pn * in.pt_stride + q];
            nqe_mbar_arrive_expect_tx(&sm.buf.full, rows * 16u);
            const ulonglong2 *src = in.pool + ((size_t)(pte - 1u) << PS_PAGE_SHIFT) + row0;
            for (unsigned off = 0; off < rows; off += 1024u) {
                const unsigned len = rows - off < 1024u ? rows - off : 1024u;
                nqe_bulk_g2s(sm.piece + off, src + off, len * 16u, &sm.buf.full, nqe_policy_evict_first());
            }
            return;
        }
    };
    if (tid == 0) issue_next();
    int p = 0;
    uint32_t it = 0;
    for (unsigned i = blockIdx.x; i < total; i += gridDim.x) {
        unsigned q, row0;
        const unsigned rows = jp_locate<PIECE>(in, sm.pstart, i, &p, &q, &row0);
        if (!rows) continue; // CTA-uniform
        nqe_mbar_wait(&sm.buf.full, it & 1u);
        unsigned long long key[K], val[K];
        uint32_t live = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const unsigned idx = j * T + tid;
            ulonglong2 r = make_ulonglong2(0ull, 0ull);
            if (idx < rows) { r = sm.piece[idx]; live |= 1u << j; }
            key[j] = r.x;
            val[j] = r.y;
        }
        __syncwarp();
        if (lane == 0) nqe_mbar_arrive(&sm.buf.empty);
        if (tid == 0) {
            nqe_mbar_wait(&sm.buf.empty, it & 1u);
            issue_next();
        }
        it++;
        unsigned long long grp[K];
        uint64_t slot[K];
        probe_first<K>(jt, key, live, grp, slot); // unique build keys: the first match is the only one
        int pid[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (grp[j] == EMPTY_ROW) live &= ~(1u << j);
            // dense group keys (their exact range is known from the build side): partition = key range, else key hash
            pid[j] = dense_width ? (int)ps_div((uint32_t)(grp[j] - (unsigned long long)dense_lo), dense_width, dense_magic)
                                 : (int)__umulhi((uint32_t)(nqe_mix64(grp[j]) >> 32), P2);
            if (!((live >> j) & 1u)) pid[j] = 0;
        }
        ps_scatter_tile<T, K>(out, sm.sc, grp, val, pid, live);
    }
}

template <int T, int K, int MINB>
int32_t ja_probe_scatter_launch_shape(nqe_ctx *ctx, const PagedStreams &in, const PagedStreams &out, const JoinTable &jt, uint32_t P2,
                                      long long dense_lo, uint32_t dense_width) {
    auto kern = ja_probe_scatter_kernel<T, K, MINB>;
    NQE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(JpSmem<T, K>)));
    kern<<<ctx->sm_count * MINB, T, sizeof(JpSmem<T, K>), ctx->stream>>>(in, out, jt, P2, dense_lo, dense_width);
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}
// knob NQE_JA_PROBE_SHAPE: 0 (default) = 256 threads x 2 rows (512-row pieces, 6 CTAs/SM at 40 registers), 1 = 512 x 4
// (half pages, 2 CTAs/SM).  Measured 1e8 x 1e7, whole operator, with the shapes that were removed again (1024 x 4 at one
// CTA/SM, 256 x 4 at 5 and at 4 CTAs/SM): 4.69 / 4.92 / 4.95 / 5.00 / 4.84 ms: the kernel waits on L2 round trips, more
// resident warps beat more rows per thread.
int32_t ja_probe_scatter_launch(nqe_ctx *ctx, const PagedStreams &in, const PagedStreams &out, const JoinTable &jt, uint32_t P2,
                                long long dense_lo, uint32_t dense_width) {
    static int shape = -1;
    if (shape < 0) {
        const char *e = getenv("NQE_JA_PROBE_SHAPE");
        shape = e ? atoi(e) : 0;
    }
    switch (shape) {
    case 1: return ja_probe_scatter_launch_shape<512, 4, 2>(ctx, in, out, jt, P2, dense_lo, dense_width);
    default: return ja_probe_scatter_launch_shape<256, 2, 6>(ctx, in, out, jt, P2, dense_lo, dense_width);
    }
}

// The same pass over a DIRECT table (dense unique build keys): the table is probed in place -- 8 bytes per key of the
// range, mostly L2-resident -- so the probe rows are read straight from their columns, no first split: a probe row
// (fk, value) becomes (group key, value as f64) and goes into the group-by's partitions.
// CM: bit 0 = the streams (probe rows in, pages out) carry evict_first, bit 1 = the table reads carry evict_last
template <int T, int K, int MINB, int CM>
__global__ void __launch_bounds__(T, MINB)
ja_direct_scatter_kernel(const __grid_constant__ PagedStreams out, const PsSplitArgs a, const JoinTable jt, uint32_t P2,
                         long long dense_lo, uint32_t dense_width) {
    constexpr int TILE = T * K;
    extern __shared__ __align__(16) unsigned char jd_smem_raw[];
    PsScatterSmem<T, K> &sm = *reinterpret_cast<PsScatterSmem<T, K> *>(jd_smem_raw);
    ps_scatter_init(sm);
    const uint32_t dense_magic = dense_width ? ps_div_magic(dense_width) : 0u;
    const int64_t num_tiles = (a.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE + threadIdx.x;
        unsigned long long key[K], val[K], grp[K];
        uint64_t slot[K];
        int pid[K];
        uint32_t live = 0;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (e0 + (int64_t)j * T < a.n) live |= 1u << j;
        const unsigned long long ef = (CM & 1) ? pj_policy() : 0ull;
        auto ld_in = [&](const unsigned long long *p) { return (CM & 1) ? ld_ef(p, ef) : (unsigned long long)ld_stream_u64(p); };
#pragma unroll
        for (int j = 0; j < K; j++) key[j] = ((live >> j) & 1u) ? ld_in(a.keys + e0 + (int64_t)j * T) : 0ull;
        probe_direct<K, (CM & 2) != 0>(jt, key, live, grp, slot);
#pragma unroll
        for (int j = 0; j < K; j++) val[j] = ((live >> j) & 1u) ? ps_as_f64_bits(a.val_dtype, ld_in(a.vals + e0 + (int64_t)j * T)) : 0ull;
#pragma unroll
        for (int j = 0; j < K; j++) {
            if (grp[j] == EMPTY_ROW) live &= ~(1u << j);
            pid[j] = dense_width ? (int)ps_div((uint32_t)(grp[j] - (unsigned long long)dense_lo), dense_width, dense_magic)
                                 : (int)__umulhi((uint32_t)(nqe_mix64(grp[j]) >> 32), P2);
            if (!((live >> j) & 1u)) pid[j] = 0;
        }
        ps_scatter_tile<T, K, PsNoHook, (CM & 1) != 0>(out, sm, grp, val, pid, live);
    }
}

// (A software-pipelined variant -- probe rows through a ring of cp.async.bulk stages, the next tile's table reads issued
// before this tile's page stores -- was measured and is slower in every shape tried: whole operator 2.71 .. 3.62 ms
// against 2.14 ms; the smaller tiles it needs cost more cursor atomics and barriers than the overlap wins back.)
template <int T, int K, int MINB, int CM = 3>
int32_t ja_direct_scatter_launch_shape(nqe_ctx *ctx, const PsSplitArgs &a, const PagedStreams &out, const JoinTable &jt, uint32_t P2,
                                       long long dense_lo, uint32_t dense_width) {
    auto kern = ja_direct_scatter_kernel<T, K, MINB, CM>;
    const size_t smem = sizeof(PsScatterSmem<T, K>);
    NQE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (a.n + (int64_t)T * K - 1) / ((int64_t)T * K);
    int grid = ctx->sm_count * MINB;
    if (grid > tiles) grid = (int)tiles;
    if (grid < 1) return NQE_OK;
    kern<<<grid, T, smem, ctx->stream>>>(out, a, jt, P2, dense_lo, dense_width);
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}
// one shape: 256 threads x 8 rows, 4 CTAs per SM (the plain split's); 256 x 4 at 6 CTAs, 512 x 8 at 2 and 256 x 8 at 3 CTAs per
// SM measured within noise of it (2.40 .. 2.50 ms for the whole operator), as did the L2 policy combinations
int32_t ja_direct_scatter_launch(nqe_ctx *ctx, const PsSplitArgs &a, const PagedStreams &out, const JoinTable &jt, uint32_t P2,
                                 long long dense_lo, uint32_t dense_width) {
    return ja_direct_scatter_launch_shape<256, 8, 4, 3>(ctx, a, out, jt, P2, dense_lo, dense_width);
}


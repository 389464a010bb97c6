// paged_split.cuh -- "paged partition streams": one pass that splits 16-byte rows {key, value as f64} into P partitions
// without knowing the partition sizes beforehand (no count pass, no scan pass).
//
// A partition is a VIRTUAL row stream: a 64-bit append cursor per partition plus a page table that maps its 4096-row
// (64 KB) virtual pages onto pages of one shared pool.  A tile of rows is counting-sorted by partition in shared
// memory (one shared-memory atomic per row: the value it returns is the row's rank inside its partition's run), every
// run then takes its place in its partition with ONE global atomic on the cursor and leaves the CTA as one contiguous
// run of 16-byte stores.  The run that contains the first row of a virtual page maps it (pool page from a bump
// allocator, published in the page table); runs that land in the middle of a page wait for that entry -- the only
// inter-CTA dependency, always on a run that already holds a lower position.  Consumers walk a partition page by page
// (page q of partition p = pool page pt[p][q] - 1, rows = min(4096, cursor[p] - 4096 q)), each page one contiguous
// cp.async.bulk source.  Pool = rows/4096 + P pages: no slack for skew, no partially filled pages but each
// partition's last.
//
// Users: the partitioned shared-memory group-by (hash_aggregate.cu) and the fused join -> group-by (hash_join.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "expr_eval.cuh"
#include "hash_common.cuh"
#include "nqe_internal.cuh"
#include "tma_utils.cuh"

constexpr int PS_PAGE_ROWS = 4096;                  // 64 KB pages of 16-byte rows
constexpr int PS_PAGE_SHIFT = 12;
constexpr int PS_MAX_PARTS = 160;                   // five warps of partition bookkeeping (>= 148 SMs)

struct PagedStreams {
    ulonglong2 *pool;            // [max_pages][PS_PAGE_ROWS]
    unsigned long long *cursor;  // [P] rows appended to partition p so far
    unsigned int *pt;            // [P][pt_stride]: pool page + 1 of virtual page q (0: not mapped yet)
    unsigned int *pool_next;     // bump allocator of the pool
    uint32_t *status;            // DEV_ERR_CAPACITY if the pool ran out (cannot happen for max_rows rows)
    uint32_t max_pages, pt_stride;
    int32_t P, pad;
};

__device__ __forceinline__ unsigned int ps_ld_relaxed_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void ps_st_relaxed_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// shared memory of one scattering CTA (THREADS * K rows per tile, at most PS_PAGE_ROWS)
template <int THREADS, int K>
struct PsScatterSmem {
    ulonglong2 rows[THREADS * K];
    unsigned char pid[THREADS * K];
    unsigned int cnt[PS_MAX_PARTS];
    unsigned int start[PS_MAX_PARTS + 1];
    unsigned long long vpos[PS_MAX_PARTS];
    unsigned int phys_a[PS_MAX_PARTS], phys_b[PS_MAX_PARTS];
    unsigned int warp_tot[PS_MAX_PARTS / 32];
};

template <int THREADS, int K>
__device__ __forceinline__ void ps_scatter_init(PsScatterSmem<THREADS, K> &sm) {
    static_assert(THREADS * K <= PS_PAGE_ROWS, "a tile's run may span at most two pages");
    static_assert(THREADS >= PS_MAX_PARTS, "one bookkeeping thread per partition");
    if (threadIdx.x < PS_MAX_PARTS) sm.cnt[threadIdx.x] = 0;
    __syncthreads();
}

// pool page for virtual page q of partition p, mapped by the calling thread
__device__ __forceinline__ unsigned int ps_map_page(const PagedStreams &ps, int p, unsigned int q) {
    unsigned int page = atomicAdd(ps.pool_next, 1u);
    if (page >= ps.max_pages || q >= ps.pt_stride) {
        atomicOr(ps.status, DEV_ERR_CAPACITY);
        return 0u;
    }
    ps_st_relaxed_u32(ps.pt + (size_t)p * ps.pt_stride + q, page + 1u);
    return page;
}

// Append the live rows of one tile (K per thread: key[j], val[j] -> partition pid[j], bit j of `live`) to their
// partitions' streams.  CTA-wide: four barriers; cnt[] must be zero on entry (ps_scatter_init) and is zero on exit.
struct PsNoHook {
    __device__ __forceinline__ void operator()() const {}
};
// `after_reads`: called by every thread right after the tile's first barrier, i.e. once the whole CTA has its rows in
// registers -- where a caller that staged the tile in shared memory hands the buffer back to its producer.
// EF: the page stores carry an L2 evict_first policy (for callers that keep something else resident in the L2).
template <int THREADS, int K, typename Hook = PsNoHook, bool EF = false>
__device__ __forceinline__ void ps_scatter_tile(const PagedStreams &ps, PsScatterSmem<THREADS, K> &sm,
                                                const unsigned long long (&key)[K], const unsigned long long (&val)[K],
                                                const int (&pid)[K], uint32_t live, const Hook &after_reads = Hook()) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned rank[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        rank[j] = 0;
        if ((live >> j) & 1u) rank[j] = atomicAdd(&sm.cnt[pid[j]], 1u);
    }
    __syncthreads();
    after_reads();
    // local starts of the runs: exclusive scan of the counts over the partitions
    unsigned c = 0, incl = 0;
    if (tid < PS_MAX_PARTS) {
        c = tid < ps.P ? sm.cnt[tid] : 0u;
        incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) sm.warp_tot[warp] = incl;
    }
    __syncthreads();
    if (tid < PS_MAX_PARTS) {
        unsigned base = 0;
#pragma unroll
        for (int w = 0; w < PS_MAX_PARTS / 32; w++)
            if (w < warp) base += sm.warp_tot[w];
        sm.start[tid] = base + incl - c;
        if (tid == PS_MAX_PARTS - 1) sm.start[PS_MAX_PARTS] = base + incl;
        sm.cnt[tid] = 0; // ready for the next tile
    }
    __syncthreads();
    // global positions of the runs (first warps) while everybody stages its rows in run order
    if (tid < ps.P && c) {
        const unsigned long long v = atomicAdd(ps.cursor + tid, (unsigned long long)c);
        const unsigned q0 = (unsigned)(v >> PS_PAGE_SHIFT), q1 = (unsigned)((v + c - 1) >> PS_PAGE_SHIFT);
        unsigned pa = 0, pb = 0;
        bool have_a = false;
        if ((v & (PS_PAGE_ROWS - 1)) == 0) { // this run holds the first row of page q0: it maps the page
            pa = ps_map_page(ps, tid, q0);
            have_a = true;
        }
        if (q1 != q0) pb = ps_map_page(ps, tid, q1); // ... and of page q1
        if (!have_a) {
            const unsigned int *e = ps.pt + (size_t)tid * ps.pt_stride + q0;
            unsigned x;
            do { x = ps_ld_relaxed_u32(e); } while (x == 0u); // mapped by the run that holds the page's first row
            pa = x - 1u;
        }
        sm.vpos[tid] = v;
        sm.phys_a[tid] = pa;
        sm.phys_b[tid] = pb;
    }
#pragma unroll
    for (int j = 0; j < K; j++)
        if ((live >> j) & 1u) {
            const unsigned local = sm.start[pid[j]] + rank[j];
            sm.rows[local] = make_ulonglong2(key[j], val[j]);
            sm.pid[local] = (unsigned char)pid[j];
        }
    __syncthreads();
    const unsigned total = sm.start[PS_MAX_PARTS];
    const unsigned long long pol = EF ? nqe_policy_evict_first() : 0ull;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const unsigned idx = tid + j * THREADS;
        if (idx < total) {
            const int p = sm.pid[idx];
            const unsigned long long v0 = sm.vpos[p], v = v0 + (idx - sm.start[p]);
            const unsigned phys = (v >> PS_PAGE_SHIFT) == (v0 >> PS_PAGE_SHIFT) ? sm.phys_a[p] : sm.phys_b[p];
            ulonglong2 *dst = ps.pool + ((size_t)phys << PS_PAGE_SHIFT) + (v & (PS_PAGE_ROWS - 1));
            const ulonglong2 r = sm.rows[idx];
            if (EF) asm volatile("st.global.L2::cache_hint.v2.b64 [%0], {%1, %2}, %3;" ::"l"(dst), "l"(r.x), "l"(r.y), "l"(pol) : "memory");
            else *dst = r;
        }
    }
    // no barrier here: the next tile's first barrier orders these reads before the next writes of start/vpos/rows
}

// ---- splitting two 8-byte columns (key, value) into paged streams -------------------------------------------------
// Part(key) -> partition.  KEYEXPR: the key is program 0 of `prog` (rows whose key is NULL are dropped, aggregate/mod.rs:63-71),
// else the bare column `keys`.
struct PsSplitArgs {
    const unsigned long long *keys, *vals;
    int64_t n;
    int32_t val_dtype; // the value travels "as f64" (sum.rs:44 `val as f64`): Int64 / UInt64 values are converted here, once
};
__device__ __forceinline__ unsigned long long ps_as_f64_bits(int dtype, unsigned long long bits) {
    if (dtype == NQE_INT64) return (unsigned long long)__double_as_longlong(__ll2double_rn((long long)bits));
    if (dtype == NQE_UINT64) return (unsigned long long)__double_as_longlong(__ull2double_rn(bits));
    return bits;
}
struct PartByHash { // partition = high 32 bits of the key's hash, range-reduced
    uint32_t P;
    __device__ __forceinline__ int operator()(unsigned long long key) const {
        return (int)__umulhi((uint32_t)(nqe_mix64(key) >> 32), P);
    }
};
// Dense integer keys (surrogate ids, dictionary codes, `x % n`): partition = key RANGE, so that a partition's keys are
// the consecutive integers [lo + p * width, lo + (p + 1) * width) and its groups can be aggregated in a directly
// indexed table.  `range` = hi - lo + 1 as the caller believes it to be: a key outside raises DEV_ERR_RANGE (the caller
// then falls back to hashing) and is parked in partition 0, where the aggregator sends it to the global table.
// d / width for d < 2^32 without a division: magic = ceil(2^32 / width) overshoots the quotient by at most one
__host__ __device__ __forceinline__ uint32_t ps_div_magic(uint32_t width) { return (uint32_t)((0x100000000ull + width - 1) / width); }
__device__ __forceinline__ uint32_t ps_div(uint32_t d, uint32_t width, uint32_t magic) {
    uint32_t q = __umulhi(d, magic);
    q -= (unsigned long long)q * width > d ? 1u : 0u;
    return q;
}
struct PartByRange {
    long long lo;
    unsigned long long range;
    uint32_t width, magic; // magic = ps_div_magic(width)
    uint32_t *status;
    __device__ __forceinline__ int operator()(unsigned long long key) const {
        const unsigned long long d = key - (unsigned long long)lo;
        if (d >= range) {
            atomicOr(status, DEV_ERR_RANGE);
            return 0;
        }
        return (int)ps_div((uint32_t)d, width, magic); // range < 2^32 (checked on the host)
    }
};

// Tile shapes of the split kernel (knob NQE_PS_SPLIT_SHAPE): 0 (default) = 256 threads x 8 rows (2048-row tiles,
// 4 CTAs/SM), 1 = 512 x 8 (4096-row tiles, 2 CTAs/SM).  Measured, group-by of 1e8 rows into 148 partitions, whole
// operator, with the shapes that were removed again (1024 x 4 at one CTA/SM, 256 x 8 at 5 CTAs/SM, 256 x 4 at 6):
// 1.95-1.98 / 1.98 / 2.24 / 2.08 / 2.46 ms.
template <bool KEYEXPR, typename Part, int T, int K, int MINB>
__global__ void __launch_bounds__(T, MINB)
ps_split_kernel(const __grid_constant__ PagedStreams ps, const PsSplitArgs a, const Part part,
                const __grid_constant__ DevProgramSet prog, uint32_t *status) {
    constexpr int TILE = T * K;
    extern __shared__ __align__(16) unsigned char ps_smem_raw[];
    PsScatterSmem<T, K> &sm = *reinterpret_cast<PsScatterSmem<T, K> *>(ps_smem_raw);
    ps_scatter_init(sm);
    const int64_t num_tiles = (a.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE + threadIdx.x;
        unsigned long long key[K], val[K];
        int pid[K];
        uint32_t live = 0;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (e0 + (int64_t)j * T < a.n) live |= 1u << j;
        if (KEYEXPR) {
            RowRegs<K> kr;
            run_program<K>(prog, 0, e0, T, live, live, 0u, kr, status);
#pragma unroll
            for (int j = 0; j < K; j++) key[j] = kr.v[j];
            live &= kr.valid;
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) key[j] = ((live >> j) & 1u) ? ld_stream_u64(a.keys + e0 + (int64_t)j * T) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < K; j++) val[j] = ((live >> j) & 1u) ? ps_as_f64_bits(a.val_dtype, ld_stream_u64(a.vals + e0 + (int64_t)j * T)) : 0ull;
#pragma unroll
        for (int j = 0; j < K; j++) pid[j] = part(key[j]);
        ps_scatter_tile<T, K>(ps, sm, key, val, pid, live);
    }
}

// (A variant with the input tile staged by cp.async.bulk copies was measured -- NQE_PS_SPLIT_TMA in round 2 -- and gave
// nothing, 1.50 vs 1.50 ms for the whole group-by: the split is not latency-bound.  Removed.)
int nqe_ps_split_shape(); // paged_split.cu: knob NQE_PS_SPLIT_SHAPE
template <bool KEYEXPR, typename Part, int T, int K, int MINB>
static int32_t ps_split_launch_shape(nqe_ctx *ctx, const PagedStreams &ps, const PsSplitArgs &a, const Part &part,
                                     const DevProgramSet &prog, uint32_t *status) {
    auto kern = ps_split_kernel<KEYEXPR, Part, T, K, MINB>;
    const size_t smem = sizeof(PsScatterSmem<T, K>);
    NQE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (a.n + (int64_t)T * K - 1) / ((int64_t)T * K);
    int grid = ctx->sm_count * MINB;
    if (grid > tiles) grid = (int)tiles;
    if (grid < 1) return NQE_OK;
    kern<<<grid, T, smem, ctx->stream>>>(ps, a, part, prog, status);
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}
// split (key | key expression, value) into the streams `ps` on ctx->stream
template <bool KEYEXPR, typename Part>
static int32_t ps_split_launch(nqe_ctx *ctx, const PagedStreams &ps, const PsSplitArgs &a, const Part &part,
                               const DevProgramSet &prog, uint32_t *status) {
    switch (nqe_ps_split_shape()) {
    case 1: return ps_split_launch_shape<KEYEXPR, Part, 512, 8, 2>(ctx, ps, a, part, prog, status);
    default: return ps_split_launch_shape<KEYEXPR, Part, 256, 8, 4>(ctx, ps, a, part, prog, status);
    }
}

// ---- consumer side: one page at a time through a 64 KB shared-memory buffer filled by cp.async.bulk ---------------
// Thread 0 of the CTA is the elected producer: as soon as every warp has copied its rows of the current page into
// registers (empty barrier) it issues the bulk copy of the next page, which then flies while the CTA works on the rows.
struct PsPageBuf {
    unsigned long long full, empty;
};
__device__ __forceinline__ void ps_pagebuf_init(PsPageBuf &b, int warps) {
    if (threadIdx.x == 0) {
        nqe_mbar_init(&b.full, 1);
        nqe_mbar_init(&b.empty, (uint32_t)warps);
        nqe_mbar_init_fence();
    }
}
// rows of virtual page q of partition p
__device__ __forceinline__ unsigned ps_page_rows(const PagedStreams &ps, int p, unsigned q) {
    const unsigned long long total = ps.cursor[p], before = (unsigned long long)q << PS_PAGE_SHIFT;
    const unsigned long long left = total - before;
    return left < (unsigned long long)PS_PAGE_ROWS ? (unsigned)left : (unsigned)PS_PAGE_ROWS;
}
__device__ __forceinline__ void ps_issue_page(const PagedStreams &ps, PsPageBuf &b, ulonglong2 *dst, unsigned phys, unsigned rows,
                                              unsigned long long policy) {
    const uint32_t bytes = rows * 16u;
    nqe_mbar_arrive_expect_tx(&b.full, bytes);
    const ulonglong2 *src = ps.pool + ((size_t)phys << PS_PAGE_SHIFT);
    for (uint32_t off = 0; off < bytes; off += 16384u) {
        const uint32_t len = bytes - off < 16384u ? bytes - off : 16384u;
        nqe_bulk_g2s((unsigned char *)dst + off, (const unsigned char *)src + off, len, &b.full, policy);
    }
}

// host side (paged_split.cu)
int32_t nqe_ps_create(nqe_ctx *ctx, int64_t max_rows, int P, PagedStreams *ps);
void nqe_ps_destroy(nqe_ctx *ctx, PagedStreams *ps);

// hash_aggregate.cu -- PhysicalAggregatePlan::execute (aggregate/mod.rs:113-222)
// and the five AggregateOperators (aggregate/{count,sum,avg,min,max}.rs).
//
// Group-by: one open-addressing table in HBM/L2 with array-of-struct records
//   [ key | state words ... ]   (8-byte words)
// whose state words are pre-initialised to the operators' identities, so a row
// only needs (1) a probe that claims or finds its key slot with one CAS and
// (2) one fire-and-forget reduction per state word:
//   count -> u64 add; sum/avg -> f64 add (+ u64 add); min/max -> u64 min/max on
//   the OrderedFloat order-preserving encoding (hash_common.cuh), skipped when
//   a plain load shows the value cannot improve the state.
// All value columns are accumulated "as f64" exactly like the reference
// (sum.rs:44 `val as f64`), NULL keys are dropped (mod.rs:63-71), NULL values
// skipped, and the output carries no key column.
#include <atomic>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <vector>

#include "agg_device.cuh"
#include "paged_split.cuh"

int32_t nqe_utf8_key_ids(nqe_ctx *ctx, const DevColumn &dict, const DevColumn *probe, DevColumn *dict_ids, DevColumn *probe_ids);
static const char *agg_fn_name(int op);

namespace {

__global__ void agg_init_kernel(AggParams ap, uint64_t n_records) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_records) return;
    unsigned long long *rec = ap.table + (r << ap.rec_shift);
    rec[0] = EMPTY_KEY;
    for (int s = 0; s < ap.n_states; s++) {
        unsigned long long *w = rec + ap.st_off[s];
        switch (ap.st_kind[s]) {
        case ST_CNT: *w = 0; break;
        case ST_SUM: *w = 0; break; // +0.0
        case ST_MIN: *w = nqe_f64_to_ord(DBL_MAX); break;  // f64::MAX, min.rs:38
        default: *w = nqe_f64_to_ord(-DBL_MAX); break;     // f64::MIN, max.rs:38
        }
    }
}

// argument values of one input row, read from the columns registered in DevProgramSet
template <bool SIMPLE>
struct RowSource {
    const DevProgramSet &ps;
    int64_t e;
    __device__ __forceinline__ bool operator()(int slot, int *dtype, uint64_t *bits) const {
        const DevColRef &c = ps.cols[slot];
        if (!SIMPLE && c.validity && !((__ldg(c.validity + (e >> 5)) >> (e & 31)) & 1u)) return false;
        *dtype = c.dtype;
        *bits = (c.dtype == NQE_BOOL || c.dtype == NQE_UTF8) ? 0ull : ld_stream_u64((const uint64_t *)c.values + e);
        return true;
    }
};

// program 0 of ps = group key expression.  SIMPLE: the key is a bare NULL-free column and no
// argument column has a validity bitmap (the BASELINE shape): keys and values are streamed with
// plain coalesced loads, no interpreter.
template <bool SIMPLE, int AGK>
__global__ void __launch_bounds__(AG_THREADS, SIMPLE ? (AGK <= 2 ? 6 : AGK <= 4 ? 4 : 2) : 3)
group_aggregate_kernel(const __grid_constant__ DevProgramSet ps, const __grid_constant__ AggParams ap, int key_slot) {
    constexpr int TILE = AGK * AG_THREADS;
    const int64_t num_tiles = (ap.n_rows + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE + threadIdx.x;
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < AGK; j++)
            if (e0 + (int64_t)j * AG_THREADS < ap.n_rows) inrange |= 1u << j;
        RowRegs<AGK> key;
        if (SIMPLE) {
            const uint64_t *kc = (const uint64_t *)ps.cols[key_slot].values;
#pragma unroll
            for (int j = 0; j < AGK; j++) key.v[j] = ((inrange >> j) & 1u) ? ld_stream_u64(kc + e0 + (int64_t)j * AG_THREADS) : 0ull;
            key.valid = inrange;
        } else {
            run_program<AGK>(ps, 0, e0, AG_THREADS, inrange, inrange, 0u, key, ap.status);
        }
        unsigned long long *rec[AGK];
        Sector0 s0[AGK];
        find_slots<AGK>(ap, key.v, key.valid, rec, s0); // NULL keys are dropped (valid bit clear)
#pragma unroll
        for (int j = 0; j < AGK; j++)
            if (rec[j]) update_states(ap, rec[j], s0[j], RowSource<SIMPLE>{ps, e0 + (int64_t)j * AG_THREADS});
    }
}

// ---------------------------------------------------------------------------
// Partitioned shared-memory group-by.
//
// The table path above costs one L2 sector read plus one `red` per additive state per ROW, and the L2 retires only
// ~1.5e11 of those per second (profiles/README_r01.md 2): 2.25 ms per 1e8 rows whatever the kernel does.  Shared
// memory takes ~20x that rate, but holds only a few thousand groups per SM.  So the (key, value as f64) rows are first
// split by key hash into P partitions of <= ~700 groups (paged_split.cuh: one pass, no count pass), and every
// partition is then aggregated by sm_count / P CTAs, each in its own 3072-slot table in SHARED memory:
//      keys[slot]                     two-slot buckets, a key lives in one of TWO buckets (four candidate slots)
//      filt[slot] = high halves of the current {min, max} in the ordered encoding
//      sum[slot]  = f64 sum           cnt[slot] = u32 count           mm[slot] = {min, max}, full 64-bit
// SIMT shapes the table: with open addressing a warp walks as many probe steps as its unluckiest lane (2.7 extra
// steps per 32 rows at load 0.33, measured), so the lookup is made branch-free instead -- both buckets are read with
// two unconditional 16-byte loads and the slot is selected with compares; only the first row of a group per CTA
// takes the (divergent) insert path.
// Data movement: every WARP owns a 2 KB stage; its lanes copy their four rows into registers, lane 0 at once issues
// the cp.async.bulk of the warp's next 128 rows (one 2 KB slice of the CTA's next page), which flies while the warp
// updates the table -- no CTA-wide barrier anywhere in the main loop, warps drift freely.  Per row: the two bucket
// reads, a native 32-bit shared atomic for the count, an 8-byte read + one 64-bit CAS for the sum (shared memory has
// no f64 add; the read is issued right before the CAS so that retries are rare), an 8-byte read of the filter and
// NOTHING more for min/max unless the value's high half reaches the filter -- then the exact 64-bit compare + CAS run
// on mm[] and the filter is tightened.  The four rows of a lane are kept in lock step so that their shared-memory
// round trips overlap, and the warp is re-converged explicitly after every data-dependent loop (without
// __syncwarp() the compiler lets the stragglers run the rest of the loop body on their own: measured 1.64 -> 1.38 ms).
// At the end every CTA folds its partial states into the global table (one probe + one `red` per state per GROUP).
// Rows that do not fit (all four candidate slots taken by other keys; the key i64::MIN) go to the global table one by
// one, so any key distribution stays correct.
// Eligible: all aggregates over ONE NULL-free 8-byte column; the key may be any Int64/UInt64 expression.
constexpr int GA_THREADS = 1024, GA_WARPS = GA_THREADS / 32, GA_K = 4;
constexpr int GA_SLOTS = 3072, GA_BUCKETS = GA_SLOTS / 2;
constexpr int GA_CHUNK = 32 * GA_K;                 // rows per warp and stage
constexpr unsigned GA_NONE = 0xFFFFFFFFu;
static_assert(GA_WARPS * GA_CHUNK == PS_PAGE_ROWS, "a page is one chunk per warp");
struct GaDense {
    long long lo;      // smallest key of partition 0
    unsigned width;    // keys per partition (<= GA_SLOTS)
    unsigned pad;
};
struct GaSmem {
    ulonglong2 stage[GA_WARPS][GA_CHUNK];           // 64 KB
    ulonglong2 keys2[GA_BUCKETS];                   // keys[2 b], keys[2 b + 1]
    ulonglong2 mm[GA_SLOTS];
    ulonglong2 sf[GA_SLOTS];                        // x: f64 sum; y: filter, low half min_hi (>= high half of the true min), high half max_hi (<=)
    unsigned int cnt[GA_SLOTS];
    unsigned long long full[GA_WARPS];
};
constexpr size_t GA_SMEM = sizeof(GaSmem);

// one row straight into the global table (keys that cannot live in the shared table); the value is already an f64
__device__ __noinline__ void gp_global_row(const AggParams &ap, uint64_t key, uint64_t f64bits) {
    struct Src {
        uint64_t bits;
        __device__ __forceinline__ bool operator()(int, int *dt, uint64_t *b) const { *dt = NQE_FLOAT64; *b = bits; return true; }
    };
    Sector0 s0;
    unsigned long long *rec = find_slot(ap, key, &s0);
    if (rec) update_states(ap, rec, s0, Src{f64bits});
}

// the two buckets of a key inside a partition: multiply-shift on the folded key (the partition was chosen by the
// murmur-style mix of the same key, so the three are independent enough)
__device__ __forceinline__ void ga_buckets_of(unsigned long long key, unsigned *b1, unsigned *b2) {
    const unsigned f = (unsigned)key ^ (unsigned)(key >> 32) * 0x85EBCA6Bu;
    const unsigned x = __umulhi(f * 0x9E3779B1u, (unsigned)GA_BUCKETS), y = __umulhi(f * 0xC2B2AE35u + 0x27D4EB2Fu, (unsigned)GA_BUCKETS - 1u);
    *b1 = x;
    *b2 = y + (y >= x); // one of the other GA_BUCKETS - 1 buckets
}

// claim a slot for `key` among its four candidates (first free one in a fixed order, so that concurrent inserters of
// the same key agree); GA_NONE: all four belong to other keys
__device__ __noinline__ unsigned ga_insert(unsigned long long *keys, unsigned long long key) {
    unsigned b1, b2;
    ga_buckets_of(key, &b1, &b2);
    for (int c = 0; c < 4; c++) {
        const unsigned slot = (c < 2 ? 2 * b1 : 2 * b2) + (c & 1);
        unsigned long long cur = *(volatile unsigned long long *)&keys[slot];
        if (cur == EMPTY_KEY) {
            cur = atomicCAS(&keys[slot], (unsigned long long)EMPTY_KEY, key);
            if (cur == EMPTY_KEY) return slot;
        }
        if (cur == key) return slot;
    }
    return GA_NONE;
}

// need: bit ST_CNT / ST_SUM / ST_MIN / ST_MAX set when the plan has such a state.
// DENSE: the keys of partition p are the integers [dense.lo + p * dense.width, + dense.width) (the split was by key
// RANGE, PartByRange): the slot of a key is its offset in that range -- no key array, no lookup, no inserts.
template <bool DENSE>
__global__ void __launch_bounds__(GA_THREADS, 1)
gp2_aggregate_kernel(const __grid_constant__ PagedStreams st, const __grid_constant__ AggParams ap, int m, int need,
                     const GaDense dense) {
    extern __shared__ __align__(128) unsigned char ga_smem[];
    GaSmem &sm = *reinterpret_cast<GaSmem *>(ga_smem);
    unsigned long long *const keys = (unsigned long long *)sm.keys2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p = blockIdx.x / m, sub = blockIdx.x % m;
    const unsigned long long rows_p = st.cursor[p];
    const unsigned npg = (unsigned)((rows_p + PS_PAGE_ROWS - 1) >> PS_PAGE_SHIFT);
    const unsigned last_fill = (unsigned)(rows_p - ((unsigned long long)(npg ? npg - 1 : 0) << PS_PAGE_SHIFT)); // rows of the last page
    const unsigned long long dense_base = (unsigned long long)dense.lo + (unsigned long long)p * dense.width; // key of slot 0
    // DENSE: a partition has only `width` groups but ~4096 rows in flight per CTA, i.e. several concurrent updates per
    // group (measured: ~40 % of the sum CASes lose).  The table is therefore replicated GA_SLOTS / width times and the
    // warps spread over the copies; the copies are merged in the flush.
    const unsigned copies = DENSE ? (dense.width ? (unsigned)GA_SLOTS / dense.width : 1u) : 1u;
    const unsigned copy_off = DENSE ? ((unsigned)warp % copies) * dense.width : 0u;
    for (int i = tid; i < GA_SLOTS; i += GA_THREADS) {
        if (!DENSE) keys[i] = EMPTY_KEY;
        sm.mm[i] = make_ulonglong2(~0ull, 0ull);   // identities of min / max in the ordered encoding
        sm.sf[i] = make_ulonglong2(0ull, 0x00000000FFFFFFFFull); // sum = +0.0; min_hi = ~0, max_hi = 0: everything passes
        sm.cnt[i] = 0u;
    }
    if (lane == 0) nqe_mbar_init(&sm.full[warp], 1);
    if (tid == 0) nqe_mbar_init_fence();
    __syncthreads();
    const unsigned int *pt = st.pt + (size_t)p * st.pt_stride;
    // rows of page q that belong to this warp: [128 warp, 128 warp + 128) clipped to the page's fill
    auto chunk_rows = [&](unsigned q) -> unsigned {
        const unsigned fill = q + 1 == npg ? last_fill : (unsigned)PS_PAGE_ROWS, lo = warp * GA_CHUNK;
        return fill <= lo ? 0u : (fill - lo < (unsigned)GA_CHUNK ? fill - lo : (unsigned)GA_CHUNK);
    };
    auto issue = [&](unsigned q, unsigned pte) { // lane 0; pte = page table entry (pool page + 1)
        const unsigned rows = chunk_rows(q);
        if (!rows) return;
        nqe_mbar_arrive_expect_tx(&sm.full[warp], rows * 16u);
        nqe_bulk_g2s(sm.stage[warp], st.pool + ((size_t)(pte - 1u) << PS_PAGE_SHIFT) + warp * GA_CHUNK, rows * 16u, &sm.full[warp],
                     nqe_policy_evict_first());
    };
    unsigned next_pte = 0; // lane 0: loaded one iteration before it is used, so that nobody waits for the L2 round trip
    if (lane == 0) {
        if ((unsigned)sub < npg) issue(sub, pt[sub]);
        if ((unsigned)(sub + m) < npg) next_pte = pt[sub + m];
    }
    uint32_t it = 0; // stages this warp has consumed
    for (unsigned q = sub; q < npg; q += m) {
        const unsigned rows = chunk_rows(q);
        unsigned long long key[GA_K], bits[GA_K];
        uint32_t live = 0;
        if (rows) {
            nqe_mbar_wait(&sm.full[warp], it & 1u);
            it++;
#pragma unroll
            for (int j = 0; j < GA_K; j++) {
                const unsigned idx = j * 32 + lane;
                ulonglong2 r = make_ulonglong2(EMPTY_KEY, 0ull);
                if (idx < rows) { r = sm.stage[warp][idx]; live |= 1u << j; }
                key[j] = r.x;
                bits[j] = r.y;
            }
        }
        __syncwarp(); // the stage is in registers: refill it while the warp works
        if (lane == 0 && q + m < npg) {
            issue(q + m, next_pte);
            if (q + 2 * m < npg) next_pte = pt[q + 2 * m];
        }
        if (!rows) continue;
        unsigned slot[GA_K];
        if (DENSE) {
#pragma unroll
            for (int j = 0; j < GA_K; j++) {
                const unsigned long long d = key[j] - dense_base;
                slot[j] = ((live >> j) & 1u) && d < (unsigned long long)dense.width ? (unsigned)d + copy_off : GA_NONE; // outside: global table
            }
        } else {
            // ---- branch-free lookup: both buckets of a row, then compares (two rows at a time: registers)
            uint32_t miss = 0;
    #pragma unroll
            for (int h = 0; h < GA_K; h += 2) {
                ulonglong2 A[2], B[2];
                unsigned b1[2], b2[2];
    #pragma unroll
                for (int j = h; j < h + 2; j++) {
                    ga_buckets_of(key[j], &b1[j - h], &b2[j - h]);
                    A[j - h] = sm.keys2[b1[j - h]];
                }
    #pragma unroll
                for (int j = h; j < h + 2; j++) { // the second bucket only if some lane needs it (most keys sit in their first)
                    const bool inA = A[j - h].x == key[j] || A[j - h].y == key[j] || key[j] == EMPTY_KEY;
                    B[j - h] = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
                    if (__any_sync(0xffffffffu, !inA)) B[j - h] = sm.keys2[b2[j - h]];
                }
    #pragma unroll
                for (int j = h; j < h + 2; j++) {
                    const unsigned long long k = key[j];
                    const ulonglong2 a = A[j - h], b = B[j - h];
                    const unsigned s1 = 2 * b1[j - h], s2 = 2 * b2[j - h];
                    unsigned s = GA_NONE; // selects, not branches
                    s = b.y == k ? s2 + 1 : s;
                    s = b.x == k ? s2 : s;
                    s = a.y == k ? s1 + 1 : s;
                    s = a.x == k ? s1 : s;
                    s = k == EMPTY_KEY ? GA_NONE : s; // i64::MIN marks free slots (and dead lanes): that key lives in the global table only
                    slot[j] = s;
                    miss |= (k != EMPTY_KEY && s == GA_NONE) ? 1u << j : 0u;
                }
            }
            if (miss) { // first row of a group in this CTA (or a row that raced with it)
    #pragma unroll
                for (int j = 0; j < GA_K; j++)
                    if ((miss >> j) & 1u) slot[j] = ga_insert(keys, key[j]);
            }
            __syncwarp();
        }
        uint32_t found = 0;
#pragma unroll
        for (int j = 0; j < GA_K; j++)
            if (slot[j] != GA_NONE) found |= 1u << j;
        const uint32_t lost = live & ~found; // no room in the shared table / i64::MIN key: straight to the global table
        if (lost) {
#pragma unroll
            for (int j = 0; j < GA_K; j++)
                if ((lost >> j) & 1u) gp_global_row(ap, key[j], bits[j]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < GA_K; j++)
            if (!((found >> j) & 1u)) slot[j] = 0; // harmless address for the unconditional reads below
        // ---- updates, the four rows in lock step (key[j] is dead from here on)
        if (DENSE || (need & (1 << ST_CNT))) { // the dense table has no key array: a non-zero count marks a slot as used
#pragma unroll
            for (int j = 0; j < GA_K; j++)
                if ((found >> j) & 1u) atomicAdd(&sm.cnt[slot[j]], 1u);
        }
        unsigned long long flt[GA_K];
        if (need & (1 << ST_SUM)) {
            unsigned long long seen[GA_K], old[GA_K];
#pragma unroll
            for (int j = 0; j < GA_K; j++) {
                const ulonglong2 r = sm.sf[slot[j]]; // the sum right before its CAS (retries stay rare) and the min/max filter
                seen[j] = r.x;
                flt[j] = r.y;
            }
#pragma unroll
            for (int j = 0; j < GA_K; j++) {
                old[j] = seen[j];
                if ((found >> j) & 1u)
                    old[j] = atomicCAS(&sm.sf[slot[j]].x, seen[j],
                                       (unsigned long long)__double_as_longlong(__longlong_as_double((long long)seen[j]) +
                                                                                __longlong_as_double((long long)bits[j])));
            }
#pragma unroll
            for (int j = 0; j < GA_K; j++) {
#pragma unroll 1
                while (old[j] != seen[j]) { // another row of this group got in between: retry on the value it left
                    seen[j] = old[j];
                    old[j] = atomicCAS(&sm.sf[slot[j]].x, seen[j],
                                       (unsigned long long)__double_as_longlong(__longlong_as_double((long long)seen[j]) +
                                                                                __longlong_as_double((long long)bits[j])));
                }
                __syncwarp();
            }
        }
        if (need & ((1 << ST_MIN) | (1 << ST_MAX))) {
            uint32_t maybe = 0;
            if (!(need & (1 << ST_SUM))) {
#pragma unroll
                for (int j = 0; j < GA_K; j++) flt[j] = sm.sf[slot[j]].y;
            }
#pragma unroll
            for (int j = 0; j < GA_K; j++) {
                // high half of the ordered encoding (hash_common.cuh) -- NaN encodes as all ones
                const int hi = (int)(bits[j] >> 32);
                unsigned ohi = hi < 0 ? ~(unsigned)hi : (unsigned)hi | 0x80000000u;
                const double v = __longlong_as_double((long long)bits[j]);
                if (v != v) ohi = 0xFFFFFFFFu;
                if (((found >> j) & 1u) && (ohi <= (unsigned)flt[j] || ohi >= (unsigned)(flt[j] >> 32))) maybe |= 1u << j;
            }
            if (maybe) {
#pragma unroll
                for (int j = 0; j < GA_K; j++) {
                    if (!((maybe >> j) & 1u)) continue;
                    const double v = __longlong_as_double((long long)bits[j]);
                    const unsigned long long o = nqe_f64_to_ord(v);
                    const ulonglong2 cur = sm.mm[slot[j]];
                    unsigned int *filt = (unsigned int *)&sm.sf[slot[j]].y;
                    if ((need & (1 << ST_MAX)) && o > cur.y) {
                        atomicMax(&sm.mm[slot[j]].y, o);
                        atomicMax(filt + 1, (unsigned)(o >> 32));
                    }
                    if ((need & (1 << ST_MIN)) && v == v && o < cur.x) { // min never picks NaN (min.rs:49)
                        atomicMin(&sm.mm[slot[j]].x, o);
                        atomicMin(filt, (unsigned)(o >> 32));
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // partial states of this CTA -> global table
    if (DENSE) {
        for (unsigned i = tid; i < dense.width; i += GA_THREADS) {
            unsigned long long c = 0, mn = ~0ull, mx = 0ull;
            double sum = 0.0;
            for (unsigned k = 0; k < copies; k++) { // merge the copies
                const unsigned sl = k * dense.width + i;
                const unsigned ck = sm.cnt[sl];
                if (!ck) continue;
                c += ck;
                sum += __longlong_as_double((long long)sm.sf[sl].x);
                const ulonglong2 x = sm.mm[sl];
                mn = x.x < mn ? x.x : mn;
                mx = x.y > mx ? x.y : mx;
            }
            if (!c) continue;
            Sector0 s0;
            unsigned long long *rec = find_slot(ap, dense_base + (unsigned long long)i, &s0);
            if (!rec) continue; // table full: flagged, the host grows the table and repeats this kernel
            for (int s = 0; s < ap.n_states; s++) {
                unsigned long long *w = rec + ap.st_off[s];
                switch (ap.st_kind[s]) {
                case ST_CNT: red_add_u64(w, c); break;
                case ST_SUM: red_add_f64(w, sum); break;
                case ST_MIN: red_min_u64(w, mn); break;
                default: red_max_u64(w, mx); break;
                }
            }
        }
        return;
    }
    for (int i = tid; i < GA_SLOTS; i += GA_THREADS) {
        const unsigned long long key = keys[i];
        if (key == EMPTY_KEY) continue;
        Sector0 s0;
        unsigned long long *rec = find_slot(ap, key, &s0);
        if (!rec) continue; // table full: flagged, the host grows the table and repeats this kernel
        const ulonglong2 x = sm.mm[i];
        for (int s = 0; s < ap.n_states; s++) {
            unsigned long long *w = rec + ap.st_off[s];
            switch (ap.st_kind[s]) {
            case ST_CNT: red_add_u64(w, (unsigned long long)sm.cnt[i]); break;
            case ST_SUM: red_add_f64(w, __longlong_as_double((long long)sm.sf[i].x)); break;
            case ST_MIN: red_min_u64(w, x.x); break;
            default: red_max_u64(w, x.y); break;
            }
        }
    }
}

// Global (no GROUP BY) path, mod.rs:123-139: per-thread partial states, warp
// shuffle reduction, then one reduction per warp into the single record.
__global__ void __launch_bounds__(AG_THREADS)
global_aggregate_kernel(const __grid_constant__ DevProgramSet ps, const __grid_constant__ AggParams ap) {
    const int lane = threadIdx.x & 31;
    for (int s = 0; s < ap.n_states; s++) {
        const DevColRef &c = ps.cols[ap.st_src[s]];
        const int kind = ap.st_kind[s];
        double sum = 0.0;
        unsigned long long cnt = 0;
        unsigned long long ext = kind == ST_MIN ? nqe_f64_to_ord(DBL_MAX) : nqe_f64_to_ord(-DBL_MAX);
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ap.n_rows; e += (int64_t)gridDim.x * blockDim.x) {
            if (c.validity && !((__ldg(c.validity + (e >> 5)) >> (e & 31)) & 1u)) continue;
            cnt++;
            if (kind == ST_CNT) continue;
            const double v = value_as_f64(c.dtype, ld_cached_u64((const uint64_t *)c.values + e));
            if (kind == ST_SUM) sum += v;
            else if (kind == ST_MAX) { const unsigned long long k = nqe_f64_to_ord(v); if (k > ext) ext = k; }
            else if (v == v) { const unsigned long long k = nqe_f64_to_ord(v); if (k < ext) ext = k; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, ext, o);
            if (kind == ST_MIN ? other < ext : other > ext) ext = other;
        }
        if (lane == 0) {
            unsigned long long *w = ap.table + ap.st_off[s];
            if (kind == ST_CNT) red_add_u64(w, cnt);
            else if (kind == ST_SUM) red_add_f64(w, sum);
            else if (kind == ST_MAX) red_max_u64(w, ext);
            else red_min_u64(w, ext);
        }
    }
}

struct ExtractParams {
    void *out[AG_MAX];
    unsigned long long *out_count;
    int32_t is_global;
};

// occupied records -> dense output rows (order unspecified, as in the reference)
__global__ void agg_extract_kernel(AggParams ap, ExtractParams xp, uint64_t n_records) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool occ = false;
    const unsigned long long *rec = nullptr;
    if (r < n_records) {
        rec = ap.table + (r << ap.rec_shift);
        if (xp.is_global) occ = true;
        else if (r == n_records - 1) {
            // the record of key == i64::MIN is occupied iff any state moved off its identity;
            // word rec[0] is set to a non-EMPTY marker by the first updater (see mark below)
            occ = rec[0] != EMPTY_KEY;
        } else occ = rec[0] != EMPTY_KEY;
    }
    const unsigned m = __ballot_sync(0xffffffffu, occ);
    unsigned long long base = 0;
    if (lane == 0 && m) base = atomicAdd(xp.out_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!occ) return;
    const unsigned long long row = base + __popc(m & ((1u << lane) - 1u));
    for (int a = 0; a < ap.n_aggs; a++) {
        const unsigned long long w = ap.agg_op[a] == NQE_AGG_GROUP_KEY ? 0ull : rec[ap.st_off[ap.agg_state[a]]];
        switch (ap.agg_op[a]) {
        case NQE_AGG_GROUP_KEY: ((unsigned long long *)xp.out[a])[row] = r == n_records - 1 ? EMPTY_KEY : rec[0]; break; // last record: the key i64::MIN
        case NQE_AGG_COUNT: ((unsigned long long *)xp.out[a])[row] = w; break;
        case NQE_AGG_SUM: ((unsigned long long *)xp.out[a])[row] = w; break;
        case NQE_AGG_AVG: { // avg.rs:118: sum / cnt as f64, cnt is u32 (wraps in release builds)
            const unsigned long long c = rec[ap.st_off[ap.agg_state2[a]]];
            ((double *)xp.out[a])[row] = __longlong_as_double((long long)w) / (double)(uint32_t)c;
            break;
        }
        default: ((double *)xp.out[a])[row] = nqe_ord_to_f64(w); break;
        }
    }
}

// cardinality estimate by linear counting over a strided sample
// ... and a 256-bin histogram of the sampled keys over the hash bits the partitioned path splits on: a bin far above the
// mean is a hot key (all its rows land in one partition, i.e. on one SM)
constexpr int AGG_SKEW_BINS = 256;
__global__ void agg_sample_kernel(const __grid_constant__ DevProgramSet ps, int64_t n_rows, int64_t stride,
                                  int64_t n_sample, uint32_t *bitmap, uint32_t bits_mask, uint32_t *status, unsigned int *bins,
                                  long long *minmax) {
    __shared__ unsigned int s_bins[AGG_SKEW_BINS];
    __shared__ long long s_min[8], s_max[8];
    long long kmin = LLONG_MAX, kmax = LLONG_MIN;
    for (int b = threadIdx.x; b < AGG_SKEW_BINS; b += blockDim.x) s_bins[b] = 0;
    __syncthreads();
    // grid-stride over the sample: a few hundred CTAs, so that the per-CTA histogram / min / max reach global memory
    // with a few thousand atomics instead of one set per 256 samples (that was 0.11 ms of same-address atomics)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sample; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i * stride;
        RowRegs<1> key;
        run_program<1>(ps, 0, e, 1, e < n_rows ? 1u : 0u, 0u, 0u, key, status);
        if (key.valid & 1u) {
            const uint64_t hh = nqe_mix64(key.v[0]);
            const uint32_t h = (uint32_t)(hh >> 32) & bits_mask;
            atomicOr(bitmap + (h >> 5), 1u << (h & 31));
            atomicAdd(&s_bins[hh >> 56], 1u);
            const long long k = (long long)key.v[0];
            kmin = k < kmin ? k : kmin;
            kmax = k > kmax ? k : kmax;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, kmin, o), b = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = kmin;
        s_max[threadIdx.x >> 5] = kmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
            kmin = s_min[w] < kmin ? s_min[w] : kmin;
            kmax = s_max[w] > kmax ? s_max[w] : kmax;
        }
        if (kmin <= kmax) {
            atomicMin(minmax, kmin);
            atomicMax(minmax + 1, kmax);
        }
    }
    for (int b = threadIdx.x; b < AGG_SKEW_BINS; b += blockDim.x)
        if (s_bins[b]) atomicAdd(bins + b, s_bins[b]);
}
__global__ void popcount_kernel(const uint32_t *bitmap, int n_words, unsigned long long *out) {
    unsigned int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) c += __popc(bitmap[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// distinct values of an 8-byte column, estimated by linear counting over a strided sample
__global__ void distinct_sample_kernel(const unsigned long long *col, int64_t stride, int64_t n_sample, uint32_t *bitmap, uint32_t bits_mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_sample) return;
    const uint32_t h = (uint32_t)(nqe_mix64(col[i * stride]) >> 32) & bits_mask;
    atomicOr(bitmap + (h >> 5), 1u << (h & 31));
}

// smallest / largest value of an 8-byte column in signed order
__global__ void minmax_i64_kernel(const long long *col, int64_t n, long long *out) {
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const long long v = col[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

} // namespace

int nqe_dense_parts(nqe_ctx *ctx, unsigned long long range) {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("NQE_DENSE_PARTS");
        v = e ? atoi(e) : 0;
    }
    if (v > 0 && v <= ctx->sm_count) return v;
    const unsigned long long least = (range + NQE_GP2_DENSE_MAX_WIDTH - 1) / NQE_GP2_DENSE_MAX_WIDTH;
    int m = least ? (int)((unsigned long long)ctx->sm_count / least) : 4;
    if (m > 4) m = 4;
    if (m < 1) m = 1;
    return ctx->sm_count / m;
}

int32_t nqe_minmax_i64(nqe_ctx *ctx, const unsigned long long *col, int64_t n, long long *lo, long long *hi) {
    long long mm[4];
    NQE_TRY(nqe_minmax2_i64(ctx, col, nullptr, n, mm));
    *lo = mm[0];
    *hi = mm[1];
    return NQE_OK;
}

// signed min / max of one or two 8-byte columns of the same length with ONE host synchronisation: mm = {lo a, hi a, lo b, hi b}
int32_t nqe_minmax2_i64(nqe_ctx *ctx, const unsigned long long *a, const unsigned long long *b, int64_t n, long long *mm) {
    mm[0] = mm[2] = 0;
    mm[1] = mm[3] = -1;
    if (n <= 0) return NQE_OK;
    long long *d = (long long *)(ctx->d_scratch + 10);
    const long long init[4] = {LLONG_MAX, LLONG_MIN, LLONG_MAX, LLONG_MIN};
    NQE_CUDA(ctx, cudaMemcpyAsync(d, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    int grid = ctx->sm_count * 8;
    if ((int64_t)grid * 256 > n) grid = (int)((n + 255) / 256);
    minmax_i64_kernel<<<grid, 256, 0, ctx->stream>>>((const long long *)a, n, d);
    ctx->launches++;
    if (b) {
        minmax_i64_kernel<<<grid, 256, 0, ctx->stream>>>((const long long *)b, n, d + 2);
        ctx->launches++;
    }
    long long h[4];
    NQE_CUDA(ctx, cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    NQE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < (b ? 4 : 2); i++) mm[i] = h[i];
    return NQE_OK;
}

static double linear_count(uint64_t ones, uint32_t bits, int64_t n, int64_t n_sample) {
    const double m = (double)bits, z = m - (double)ones;
    const double distinct = z > 0 ? -m * log(z / m) : m * 16;
    double est = distinct;
    if (distinct > 0.5 * (double)n_sample) est = distinct * ((double)n / (double)n_sample); // still growing
    return est > (double)n ? (double)n : est;
}

int32_t nqe_estimate_distinct_u64(nqe_ctx *ctx, const unsigned long long *col, int64_t n, double *est) {
    *est = 0;
    if (n <= 0) return NQE_OK;
    const int64_t n_sample = n < (1 << 20) ? n : (1 << 20), stride = n / n_sample;
    const uint32_t bits = 1u << 23;
    void *bm = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &bm, bits / 8));
    cudaMemsetAsync(bm, 0, bits / 8, ctx->stream);
    cudaMemsetAsync(ctx->d_scratch + 8, 0, sizeof(uint64_t), ctx->stream);
    distinct_sample_kernel<<<(unsigned)((n_sample + 255) / 256), 256, 0, ctx->stream>>>(col, stride, n_sample, (uint32_t *)bm, bits - 1);
    popcount_kernel<<<64, 256, 0, ctx->stream>>>((const uint32_t *)bm, bits / 32, (unsigned long long *)(ctx->d_scratch + 8));
    ctx->launches += 2;
    cudaMemcpyAsync(ctx->h_scratch + 8, ctx->d_scratch + 8, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e = cudaStreamSynchronize(ctx->stream);
    nqe_dev_free(ctx, bm);
    if (e != cudaSuccess) return nqe_fail(ctx, NQE_ERR_CUDA, "distinct estimate failed: %s", cudaGetErrorString(e));
    *est = linear_count(ctx->h_scratch[8], bits, n, n_sample);
    return NQE_OK;
}

// partitions for the shared-memory group-by: <= 0.3 * GA_SLOTS expected groups per partition, P * m CTAs, one per SM
bool nqe_gp2_plan(nqe_ctx *ctx, double est_groups, int *P, int *m) {
    const int want = (int)(est_groups * 1.15 / (0.3 * GA_SLOTS)) + 1;
    if (want > ctx->sm_count || want > PS_MAX_PARTS) return false;
    *m = ctx->sm_count / want;
    *P = ctx->sm_count / *m;
    if (*P > PS_MAX_PARTS) *P = PS_MAX_PARTS;
    return true;
}

int32_t nqe_gp2_aggregate(nqe_ctx *ctx, const PagedStreams &streams, const AggParams &ap, int m, int need, long long dense_lo,
                          unsigned dense_width) {
    const GaDense dense{dense_lo, dense_width, 0u};
    if (dense_width) {
        NQE_CUDA(ctx, cudaFuncSetAttribute(gp2_aggregate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GA_SMEM));
        gp2_aggregate_kernel<true><<<streams.P * m, GA_THREADS, GA_SMEM, ctx->stream>>>(streams, ap, m, need, dense);
    } else {
        NQE_CUDA(ctx, cudaFuncSetAttribute(gp2_aggregate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GA_SMEM));
        gp2_aggregate_kernel<false><<<streams.P * m, GA_THREADS, GA_SMEM, ctx->stream>>>(streams, ap, m, need, dense);
    }
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}

static const char *agg_fn_name(int op) {
    static const char *n[] = {"Count", "Sum", "Avg", "min", "Max"};
    return n[op];
}
static const char *dtype_name2(int d) {
    switch (d) {
    case NQE_BOOL: return "Boolean"; case NQE_INT64: return "Int64"; case NQE_UINT64: return "UInt64";
    case NQE_FLOAT64: return "Float64"; default: return "Utf8";
    }
}

// validate the aggregates against their argument dtypes and lay out the (shared) state words
int32_t nqe_agg_layout(nqe_ctx *ctx, const nqe_agg *aggs, int32_t n_aggs, const int32_t *col_dtypes,
                       const int32_t *src_ids, bool grouped, AggParams *ap) {
    if (n_aggs < 0 || n_aggs > AG_MAX) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "at most %d aggregates per plan", AG_MAX);
    ap->n_aggs = n_aggs;
    ap->n_states = 0;
    auto state_of = [&](int kind, int src) {
        for (int s = 0; s < ap->n_states; s++)
            if (ap->st_kind[s] == kind && ap->st_src[s] == src) return s;
        ap->st_kind[ap->n_states] = kind;
        ap->st_src[ap->n_states] = src;
        return ap->n_states++;
    };
    for (int a = 0; a < n_aggs; a++) {
        const int op = aggs[a].op;
        if (op < NQE_AGG_COUNT || op > NQE_AGG_GROUP_KEY) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "bad aggregate op %d", op);
        if (op == NQE_AGG_GROUP_KEY) { // extension: the group key as an output column; no state
            if (!grouped) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "NQE_AGG_GROUP_KEY needs a GROUP BY");
            ap->agg_op[a] = op;
            ap->agg_state[a] = ap->agg_state2[a] = 0;
            continue;
        }
        const int dt = col_dtypes[a];
        if (op != NQE_AGG_COUNT && (dt == NQE_BOOL || dt == NQE_UTF8)) {
            // update_batch: Err(NotSupported) (sum.rs:91-96); update(row): unimplemented!() (sum.rs:108)
            return nqe_fail(ctx, grouped ? NQE_ERR_PANIC : NQE_ERR_NOT_SUPPORTED, "%s func for %s is not supported",
                            agg_fn_name(op), dtype_name2(dt));
        }
        ap->agg_op[a] = op;
        ap->agg_state2[a] = 0;
        switch (op) {
        case NQE_AGG_COUNT: ap->agg_state[a] = state_of(ST_CNT, src_ids[a]); break;
        case NQE_AGG_SUM: ap->agg_state[a] = state_of(ST_SUM, src_ids[a]); break;
        case NQE_AGG_AVG:
            ap->agg_state[a] = state_of(ST_SUM, src_ids[a]);
            ap->agg_state2[a] = state_of(ST_CNT, src_ids[a]);
            break;
        case NQE_AGG_MIN: ap->agg_state[a] = state_of(ST_MIN, src_ids[a]); break;
        default: ap->agg_state[a] = state_of(ST_MAX, src_ids[a]); break;
        }
    }
    // sort states by source id (stable) so that update_states fetches each column once
    int order[AG_MAXS], inv[AG_MAXS];
    for (int s = 0; s < ap->n_states; s++) order[s] = s;
    std::stable_sort(order, order + ap->n_states, [&](int x, int y) { return ap->st_src[x] < ap->st_src[y]; });
    int kind2[AG_MAXS], src2[AG_MAXS];
    for (int i = 0; i < ap->n_states; i++) {
        kind2[i] = ap->st_kind[order[i]];
        src2[i] = ap->st_src[order[i]];
        inv[order[i]] = i;
    }
    // word offsets: MIN/MAX states first, so that they share sector 0 (the first 32 bytes) with the key
    int next_off = 1;
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < ap->n_states; i++) {
            const bool ext = kind2[i] == ST_MIN || kind2[i] == ST_MAX;
            if (ext == (pass == 0)) ap->st_off[i] = next_off++;
        }
    for (int i = 0; i < ap->n_states; i++) {
        ap->st_kind[i] = kind2[i];
        ap->st_src[i] = src2[i];
    }
    for (int a = 0; a < n_aggs; a++) {
        if (ap->agg_op[a] == NQE_AGG_GROUP_KEY) continue;
        ap->agg_state[a] = inv[ap->agg_state[a]];
        ap->agg_state2[a] = inv[ap->agg_state2[a]];
    }
    ap->rec_words = 4; // power of two >= 32 bytes: a record never straddles more sectors than needed
    ap->rec_shift = 2;
    while (ap->rec_words < 1 + ap->n_states) {
        ap->rec_words <<= 1;
        ap->rec_shift++;
    }
    return NQE_OK;
}

int32_t nqe_agg_table_create(nqe_ctx *ctx, AggParams *ap, uint64_t capacity) {
    const uint64_t n_records = capacity + 1;
    void *table = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &table, (size_t)n_records * ap->rec_words * 8));
    ap->table = (unsigned long long *)table;
    ap->mask = capacity - 1;
    ap->hash_shift = 64;
    for (uint64_t c = capacity; c > 1; c >>= 1) ap->hash_shift--;
    if (capacity < 2) ap->hash_shift = 63; // single-record (global) table: slot 0/1 never used
    agg_init_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, ctx->stream>>>(*ap, n_records);
    ctx->launches++;
    NQE_CUDA(ctx, cudaGetLastError());
    return NQE_OK;
}

static int32_t read_scratch(nqe_ctx *ctx, int words_n) {
    cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, words_n * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return nqe_fail(ctx, NQE_ERR_CUDA, "aggregate kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    return NQE_OK;
}

// occupied records -> output table columns (count: UInt64, others Float64)
int32_t nqe_agg_extract(nqe_ctx *ctx, const AggParams &ap, bool is_global, int64_t max_groups, nqe_table *t) {
    ExtractParams xp;
    memset(&xp, 0, sizeof xp);
    const uint64_t n_records = is_global ? 1 : ap.mask + 2;
    int64_t out_cap = is_global ? 1 : (max_groups < (int64_t)n_records ? max_groups : (int64_t)n_records);
    if (out_cap < 1) out_cap = 1;
    t->cols.resize(ap.n_aggs);
    for (int a = 0; a < ap.n_aggs; a++) {
        const int32_t odt = ap.agg_op[a] == NQE_AGG_COUNT ? NQE_UINT64 : ap.agg_op[a] == NQE_AGG_GROUP_KEY ? NQE_INT64 : NQE_FLOAT64;
        NQE_TRY(nqe_column_alloc(ctx, odt, out_cap, false, &t->cols[a]));
        xp.out[a] = t->cols[a].values;
    }
    NQE_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, sizeof(uint64_t), ctx->stream));
    xp.out_count = (unsigned long long *)ctx->d_scratch;
    xp.is_global = is_global ? 1 : 0;
    agg_extract_kernel<<<(unsigned)((n_records + 255) / 256), 256, 0, ctx->stream>>>(ap, xp, n_records);
    ctx->launches++;
    NQE_TRY(read_scratch(ctx, 1));
    t->nrows = (int64_t)ctx->h_scratch[0];
    for (auto &c : t->cols) c.length = t->nrows;
    return NQE_OK;
}

// capacity for `est` expected groups
uint64_t nqe_agg_capacity(double est) {
    if (const char *e = getenv("NQE_AGG_CAP")) return nqe_next_pow2((uint64_t)atoll(e)); // tuning/debug knob
    return nqe_next_pow2((uint64_t)(est * 2.0) + 1024);
}

extern "C" int32_t nqe_hash_aggregate(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *group_expr,
                                      const nqe_agg *aggs, int32_t n_aggs, nqe_table **out) {
    if (!ctx || !in || !out || (!aggs && n_aggs > 0)) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    const int64_t n = in->nrows;

    // Utf8 group key (aggregate/mod.rs:170-216): a bare string column is replaced by its dictionary ids
    if (group_expr && group_expr->n_nodes == 1 && group_expr->nodes[0].kind == NQE_NODE_COLUMN &&
        group_expr->nodes[0].column >= 0 && group_expr->nodes[0].column < (int)in->cols.size() &&
        in->cols[group_expr->nodes[0].column].dtype == NQE_UTF8) {
        const int kc = group_expr->nodes[0].column;
        for (int a = 0; a < n_aggs; a++) // update(row) on a Utf8 argument: unimplemented!() (sum.rs:108)
            if (aggs[a].column == kc && aggs[a].op != NQE_AGG_COUNT && aggs[a].op != NQE_AGG_GROUP_KEY)
                return nqe_fail(ctx, NQE_ERR_PANIC, "%s func for Utf8 is not supported", agg_fn_name(aggs[a].op));
        for (int a = 0; a < n_aggs; a++)
            if (aggs[a].op == NQE_AGG_GROUP_KEY) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "NQE_AGG_GROUP_KEY over a Utf8 key");
        DevColumn ids;
        NQE_TRY(nqe_utf8_key_ids(ctx, in->cols[kc], nullptr, &ids, nullptr));
        nqe_table view;
        view.ctx = ctx;
        view.nrows = in->nrows;
        view.cols = in->cols;
        for (auto &c : view.cols) c.owned = false;
        view.cols[kc] = ids;
        view.cols[kc].owned = false;
        const int32_t rc = nqe_hash_aggregate(ctx, &view, group_expr, aggs, n_aggs, out);
        nqe_dev_free(ctx, ids.values);
        return rc;
    }

    DevProgramSet ps;
    memset(&ps, 0, sizeof ps);
    ExprInfo info[2];
    if (group_expr) { // program 0: group key
        const nqe_expr *list[1] = {group_expr};
        NQE_TRY(nqe_compile_exprs(ctx, in, list, 1, &ps, info));
        if (info[0].result_dtype != NQE_INT64 && info[0].result_dtype != NQE_UINT64) { // aggregate/mod.rs:217-219
            if (info[0].result_dtype == NQE_UTF8)
                return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "Utf8 group keys are not implemented on the CUDA path yet");
            return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "group by only support by `Int64`, `UInt64`, `String`");
        }
    }
    AggParams ap;
    memset(&ap, 0, sizeof ap);
    ap.n_rows = n;
    int32_t dts[AG_MAX];
    if (n_aggs > AG_MAX) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "at most %d aggregates per plan", AG_MAX);
    for (int a = 0; a < n_aggs; a++) {
        const int col = aggs[a].column;
        if (aggs[a].op == NQE_AGG_GROUP_KEY) { dts[a] = NQE_INT64; continue; } // no argument
        if (col < 0 || col >= (int)in->cols.size()) return nqe_fail(ctx, NQE_ERR_PANIC, "aggregate column index %d out of range", col);
        dts[a] = in->cols[col].dtype;
    }
    int32_t slots[AG_MAX];
    for (int a = 0; a < n_aggs; a++) { // register the argument columns as column slots
        if (aggs[a].op == NQE_AGG_GROUP_KEY) { slots[a] = 0; continue; }
        const DevColumn &c = in->cols[aggs[a].column];
        int slot = -1;
        for (int q = 0; q < ps.n_cols; q++)
            if (ps.cols[q].values == c.values && ps.cols[q].dtype == c.dtype) slot = q;
        if (slot < 0) {
            if (ps.n_cols >= NQE_MAX_COLS) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "too many distinct columns");
            slot = ps.n_cols++;
            ps.cols[slot].values = c.values;
            ps.cols[slot].validity = (const uint32_t *)c.validity;
            ps.cols[slot].dtype = c.dtype;
        }
        slots[a] = slot;
    }
    NQE_TRY(nqe_agg_layout(ctx, aggs, n_aggs, dts, slots, group_expr != nullptr, &ap));
    ap.status = (uint32_t *)(ctx->d_scratch + 1);

    nqe_table *t;
    nqe_table_new(ctx, 0, &t);
    int32_t rc = NQE_OK;
    OpTimer timer(ctx);

    if (!group_expr) {
        cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
        rc = nqe_agg_table_create(ctx, &ap, 1); // record 0 is the single global row
        if (rc == NQE_OK && n > 0 && n_aggs > 0) {
            int grid = ctx->sm_count * 8;
            const int64_t need = (n + AG_THREADS - 1) / AG_THREADS;
            if (grid > need) grid = (int)need;
            global_aggregate_kernel<<<grid, AG_THREADS, 0, ctx->stream>>>(ps, ap);
            ctx->launches++;
        }
    } else {
        // fast path: bare NULL-free key column, NULL-free arguments
        int key_slot = 0;
        bool simple = ps.prog_begin[1] - ps.prog_begin[0] == 1 && ps.ops[ps.prog_begin[0]].code == UOP_LOAD &&
                      ps.ops[ps.prog_begin[0]].src == SRC_COL;
        if (simple) {
            key_slot = ps.ops[ps.prog_begin[0]].slot;
            for (int q = 0; q < ps.n_cols; q++)
                if (ps.cols[q].validity) simple = false;
        }
        // --- size the table from a sampled cardinality estimate, grow on overflow
        uint64_t capacity = 1024;
        double est_groups = 1e18;
        long long key_minmax[2] = {0, -1}; // smallest / largest key of the sample (signed order)
        bool skewed = false; // a hot key: the partitioned path would put all its rows on one SM (measured, Zipf-1.0: 52 vs 21 ms)
        if (n > 0) {
            const int64_t n_sample = n < (1 << 20) ? n : (1 << 20);
            const int64_t stride = n / n_sample;
            const uint32_t bits = 1u << 23; // 8 Mi bits = 1 MiB bitmap
            void *bm = nullptr;
            rc = nqe_dev_alloc(ctx, &bm, bits / 8 + AGG_SKEW_BINS * 4 + 16);
            if (rc == NQE_OK) {
                unsigned int *bins = (unsigned int *)((uint8_t *)bm + bits / 8);
                long long *d_minmax = (long long *)(bins + AGG_SKEW_BINS);
                unsigned int h_bins[AGG_SKEW_BINS];
                const long long mm_init[2] = {LLONG_MAX, LLONG_MIN};
                cudaMemsetAsync(bm, 0, bits / 8 + AGG_SKEW_BINS * 4, ctx->stream);
                cudaMemcpyAsync(d_minmax, mm_init, sizeof mm_init, cudaMemcpyHostToDevice, ctx->stream);
                cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
                unsigned sgrid = (unsigned)((n_sample + 255) / 256);
                if (sgrid > (unsigned)ctx->sm_count * 4) sgrid = (unsigned)ctx->sm_count * 4;
                agg_sample_kernel<<<sgrid, 256, 0, ctx->stream>>>(
                    ps, n, stride, n_sample, (uint32_t *)bm, bits - 1, ap.status, bins, d_minmax);
                popcount_kernel<<<64, 256, 0, ctx->stream>>>((const uint32_t *)bm, bits / 32, (unsigned long long *)ctx->d_scratch);
                ctx->launches += 2;
                cudaMemcpyAsync(h_bins, bins, sizeof h_bins, cudaMemcpyDeviceToHost, ctx->stream);
                cudaMemcpyAsync(key_minmax, d_minmax, sizeof key_minmax, cudaMemcpyDeviceToHost, ctx->stream);
                rc = read_scratch(ctx, 2);
                nqe_dev_free(ctx, bm);
                if (rc == NQE_OK) {
                    uint64_t tot = 0, mx = 0;
                    for (unsigned int c : h_bins) { tot += c; mx = c > mx ? c : mx; }
                    // a partition of the split holds ~1/148 of the keys; a bin 4x the mean holds a key with > ~1 % of the rows
                    skewed = tot >= 4096 && mx * AGG_SKEW_BINS > 4 * tot;
                }
            }
            if (rc == NQE_OK) {
                const double m = (double)bits, z = m - (double)ctx->h_scratch[0];
                const double distinct = z > 0 ? -m * log(z / m) : m * 16; // linear counting
                double est = distinct;
                if (distinct > 0.5 * (double)n_sample) est = distinct * ((double)n / (double)n_sample); // still growing
                if (est > (double)n) est = (double)n;
                capacity = nqe_agg_capacity(est);
                est_groups = est;
            }
        }
        // --- partitioned shared-memory path (gp2_aggregate_kernel): split once, aggregate inside the retry loop
        static std::atomic<int> agg_part{-1}; // guard published last: operators may be called from several host threads (multi.cu)
        static int64_t agg_part_min_rows = 0;
        if (agg_part.load(std::memory_order_acquire) < 0) {
            const char *e = getenv("NQE_AGG_PART_MIN_ROWS");
            agg_part_min_rows = e ? atoll(e) : ((int64_t)1 << 22);
            e = getenv("NQE_AGG_PART");
            agg_part.store(e ? atoi(e) != 0 : 1, std::memory_order_release);
        }
        bool use_part = false;
        PagedStreams streams;
        memset(&streams, 0, sizeof streams);
        int gp_need = 0, gp_m = 1, gp_dtype = 0;
        long long dense_lo = 0;
        unsigned dense_width = 0;
        static int allow_dense = -1; // knob NQE_AGG_DENSE=0: always hash the keys
        if (allow_dense < 0) {
            const char *e = getenv("NQE_AGG_DENSE");
            allow_dense = e ? atoi(e) : 1;
        }
        bool try_dense = allow_dense != 0;
    split_again:
        if (rc == NQE_OK && agg_part && !skewed && n >= agg_part_min_rows && ap.n_states > 0 && est_groups >= 2048.0) {
            bool one_src = true;
            for (int q = 0; q < ap.n_states; q++) {
                if (ap.st_src[q] != ap.st_src[0]) one_src = false;
                gp_need |= 1 << ap.st_kind[q];
            }
            const DevColRef &vc = ps.cols[ap.st_src[0]];
            int P = 0;
            if (!nqe_gp2_plan(ctx, est_groups, &P, &gp_m)) P = 0;
            // dense keys (the sampled range is not much larger than the number of groups): split by key RANGE, aggregate in
            // directly indexed tables; a key outside the sampled range is detected by the split and sends us back to hashing
            dense_width = 0;
            if (try_dense && key_minmax[0] <= key_minmax[1]) {
                const unsigned long long range = (unsigned long long)key_minmax[1] - (unsigned long long)key_minmax[0] + 1ull;
                const unsigned long long per_part = (range + nqe_dense_parts(ctx, range) - 1) / nqe_dense_parts(ctx, range);
                if (range < (1ull << 32) && (double)range <= 4.0 * est_groups + 1024.0 && per_part <= NQE_GP2_DENSE_MAX_WIDTH &&
                    ctx->sm_count <= PS_MAX_PARTS) {
                    dense_lo = key_minmax[0];
                    dense_width = (unsigned)(per_part < 64 ? 64 : per_part);
                    P = (int)((range + dense_width - 1) / dense_width);
                    gp_m = ctx->sm_count / P;
                }
            }
            if (P > 0 && one_src && !vc.validity && (vc.dtype == NQE_INT64 || vc.dtype == NQE_UINT64 || vc.dtype == NQE_FLOAT64)) {
                gp_dtype = vc.dtype;
                rc = nqe_ps_create(ctx, n, P, &streams);
                if (rc == NQE_OK) {
                    PsSplitArgs sa{simple ? (const unsigned long long *)ps.cols[key_slot].values : nullptr,
                                   (const unsigned long long *)vc.values, n, gp_dtype};
                    if (dense_width) {
                        const PartByRange part{dense_lo, (unsigned long long)key_minmax[1] - (unsigned long long)key_minmax[0] + 1ull, dense_width,
                                               ps_div_magic(dense_width), ap.status};
                        if (simple) rc = ps_split_launch<false, PartByRange>(ctx, streams, sa, part, ps, ap.status);
                        else rc = ps_split_launch<true, PartByRange>(ctx, streams, sa, part, ps, ap.status);
                    } else {
                        const PartByHash part{(uint32_t)P};
                        if (simple) rc = ps_split_launch<false, PartByHash>(ctx, streams, sa, part, ps, ap.status);
                        else rc = ps_split_launch<true, PartByHash>(ctx, streams, sa, part, ps, ap.status);
                    }
                    use_part = rc == NQE_OK;
                }
            }
        }
        uint32_t split_flags = 0; // status bits raised by the split (key expression errors, keys outside the dense range)
        for (int attempt = 0; rc == NQE_OK && attempt < 8; attempt++) {
            // the split ran before this loop and wrote its flags into the same status word: they are read with the first attempt
            if (!(use_part && attempt == 0)) cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
            rc = nqe_agg_table_create(ctx, &ap, capacity);
            if (rc != NQE_OK) break;
            if (n > 0) {
                static int agk = 0; // knob NQE_AGG_K: rows per thread of the fast path (2, 4 or 8)
                if (!agk) {
                    const char *e = getenv("NQE_AGG_K");
                    agk = e ? atoi(e) : 4;
                    if (agk != 2 && agk != 8) agk = 4;
                }
                const int kk = simple ? agk : AG_K;
                const int64_t tiles = (n + kk * AG_THREADS - 1) / (kk * AG_THREADS);
                int grid = ctx->sm_count * 8;
                if (grid > tiles) grid = (int)tiles;
                if (use_part) rc = nqe_gp2_aggregate(ctx, streams, ap, gp_m, gp_need, dense_lo, dense_width);
                else if (simple && agk == 2) group_aggregate_kernel<true, 2><<<grid, AG_THREADS, 0, ctx->stream>>>(ps, ap, key_slot);
                else if (simple && agk == 8) group_aggregate_kernel<true, 8><<<grid, AG_THREADS, 0, ctx->stream>>>(ps, ap, key_slot);
                else if (simple) group_aggregate_kernel<true, 4><<<grid, AG_THREADS, 0, ctx->stream>>>(ps, ap, key_slot);
                else group_aggregate_kernel<false, AG_K><<<grid, AG_THREADS, 0, ctx->stream>>>(ps, ap, 0);
                if (!use_part) ctx->launches++;
                if (rc != NQE_OK) break;
            }
            rc = read_scratch(ctx, 2);
            if (rc != NQE_OK) break;
            if (use_part && attempt == 0) split_flags = (uint32_t)ctx->h_scratch[1] & (DEV_ERR_DIV0 | DEV_ERR_OVERFLOW | DEV_ERR_RANGE | DEV_ERR_CAPACITY);
            const uint32_t st = (uint32_t)ctx->h_scratch[1] | split_flags;
            if (st & DEV_ERR_DIV0) { rc = nqe_fail(ctx, NQE_ERR_DIVIDE_BY_ZERO, "Divide by zero error"); break; }
            if (st & DEV_ERR_OVERFLOW) { rc = nqe_fail(ctx, NQE_ERR_PANIC, "attempt to divide with overflow"); break; }
            if (st & DEV_ERR_CAPACITY) { rc = nqe_fail(ctx, NQE_ERR_CUDA, "internal: paged stream pool exhausted"); break; }
            if (use_part && dense_width && (st & DEV_ERR_RANGE)) { // a key outside the sampled range: hash the keys instead
                nqe_dev_free(ctx, ap.table);
                ap.table = nullptr;
                nqe_ps_destroy(ctx, &streams);
                use_part = false;
                try_dense = false;
                cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
                goto split_again;
            }
            if (!(st & DEV_ERR_TABLE_FULL)) break;
            nqe_dev_free(ctx, ap.table);
            ap.table = nullptr;
            capacity *= 8;
            if (attempt == 7) rc = nqe_fail(ctx, NQE_ERR_OOM, "group-by table kept overflowing");
        }
        nqe_ps_destroy(ctx, &streams);
    }
    if (rc == NQE_OK) rc = nqe_agg_extract(ctx, ap, group_expr == nullptr, n, t);
    timer.stop();
    nqe_dev_free(ctx, ap.table);
    if (rc != NQE_OK) {
        nqe_table_free(t);
        return rc;
    }
    *out = t;
    return NQE_OK;
}

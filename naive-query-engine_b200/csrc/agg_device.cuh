// agg_device.cuh -- group-by table records and per-row state updates shared by
// hash_aggregate.cu and the fused join->aggregate kernel in hash_join.cu.
#pragma once
#include <cfloat>

#include "expr_eval.cuh"
#include "hash_common.cuh"
#include "nqe_internal.cuh"

constexpr int AG_THREADS = 256;
constexpr int AG_K = 4;
constexpr int AG_MAX = 16;
constexpr uint64_t EMPTY_KEY = 0x8000000000000000ULL; // i64::MIN; its group lives in the extra last record
constexpr int MAX_PROBE = 256;

constexpr int AG_MAXS = 2 * AG_MAX; // state words (an AVG needs a SUM and a CNT state)
enum : int32_t { ST_CNT = 0, ST_SUM = 1, ST_MIN = 2, ST_MAX = 3 };

// Aggregates over the same argument share state words: count(v), sum(v), avg(v), min(v),
// max(v) need {CNT(v), SUM(v), MIN(v), MAX(v)} -- two reductions and two compare-loads per row.
struct AggParams {
    int64_t n_rows;
    int32_t n_aggs, n_states, rec_words, pad;
    int32_t agg_op[AG_MAX];
    int32_t agg_state[AG_MAX];  // primary state of the aggregate
    int32_t agg_state2[AG_MAX]; // AVG: its CNT state
    int32_t st_kind[AG_MAXS];
    int32_t st_src[AG_MAXS];    // caller-defined source id of the argument column (sorted)
    int32_t st_off[AG_MAXS];    // word offset of the state inside a record
    unsigned long long *table;  // (capacity + 1) records: [key | states...]
    uint64_t mask;              // capacity - 1
    uint32_t *status;
    int32_t rec_shift;          // log2(rec_words): records are power-of-two sized
    int32_t hash_shift;         // 64 - log2(capacity)
};

__device__ __forceinline__ double value_as_f64(int dtype, uint64_t bits) {
    if (dtype == NQE_INT64) return __ll2double_rn((long long)bits);
    if (dtype == NQE_UINT64) return __ull2double_rn(bits);
    return __longlong_as_double((long long)bits);
}

__device__ __forceinline__ void red_add_f64(unsigned long long *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_max_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_min_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.global.min.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// slot of a key: Fibonacci (multiply-shift) hashing, one 64-bit multiply
__device__ __forceinline__ uint64_t agg_slot_of(const AggParams &ap, uint64_t key) {
    return (key * 0x9E3779B97F4A7C15ULL) >> ap.hash_shift;
}
__device__ __forceinline__ unsigned long long *agg_rec(const AggParams &ap, uint64_t slot) {
    return ap.table + (slot << ap.rec_shift);
}

// Sector 0 of a record = its first 32 bytes: the key and the first three state words.  The
// layout (nqe_agg_layout) puts MIN/MAX states there, so ONE 256-bit load (LDG.256, L2 only:
// an L1 miss would fetch the whole 128-byte line) answers "is this my key?" and "can my
// value improve min/max?" at once -- a row costs one random sector read plus its reductions.
struct Sector0 {
    unsigned long long w[4];
};
__device__ __forceinline__ Sector0 ld_sector0(const AggParams &, const unsigned long long *rec) {
    Sector0 s;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(s.w[0]), "=l"(s.w[1]), "=l"(s.w[2]), "=l"(s.w[3]) : "l"(rec));
    return s;
}

// one probe step at `slot`: 1 = found/claimed, 0 = occupied by another key
__device__ __forceinline__ int agg_probe_step(unsigned long long *r, unsigned long long seen, uint64_t key) {
    if (seen == EMPTY_KEY) seen = atomicCAS(r, (unsigned long long)EMPTY_KEY, (unsigned long long)key);
    return seen == key || seen == EMPTY_KEY;
}

// find-or-claim the records of K keys.  The first probe of all K keys is issued together
// (K independent loads in flight: most keys resolve there at load factor <= 0.5); the rare
// collisions are then walked one key at a time with a tight loop.  rec[j] == nullptr afterwards:
// key j was not looked up (want bit clear) or the table is full (status flagged).  s0[j] is a
// (possibly stale, which is safe: MIN/MAX states only move one way) copy of the record's sector 0.
template <int K>
__device__ __forceinline__ void find_slots(const AggParams &ap, const uint64_t (&key)[K], uint32_t want,
                                           unsigned long long *(&rec)[K], Sector0 (&s0)[K]) {
    uint64_t slot[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        slot[j] = agg_slot_of(ap, key[j]);
        if ((want >> j) & 1u) s0[j] = ld_sector0(ap, agg_rec(ap, slot[j]));
    }
#pragma unroll
    for (int j = 0; j < K; j++) {
        rec[j] = nullptr;
        if (!((want >> j) & 1u)) continue;
        if (key[j] == EMPTY_KEY) { // i64::MIN has its own record; word 0 != EMPTY marks it occupied
            rec[j] = agg_rec(ap, ap.mask + 1);
            *(volatile unsigned long long *)rec[j] = 0ull;
            s0[j] = ld_sector0(ap, rec[j]);
            continue;
        }
        unsigned long long *r = agg_rec(ap, slot[j]);
        if (agg_probe_step(r, s0[j].w[0], key[j])) { rec[j] = r; continue; }
        uint64_t s = slot[j];
        for (int probe = 1; probe < MAX_PROBE; probe++) {
            s = (s + 1) & ap.mask;
            r = agg_rec(ap, s);
            s0[j] = ld_sector0(ap, r);
            if (agg_probe_step(r, s0[j].w[0], key[j])) { rec[j] = r; break; }
        }
        if (!rec[j]) atomicOr(ap.status, DEV_ERR_TABLE_FULL);
    }
}

__device__ __forceinline__ unsigned long long *find_slot(const AggParams &ap, uint64_t key, Sector0 *s0) {
    const uint64_t k1[1] = {key};
    unsigned long long *r1[1];
    Sector0 s1[1];
    find_slots<1>(ap, k1, 1u, r1, s1);
    *s0 = s1[0];
    return r1[0];
}

// Fold one input row into its group's record.  src(id, &dtype, &bits) -> the argument value
// of source id for this row is non-NULL.  States are sorted by source id, so a column that
// feeds several states is fetched once.  s0 = the record's sector 0 as seen by the probe.
template <typename Src>
__device__ __forceinline__ void update_states(const AggParams &ap, unsigned long long *rec, const Sector0 &s0, const Src &src) {
    int last = -1, dtype = 0;
    bool valid = false;
    uint64_t bits = 0;
    for (int s = 0; s < ap.n_states; s++) {
        if (ap.st_src[s] != last) {
            last = ap.st_src[s];
            valid = src(last, &dtype, &bits);
        }
        if (!valid) continue; // NULL argument: the row does not touch this state
        const int off = ap.st_off[s];
        unsigned long long *w = rec + off;
        const int kind = ap.st_kind[s];
        if (kind == ST_CNT) { red_add_u64(w, 1ull); continue; }
        const double v = value_as_f64(dtype, bits);
        if (kind == ST_SUM) { red_add_f64(w, v); continue; }
        const unsigned long long k = nqe_f64_to_ord(v);
        // current state: from the probe's sector-0 copy when the word lives there (layout puts MIN/MAX first)
        const unsigned long long cur = off == 1 ? s0.w[1] : off == 2 ? s0.w[2] : off == 3 ? s0.w[3] : ld_relaxed_u64(w);
        if (kind == ST_MAX) {
            if (k > cur) red_max_u64(w, k);
        } else if (v == v) { // MIN: `val < self.val` is never true for NaN (min.rs:49)
            if (k < cur) red_min_u64(w, k);
        }
    }
}

// host helpers implemented in hash_aggregate.cu
int32_t nqe_agg_layout(nqe_ctx *ctx, const nqe_agg *aggs, int32_t n_aggs, const int32_t *col_dtypes,
                       const int32_t *src_ids, bool grouped, AggParams *ap);
int32_t nqe_agg_table_create(nqe_ctx *ctx, AggParams *ap, uint64_t capacity);
int32_t nqe_agg_extract(nqe_ctx *ctx, const AggParams &ap, bool is_global, int64_t max_groups, nqe_table *t);
uint64_t nqe_agg_capacity(double est);
// partitioned shared-memory group-by over paged streams (hash_aggregate.cu; paged_split.cuh)
struct PagedStreams;
int32_t nqe_estimate_distinct_u64(nqe_ctx *ctx, const unsigned long long *col, int64_t n, double *est);
bool nqe_gp2_plan(nqe_ctx *ctx, double est_groups, int *P, int *m);
// dense_width != 0: the streams were split by key RANGE (PartByRange in paged_split.cuh): partition p holds the keys
// [dense_lo + p * dense_width, + dense_width) and is aggregated in a directly indexed table
int32_t nqe_gp2_aggregate(nqe_ctx *ctx, const PagedStreams &streams, const AggParams &ap, int m, int need, long long dense_lo = 0,
                          unsigned dense_width = 0);
constexpr unsigned NQE_GP2_DENSE_MAX_WIDTH = 3072; // = GA_SLOTS
// partitions of the dense (key-range) split: the fewest whose per-partition range fits the table, but at least a quarter
// of the SMs (fewer partitions = longer runs and fewer cursor atomics per tile in the split; the aggregator then runs
// sm_count / P CTAs per partition).  Knob NQE_DENSE_PARTS overrides.
int nqe_dense_parts(nqe_ctx *ctx, unsigned long long range);
int32_t nqe_minmax_i64(nqe_ctx *ctx, const unsigned long long *col, int64_t n, long long *lo, long long *hi);
int32_t nqe_minmax2_i64(nqe_ctx *ctx, const unsigned long long *a, const unsigned long long *b, int64_t n, long long *mm);

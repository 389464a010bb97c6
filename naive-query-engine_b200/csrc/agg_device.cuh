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

struct AggParams {
    int64_t n_rows;
    int32_t n_aggs;
    int32_t rec_words;
    int32_t op[AG_MAX];
    int32_t col_slot[AG_MAX];  // index into DevProgramSet::cols
    int32_t state_off[AG_MAX]; // word offset of the op's state inside a record
    unsigned long long *table; // (capacity + 1) records
    uint64_t mask;             // capacity - 1
    uint32_t *status;
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

// find-or-claim the record of `key`; returns nullptr when the table is full
__device__ __forceinline__ unsigned long long *find_slot(const AggParams &ap, uint64_t key) {
    if (key == EMPTY_KEY) { // i64::MIN has its own record; word 0 != EMPTY marks it occupied
        unsigned long long *rec = ap.table + (ap.mask + 1) * ap.rec_words;
        *(volatile unsigned long long *)rec = 0ull;
        return rec;
    }
    uint64_t slot = nqe_mix64(key) & ap.mask;
    for (int probe = 0; probe < MAX_PROBE; probe++) {
        unsigned long long *rec = ap.table + slot * ap.rec_words;
        unsigned long long k = ld_relaxed_u64(rec);
        if (k == key) return rec;
        if (k == EMPTY_KEY) {
            k = atomicCAS(rec, (unsigned long long)EMPTY_KEY, (unsigned long long)key);
            if (k == EMPTY_KEY || k == key) return rec;
        }
        slot = (slot + 1) & ap.mask;
    }
    return nullptr;
}

// one value (already known non-NULL) folded into state word(s) of aggregate a
__device__ __forceinline__ void update_state(const AggParams &ap, int a, unsigned long long *rec, int dtype,
                                             uint64_t bits) {
    unsigned long long *s = rec + ap.state_off[a];
    const int op = ap.op[a];
    if (op == NQE_AGG_COUNT) { red_add_u64(s, 1ull); return; }
    const double v = value_as_f64(dtype, bits);
    if (op == NQE_AGG_SUM) red_add_f64(s, v);
    else if (op == NQE_AGG_AVG) { red_add_f64(s, v); red_add_u64(s + 1, 1ull); }
    else if (op == NQE_AGG_MAX) {
        const unsigned long long k = nqe_f64_to_ord(v);
        if (k > ld_relaxed_u64(s)) red_max_u64(s, k);
    } else if (v == v) { // MIN: `val < self.val` is never true for NaN (min.rs:49)
        const unsigned long long k = nqe_f64_to_ord(v);
        if (k < ld_relaxed_u64(s)) red_min_u64(s, k);
    }
}

__device__ __forceinline__ void update_record(const AggParams &ap, const DevProgramSet &ps, unsigned long long *rec,
                                              int64_t e) {
    for (int a = 0; a < ap.n_aggs; a++) {
        const DevColRef &c = ps.cols[ap.col_slot[a]];
        if (c.validity && !((__ldg(c.validity + (e >> 5)) >> (e & 31)) & 1u)) continue;
        const uint64_t bits = ap.op[a] == NQE_AGG_COUNT ? 0ull : ld_cached_u64((const uint64_t *)c.values + e);
        update_state(ap, a, rec, c.dtype, bits);
    }
}

// host helpers implemented in hash_aggregate.cu
int32_t nqe_agg_layout(nqe_ctx *ctx, const nqe_agg *aggs, int32_t n_aggs, const int32_t *col_dtypes, bool grouped,
                       AggParams *ap);
int32_t nqe_agg_table_create(nqe_ctx *ctx, AggParams *ap, uint64_t capacity);
int32_t nqe_agg_extract(nqe_ctx *ctx, const AggParams &ap, bool is_global, int64_t max_groups, nqe_table *t);
uint64_t nqe_agg_capacity(double est);

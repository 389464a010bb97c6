// expr_eval.cuh -- per-row evaluation of lowered PhysicalExpr programs.
//
// Each thread evaluates K rows (row index e0 + j*stride) with the running value
// in registers.  Semantics follow the reference / arrow 13 kernels:
//   compare  : NULL if either side NULL; IEEE partial order for Float64
//   and/or   : Kleene logic (and_kleene / or_kleene)
//   + - *    : wrapping for Int64/UInt64, IEEE for Float64
//   / %      : any valid zero divisor (floats included) raises DivideByZero;
//              i64::MIN / -1 and % -1 are overflow panics; % is truncated
//   abs/sin/cos on Float64 (Tan lowered to cos on the host)
#pragma once

#include "nqe_internal.cuh"

__device__ __forceinline__ uint64_t ld_stream_u64(const void *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t ld_cached_u64(const void *p) { return __ldg((const unsigned long long *)p); }

template <int K>
struct RowRegs {
    uint64_t v[K];
    uint32_t valid; // bit j: v[j] is non-NULL
};

// Row sources.  GlobalRows reads column j-th row straight from HBM (row index
// e0 + j*stride); SmemRows reads a tile staged in shared memory by TMA bulk copies
// (values of column slot s at byte offset voff[s], bitmap words at boff[s]).
struct GlobalRows {
    int64_t e0, stride;
    __device__ __forceinline__ uint64_t value(const DevColRef &c, int, int j) const {
        return ld_cached_u64((const uint64_t *)c.values + e0 + j * stride);
    }
    __device__ __forceinline__ uint32_t boolbit(const DevColRef &c, int, int j) const {
        const int64_t e = e0 + j * stride;
        return (__ldg((const uint32_t *)c.values + (e >> 5)) >> (e & 31)) & 1u;
    }
    __device__ __forceinline__ uint32_t validbit(const DevColRef &c, int, int j) const {
        const int64_t e = e0 + j * stride;
        return (__ldg(c.validity + (e >> 5)) >> (e & 31)) & 1u;
    }
};

struct SmemRows {
    const uint8_t *stage;    // this tile's stage buffer
    const uint16_t *voff16;  // per column slot: values offset / 16
    const uint16_t *boff16;  // per column slot: validity bitmap offset / 16
    int r0, stride;          // row in tile = r0 + j*stride
    __device__ __forceinline__ uint64_t value(const DevColRef &, int slot, int j) const {
        return *(const uint64_t *)(stage + (size_t)voff16[slot] * 16 + (size_t)(r0 + j * stride) * 8);
    }
    __device__ __forceinline__ uint32_t boolbit(const DevColRef &, int slot, int j) const {
        const int r = r0 + j * stride;
        return (*(const uint32_t *)(stage + (size_t)voff16[slot] * 16 + (r >> 5) * 4) >> (r & 31)) & 1u;
    }
    __device__ __forceinline__ uint32_t validbit(const DevColRef &, int slot, int j) const {
        const int r = r0 + j * stride;
        return (*(const uint32_t *)(stage + (size_t)boff16[slot] * 16 + (r >> 5) * 4) >> (r & 31)) & 1u;
    }
};

template <int K, typename Rows>
__device__ __forceinline__ void load_operand(const DevProgramSet &ps, const DevOp &op, const Rows &rows,
                                             uint32_t inrange, uint32_t rownull, RowRegs<K> &b) {
    if (op.src == SRC_COL) {
        const DevColRef &c = ps.cols[op.slot];
        uint32_t valid = inrange;
        if (c.dtype == NQE_BOOL) {
#pragma unroll
            for (int j = 0; j < K; j++) b.v[j] = ((inrange >> j) & 1u) ? rows.boolbit(c, op.slot, j) : 0ull;
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) b.v[j] = ((inrange >> j) & 1u) ? rows.value(c, op.slot, j) : 0ull;
        }
        if (c.validity) {
#pragma unroll
            for (int j = 0; j < K; j++)
                if ((inrange >> j) & 1u)
                    if (!rows.validbit(c, op.slot, j)) valid &= ~(1u << j);
        }
        b.valid = valid & ~rownull;
    } else if (op.src == SRC_LIT) {
#pragma unroll
        for (int j = 0; j < K; j++) b.v[j] = op.imm;
        b.valid = inrange;
    } else { // SRC_NULL
#pragma unroll
        for (int j = 0; j < K; j++) b.v[j] = 0;
        b.valid = 0;
    }
}

template <typename T>
__device__ __forceinline__ bool cmp_op(uint8_t code, T x, T y) {
    switch (code) {
    case NQE_OP_EQ: return x == y;
    case NQE_OP_NOT_EQ: return x != y;
    case NQE_OP_LT: return x < y;
    case NQE_OP_LT_EQ: return x <= y;
    case NQE_OP_GT: return x > y;
    default: return x >= y;
    }
}

// a (op) b -> a.  `active` = rows whose errors count (kept, in range).
template <int K>
__device__ __forceinline__ void apply_binary(uint8_t code, uint8_t type, RowRegs<K> &a, const RowRegs<K> &b,
                                             uint32_t active, uint32_t *status) {
    const uint32_t both = a.valid & b.valid;
    if (code <= NQE_OP_GT_EQ) {
        if (type == T_I64) {
#pragma unroll
            for (int j = 0; j < K; j++) a.v[j] = cmp_op<long long>(code, (long long)a.v[j], (long long)b.v[j]);
        } else if (type == T_F64) {
#pragma unroll
            for (int j = 0; j < K; j++)
                a.v[j] = cmp_op<double>(code, __longlong_as_double((long long)a.v[j]), __longlong_as_double((long long)b.v[j]));
        } else { // U64, BOOL (0/1)
#pragma unroll
            for (int j = 0; j < K; j++) a.v[j] = cmp_op<unsigned long long>(code, a.v[j], b.v[j]);
        }
        a.valid = both;
        return;
    }
    if (code == NQE_OP_AND || code == NQE_OP_OR) {
        uint32_t at = 0, bt = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            at |= (uint32_t)(a.v[j] & 1) << j;
            bt |= (uint32_t)(b.v[j] & 1) << j;
        }
        at &= a.valid; bt &= b.valid;
        const uint32_t af = a.valid & ~at, bf = b.valid & ~bt;
        uint32_t val, ok;
        if (code == NQE_OP_AND) { val = at & bt; ok = both | af | bf; }
        else { val = at | bt; ok = both | at | bt; }
#pragma unroll
        for (int j = 0; j < K; j++) a.v[j] = (val >> j) & 1u;
        a.valid = ok;
        return;
    }
    // arithmetic
    if (type == T_F64) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            double x = __longlong_as_double((long long)a.v[j]), y = __longlong_as_double((long long)b.v[j]), r;
            switch (code) {
            case NQE_OP_PLUS: r = __dadd_rn(x, y); break;
            case NQE_OP_MINUS: r = __dsub_rn(x, y); break;
            case NQE_OP_MULTIPLY: r = __dmul_rn(x, y); break;
            case NQE_OP_DIVIDE:
                if (y == 0.0 && ((both & active) >> j & 1u)) atomicOr(status, DEV_ERR_DIV0);
                r = x / y; break;
            default:
                if (y == 0.0 && ((both & active) >> j & 1u)) atomicOr(status, DEV_ERR_DIV0);
                r = fmod(x, y); break;
            }
            a.v[j] = (uint64_t)__double_as_longlong(r);
        }
    } else if (type == T_I64) {
#pragma unroll
        for (int j = 0; j < K; j++) {
            long long x = (long long)a.v[j], y = (long long)b.v[j];
            unsigned long long r;
            switch (code) {
            case NQE_OP_PLUS: r = (unsigned long long)x + (unsigned long long)y; break;
            case NQE_OP_MINUS: r = (unsigned long long)x - (unsigned long long)y; break;
            case NQE_OP_MULTIPLY: r = (unsigned long long)x * (unsigned long long)y; break;
            default: {
                const bool live = (both & active) >> j & 1u;
                if (y == 0) { if (live) atomicOr(status, DEV_ERR_DIV0); r = 0; }
                else if (y == -1) {
                    if (x == LLONG_MIN) { if (live) atomicOr(status, DEV_ERR_OVERFLOW); r = 0; }
                    else r = code == NQE_OP_DIVIDE ? (unsigned long long)(-x) : 0ull;
                } else r = (unsigned long long)(code == NQE_OP_DIVIDE ? x / y : x % y);
            }
            }
            a.v[j] = r;
        }
    } else { // U64
#pragma unroll
        for (int j = 0; j < K; j++) {
            unsigned long long x = a.v[j], y = b.v[j], r;
            switch (code) {
            case NQE_OP_PLUS: r = x + y; break;
            case NQE_OP_MINUS: r = x - y; break;
            case NQE_OP_MULTIPLY: r = x * y; break;
            default:
                if (y == 0) { if ((both & active) >> j & 1u) atomicOr(status, DEV_ERR_DIV0); r = 0; }
                else r = code == NQE_OP_DIVIDE ? x / y : x % y;
            }
            a.v[j] = r;
        }
    }
    a.valid = both;
}

// Run program p for K rows; result in acc.  rownull = rows whose inputs are
// forced NULL (predicate was NULL, selection.rs:46).
template <int K, typename Rows>
__device__ __forceinline__ void run_program_on(const DevProgramSet &ps, int p, const Rows &rows, uint32_t inrange,
                                               uint32_t active, uint32_t rownull, RowRegs<K> &acc, uint32_t *status) {
    RowRegs<K> stack[NQE_STACK];
    const int end = ps.prog_begin[p + 1];
    for (int i = ps.prog_begin[p]; i < end; i++) {
        const DevOp op = ps.ops[i];
        if (op.code == UOP_LOAD) {
            load_operand<K>(ps, op, rows, inrange, rownull, acc);
        } else if (op.code == UOP_PUSH) {
            stack[op.slot] = acc;
        } else if (op.code >= UOP_ABS) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                double x = __longlong_as_double((long long)acc.v[j]);
                double r = op.code == UOP_ABS ? fabs(x) : (op.code == UOP_SIN ? sin(x) : cos(x));
                acc.v[j] = (uint64_t)__double_as_longlong(r);
            }
        } else if (op.src == SRC_STACK) {
            RowRegs<K> l = stack[op.slot];
            apply_binary<K>(op.code, op.type, l, acc, active, status);
            acc = l;
        } else {
            RowRegs<K> b;
            load_operand<K>(ps, op, rows, inrange, rownull, b);
            apply_binary<K>(op.code, op.type, acc, b, active, status);
        }
    }
}

template <int K>
__device__ __forceinline__ void run_program(const DevProgramSet &ps, int p, int64_t e0, int64_t stride,
                                            uint32_t inrange, uint32_t active, uint32_t rownull,
                                            RowRegs<K> &acc, uint32_t *status) {
    run_program_on<K>(ps, p, GlobalRows{e0, stride}, inrange, active, rownull, acc, status);
}

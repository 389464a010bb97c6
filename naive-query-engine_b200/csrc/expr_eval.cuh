// expr_eval.cuh -- per-row evaluation of lowered PhysicalExpr programs.
//
// Each thread evaluates K rows with the running value in registers; every
// micro-op is decoded once per K rows (the dispatch is warp-uniform) and its
// element loop is fully unrolled.  Two instantiations exist:
//   NULLS = false : all K rows are in range and no operand can be NULL (no
//                   validity bitmap, no NULL literal) -- no validity tracking at all
//   NULLS = true  : ragged tiles and/or nullable operands
// Semantics follow the reference / arrow 13 kernels:
//   compare  : NULL if either side NULL; IEEE partial order for Float64
//   and/or   : Kleene logic (and_kleene / or_kleene)
//   + - *    : wrapping for Int64/UInt64, IEEE for Float64
//   / %      : any valid zero divisor (floats included) raises DivideByZero;
//              i64::MIN / -1 and % -1 are overflow panics; % is truncated
//   abs/sin/cos on Float64 (Tan lowered to cos on the host)
#pragma once

#include <climits>

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
    uint32_t valid; // bit j: v[j] is non-NULL (all ones when NULLS == false)
};

// Row sources.  GlobalRows reads row e0 + j*stride straight from HBM; SmemRows
// reads a tile staged in shared memory by TMA bulk copies (values of column slot
// s at byte offset voff16[s]*16, its validity bitmap words at boff16[s]*16).
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
    const uint8_t *stage;
    const uint16_t *voff16;
    const uint16_t *boff16;
    int r0, stride; // row in tile = r0 + j*stride
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

template <int K, bool NULLS, typename Rows>
__device__ __forceinline__ void load_operand(const DevProgramSet &ps, const DevOp &op, const Rows &rows,
                                             uint32_t inrange, uint32_t rownull, RowRegs<K> &b) {
    constexpr uint32_t ALL = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    if (op.src == SRC_COL) {
        const DevColRef &c = ps.cols[op.slot];
        if (!NULLS) {
            if (c.dtype == NQE_BOOL) {
#pragma unroll
                for (int j = 0; j < K; j++) b.v[j] = rows.boolbit(c, op.slot, j);
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) b.v[j] = rows.value(c, op.slot, j);
            }
            b.valid = ALL;
            return;
        }
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
        b.valid = NULLS ? inrange : ALL;
    } else { // SRC_NULL
#pragma unroll
        for (int j = 0; j < K; j++) b.v[j] = 0;
        b.valid = 0;
    }
}

#define NQE_FOR_J _Pragma("unroll") for (int j = 0; j < K; j++)
#define NQE_AS_F64(x) __longlong_as_double((long long)(x))
#define NQE_F64_BITS(x) ((uint64_t)__double_as_longlong(x))

// right-hand operand accessors: a per-row vector, or one broadcast immediate
template <int K>
struct VecOperand {
    const RowRegs<K> &r;
    __device__ __forceinline__ uint64_t operator[](int j) const { return r.v[j]; }
    __device__ __forceinline__ uint32_t valid() const { return r.valid; }
};
struct ImmOperand {
    uint64_t imm;
    uint32_t vmask;
    __device__ __forceinline__ uint64_t operator[](int) const { return imm; }
    __device__ __forceinline__ uint32_t valid() const { return vmask; }
};

// a (op) b -> a.  `live` = rows whose errors count (kept, in range, both operands valid).
template <int K, bool NULLS, typename B>
__device__ __forceinline__ void apply_binary(uint8_t code, uint8_t type, RowRegs<K> &a, const B &b,
                                             uint32_t active, uint32_t *status) {
    const uint32_t both = a.valid & b.valid();
    if (code <= NQE_OP_GT_EQ) {
#define NQE_CMP_ALL(T, CONV)                                                                      \
    switch (code) {                                                                               \
    case NQE_OP_EQ: NQE_FOR_J a.v[j] = (T)CONV(a.v[j]) == (T)CONV(b[j]); break;                 \
    case NQE_OP_NOT_EQ: NQE_FOR_J a.v[j] = (T)CONV(a.v[j]) != (T)CONV(b[j]); break;             \
    case NQE_OP_LT: NQE_FOR_J a.v[j] = (T)CONV(a.v[j]) < (T)CONV(b[j]); break;                  \
    case NQE_OP_LT_EQ: NQE_FOR_J a.v[j] = (T)CONV(a.v[j]) <= (T)CONV(b[j]); break;              \
    case NQE_OP_GT: NQE_FOR_J a.v[j] = (T)CONV(a.v[j]) > (T)CONV(b[j]); break;                  \
    default: NQE_FOR_J a.v[j] = (T)CONV(a.v[j]) >= (T)CONV(b[j]); break;                        \
    }
        if (type == T_I64) { NQE_CMP_ALL(long long, ) }
        else if (type == T_F64) { NQE_CMP_ALL(double, NQE_AS_F64) }
        else { NQE_CMP_ALL(unsigned long long, ) } // U64, BOOL (0/1)
#undef NQE_CMP_ALL
        a.valid = both;
        return;
    }
    if (code == NQE_OP_AND || code == NQE_OP_OR) {
        uint32_t at = 0, bt = 0;
        NQE_FOR_J {
            at |= (uint32_t)(a.v[j] & 1) << j;
            bt |= (uint32_t)(b[j] & 1) << j;
        }
        at &= a.valid; bt &= b.valid();
        const uint32_t af = a.valid & ~at, bf = b.valid() & ~bt;
        uint32_t val, ok;
        if (code == NQE_OP_AND) { val = at & bt; ok = both | af | bf; }
        else { val = at | bt; ok = both | at | bt; }
        NQE_FOR_J a.v[j] = (val >> j) & 1u;
        a.valid = ok;
        return;
    }
    const uint32_t live = both & active;
    if (type == T_F64) {
        switch (code) {
        case NQE_OP_PLUS: NQE_FOR_J a.v[j] = NQE_F64_BITS(__dadd_rn(NQE_AS_F64(a.v[j]), NQE_AS_F64(b[j]))); break;
        case NQE_OP_MINUS: NQE_FOR_J a.v[j] = NQE_F64_BITS(__dsub_rn(NQE_AS_F64(a.v[j]), NQE_AS_F64(b[j]))); break;
        case NQE_OP_MULTIPLY: NQE_FOR_J a.v[j] = NQE_F64_BITS(__dmul_rn(NQE_AS_F64(a.v[j]), NQE_AS_F64(b[j]))); break;
        case NQE_OP_DIVIDE:
            NQE_FOR_J {
                const double y = NQE_AS_F64(b[j]);
                if (y == 0.0 && ((live >> j) & 1u)) atomicOr(status, DEV_ERR_DIV0);
                a.v[j] = NQE_F64_BITS(NQE_AS_F64(a.v[j]) / y);
            }
            break;
        default:
            NQE_FOR_J {
                const double y = NQE_AS_F64(b[j]);
                if (y == 0.0 && ((live >> j) & 1u)) atomicOr(status, DEV_ERR_DIV0);
                a.v[j] = NQE_F64_BITS(fmod(NQE_AS_F64(a.v[j]), y));
            }
            break;
        }
    } else {
        switch (code) {
        case NQE_OP_PLUS: NQE_FOR_J a.v[j] = a.v[j] + b[j]; break; // two's complement: same bits for i64/u64
        case NQE_OP_MINUS: NQE_FOR_J a.v[j] = a.v[j] - b[j]; break;
        case NQE_OP_MULTIPLY: NQE_FOR_J a.v[j] = a.v[j] * b[j]; break;
        default:
            if (type == T_I64) {
                NQE_FOR_J {
                    const long long x = (long long)a.v[j], y = (long long)b[j];
                    const bool lv = (live >> j) & 1u;
                    unsigned long long r = 0;
                    if (y == 0) { if (lv) atomicOr(status, DEV_ERR_DIV0); }
                    else if (y == -1) {
                        if (x == LLONG_MIN) { if (lv) atomicOr(status, DEV_ERR_OVERFLOW); }
                        else r = code == NQE_OP_DIVIDE ? (unsigned long long)(-x) : 0ull;
                    } else r = (unsigned long long)(code == NQE_OP_DIVIDE ? x / y : x % y);
                    a.v[j] = r;
                }
            } else {
                NQE_FOR_J {
                    const unsigned long long x = a.v[j], y = b[j];
                    unsigned long long r = 0;
                    if (y == 0) { if ((live >> j) & 1u) atomicOr(status, DEV_ERR_DIV0); }
                    else r = code == NQE_OP_DIVIDE ? x / y : x % y;
                    a.v[j] = r;
                }
            }
            break;
        }
    }
    a.valid = both;
}

// Run program p for K rows; result in acc.  rownull = rows whose column inputs are
// forced NULL (the selection predicate was NULL, selection.rs:46).
template <int K, bool NULLS, typename Rows>
__device__ __forceinline__ void run_program_on(const DevProgramSet &ps, int p, const Rows &rows, uint32_t inrange,
                                               uint32_t active, uint32_t rownull, RowRegs<K> &acc, uint32_t *status) {
    RowRegs<K> stack[NQE_STACK];
    const int end = ps.prog_begin[p + 1];
    for (int i = ps.prog_begin[p]; i < end; i++) {
        const DevOp op = ps.ops[i];
        if (op.code == UOP_LOAD) {
            load_operand<K, NULLS>(ps, op, rows, inrange, rownull, acc);
        } else if (op.code == UOP_PUSH) {
            stack[op.slot] = acc;
        } else if (op.code >= UOP_ABS) {
            if (op.code == UOP_ABS) { NQE_FOR_J acc.v[j] &= 0x7FFFFFFFFFFFFFFFull; }
            else if (op.code == UOP_SIN) { NQE_FOR_J acc.v[j] = NQE_F64_BITS(sin(NQE_AS_F64(acc.v[j]))); }
            else { NQE_FOR_J acc.v[j] = NQE_F64_BITS(cos(NQE_AS_F64(acc.v[j]))); }
        } else if (op.src == SRC_STACK) {
            RowRegs<K> l = stack[op.slot];
            apply_binary<K, NULLS>(op.code, op.type, l, VecOperand<K>{acc}, active, status);
            acc = l;
        } else if (op.src == SRC_LIT) {
            constexpr uint32_t ALL = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
            apply_binary<K, NULLS>(op.code, op.type, acc, ImmOperand{op.imm, NULLS ? inrange : ALL}, active, status);
        } else {
            RowRegs<K> b;
            load_operand<K, NULLS>(ps, op, rows, inrange, rownull, b);
            apply_binary<K, NULLS>(op.code, op.type, acc, VecOperand<K>{b}, active, status);
        }
    }
}

// general entry (row source = HBM, nullable / ragged allowed)
template <int K>
__device__ __forceinline__ void run_program(const DevProgramSet &ps, int p, int64_t e0, int64_t stride,
                                            uint32_t inrange, uint32_t active, uint32_t rownull,
                                            RowRegs<K> &acc, uint32_t *status) {
    run_program_on<K, true>(ps, p, GlobalRows{e0, stride}, inrange, active, rownull, acc, status);
}

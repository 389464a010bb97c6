// nqe_internal.cuh -- shared host/device declarations of libnqe_b200.so
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "nqe.h"

// ---------------------------------------------------------------------------
// context / table
// ---------------------------------------------------------------------------
struct DevColumn {
    int32_t dtype = 0;
    int64_t length = 0;
    int64_t null_count = 0;      // 0 => validity == nullptr
    void *values = nullptr;      // 8-byte values | bool bitmap | utf8 offsets
    uint8_t *validity = nullptr; // LSB-first bitmap
    uint8_t *data = nullptr;     // utf8 bytes
    int64_t data_bytes = 0;
    bool owned = true;
    // library-internal: a GATHERED column.  via >= 0: row r of this column is values[ pos ] where pos is row r of column
    // `via` of the same table, a column of dtype NQE_POS32 (4-byte unsigned positions).  Only the shape-specialised
    // filter/project kernel (jit.cu) evaluates such columns; the partitioned join hands it its results this way.
    int32_t via = -1;
};
constexpr int32_t NQE_POS32 = 64; // internal dtype: 4-byte unsigned row positions (never leaves the library)

struct nqe_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr; // copy streams of the host pipeline (created on first use)
    cudaStream_t s_aux = nullptr;                  // second compute stream: independent passes of one operator overlap
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string last_error;
    int64_t launches = 0;
    double last_op_ms = 0.0;
    int timer_depth = 0;
    // small pinned scratch for result words read back after an operator
    uint64_t *h_scratch = nullptr; // pinned, 64 words
    uint64_t *d_scratch = nullptr; // device, 64 words
    // pinned staging ring for uploads/downloads
    uint8_t *h_stage = nullptr;
    size_t stage_bytes = 0;
};

struct nqe_table {
    nqe_ctx *ctx = nullptr;
    int64_t nrows = 0;
    std::vector<DevColumn> cols;
};

int32_t nqe_fail(nqe_ctx *ctx, int32_t code, const char *fmt, ...);

#define NQE_CUDA(ctx, call)                                                                     \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return nqe_fail((ctx), _e == cudaErrorMemoryAllocation ? NQE_ERR_OOM : NQE_ERR_CUDA, \
                            "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,       \
                            __LINE__, cudaGetErrorString(_e));                                  \
    } while (0)

#define NQE_TRY(call)                  \
    do {                               \
        int32_t _s = (call);           \
        if (_s != NQE_OK) return _s;   \
    } while (0)

// device memory: stream-ordered pool (cudaMallocAsync) so that operator outputs
// sized for the worst case do not pay cudaMalloc/cudaFree synchronisation.
int32_t nqe_dev_alloc(nqe_ctx *ctx, void **p, size_t bytes);
void nqe_dev_free(nqe_ctx *ctx, void *p);

static inline size_t nqe_bitmap_bytes(int64_t n) { return (size_t)(((n + 63) / 64) * 8 + 64); }
static inline size_t nqe_values_bytes(int32_t dtype, int64_t n) {
    if (dtype == NQE_BOOL) return nqe_bitmap_bytes(n);
    if (dtype == NQE_UTF8) return (size_t)(n + 1) * 4 + 64;
    return (size_t)(n > 0 ? n : 1) * 8;
}

int32_t nqe_table_new(nqe_ctx *ctx, int64_t nrows, nqe_table **out);
// allocate an owned output column of capacity n rows (validity optional)
int32_t nqe_column_alloc(nqe_ctx *ctx, int32_t dtype, int64_t n, bool with_validity, DevColumn *c);
void nqe_column_release(nqe_ctx *ctx, DevColumn *c);

// Times the kernels of one operator call.  Operators may call each other (the partitioned join
// emits through nqe_filter_project): only the outermost timer records.
struct OpTimer {
    nqe_ctx *ctx;
    explicit OpTimer(nqe_ctx *c) : ctx(c) {
        if (c->timer_depth++ == 0) cudaEventRecord(c->ev0, c->stream);
    }
    bool marked = false, stopped = false;
    ~OpTimer() { // an early return (NQE_CUDA / NQE_TRY) must not leave the context's timer nesting depth raised
        if (!stopped) --ctx->timer_depth;
    }
    // end of the timed kernels, recorded without blocking: a caller that synchronises the stream anyway
    // (to read back a row count) marks first, so that stop() finds the event complete
    void mark_end() {
        if (ctx->timer_depth == 1 && !marked) {
            cudaEventRecord(ctx->ev1, ctx->stream);
            marked = true;
        }
    }
    void stop() {
        if (stopped) return;
        stopped = true;
        if (--ctx->timer_depth > 0) return;
        if (!marked) cudaEventRecord(ctx->ev1, ctx->stream);
        cudaEventSynchronize(ctx->ev1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->last_op_ms = ms;
    }
};

// ---------------------------------------------------------------------------
// device expression programs ("accumulator machine")
//
// A PhysicalExpr tree is lowered on the host (expr_compile.cu) to a short list
// of micro-ops executed per row with the running value in registers:
//     LOAD  src          acc = src
//     PUSH  slot         stack[slot] = acc           (only for non-leaf right operands)
//     <bin> src          acc = acc (op) src          (src = column | literal)
//     <bin> STACK slot   acc = stack[slot] (op) acc
//     <un>               acc = fn(acc)
// Left-deep chains (`age + 100`, `(id + 1) > 5`) never touch the stack.
// ---------------------------------------------------------------------------
enum : uint8_t { T_BOOL = 1, T_I64 = 2, T_U64 = 3, T_F64 = 4 };
enum : uint8_t { SRC_COL = 0, SRC_LIT = 1, SRC_NULL = 2, SRC_STACK = 3 };
enum : uint8_t {
    // binary opcodes follow nqe_operator numbering 0..12
    UOP_LOAD = 32,
    UOP_PUSH = 33,
    UOP_ABS = 40,
    UOP_SIN = 41,
    UOP_COS = 42
};

struct DevOp {
    uint8_t code;
    uint8_t type; // operand type of the op (for compare/arith), T_*
    uint8_t src;
    uint8_t slot; // column slot (SRC_COL) or stack slot (SRC_STACK / PUSH)
    uint32_t pad;
    uint64_t imm;
};

constexpr int NQE_MAX_OPS = 48;   // micro-ops over all programs of one launch
constexpr int NQE_MAX_PROGS = 17; // predicate + up to 16 outputs
constexpr int NQE_MAX_COLS = 16;  // distinct input columns referenced
constexpr int NQE_STACK = 4;

struct DevColRef {
    const void *values;       // 8-byte values or bool bitmap
    const uint32_t *validity; // bitmap words or nullptr
    int32_t dtype;
    int32_t pad;
};

struct DevProgramSet {
    DevOp ops[NQE_MAX_OPS];
    DevColRef cols[NQE_MAX_COLS];
    int16_t prog_begin[NQE_MAX_PROGS + 1]; // ops of program p = [begin[p], begin[p+1])
    uint8_t prog_type[NQE_MAX_PROGS];      // result type T_*
    int32_t n_progs;
    int32_t n_cols;
    int32_t any_nulls; // some referenced column has a validity bitmap or a NULL literal is used
};

struct ExprInfo {
    int32_t result_dtype;     // nqe_dtype
    bool nullable;            // may produce NULLs
    int32_t passthrough_col;  // >= 0 if the expression is a bare column reference
};

// host: type-check (reference error behaviour) and lower expressions.
// programs[0..n) are appended to `set`; info[i] describes each.
int32_t nqe_compile_exprs(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *const *exprs, int32_t n,
                          DevProgramSet *set, ExprInfo *info);

// kernels' device status word bits
enum : uint32_t { DEV_ERR_DIV0 = 1u, DEV_ERR_OVERFLOW = 2u, DEV_ERR_TABLE_FULL = 4u, DEV_ERR_CAPACITY = 8u, DEV_ERR_RANGE = 16u };

uint64_t nqe_next_pow2(uint64_t x);

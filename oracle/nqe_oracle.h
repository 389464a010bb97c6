/*
 * nqe_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, single thread) of the physical_plan hot path of
 * Veeupup/naive-query-engine.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may link or call this.
 * The shipped product (naive-query-engine_b200/) never does.
 *
 * The reference is Rust and cannot be compiled in this image (no cargo/rustc),
 * so this file follows the reference sources function by function; each
 * function cites the reference file:line it restates.  Third-party arithmetic
 * that is not under /root/reference (arrow 13.0.0 compute kernels, twox-hash
 * 1.6.3 XxHash64, ordered-float 3.0.0) is restated from the published
 * algorithms; parity is pinned by the reference's own asserted test vectors and
 * README known answers (tests/test_oracle_golden.py).  Everything the
 * reference's tests do not assert is "parity unpinned" (see DESIGN.md).
 */
#ifndef NQE_ORACLE_H
#define NQE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { NQO_BOOL = 1, NQO_INT64 = 2, NQO_UINT64 = 3, NQO_FLOAT64 = 4 };

/* expression node kinds (postfix program) */
enum { NQO_N_COL = 0, NQO_N_LIT = 1, NQO_N_BIN = 2, NQO_N_UN = 3 };

/* Operator, reference src/logical_plan/expression.rs:335-362 (same order) */
enum {
    NQO_EQ = 0, NQO_NEQ, NQO_LT, NQO_LTEQ, NQO_GT, NQO_GTEQ,
    NQO_PLUS, NQO_MINUS, NQO_MUL, NQO_DIV, NQO_MOD, NQO_AND, NQO_OR
};
/* UnaryOperator subset that has an implementation, unary.rs:92-96 */
enum { NQO_ABS = 0, NQO_SIN = 1, NQO_COS = 2, NQO_TAN = 3 };

/* aggregate ops, aggregate/{count,sum,avg,min,max}.rs */
enum { NQO_COUNT = 0, NQO_SUM = 1, NQO_AVG = 2, NQO_MIN = 3, NQO_MAX = 4 };

/* status codes (mirror ErrorCode variants that the hot path can raise) */
enum {
    NQO_OK = 0,
    NQO_ERR_DIVIDE_BY_ZERO = 1, /* ErrorCode::ArrowError(DivideByZero) */
    NQO_ERR_INTERVAL = 2,       /* ErrorCode::IntervalError (dtype mismatch) */
    NQO_ERR_NOT_SUPPORTED = 3,  /* ErrorCode::NotSupported */
    NQO_ERR_NOT_IMPLEMENTED = 4,/* ErrorCode::NotImplemented */
    NQO_ERR_PANIC = 5           /* a Rust panic (unwrap/unimplemented!/overflow) */
};

/* A column: 8-byte values (bool: 1 byte per row), valid = byte per row or NULL */
typedef struct {
    int32_t dtype;
    int32_t _pad;
    int64_t len;
    void *values;
    uint8_t *valid;
} nqo_col;

typedef struct {
    int32_t kind;   /* NQO_N_* */
    int32_t op;     /* operator / unary fn */
    int32_t col;    /* column index for NQO_N_COL */
    int32_t dtype;  /* literal dtype */
    int32_t is_null;/* literal is None */
    int32_t _pad;
    union { int64_t i; uint64_t u; double f; } lit;
} nqo_node;

typedef struct { int32_t op; int32_t col; } nqo_agg;

uint64_t nqo_xxh64_u64(uint64_t v);
uint64_t nqo_splitmix(uint64_t seed, uint64_t i);
void nqo_gen_mod_i64(uint64_t seed, int64_t start, int64_t n, uint64_t mod, int64_t *out);
void nqo_gen_unif_f64(uint64_t seed, int64_t start, int64_t n, double scale, double *out);
void nqo_gen_perm_i64(int64_t start, int64_t n, uint64_t mul, uint64_t mod, int64_t *out);

void nqo_free_col(nqo_col *c);

int nqo_eval_expr(const nqo_col *cols, int ncols, int64_t nrows,
                  const nqo_node *prog, int nprog, nqo_col *out, char *err, int errlen);

int nqo_selection(const nqo_col *cols, int ncols, int64_t nrows, const nqo_col *mask,
                  nqo_col *out_cols, int64_t *out_rows);

int nqo_hash_join(const nqo_col *left, int nl, int64_t lrows,
                  const nqo_col *right, int nr, int64_t rrows,
                  int lkey, int rkey, nqo_col *out_cols, int64_t *out_rows);

int nqo_aggregate(const nqo_col *cols, int ncols, int64_t nrows, const nqo_col *key,
                  const nqo_agg *aggs, int naggs, nqo_col *out_cols, int64_t *out_groups,
                  char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif

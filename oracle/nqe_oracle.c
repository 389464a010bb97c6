/*
 * nqe_oracle.c -- TEST INFRASTRUCTURE ONLY (see nqe_oracle.h).
 *
 * Single-threaded CPU restatement of naive-query-engine's physical_plan
 * operators.  Structure deliberately mirrors the reference (column-at-a-time
 * expression temporaries with materialised literals, row-at-a-time builder
 * compaction, hash-of-key chained join table, per-group row-index lists and
 * per-row per-op aggregate updates) so it doubles as the timed CPU baseline.
 */
#include "nqe_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* synthetic data: SURVEY.md 8(d) splitmix64 finaliser                       */
/* ------------------------------------------------------------------------ */
static inline uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
uint64_t nqo_splitmix(uint64_t seed, uint64_t i) { return mix64(seed + i); }

void nqo_gen_mod_i64(uint64_t seed, int64_t start, int64_t n, uint64_t mod, int64_t *out) {
    for (int64_t i = 0; i < n; i++) out[i] = (int64_t)(mix64(seed + (uint64_t)(start + i)) % mod);
}
void nqo_gen_unif_f64(uint64_t seed, int64_t start, int64_t n, double scale, double *out) {
    for (int64_t i = 0; i < n; i++)
        out[i] = scale * ((double)(mix64(seed + (uint64_t)(start + i)) >> 11) * 0x1.0p-53);
}
void nqo_gen_perm_i64(int64_t start, int64_t n, uint64_t mul, uint64_t mod, int64_t *out) {
    for (int64_t i = 0; i < n; i++)
        out[i] = (int64_t)((((unsigned __int128)(uint64_t)(start + i)) * mul) % mod);
}

/* ------------------------------------------------------------------------ */
/* XXH64 of one 8-byte little-endian word, seed 0.                           */
/* twox-hash 1.6.3 XxHash64::default() + write_i64/write_u64 + finish(),     */
/* as used at hash_join.rs:68-70 and :88-90.  (Published xxHash algorithm;   */
/* cross-checked against python-xxhash in tests/test_oracle_golden.py.)      */
/* ------------------------------------------------------------------------ */
#define XP1 0x9E3779B185EBCA87ULL
#define XP2 0xC2B2AE3D27D4EB4FULL
#define XP3 0x165667B19E3779F9ULL
#define XP4 0x85EBCA77C2B2AE63ULL
#define XP5 0x27D4EB2F165667C5ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
uint64_t nqo_xxh64_u64(uint64_t v) {
    uint64_t h = 0 /*seed*/ + XP5 + 8;
    uint64_t k1 = rotl64(v * XP2, 31) * XP1;
    h ^= k1;
    h = rotl64(h, 27) * XP1 + XP4;
    h ^= h >> 33; h *= XP2;
    h ^= h >> 29; h *= XP3;
    h ^= h >> 32;
    return h;
}

/* SipHash-1-3 of one u64 with zero keys: the cost model of std HashMap's
 * default hasher (keys are random per process in Rust; results never depend
 * on them). */
static inline uint64_t siphash13_u64(uint64_t m) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
#define SIPROUND do { v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
    v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2; v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0; \
    v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); } while (0)
    v3 ^= m; SIPROUND; v0 ^= m;
    uint64_t b = 8ULL << 56;
    v3 ^= b; SIPROUND; v0 ^= b;
    v2 ^= 0xff; SIPROUND; SIPROUND; SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

/* ------------------------------------------------------------------------ */
/* helpers                                                                   */
/* ------------------------------------------------------------------------ */
static size_t esize(int dtype) { return dtype == NQO_BOOL ? 1 : 8; }

void nqo_free_col(nqo_col *c) {
    if (!c) return;
    free(c->values); free(c->valid);
    c->values = NULL; c->valid = NULL; c->len = 0;
}

static int col_alloc(nqo_col *c, int dtype, int64_t n, int with_valid) {
    c->dtype = dtype; c->len = n; c->_pad = 0;
    c->values = malloc((size_t)(n > 0 ? n : 1) * esize(dtype));
    c->valid = with_valid ? (uint8_t *)malloc((size_t)(n > 0 ? n : 1)) : NULL;
    return (c->values && (!with_valid || c->valid)) ? 0 : -1;
}

static const char *dtype_name(int d) {
    switch (d) {
    case NQO_BOOL: return "Boolean"; case NQO_INT64: return "Int64";
    case NQO_UINT64: return "UInt64"; case NQO_FLOAT64: return "Float64";
    default: return "Null";
    }
}
static const char *op_name(int op) {
    static const char *n[] = {"Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq", "Plus", "Minus",
                              "Multiply", "Divide", "Modulos", "And", "Or"};
    return (op >= 0 && op <= NQO_OR) ? n[op] : "?";
}

/* ------------------------------------------------------------------------ */
/* PhysicalExpr::evaluate -- expression/{column,literal,binary,unary}.rs     */
/* ------------------------------------------------------------------------ */

/* ColumnExpr::evaluate, column.rs:39-57: a shared reference to the input
 * column (Arc clone).  Here: a borrowed view (owned = 0). */
typedef struct { nqo_col c; int owned; } tmpcol;

static void tmp_release(tmpcol *t) { if (t->owned) nqo_free_col(&t->c); t->owned = 0; }

/* PhysicalLiteralExpr::evaluate -> ColumnValue::Const(v, n) (literal.rs:32-34)
 * then into_array (binary.rs:123-124; logical_plan/expression.rs:210-222):
 * n materialised copies; None => all-null array. */
static int materialise_literal(const nqo_node *nd, int64_t n, tmpcol *out) {
    if (col_alloc(&out->c, nd->dtype, n, nd->is_null)) return -1;
    out->owned = 1;
    if (nd->is_null) {
        memset(out->c.valid, 0, (size_t)n);
        memset(out->c.values, 0, (size_t)n * esize(nd->dtype));
        return 0;
    }
    if (nd->dtype == NQO_BOOL) {
        memset(out->c.values, nd->lit.u ? 1 : 0, (size_t)n);
    } else {
        uint64_t *v = (uint64_t *)out->c.values;
        for (int64_t i = 0; i < n; i++) v[i] = nd->lit.u;
    }
    return 0;
}

#define VALID(c, i) (!(c)->valid || (c)->valid[i])

/* arrow 13 comparison kernels eq_dyn..gt_eq_dyn (binary.rs:127-132): result
 * NULL where either side NULL; floats use IEEE partial order. */
#define CMP_LOOP(T, EXPR) do { const T *a = (const T *)l->values, *b = (const T *)r->values; \
    for (int64_t i = 0; i < n; i++) { T x = a[i], y = b[i]; o[i] = (uint8_t)(EXPR); } } while (0)

static void cmp_typed(int op, const nqo_col *l, const nqo_col *r, int64_t n, uint8_t *o) {
#define CMP_BY_OP(T) switch (op) { \
    case NQO_EQ: CMP_LOOP(T, x == y); break; case NQO_NEQ: CMP_LOOP(T, x != y); break; \
    case NQO_LT: CMP_LOOP(T, x < y); break; case NQO_LTEQ: CMP_LOOP(T, x <= y); break; \
    case NQO_GT: CMP_LOOP(T, x > y); break; case NQO_GTEQ: CMP_LOOP(T, x >= y); break; }
    switch (l->dtype) {
    case NQO_INT64: CMP_BY_OP(int64_t); break;
    case NQO_UINT64: CMP_BY_OP(uint64_t); break;
    case NQO_FLOAT64: CMP_BY_OP(double); break;
    case NQO_BOOL: CMP_BY_OP(uint8_t); break;
    }
}

static int eval_binary(int op, const nqo_col *l, const nqo_col *r, int64_t n, tmpcol *out,
                       char *err, int errlen) {
    /* binary.rs:112-119: dtypes must be identical, no coercion */
    if (l->dtype != r->dtype) {
        snprintf(err, errlen, "Cannot evaluate binary expression %s with types %s and %s",
                 op_name(op), dtype_name(l->dtype), dtype_name(r->dtype));
        return NQO_ERR_INTERVAL;
    }
    int has_null = l->valid || r->valid;
    if (op <= NQO_GTEQ) {
        if (col_alloc(&out->c, NQO_BOOL, n, has_null)) return NQO_ERR_PANIC;
        out->owned = 1;
        cmp_typed(op, l, r, n, (uint8_t *)out->c.values);
        if (has_null)
            for (int64_t i = 0; i < n; i++) {
                out->c.valid[i] = VALID(l, i) && VALID(r, i);
                if (!out->c.valid[i]) ((uint8_t *)out->c.values)[i] = 0;
            }
        return NQO_OK;
    }
    if (op == NQO_AND || op == NQO_OR) {
        /* binary_op! macro, binary.rs:30-44: Boolean x Boolean only */
        if (l->dtype != NQO_BOOL) {
            snprintf(err, errlen, "Cannot evaluate binary expression %s with types %s and %s",
                     op_name(op), dtype_name(l->dtype), dtype_name(r->dtype));
            return NQO_ERR_INTERVAL;
        }
        if (col_alloc(&out->c, NQO_BOOL, n, has_null)) return NQO_ERR_PANIC;
        out->owned = 1;
        const uint8_t *a = (const uint8_t *)l->values, *b = (const uint8_t *)r->values;
        uint8_t *o = (uint8_t *)out->c.values;
        /* arrow and_kleene / or_kleene: SQL three-valued logic */
        for (int64_t i = 0; i < n; i++) {
            int av = VALID(l, i), bv = VALID(r, i);
            int at = av && a[i], af = av && !a[i], bt = bv && b[i], bf = bv && !b[i];
            int val, ok;
            if (op == NQO_AND) { ok = (av && bv) || af || bf; val = at && bt; }
            else { ok = (av && bv) || at || bt; val = at || bt; }
            o[i] = (uint8_t)(ok ? val : 0);
            if (has_null) out->c.valid[i] = (uint8_t)ok;
        }
        return NQO_OK;
    }
    /* arithemic_op! macro, binary.rs:46-88: Int64/UInt64/Float64 else
     * unimplemented!() panic */
    if (l->dtype == NQO_BOOL) {
        snprintf(err, errlen, "not implemented");
        return NQO_ERR_PANIC;
    }
    if (col_alloc(&out->c, l->dtype, n, has_null)) return NQO_ERR_PANIC;
    out->owned = 1;
    if (has_null)
        for (int64_t i = 0; i < n; i++) out->c.valid[i] = VALID(l, i) && VALID(r, i);
    const uint8_t *ov = out->c.valid;
    /* arrow 13 arithmetic kernels: add/subtract/multiply wrap for integers;
     * divide/modulus return DivideByZero if any valid divisor is zero (floats
     * included); the remainder is truncated (sign of dividend);
     * i64::MIN / -1 and % -1 are Rust overflow panics. */
    if (l->dtype == NQO_FLOAT64) {
        const double *a = (const double *)l->values, *b = (const double *)r->values;
        double *o = (double *)out->c.values;
        for (int64_t i = 0; i < n; i++) {
            if (ov && !ov[i]) { o[i] = 0.0; continue; }
            switch (op) {
            case NQO_PLUS: o[i] = a[i] + b[i]; break;
            case NQO_MINUS: o[i] = a[i] - b[i]; break;
            case NQO_MUL: o[i] = a[i] * b[i]; break;
            case NQO_DIV:
                if (b[i] == 0.0) { snprintf(err, errlen, "Divide by zero error"); return NQO_ERR_DIVIDE_BY_ZERO; }
                o[i] = a[i] / b[i]; break;
            case NQO_MOD:
                if (b[i] == 0.0) { snprintf(err, errlen, "Divide by zero error"); return NQO_ERR_DIVIDE_BY_ZERO; }
                o[i] = fmod(a[i], b[i]); break;
            }
        }
    } else if (l->dtype == NQO_INT64) {
        const int64_t *a = (const int64_t *)l->values, *b = (const int64_t *)r->values;
        int64_t *o = (int64_t *)out->c.values;
        for (int64_t i = 0; i < n; i++) {
            if (ov && !ov[i]) { o[i] = 0; continue; }
            switch (op) {
            case NQO_PLUS: o[i] = (int64_t)((uint64_t)a[i] + (uint64_t)b[i]); break;
            case NQO_MINUS: o[i] = (int64_t)((uint64_t)a[i] - (uint64_t)b[i]); break;
            case NQO_MUL: o[i] = (int64_t)((uint64_t)a[i] * (uint64_t)b[i]); break;
            case NQO_DIV: case NQO_MOD:
                if (b[i] == 0) { snprintf(err, errlen, "Divide by zero error"); return NQO_ERR_DIVIDE_BY_ZERO; }
                if (a[i] == INT64_MIN && b[i] == -1) {
                    snprintf(err, errlen, "attempt to %s with overflow",
                             op == NQO_DIV ? "divide" : "calculate the remainder");
                    return NQO_ERR_PANIC;
                }
                o[i] = op == NQO_DIV ? a[i] / b[i] : a[i] % b[i]; break;
            }
        }
    } else {
        const uint64_t *a = (const uint64_t *)l->values, *b = (const uint64_t *)r->values;
        uint64_t *o = (uint64_t *)out->c.values;
        for (int64_t i = 0; i < n; i++) {
            if (ov && !ov[i]) { o[i] = 0; continue; }
            switch (op) {
            case NQO_PLUS: o[i] = a[i] + b[i]; break;
            case NQO_MINUS: o[i] = a[i] - b[i]; break;
            case NQO_MUL: o[i] = a[i] * b[i]; break;
            case NQO_DIV: case NQO_MOD:
                if (b[i] == 0) { snprintf(err, errlen, "Divide by zero error"); return NQO_ERR_DIVIDE_BY_ZERO; }
                o[i] = op == NQO_DIV ? a[i] / b[i] : a[i] % b[i]; break;
            }
        }
    }
    return NQO_OK;
}

/* PhysicalUnaryExpr::evaluate, unary.rs:85-108: abs/sin/cos on Float64 via
 * arity::unary (validity carried over); Tan calls cos (unary.rs:96);
 * non-float dtype => unimplemented!() panic. */
static int eval_unary(int fn, const nqo_col *c, int64_t n, tmpcol *out, char *err, int errlen) {
    if (c->dtype != NQO_FLOAT64) { snprintf(err, errlen, "not implemented"); return NQO_ERR_PANIC; }
    if (col_alloc(&out->c, NQO_FLOAT64, n, c->valid != NULL)) return NQO_ERR_PANIC;
    out->owned = 1;
    const double *a = (const double *)c->values;
    double *o = (double *)out->c.values;
    for (int64_t i = 0; i < n; i++) {
        switch (fn) {
        case NQO_ABS: o[i] = fabs(a[i]); break;
        case NQO_SIN: o[i] = sin(a[i]); break;
        default: o[i] = cos(a[i]); break; /* Cos and Tan */
        }
    }
    if (c->valid) memcpy(out->c.valid, c->valid, (size_t)n);
    return NQO_OK;
}

int nqo_eval_expr(const nqo_col *cols, int ncols, int64_t nrows, const nqo_node *prog,
                  int nprog, nqo_col *out, char *err, int errlen) {
    tmpcol stack[64];
    int sp = 0, rc = NQO_OK;
    if (errlen > 0) err[0] = 0;
    for (int p = 0; p < nprog && rc == NQO_OK; p++) {
        const nqo_node *nd = &prog[p];
        if (sp >= 63) { rc = NQO_ERR_PANIC; break; }
        switch (nd->kind) {
        case NQO_N_COL:
            if (nd->col < 0 || nd->col >= ncols) { snprintf(err, errlen, "column index out of range"); rc = NQO_ERR_PANIC; break; }
            stack[sp].c = cols[nd->col]; stack[sp].owned = 0; sp++;
            break;
        case NQO_N_LIT:
            if (materialise_literal(nd, nrows, &stack[sp])) rc = NQO_ERR_PANIC; else sp++;
            break;
        case NQO_N_BIN: {
            tmpcol r = stack[--sp], l = stack[--sp], o;
            memset(&o, 0, sizeof o);
            rc = eval_binary(nd->op, &l.c, &r.c, nrows, &o, err, errlen);
            tmp_release(&l); tmp_release(&r);
            if (rc == NQO_OK) stack[sp++] = o; else tmp_release(&o);
            break;
        }
        case NQO_N_UN: {
            tmpcol a = stack[--sp], o;
            memset(&o, 0, sizeof o);
            rc = eval_unary(nd->op, &a.c, nrows, &o, err, errlen);
            tmp_release(&a);
            if (rc == NQO_OK) stack[sp++] = o; else tmp_release(&o);
            break;
        }
        }
    }
    if (rc != NQO_OK || sp != 1) {
        for (int i = 0; i < sp; i++) tmp_release(&stack[i]);
        return rc != NQO_OK ? rc : NQO_ERR_PANIC;
    }
    /* ColumnValue::into_array: a bare column reference is returned as a copy */
    if (!stack[0].owned) {
        const nqo_col *s = &stack[0].c;
        if (col_alloc(out, s->dtype, nrows, s->valid != NULL)) return NQO_ERR_PANIC;
        memcpy(out->values, s->values, (size_t)nrows * esize(s->dtype));
        if (s->valid) memcpy(out->valid, s->valid, (size_t)nrows);
    } else {
        *out = stack[0].c;
    }
    return NQO_OK;
}

/* ------------------------------------------------------------------------ */
/* SelectionPlan::execute + build_array_by_predicate!, selection.rs:34-107   */
/* A growable builder per column, one append per kept row:                   */
/*   mask Some(true)  => append value-or-null                                */
/*   mask Some(false) => skip                                                */
/*   mask None        => append NULL  (row kept as all-NULL, selection.rs:46)*/
/* ------------------------------------------------------------------------ */
typedef struct { uint8_t *vals; uint8_t *valid; int64_t len, cap; size_t es; int any_null; } builder;

static void builder_init(builder *b, size_t es, int64_t cap) {
    b->es = es; b->cap = cap > 0 ? cap : 1; b->len = 0; b->any_null = 0;
    b->vals = (uint8_t *)malloc((size_t)b->cap * es);
    b->valid = (uint8_t *)malloc((size_t)b->cap);
}
static inline void builder_append(builder *b, const void *v, int is_valid) {
    if (b->len == b->cap) {
        b->cap *= 2;
        b->vals = (uint8_t *)realloc(b->vals, (size_t)b->cap * b->es);
        b->valid = (uint8_t *)realloc(b->valid, (size_t)b->cap);
    }
    if (is_valid) memcpy(b->vals + (size_t)b->len * b->es, v, b->es);
    else { memset(b->vals + (size_t)b->len * b->es, 0, b->es); b->any_null = 1; }
    b->valid[b->len] = (uint8_t)is_valid;
    b->len++;
}

int nqo_selection(const nqo_col *cols, int ncols, int64_t nrows, const nqo_col *mask,
                  nqo_col *out_cols, int64_t *out_rows) {
    if (mask->dtype != NQO_BOOL) return NQO_ERR_PANIC; /* downcast_ref::<BooleanArray>().unwrap() */
    const uint8_t *m = (const uint8_t *)mask->values;
    int64_t n = nrows < mask->len ? nrows : mask->len; /* zip truncates */
    for (int c = 0; c < ncols; c++) {
        const nqo_col *col = &cols[c];
        size_t es = esize(col->dtype);
        builder b;
        builder_init(&b, es, nrows);
        const uint8_t *v = (const uint8_t *)col->values;
        for (int64_t i = 0; i < n; i++) {
            if (VALID(mask, i)) {
                if (m[i]) builder_append(&b, v + (size_t)i * es, VALID(col, i));
            } else {
                builder_append(&b, NULL, 0);
            }
        }
        out_cols[c].dtype = col->dtype; out_cols[c]._pad = 0; out_cols[c].len = b.len;
        out_cols[c].values = b.vals;
        if (b.any_null) out_cols[c].valid = b.valid; else { free(b.valid); out_cols[c].valid = NULL; }
        *out_rows = b.len;
    }
    if (ncols == 0) *out_rows = 0;
    return NQO_OK;
}

/* ------------------------------------------------------------------------ */
/* HashJoin::build / probe, hash_join.rs:58-103,124-254                      */
/* hashtable: HashMap<u64 = XXH64(key), Vec<usize>>.  Validity of the key  */
/* columns is ignored (left_col.value(i), :67,:86).  Probe: right-row-major, */
/* chain order = build-row ascending, real-value equality check (:95).       */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t key; int64_t *rows; int32_t len, cap; int used; } hbucket;
typedef struct { hbucket *b; uint64_t cap, count; } hmap;

static void hmap_init(hmap *m, uint64_t cap) {
    m->cap = 16; while (m->cap < cap) m->cap <<= 1;
    m->count = 0;
    m->b = (hbucket *)calloc(m->cap, sizeof(hbucket));
}
static void hmap_free(hmap *m) {
    for (uint64_t i = 0; i < m->cap; i++) if (m->b[i].used) free(m->b[i].rows);
    free(m->b);
}
static hbucket *hmap_find(hmap *m, uint64_t key, int insert);
static void hmap_grow(hmap *m) {
    hmap n; hmap_init(&n, m->cap * 2);
    for (uint64_t i = 0; i < m->cap; i++) if (m->b[i].used) {
        hbucket *d = hmap_find(&n, m->b[i].key, 1);
        d->rows = m->b[i].rows; d->len = m->b[i].len; d->cap = m->b[i].cap;
    }
    free(m->b); *m = n;
}
static hbucket *hmap_find(hmap *m, uint64_t key, int insert) {
    if (insert && (m->count + 1) * 8 > m->cap * 7) hmap_grow(m);
    uint64_t h = siphash13_u64(key), i = h & (m->cap - 1);
    for (;;) {
        hbucket *b = &m->b[i];
        if (!b->used) {
            if (!insert) return NULL;
            b->used = 1; b->key = key; b->rows = NULL; b->len = b->cap = 0; m->count++;
            return b;
        }
        if (b->key == key) return b;
        i = (i + 1) & (m->cap - 1);
    }
}
static inline void bucket_push(hbucket *b, int64_t row) {
    if (b->len == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 4; /* Rust Vec growth: 0 -> 4 -> 8 ... */
        b->rows = (int64_t *)realloc(b->rows, sizeof(int64_t) * (size_t)b->cap);
    }
    b->rows[b->len++] = row;
}

/* arrow compute::take over one column (hash_join.rs:236-246) */
static void take_col(const nqo_col *src, const int64_t *idx, int64_t n, nqo_col *dst) {
    col_alloc(dst, src->dtype, n, src->valid != NULL);
    size_t es = esize(src->dtype);
    if (es == 8) {
        const uint64_t *s = (const uint64_t *)src->values; uint64_t *d = (uint64_t *)dst->values;
        for (int64_t i = 0; i < n; i++) d[i] = s[idx[i]];
    } else {
        const uint8_t *s = (const uint8_t *)src->values; uint8_t *d = (uint8_t *)dst->values;
        for (int64_t i = 0; i < n; i++) d[i] = s[idx[i]];
    }
    if (src->valid) {
        int any = 0;
        for (int64_t i = 0; i < n; i++) { dst->valid[i] = src->valid[idx[i]]; any |= !dst->valid[i]; }
        if (!any) { free(dst->valid); dst->valid = NULL; }
    }
}

int nqo_hash_join(const nqo_col *left, int nl, int64_t lrows, const nqo_col *right, int nr,
                  int64_t rrows, int lkey, int rkey, nqo_col *out_cols, int64_t *out_rows) {
    const nqo_col *lk = &left[lkey], *rk = &right[rkey];
    /* hash_join.rs:139-162: Int64 / UInt64 (Utf8 handled by the Python oracle) */
    if (lk->dtype != NQO_INT64 && lk->dtype != NQO_UINT64) return NQO_ERR_NOT_IMPLEMENTED;
    if (rk->dtype != NQO_INT64 && rk->dtype != NQO_UINT64) return NQO_ERR_NOT_IMPLEMENTED;
    if (lk->dtype != rk->dtype) return NQO_ERR_PANIC; /* downcast .unwrap() on None */
    const uint64_t *lv = (const uint64_t *)lk->values, *rv = (const uint64_t *)rk->values;
    hmap m; hmap_init(&m, 16);
    /* build_match!, :58-78 */
    for (int64_t i = 0; i < lrows; i++) {
        uint64_t h = nqo_xxh64_u64(lv[i]);
        bucket_push(hmap_find(&m, h, 1), i);
    }
    /* probe_match!, :80-103: two Int64Builders */
    int64_t cap = lrows > 16 ? lrows : 16, n = 0;
    int64_t *outer = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    int64_t *inner = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    for (int64_t i = 0; i < rrows; i++) {
        uint64_t h = nqo_xxh64_u64(rv[i]);
        hbucket *b = hmap_find(&m, h, 0);
        if (!b) continue;
        for (int32_t j = 0; j < b->len; j++) {
            if (lv[b->rows[j]] == rv[i]) {
                if (n == cap) {
                    cap *= 2;
                    outer = (int64_t *)realloc(outer, sizeof(int64_t) * (size_t)cap);
                    inner = (int64_t *)realloc(inner, sizeof(int64_t) * (size_t)cap);
                }
                outer[n] = b->rows[j]; inner[n] = i; n++;
            }
        }
    }
    /* take all left columns then all right columns, :236-246 */
    for (int c = 0; c < nl; c++) take_col(&left[c], outer, n, &out_cols[c]);
    for (int c = 0; c < nr; c++) take_col(&right[c], inner, n, &out_cols[nl + c]);
    *out_rows = n;
    free(outer); free(inner); hmap_free(&m);
    return NQO_OK;
}

/* ------------------------------------------------------------------------ */
/* AggregateOperator impls: count.rs:61-77, sum.rs:40-47,103-112,            */
/* avg.rs (f64 sum + u32 cnt), max.rs / min.rs (OrderedFloat<f64>, initial   */
/* f64::MIN / f64::MAX).                                                     */
/* ------------------------------------------------------------------------ */
typedef struct { int op; const nqo_col *col; double f; uint64_t cnt; uint32_t cnt32; } aggstate;

static void agg_clear(aggstate *s) {
    s->cnt = 0; s->cnt32 = 0;
    s->f = s->op == NQO_MAX ? -DBL_MAX : (s->op == NQO_MIN ? DBL_MAX : 0.0);
}
static inline double as_f64(const nqo_col *c, int64_t i) {
    switch (c->dtype) { /* Rust `as f64` */
    case NQO_INT64: return (double)((const int64_t *)c->values)[i];
    case NQO_UINT64: return (double)((const uint64_t *)c->values)[i];
    default: return ((const double *)c->values)[i];
    }
}
/* OrderedFloat total order: NaN == NaN, NaN greater than everything */
static inline int of_gt(double a, double b) { return isnan(a) ? !isnan(b) : (!isnan(b) && a > b); }
static inline int of_lt(double a, double b) { return isnan(b) ? !isnan(a) : (!isnan(a) && a < b); }

static inline void agg_update(aggstate *s, int64_t i) {
    const nqo_col *c = s->col;
    if (!VALID(c, i)) return;
    switch (s->op) {
    case NQO_COUNT: s->cnt++; break;
    case NQO_SUM: s->f += as_f64(c, i); break;
    case NQO_AVG: s->f += as_f64(c, i); s->cnt32++; break;
    case NQO_MAX: { double v = as_f64(c, i); if (of_gt(v, s->f)) s->f = v; break; }
    case NQO_MIN: { double v = as_f64(c, i); if (of_lt(v, s->f)) s->f = v; break; }
    }
}
static inline void agg_eval(const aggstate *s, nqo_col *out, int64_t g) {
    if (s->op == NQO_COUNT) ((uint64_t *)out->values)[g] = s->cnt;
    else if (s->op == NQO_AVG) ((double *)out->values)[g] = s->f / (double)s->cnt32;
    else ((double *)out->values)[g] = s->f;
}

/* PhysicalAggregatePlan::execute, aggregate/mod.rs:113-222.
 * key == NULL: global path (:123-139).  Otherwise group_by_datatype! (:54-102):
 * HashMap<key, Vec<row idx>>, NULL keys dropped (:63-71), then per group, per
 * row (ascending), per op update(); evaluate; clear_state.  The reference emits
 * groups in std-HashMap order (random per process); this oracle emits them in
 * first-appearance order and tests compare as multisets. */
int nqo_aggregate(const nqo_col *cols, int ncols, int64_t nrows, const nqo_col *key,
                  const nqo_agg *aggs, int naggs, nqo_col *out_cols, int64_t *out_groups,
                  char *err, int errlen) {
    aggstate st[64];
    if (naggs > 64) return NQO_ERR_PANIC;
    if (errlen > 0) err[0] = 0;
    for (int a = 0; a < naggs; a++) {
        if (aggs[a].col < 0 || aggs[a].col >= ncols) return NQO_ERR_PANIC;
        st[a].op = aggs[a].op; st[a].col = &cols[aggs[a].col];
        if (aggs[a].op != NQO_COUNT && st[a].col->dtype == NQO_BOOL) {
            /* update_batch: Err(NotSupported) (sum.rs:91-96); update: unimplemented!() */
            static const char *fn[] = {"", "Sum", "Avg", "min", "Max"};
            snprintf(err, errlen, "%s func for Boolean is not supported", fn[aggs[a].op]);
            return key ? NQO_ERR_PANIC : NQO_ERR_NOT_SUPPORTED;
        }
        agg_clear(&st[a]);
    }
    if (!key) {
        for (int a = 0; a < naggs; a++) {
            col_alloc(&out_cols[a], aggs[a].op == NQO_COUNT ? NQO_UINT64 : NQO_FLOAT64, 1, 0);
            for (int64_t i = 0; i < nrows; i++) agg_update(&st[a], i);
            agg_eval(&st[a], &out_cols[a], 0);
        }
        *out_groups = 1;
        return NQO_OK;
    }
    if (key->dtype != NQO_INT64 && key->dtype != NQO_UINT64) {
        snprintf(err, errlen, "group by only support by `Int64`, `UInt64`, `String`");
        return NQO_ERR_NOT_SUPPORTED;
    }
    const uint64_t *kv = (const uint64_t *)key->values;
    hmap m; hmap_init(&m, 16);
    /* remember first-appearance order of groups */
    int64_t ocap = 1024, ng = 0;
    uint64_t *order = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)ocap);
    for (int64_t i = 0; i < nrows; i++) {
        if (!VALID(key, i)) continue;
        uint64_t before = m.count;
        hbucket *b = hmap_find(&m, kv[i], 1);
        if (m.count != before) {
            if (ng == ocap) { ocap *= 2; order = (uint64_t *)realloc(order, sizeof(uint64_t) * (size_t)ocap); }
            order[ng++] = kv[i];
        }
        bucket_push(b, i);
    }
    for (int a = 0; a < naggs; a++)
        col_alloc(&out_cols[a], aggs[a].op == NQO_COUNT ? NQO_UINT64 : NQO_FLOAT64, ng, 0);
    for (int64_t g = 0; g < ng; g++) {
        hbucket *b = hmap_find(&m, order[g], 0);
        for (int32_t j = 0; j < b->len; j++)
            for (int a = 0; a < naggs; a++) agg_update(&st[a], b->rows[j]);
        for (int a = 0; a < naggs; a++) { agg_eval(&st[a], &out_cols[a], g); agg_clear(&st[a]); }
    }
    *out_groups = ng;
    free(order); hmap_free(&m);
    return NQO_OK;
}

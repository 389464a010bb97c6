/*
 * nqe.h -- C ABI of libnqe_b200.so, the B200-native (sm_100a) execution layer
 * for naive-query-engine's physical_plan operator pipeline
 * (scan -> filter -> projection -> hash-join -> hash-aggregate over Arrow
 * RecordBatch columns).
 *
 * The reference (Rust, /root/reference) has no FFI of its own: the seam is the
 * trait-object interface `PhysicalPlan::execute()` (src/physical_plan/plan.rs:14-23).
 * Each entry point below is what a Rust `impl PhysicalPlan for Gpu*Plan` binds
 * (see INTEGRATION.md for the `extern "C"` block and the QueryPlanner patch);
 * the comment on each function cites the reference code it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; every call returns an nqe_status (0 = OK);
 *    the message of the last error on a context is nqe_last_error(ctx).
 *    No exception or abort crosses this boundary: reference *panics*
 *    (unwrap/unimplemented!/overflow) are reported as NQE_ERR_PANIC.
 *  - host columns use the Arrow columnar layout: little-endian 8-byte values,
 *    optional LSB-first validity bitmap (NULL pointer when null_count == 0),
 *    Boolean = LSB-first value bitmap, Utf8 = int32 offsets[length+1] + bytes.
 *  - a context is thread-compatible (one caller at a time), like the
 *    single-threaded reference; all work of a context runs on one CUDA stream.
 *  - there is NO CPU fallback: without a CUDA device every call that needs one
 *    fails with NQE_ERR_CUDA.
 */
#ifndef NQE_H
#define NQE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQE_ABI_VERSION 1

typedef struct nqe_ctx nqe_ctx;
typedef struct nqe_table nqe_table; /* a device-resident RecordBatch */

/* arrow::datatypes::DataType subset the reference's operators accept
 * (selection.rs:70-98, binary.rs:46-88, hash_join.rs:139-162) */
typedef enum nqe_dtype {
    NQE_BOOL = 1,
    NQE_INT64 = 2,
    NQE_UINT64 = 3,
    NQE_FLOAT64 = 4,
    NQE_UTF8 = 5
} nqe_dtype;

/* mirrors ErrorCode, src/error.rs:13-40, for the variants the hot path raises */
typedef enum nqe_status {
    NQE_OK = 0,
    NQE_ERR_DIVIDE_BY_ZERO = 1,  /* ErrorCode::ArrowError(ArrowError::DivideByZero) */
    NQE_ERR_INTERVAL = 2,        /* ErrorCode::IntervalError  (binary.rs:114-119)   */
    NQE_ERR_NOT_SUPPORTED = 3,   /* ErrorCode::NotSupported   (aggregate/mod.rs:217)*/
    NQE_ERR_NOT_IMPLEMENTED = 4, /* ErrorCode::NotImplemented (hash_join.rs:161)    */
    NQE_ERR_PANIC = 5,           /* a reference panic: unwrap / unimplemented! / overflow */
    NQE_ERR_PLAN = 6,            /* ErrorCode::PlanError      (hash_join.rs:125-129)*/
    NQE_ERR_LOGICAL = 7,         /* ErrorCode::LogicalError   (column.rs:24-28)     */
    NQE_ERR_INVALID_ARG = 8,
    NQE_ERR_CUDA = 9,
    NQE_ERR_OOM = 10
} nqe_status;

/* One Arrow array (host or device memory, see the call). */
typedef struct nqe_column_desc {
    int32_t dtype;           /* nqe_dtype */
    int32_t reserved;
    int64_t length;
    int64_t null_count;      /* 0 => validity may be NULL */
    const void *values;      /* 8-byte values | BOOL: bitmap | UTF8: int32 offsets[length+1] */
    const uint8_t *validity; /* LSB-first bitmap or NULL */
    const uint8_t *data;     /* UTF8 bytes, else NULL */
    int64_t data_bytes;
} nqe_column_desc;

/* ---- expressions: PhysicalExpr trees (expression/mod.rs:25-31) as postfix programs */
typedef enum nqe_node_kind {
    NQE_NODE_COLUMN = 0,  /* ColumnExpr by index            (column.rs:39-44)  */
    NQE_NODE_LITERAL = 1, /* PhysicalLiteralExpr(ScalarValue) (literal.rs:32-34) */
    NQE_NODE_BINARY = 2,  /* PhysicalBinaryExpr             (binary.rs:108-155)*/
    NQE_NODE_UNARY = 3    /* PhysicalUnaryExpr              (unary.rs:85-108)  */
} nqe_node_kind;

/* Operator, src/logical_plan/expression.rs:335-362 (same order) */
typedef enum nqe_operator {
    NQE_OP_EQ = 0, NQE_OP_NOT_EQ, NQE_OP_LT, NQE_OP_LT_EQ, NQE_OP_GT, NQE_OP_GT_EQ,
    NQE_OP_PLUS, NQE_OP_MINUS, NQE_OP_MULTIPLY, NQE_OP_DIVIDE, NQE_OP_MODULOS,
    NQE_OP_AND, NQE_OP_OR
} nqe_operator;

/* UnaryOperator, src/logical_plan/expression.rs:392-422: the four that have an
 * implementation (unary.rs:92-96; Tan evaluates cos, as the reference does). */
typedef enum nqe_unary_fn { NQE_FN_ABS = 0, NQE_FN_SIN = 1, NQE_FN_COS = 2, NQE_FN_TAN = 3 } nqe_unary_fn;

typedef struct nqe_expr_node {
    int32_t kind;    /* nqe_node_kind */
    int32_t op;      /* nqe_operator or nqe_unary_fn */
    int32_t column;  /* NQE_NODE_COLUMN: index into the input table */
    int32_t dtype;   /* NQE_NODE_LITERAL: nqe_dtype of the ScalarValue */
    int32_t is_null; /* NQE_NODE_LITERAL: ScalarValue::X(None) */
    int32_t reserved; /* NQE_NODE_LITERAL of dtype NQE_UTF8: byte length of the string */
    union { int64_t i64; uint64_t u64; double f64; } value; /* BOOL literal: u64 0/1; UTF8 literal: u64 = address of the
                                                             * (host) bytes, valid during the call.  Utf8 values may only be
                                                             * compared (Eq .. GtEq, binary.rs:127-132) in nqe_filter_project */
} nqe_expr_node;

typedef struct nqe_expr {
    const nqe_expr_node *nodes; /* post-order */
    int32_t n_nodes;
    int32_t reserved;
} nqe_expr;

/* AggregateOperator implementations, aggregate/{count,sum,avg,min,max}.rs.
 * The argument is a bare column (planner/mod.rs:104-163). */
typedef enum nqe_agg_op {
    NQE_AGG_COUNT = 0, NQE_AGG_SUM = 1, NQE_AGG_AVG = 2, NQE_AGG_MIN = 3, NQE_AGG_MAX = 4,
    /* EXTENSION (not in the reference, whose aggregate output carries no key column, aggregate/mod.rs:117-121): emits the
     * group key itself as an Int64 column (UInt64 keys: same bits); `column` is ignored.  Grouped plans only.  The
     * multi-GPU plans need it to merge partial states by key. */
    NQE_AGG_GROUP_KEY = 5
} nqe_agg_op;
typedef struct nqe_agg { int32_t op; int32_t column; } nqe_agg;

/* ---- context ------------------------------------------------------------ */
int32_t nqe_abi_version(void);
int32_t nqe_device_count(void);
int32_t nqe_ctx_create(int32_t device, nqe_ctx **out);
void nqe_ctx_destroy(nqe_ctx *ctx);
const char *nqe_last_error(const nqe_ctx *ctx);
/* run this context's work on an externally owned cudaStream_t (0 = own stream) */
int32_t nqe_ctx_set_stream(nqe_ctx *ctx, void *cuda_stream);
int32_t nqe_ctx_sync(nqe_ctx *ctx);
/* number of kernels this context has launched so far (bench.py gpu_launches) */
int64_t nqe_ctx_kernel_launches(const nqe_ctx *ctx);
/* elapsed device ms of the most recent operator call's kernels (CUDA events on the ctx stream) */
double nqe_ctx_last_op_ms(const nqe_ctx *ctx);

/* ---- tables: ScanPlan / TableSource::scan (scan.rs:34-36, memory.rs:31-41) --- */
/* Host Arrow buffers -> pinned staging -> HBM.  Replaces the Arc-clone of
 * MemTable batches with a DMA.  * The call returns after the copies have completed: the host buffers may be released or
 * modified as soon as it returns (pinned sources are DMA'd directly, pageable ones staged). */
int32_t nqe_table_upload(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, nqe_table **out);
/* Wrap device buffers owned by the caller (no copy; must outlive the table). */
int32_t nqe_table_from_device(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, nqe_table **out);
int64_t nqe_table_num_rows(const nqe_table *t);
int32_t nqe_table_num_columns(const nqe_table *t);
/* describe column i; pointers are DEVICE pointers, valid while the table lives */
int32_t nqe_table_column(const nqe_table *t, int32_t i, nqe_column_desc *out);
/* copy column i into caller-provided host buffers (any may be NULL to skip);
 * values_bytes / validity_bytes / data_bytes are the buffer capacities */
int32_t nqe_table_download_column(nqe_ctx *ctx, const nqe_table *t, int32_t i, void *values,
                                  int64_t values_bytes, uint8_t *validity, int64_t validity_bytes,
                                  uint8_t *data, int64_t data_bytes);
void nqe_table_free(nqe_table *t);
/* zero-copy row slice [offset, offset+len): PhysicalOffsetPlan / PhysicalLimitPlan
 * (offset.rs:30-51, limit.rs:32-49).  offset must be a multiple of 8 unless the
 * table has no bitmaps; otherwise a copy is made. */
int32_t nqe_table_slice(nqe_ctx *ctx, const nqe_table *t, int64_t offset, int64_t len, nqe_table **out);
/* concat_batches (hash_join.rs:258-273; used on the join's build side, hash_join.rs:131-132, and on the aggregate's
 * input, aggregate/mod.rs:143-144): the rows of tables[0..n) in order, as one new table.  Column counts and dtypes
 * must agree (else NQE_ERR_INVALID_ARG); n_tables >= 1 (an empty result needs a schema: pass an empty
 * table).  The inputs are not consumed. */
int32_t nqe_table_concat(nqe_ctx *ctx, const nqe_table *const *tables, int32_t n_tables, nqe_table **out);

/* ---- operators ----------------------------------------------------------- */
/* SelectionPlan::execute (selection.rs:58-107) fused with ProjectionPlan::execute
 * (projection.rs:43-70): out = projection(selection(in, predicate), projs).
 * predicate == NULL => pure projection; n_projs == 0 with a predicate => every
 * input column is passed through (a bare SelectionPlan).  Rows whose predicate
 * is NULL are kept as all-NULL rows (selection.rs:46).  Stable (input order). */
int32_t nqe_filter_project(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate,
                           const nqe_expr *projs, int32_t n_projs, nqe_table **out);

/* The same fused SelectionPlan + ProjectionPlan with HOST Arrow columns in and HOST buffers out -- what
 * `execute()` hands back to the reference's caller (plan.rs:14-23, db.rs:36).  Rows are streamed through HBM
 * in chunks on three CUDA streams (H2D of chunk c+1 | kernel of chunk c | D2H of chunk c-1); only the columns
 * the expressions read are uploaded.  out_cols[i].values must point to a caller-owned buffer of at least
 * cols[0].length rows (out_cols[i].length = its capacity in rows); dtype / null_count are filled in.
 * Pinned (cudaHostAlloc / cudaHostRegister) buffers are copied asynchronously; other inputs take the
 * non-overlapped upload -> operator -> download path. */
int32_t nqe_filter_project_host(nqe_ctx *ctx, const nqe_column_desc *cols, int32_t n_cols, const nqe_expr *predicate,
                                const nqe_expr *projs, int32_t n_projs, nqe_column_desc *out_cols, int64_t *out_rows);

/* HashJoin::execute = build + probe (hash_join.rs:124-254).  left = build side.
 * Inner join on one Int64/UInt64/Utf8 key pair; key validity is ignored (:67,:86);
 * output = all left columns ++ all right columns, probe-row-major, build rows
 * ascending within one probe row. */
int32_t nqe_hash_join(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right,
                      int32_t left_key, int32_t right_key, nqe_table **out);

/* PhysicalAggregatePlan::execute (aggregate/mod.rs:113-222).  group_expr == NULL
 * => one global row; otherwise group by that single expression (Int64/UInt64, or a bare Utf8 column),
 * NULL keys dropped, NO key column in the output, group order unspecified
 * (the reference's is std-HashMap order). */
int32_t nqe_hash_aggregate(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *group_expr,
                           const nqe_agg *aggs, int32_t n_aggs, nqe_table **out);

/* HashJoin feeding PhysicalAggregatePlan without materialising the join:
 * group_column / aggs[].column index the join's output schema (left ++ right). */
int32_t nqe_join_aggregate(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right,
                           int32_t left_key, int32_t right_key, int32_t group_column,
                           const nqe_agg *aggs, int32_t n_aggs, nqe_table **out);

/* ---- several GPUs of one node, ONE process (nqe_ctx_create(devices, n); 8e plans ii and iv) --------
 * The reference is a single-process engine, so this is the shape its planner can bind: the shards of a table in, one
 * result out.  A nqe_multi owns one nqe_ctx per member GPU (member i = devices[i]; the same device may be listed more
 * than once); every member's share of a plan runs on its own host thread, tables move between members as peer copies
 * over NVLink, results land on member 0.  Tables handed in must have been created through the member's own context
 * (nqe_multi_ctx).  (One process per GPU over torch.distributed / NCCL: naive-query-engine_b200/distributed.py.) */
typedef struct nqe_multi nqe_multi;
int32_t nqe_multi_create(const int32_t *devices, int32_t n, nqe_multi **out);
void nqe_multi_destroy(nqe_multi *m);
int32_t nqe_multi_size(const nqe_multi *m);
nqe_ctx *nqe_multi_ctx(nqe_multi *m, int32_t member);
const char *nqe_multi_last_error(const nqe_multi *m);
/* a copy of `src` (any member's table) on member dst_member */
int32_t nqe_multi_table_copy(nqe_multi *m, const nqe_table *src, int32_t dst_member, nqe_table **out);
/* HashJoin feeding PhysicalAggregatePlan (hash_join.rs:124-254, aggregate/mod.rs:113-222) with the probe side sharded:
 * `left` (build side, on any member) is copied to every member, member i joins it with right[i] (NULL = no shard) and
 * pre-aggregates, the partial states are merged on member 0.  Same arguments and output as nqe_join_aggregate over
 * the concatenation of the shards (sums within float re-association, group order unspecified). */
int32_t nqe_multi_join_aggregate(nqe_multi *m, const nqe_table *left, const nqe_table *const *right, int32_t left_key,
                                 int32_t right_key, int32_t group_column, const nqe_agg *aggs, int32_t n_aggs,
                                 nqe_table **out);
/* PhysicalAggregatePlan over sharded input: in[i] on member i (NULL = no shard); group_expr must not be NULL. */
int32_t nqe_multi_hash_aggregate(nqe_multi *m, const nqe_table *const *in, const nqe_expr *group_expr,
                                 const nqe_agg *aggs, int32_t n_aggs, nqe_table **out);

/* ---- multi-GPU helper: radix partition on the key ------------------------ */
/* Splits `in` into n_parts contiguous row ranges by mix64(key) % n_parts (the
 * shuffle step before an all-to-all).  out has the same schema, rows grouped by
 * destination; counts[p] receives the rows for destination p (host array). */
int32_t nqe_radix_partition(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts,
                            nqe_table **out, int64_t *counts);

/* The same partition fused with the exchange (one process per GPU, NVLink 5 / NVSwitch peer memory):
 * nqe_partition_counts returns how many rows of `in` go to each destination; after the ranks have exchanged
 * those counts (an all-gather of n_parts words) every rank knows where its rows start in each receive
 * buffer, and nqe_shuffle_scatter stores the rows straight into the destination GPUs' buffers.
 * dst_columns[c * n_parts + p] = column c's receive buffer on rank p as mapped into THIS process (CUDA IPC /
 * symmetric memory; for p == own rank a local pointer), dst_offsets[p] = first row this rank writes there.
 * The caller synchronises all ranks (stream barrier) before reading its receive buffers.  n_parts, columns <= 8. */
int32_t nqe_partition_counts(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts, int64_t *counts);
int32_t nqe_shuffle_scatter(nqe_ctx *ctx, const nqe_table *in, int32_t key_column, int32_t n_parts,
                            void *const *dst_columns, const int64_t *dst_offsets);

/* ---- synthetic benchmark columns, generated in HBM -------- */
/* kind 0: mix64(seed+i) % mod  (Int64);  kind 1: scale*unif01(mix64(seed+i)) (Float64);
 * kind 2: (i * mul) % mod (Int64), i = start .. start+n-1 */
int32_t nqe_synth_column(nqe_ctx *ctx, int32_t kind, uint64_t seed, int64_t start, int64_t n,
                         uint64_t mod_or_mul, uint64_t mod2, double scale, void *device_out);

#ifdef __cplusplus
}
#endif
#endif /* NQE_H */

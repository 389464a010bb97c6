// expr_compile.cu -- host side: type-check PhysicalExpr programs with the
// reference's error behaviour and lower them to device micro-ops.
//
// Reference semantics restated here:
//   binary.rs:112-119   operand dtypes must be identical (no coercion) else IntervalError
//   binary.rs:30-44     And/Or only on Boolean x Boolean else IntervalError
//   binary.rs:46-88     arithmetic only on Int64/UInt64/Float64 else unimplemented!() panic
//   unary.rs:20-44      abs/sin/cos only on Float64 else unimplemented!() panic; Tan -> cos (:96)
//   logical_plan/expression.rs:236-331  result dtype of arithmetic = left operand dtype
#include <cstring>

#include "nqe_internal.cuh"

static const char *dtype_name(int d) {
    switch (d) {
    case NQE_BOOL: return "Boolean";
    case NQE_INT64: return "Int64";
    case NQE_UINT64: return "UInt64";
    case NQE_FLOAT64: return "Float64";
    case NQE_UTF8: return "Utf8";
    default: return "Null";
    }
}
static const char *op_name(int op) {
    static const char *n[] = {"Eq", "NotEq", "Lt", "LtEq", "Gt", "GtEq", "Plus", "Minus",
                              "Multiply", "Divide", "Modulos", "And", "Or"};
    return (op >= 0 && op <= NQE_OP_OR) ? n[op] : "?";
}

namespace {

struct Node {
    int kind, op, dtype; // resolved dtype
    int col;             // table column
    bool is_null;
    uint64_t imm;
    int left = -1, right = -1;
    bool nullable = false;
};

struct Compiler {
    nqe_ctx *ctx;
    const nqe_table *in;
    DevProgramSet *set;
    std::vector<Node> nodes;
    int n_ops;

    int col_slot(int col) {
        const DevColumn &c = in->cols[col];
        for (int s = 0; s < set->n_cols; s++)
            if (set->cols[s].values == c.values && set->cols[s].dtype == c.dtype) return s;
        if (set->n_cols >= NQE_MAX_COLS) return -1;
        int s = set->n_cols++;
        set->cols[s].values = c.values;
        set->cols[s].validity = (const uint32_t *)c.validity;
        set->cols[s].dtype = c.dtype;
        set->cols[s].pad = 0;
        if (c.validity) set->any_nulls = 1;
        return s;
    }

    int32_t emit(uint8_t code, uint8_t type, uint8_t src, uint8_t slot, uint64_t imm) {
        if (n_ops >= NQE_MAX_OPS) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "expression too large for one kernel (%d micro-ops)", NQE_MAX_OPS);
        DevOp &o = set->ops[n_ops++];
        o.code = code; o.type = type; o.src = src; o.slot = slot; o.pad = 0; o.imm = imm;
        return NQE_OK;
    }

    bool is_leaf(int n) const { return nodes[n].kind == NQE_NODE_COLUMN || nodes[n].kind == NQE_NODE_LITERAL; }

    int32_t leaf_src(int n, uint8_t *src, uint8_t *slot, uint64_t *imm) {
        const Node &nd = nodes[n];
        *imm = 0; *slot = 0;
        if (nd.kind == NQE_NODE_COLUMN) {
            int s = col_slot(nd.col);
            if (s < 0) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "more than %d distinct columns in one kernel", NQE_MAX_COLS);
            *src = SRC_COL; *slot = (uint8_t)s;
        } else if (nd.is_null) {
            *src = SRC_NULL; set->any_nulls = 1;
        } else {
            *src = SRC_LIT; *imm = nd.imm;
        }
        return NQE_OK;
    }

    // post-order lowering; depth = number of live stack slots
    int32_t lower(int n, int depth) {
        const Node &nd = nodes[n];
        uint8_t src, slot; uint64_t imm;
        if (is_leaf(n)) {
            NQE_TRY(leaf_src(n, &src, &slot, &imm));
            return emit(UOP_LOAD, (uint8_t)nd.dtype, src, slot, imm);
        }
        if (nd.kind == NQE_NODE_UNARY) {
            NQE_TRY(lower(nd.left, depth));
            uint8_t code = nd.op == NQE_FN_ABS ? UOP_ABS : nd.op == NQE_FN_SIN ? UOP_SIN : UOP_COS; // Tan -> cos
            return emit(code, T_F64, 0, 0, 0);
        }
        uint8_t optype = (uint8_t)nodes[nd.left].dtype;
        NQE_TRY(lower(nd.left, depth));
        if (is_leaf(nd.right)) {
            NQE_TRY(leaf_src(nd.right, &src, &slot, &imm));
            return emit((uint8_t)nd.op, optype, src, slot, imm);
        }
        if (depth >= NQE_STACK) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "expression nesting deeper than %d", NQE_STACK);
        NQE_TRY(emit(UOP_PUSH, optype, 0, (uint8_t)depth, 0));
        NQE_TRY(lower(nd.right, depth + 1));
        return emit((uint8_t)nd.op, optype, SRC_STACK, (uint8_t)depth, 0);
    }
};

} // namespace

int32_t nqe_compile_exprs(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *const *exprs, int32_t n,
                          DevProgramSet *set, ExprInfo *info) {
    Compiler cc{ctx, in, set, {}, 0};
    cc.n_ops = set->n_progs ? set->prog_begin[set->n_progs] : 0;
    for (int e = 0; e < n; e++) {
        const nqe_expr *ex = exprs[e];
        if (!ex || !ex->nodes || ex->n_nodes <= 0) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "empty expression");
        if (set->n_progs >= NQE_MAX_PROGS) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "too many expressions in one kernel");
        cc.nodes.clear();
        std::vector<int> stack;
        for (int i = 0; i < ex->n_nodes; i++) {
            const nqe_expr_node &s = ex->nodes[i];
            Node nd{};
            nd.kind = s.kind; nd.op = s.op; nd.col = -1; nd.is_null = false; nd.imm = 0;
            switch (s.kind) {
            case NQE_NODE_COLUMN:
                // ColumnExpr by idx: out of range is an index panic in RecordBatch::column
                if (s.column < 0 || s.column >= (int)in->cols.size())
                    return nqe_fail(ctx, NQE_ERR_PANIC, "column index %d out of range (%zu columns)", s.column, in->cols.size());
                nd.col = s.column;
                nd.dtype = in->cols[s.column].dtype;
                nd.nullable = in->cols[s.column].validity != nullptr;
                break;
            case NQE_NODE_LITERAL:
                nd.dtype = s.dtype;
                nd.is_null = s.is_null != 0;
                nd.nullable = nd.is_null;
                nd.imm = s.dtype == NQE_BOOL ? (s.value.u64 ? 1 : 0) : s.value.u64;
                break;
            case NQE_NODE_BINARY: {
                if (stack.size() < 2) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "malformed postfix expression");
                nd.right = stack.back(); stack.pop_back();
                nd.left = stack.back(); stack.pop_back();
                int lt = cc.nodes[nd.left].dtype, rt = cc.nodes[nd.right].dtype;
                if (s.op < 0 || s.op > NQE_OP_OR) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "bad operator %d", s.op);
                if (lt != rt)
                    return nqe_fail(ctx, NQE_ERR_INTERVAL, "Cannot evaluate binary expression %s with types %s and %s",
                                    op_name(s.op), dtype_name(lt), dtype_name(rt));
                if (lt == NQE_UTF8) {
                    // comparisons of Utf8 leaves were rewritten into Boolean columns before this point (utf8.cu)
                    if (s.op == NQE_OP_AND || s.op == NQE_OP_OR) // binary.rs:30-44: and/or only on Boolean x Boolean
                        return nqe_fail(ctx, NQE_ERR_INTERVAL, "Cannot evaluate binary expression %s with types %s and %s",
                                        op_name(s.op), dtype_name(lt), dtype_name(rt));
                    if (s.op > NQE_OP_GT_EQ) return nqe_fail(ctx, NQE_ERR_PANIC, "not implemented: arithmetic on Utf8"); // binary.rs:46-88 unimplemented!()
                    return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "Utf8 comparison outside a filter/projection expression");
                }
                if (lt < NQE_BOOL || lt > NQE_FLOAT64)
                    return nqe_fail(ctx, NQE_ERR_PANIC, "binary expression on Null-typed operands");
                if (s.op <= NQE_OP_GT_EQ) nd.dtype = NQE_BOOL;
                else if (s.op == NQE_OP_AND || s.op == NQE_OP_OR) {
                    if (lt != NQE_BOOL)
                        return nqe_fail(ctx, NQE_ERR_INTERVAL, "Cannot evaluate binary expression %s with types %s and %s",
                                        op_name(s.op), dtype_name(lt), dtype_name(rt));
                    nd.dtype = NQE_BOOL;
                } else {
                    if (lt == NQE_BOOL) return nqe_fail(ctx, NQE_ERR_PANIC, "not implemented: arithmetic on Boolean");
                    nd.dtype = lt;
                }
                nd.nullable = cc.nodes[nd.left].nullable || cc.nodes[nd.right].nullable;
                break;
            }
            case NQE_NODE_UNARY: {
                if (stack.empty()) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "malformed postfix expression");
                nd.left = stack.back(); stack.pop_back();
                if (s.op < NQE_FN_ABS || s.op > NQE_FN_TAN) return nqe_fail(ctx, NQE_ERR_PANIC, "not yet implemented: unary function %d", s.op);
                if (cc.nodes[nd.left].dtype != NQE_FLOAT64)
                    return nqe_fail(ctx, NQE_ERR_PANIC, "not implemented: unary function on %s", dtype_name(cc.nodes[nd.left].dtype));
                nd.dtype = NQE_FLOAT64;
                nd.nullable = cc.nodes[nd.left].nullable;
                break;
            }
            default:
                return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "bad node kind %d", s.kind);
            }
            cc.nodes.push_back(nd);
            stack.push_back((int)cc.nodes.size() - 1);
        }
        if (stack.size() != 1) return nqe_fail(ctx, NQE_ERR_INVALID_ARG, "malformed postfix expression");
        int root = stack[0];
        const Node &r = cc.nodes[root];
        info[e].result_dtype = r.dtype;
        info[e].nullable = r.nullable;
        info[e].passthrough_col = r.kind == NQE_NODE_COLUMN ? r.col : -1;
        int p = set->n_progs;
        set->prog_begin[p] = (int16_t)cc.n_ops;
        if (r.dtype == NQE_UTF8) {
            // bare Utf8 column reference: no device program (handled by the string gather path)
            if (r.kind != NQE_NODE_COLUMN) return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "Utf8 expression");
        } else if (r.dtype < NQE_BOOL || r.dtype > NQE_FLOAT64) {
            return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "Null-typed expression");
        } else {
            NQE_TRY(cc.lower(root, 0));
        }
        set->prog_type[p] = (uint8_t)r.dtype;
        set->n_progs = p + 1;
        set->prog_begin[p + 1] = (int16_t)cc.n_ops;
    }
    return NQE_OK;
}

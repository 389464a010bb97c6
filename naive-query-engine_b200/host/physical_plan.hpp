// physical_plan.hpp -- C++ host mirror of the reference's physical_plan interface over the
// C ABI (include/nqe.h).  Header-only; link with -lnqe_b200.
//
// The reference is Rust; no Rust toolchain exists in this image, so this is the compiled-code
// host side: same type names, `create(...)` argument order, `execute()` contract and error
// kinds as
//     trait PhysicalPlan {schema, execute, children}       src/physical_plan/plan.rs:14-23
//     trait PhysicalExpr                                    src/physical_plan/expression/mod.rs:25-31
//     trait AggregateOperator                               src/physical_plan/aggregate/mod.rs:225-235
//     ScanPlan, SelectionPlan, ProjectionPlan, HashJoin, PhysicalAggregatePlan,
//     PhysicalLimitPlan, PhysicalOffsetPlan                 src/physical_plan/*.rs
// A RecordBatch here is a vector of host Arrow-layout columns (`Column`); between operators
// the data stays in HBM (`execute_device()`), `execute()` downloads like the reference returns
// `Vec<RecordBatch>`.  Two adjacent-node patterns are fused into one kernel pass:
//     ProjectionPlan(SelectionPlan(x)) -> nqe_filter_project;
//     PhysicalAggregatePlan(HashJoin)  -> nqe_join_aggregate.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "nqe.h"

namespace nqe {

// ErrorCode, src/error.rs:13-40
struct ErrorCode : std::runtime_error {
    int32_t code;
    ErrorCode(int32_t c, const std::string &m) : std::runtime_error(m), code(c) {}
    const char *kind() const {
        static const char *k[] = {"OK", "ArrowError(DivideByZero)", "IntervalError", "NotSupported", "NotImplemented",
                                  "Panic", "PlanError", "LogicalError", "InvalidArgument", "CudaError", "OutOfMemory"};
        return code >= 0 && code <= 10 ? k[code] : "Others";
    }
};

class Context {
  public:
    explicit Context(int device = 0) {
        const int32_t rc = nqe_ctx_create(device, &ctx_);
        if (rc != NQE_OK) throw ErrorCode(rc, "nqe_ctx_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { nqe_ctx_destroy(ctx_); }
    Context(const Context &) = delete;
    nqe_ctx *get() const { return ctx_; }
    void check(int32_t rc) const {
        if (rc != NQE_OK) throw ErrorCode(rc, nqe_last_error(ctx_));
    }
    static Context &instance() {
        static Context c(0);
        return c;
    }

  private:
    nqe_ctx *ctx_ = nullptr;
};

// One host Arrow array: 8-byte values (Boolean: LSB bitmap, Utf8: int32 offsets + bytes), optional validity bitmap.
struct Column {
    std::string name;
    int32_t dtype = NQE_INT64;
    int64_t length = 0;
    int64_t null_count = 0;
    std::vector<uint8_t> values, validity, data;

    template <typename T>
    static Column from(const std::string &name, int32_t dtype, const std::vector<T> &v) {
        Column c;
        c.name = name; c.dtype = dtype; c.length = (int64_t)v.size();
        c.values.resize(v.size() * sizeof(T));
        if (!v.empty()) memcpy(c.values.data(), v.data(), c.values.size());
        return c;
    }
    static Column int64(const std::string &n, const std::vector<int64_t> &v) { return from(n, NQE_INT64, v); }
    static Column uint64(const std::string &n, const std::vector<uint64_t> &v) { return from(n, NQE_UINT64, v); }
    static Column float64(const std::string &n, const std::vector<double> &v) { return from(n, NQE_FLOAT64, v); }
    static Column utf8(const std::string &n, const std::vector<std::string> &v) {
        Column c;
        c.name = n; c.dtype = NQE_UTF8; c.length = (int64_t)v.size();
        std::vector<int32_t> off(v.size() + 1, 0);
        for (size_t i = 0; i < v.size(); i++) {
            off[i + 1] = off[i] + (int32_t)v[i].size();
            c.data.insert(c.data.end(), v[i].begin(), v[i].end());
        }
        c.values.resize(off.size() * 4);
        memcpy(c.values.data(), off.data(), c.values.size());
        return c;
    }
    template <typename T> std::vector<T> as() const {
        std::vector<T> v((size_t)length);
        if (length) memcpy(v.data(), values.data(), (size_t)length * sizeof(T));
        return v;
    }
    std::vector<std::string> strings() const {
        std::vector<std::string> v;
        const int32_t *off = (const int32_t *)values.data();
        for (int64_t i = 0; i < length; i++) v.emplace_back((const char *)data.data() + off[i], (size_t)(off[i + 1] - off[i]));
        return v;
    }
    bool is_valid(int64_t i) const { return validity.empty() || ((validity[(size_t)i >> 3] >> (i & 7)) & 1); }
};
using RecordBatch = std::vector<Column>;

// A RecordBatch resident in HBM.
class DeviceTable {
  public:
    DeviceTable(nqe_table *t, std::vector<std::string> names) : t_(t), names_(std::move(names)) {}
    ~DeviceTable() { nqe_table_free(t_); }
    DeviceTable(const DeviceTable &) = delete;
    nqe_table *get() const { return t_; }
    const std::vector<std::string> &names() const { return names_; }
    int64_t num_rows() const { return nqe_table_num_rows(t_); }

    static std::shared_ptr<DeviceTable> upload(const RecordBatch &b) {
        Context &ctx = Context::instance();
        std::vector<nqe_column_desc> d(b.size());
        std::vector<std::string> names;
        for (size_t i = 0; i < b.size(); i++) {
            const Column &c = b[i];
            d[i] = nqe_column_desc{c.dtype, 0, c.length, c.null_count, c.values.data(),
                                   c.null_count ? c.validity.data() : nullptr, c.data.data(), (int64_t)c.data.size()};
            names.push_back(c.name);
        }
        nqe_table *t = nullptr;
        ctx.check(nqe_table_upload(ctx.get(), d.data(), (int32_t)d.size(), &t));
        return std::make_shared<DeviceTable>(t, names);
    }
    RecordBatch download() const {
        Context &ctx = Context::instance();
        RecordBatch b;
        const int64_t n = num_rows();
        for (int32_t i = 0; i < nqe_table_num_columns(t_); i++) {
            nqe_column_desc d;
            ctx.check(nqe_table_column(t_, i, &d));
            Column c;
            c.name = names_[i]; c.dtype = d.dtype; c.length = n; c.null_count = d.validity ? d.null_count : 0;
            const size_t vb = d.dtype == NQE_BOOL ? (size_t)((n + 7) / 8) : d.dtype == NQE_UTF8 ? (size_t)(n + 1) * 4 : (size_t)n * 8;
            c.values.resize(vb ? vb : 1);
            if (c.null_count) c.validity.resize((size_t)((n + 7) / 8));
            c.data.resize((size_t)(d.data_bytes > 0 ? d.data_bytes : 1));
            ctx.check(nqe_table_download_column(ctx.get(), t_, i, c.values.data(), (int64_t)c.values.size(),
                                                c.null_count ? c.validity.data() : nullptr, (int64_t)c.validity.size(),
                                                d.dtype == NQE_UTF8 ? c.data.data() : nullptr, (int64_t)c.data.size()));
            b.push_back(std::move(c));
        }
        return b;
    }

  private:
    nqe_table *t_;
    std::vector<std::string> names_;
};
using DeviceTableRef = std::shared_ptr<DeviceTable>;

// ScalarValue, src/logical_plan/expression.rs:174-187
struct ScalarValue {
    int32_t dtype = 0; // 0 = Null
    bool is_null = false;
    nqe_expr_node node{};
    static ScalarValue Int64(int64_t v) { ScalarValue s; s.dtype = NQE_INT64; s.node.value.i64 = v; return s; }
    static ScalarValue UInt64(uint64_t v) { ScalarValue s; s.dtype = NQE_UINT64; s.node.value.u64 = v; return s; }
    static ScalarValue Float64(double v) { ScalarValue s; s.dtype = NQE_FLOAT64; s.node.value.f64 = v; return s; }
    static ScalarValue Boolean(bool v) { ScalarValue s; s.dtype = NQE_BOOL; s.node.value.u64 = v; return s; }
    static ScalarValue Null(int32_t dtype) { ScalarValue s; s.dtype = dtype; s.is_null = true; return s; }
};

enum class Operator { Eq, NotEq, Lt, LtEq, Gt, GtEq, Plus, Minus, Multiply, Divide, Modulos, And, Or }; // expression.rs:335-362
enum class UnaryOperator { Abs, Sin, Cos, Tan };                                                        // unary.rs:92-96

// ---- PhysicalExpr
struct PhysicalExpr {
    virtual ~PhysicalExpr() = default;
    virtual void lower(const std::vector<std::string> &names, std::vector<nqe_expr_node> &out) const = 0;
};
using PhysicalExprRef = std::shared_ptr<PhysicalExpr>;

struct ColumnExpr : PhysicalExpr { // column.rs:18-57
    std::string name;
    int idx = -1;
    static std::shared_ptr<ColumnExpr> try_create(const std::string *name, const int *idx) {
        if (!name && !idx) throw ErrorCode(NQE_ERR_LOGICAL, "ColumnExpr must has name or idx");
        auto c = std::make_shared<ColumnExpr>();
        if (name) c->name = *name;
        if (idx) c->idx = *idx;
        return c;
    }
    static std::shared_ptr<ColumnExpr> by_idx(int i) { return try_create(nullptr, &i); }
    static std::shared_ptr<ColumnExpr> by_name(const std::string &n) { return try_create(&n, nullptr); }
    int resolve(const std::vector<std::string> &names) const {
        if (idx >= 0) return idx;
        for (size_t i = 0; i < names.size(); i++)
            if (names[i] == name) return (int)i;
        throw ErrorCode(NQE_ERR_LOGICAL, "ColumnExpr must has name or idx");
    }
    void lower(const std::vector<std::string> &names, std::vector<nqe_expr_node> &out) const override {
        nqe_expr_node n{};
        n.kind = NQE_NODE_COLUMN; n.column = resolve(names);
        out.push_back(n);
    }
};

struct PhysicalLiteralExpr : PhysicalExpr { // literal.rs:17-35
    ScalarValue literal;
    static PhysicalExprRef create(const ScalarValue &v) { auto p = std::make_shared<PhysicalLiteralExpr>(); p->literal = v; return p; }
    void lower(const std::vector<std::string> &, std::vector<nqe_expr_node> &out) const override {
        nqe_expr_node n = literal.node;
        n.kind = NQE_NODE_LITERAL; n.dtype = literal.dtype; n.is_null = literal.is_null;
        out.push_back(n);
    }
};

struct PhysicalBinaryExpr : PhysicalExpr { // binary.rs:91-155
    PhysicalExprRef left, right;
    Operator op;
    static PhysicalExprRef create(PhysicalExprRef l, Operator op, PhysicalExprRef r) {
        auto p = std::make_shared<PhysicalBinaryExpr>(); p->left = std::move(l); p->op = op; p->right = std::move(r); return p;
    }
    void lower(const std::vector<std::string> &names, std::vector<nqe_expr_node> &out) const override {
        left->lower(names, out);
        right->lower(names, out);
        nqe_expr_node n{};
        n.kind = NQE_NODE_BINARY; n.op = (int32_t)op;
        out.push_back(n);
    }
};

struct PhysicalUnaryExpr : PhysicalExpr { // unary.rs:46-108
    PhysicalExprRef expr;
    UnaryOperator func;
    static PhysicalExprRef create(PhysicalExprRef e, UnaryOperator f, const std::string & /*name*/, int32_t /*return_type*/) {
        auto p = std::make_shared<PhysicalUnaryExpr>(); p->expr = std::move(e); p->func = f; return p;
    }
    void lower(const std::vector<std::string> &names, std::vector<nqe_expr_node> &out) const override {
        expr->lower(names, out);
        nqe_expr_node n{};
        n.kind = NQE_NODE_UNARY; n.op = (int32_t)func;
        out.push_back(n);
    }
};

// ---- AggregateOperator, aggregate/{count,sum,avg,min,max}.rs
struct AggregateOperator {
    int32_t op;
    const char *fn;
    std::shared_ptr<ColumnExpr> col_expr;
    std::string field_name(const std::vector<std::string> &names) const {
        return std::string(fn) + "(" + names[col_expr->resolve(names)] + ")";
    }
};
#define NQE_AGG_CLASS(NAME, OP, FN) \
    struct NAME { static AggregateOperator create(std::shared_ptr<ColumnExpr> c) { return AggregateOperator{OP, FN, std::move(c)}; } };
NQE_AGG_CLASS(Count, NQE_AGG_COUNT, "count")
NQE_AGG_CLASS(Sum, NQE_AGG_SUM, "sum")
NQE_AGG_CLASS(Avg, NQE_AGG_AVG, "avg")
NQE_AGG_CLASS(Min, NQE_AGG_MIN, "min")
NQE_AGG_CLASS(Max, NQE_AGG_MAX, "max")
#undef NQE_AGG_CLASS

// ---- PhysicalPlan
struct PhysicalPlan {
    virtual ~PhysicalPlan() = default;
    virtual std::vector<std::string> schema() const = 0; // field names
    virtual std::vector<std::shared_ptr<PhysicalPlan>> children() const = 0;
    virtual DeviceTableRef execute_device() const = 0;
    std::vector<RecordBatch> execute() const { return {execute_device()->download()}; }
};
using PhysicalPlanRef = std::shared_ptr<PhysicalPlan>;

struct MemTable { // datasource/memory.rs:17-57
    RecordBatch batch;
    mutable DeviceTableRef device;
    static std::shared_ptr<MemTable> try_create(RecordBatch b) { auto m = std::make_shared<MemTable>(); m->batch = std::move(b); return m; }
    DeviceTableRef device_table() const {
        if (!device) device = DeviceTable::upload(batch);
        return device;
    }
};

inline DeviceTableRef filter_project(const DeviceTableRef &t, const PhysicalExpr *pred, const std::vector<PhysicalExprRef> &exprs,
                                     std::vector<std::string> out_names) {
    Context &ctx = Context::instance();
    std::vector<nqe_expr_node> pn;
    nqe_expr pe{};
    if (pred) {
        pred->lower(t->names(), pn);
        pe = nqe_expr{pn.data(), (int32_t)pn.size(), 0};
    }
    std::vector<std::vector<nqe_expr_node>> nodes(exprs.size());
    std::vector<nqe_expr> es(exprs.size());
    for (size_t i = 0; i < exprs.size(); i++) {
        exprs[i]->lower(t->names(), nodes[i]);
        es[i] = nqe_expr{nodes[i].data(), (int32_t)nodes[i].size(), 0};
    }
    nqe_table *out = nullptr;
    ctx.check(nqe_filter_project(ctx.get(), t->get(), pred ? &pe : nullptr, es.data(), (int32_t)es.size(), &out));
    return std::make_shared<DeviceTable>(out, exprs.empty() ? t->names() : std::move(out_names));
}

struct ScanPlan : PhysicalPlan { // scan.rs:17-49
    std::shared_ptr<MemTable> source;
    static PhysicalPlanRef create(std::shared_ptr<MemTable> s) { auto p = std::make_shared<ScanPlan>(); p->source = std::move(s); return p; }
    std::vector<std::string> schema() const override {
        std::vector<std::string> n;
        for (auto &c : source->batch) n.push_back(c.name);
        return n;
    }
    std::vector<PhysicalPlanRef> children() const override { return {}; }
    DeviceTableRef execute_device() const override { return source->device_table(); }
};

struct SelectionPlan : PhysicalPlan { // selection.rs:22-112
    PhysicalPlanRef input;
    PhysicalExprRef expr;
    static PhysicalPlanRef create(PhysicalPlanRef in, PhysicalExprRef e) {
        auto p = std::make_shared<SelectionPlan>(); p->input = std::move(in); p->expr = std::move(e); return p;
    }
    std::vector<std::string> schema() const override { return input->schema(); }
    std::vector<PhysicalPlanRef> children() const override { return {input}; }
    DeviceTableRef execute_device() const override { return filter_project(input->execute_device(), expr.get(), {}, {}); }
};

struct ProjectionPlan : PhysicalPlan { // projection.rs:17-74
    PhysicalPlanRef input;
    std::vector<std::string> names;
    std::vector<PhysicalExprRef> expr;
    static PhysicalPlanRef create(PhysicalPlanRef in, std::vector<std::string> schema, std::vector<PhysicalExprRef> e) {
        auto p = std::make_shared<ProjectionPlan>(); p->input = std::move(in); p->names = std::move(schema); p->expr = std::move(e); return p;
    }
    std::vector<std::string> schema() const override { return names; }
    std::vector<PhysicalPlanRef> children() const override { return {input}; }
    DeviceTableRef execute_device() const override {
        if (names.empty()) return input->execute_device(); // aggregate results pass through (projection.rs:46-48)
        if (auto sel = dynamic_cast<const SelectionPlan *>(input.get())) // fused filter -> project
            return filter_project(sel->input->execute_device(), sel->expr.get(), expr, names);
        return filter_project(input->execute_device(), nullptr, expr, names);
    }
};

struct HashJoin : PhysicalPlan { // hash_join.rs:43-285 (only on[0] is used, join_type ignored: always INNER)
    PhysicalPlanRef left, right;
    std::vector<std::pair<std::string, std::string>> on;
    static PhysicalPlanRef create(PhysicalPlanRef l, PhysicalPlanRef r, std::vector<std::pair<std::string, std::string>> on) {
        auto p = std::make_shared<HashJoin>(); p->left = std::move(l); p->right = std::move(r); p->on = std::move(on); return p;
    }
    std::vector<std::string> schema() const override {
        auto n = left->schema();
        for (auto &x : right->schema()) n.push_back(x);
        return n;
    }
    std::vector<PhysicalPlanRef> children() const override { return {left, right}; }
    std::pair<int, int> keys(const DeviceTable &l, const DeviceTable &r) const {
        if (on.empty()) throw ErrorCode(NQE_ERR_PLAN, "Inner Join on Conditions can't not be empty");
        return {ColumnExpr::by_name(on[0].first)->resolve(l.names()), ColumnExpr::by_name(on[0].second)->resolve(r.names())};
    }
    DeviceTableRef execute_device() const override {
        if (on.empty()) throw ErrorCode(NQE_ERR_PLAN, "Inner Join on Conditions can't not be empty");
        Context &ctx = Context::instance();
        auto l = left->execute_device(), r = right->execute_device();
        auto k = keys(*l, *r);
        nqe_table *out = nullptr;
        ctx.check(nqe_hash_join(ctx.get(), l->get(), r->get(), k.first, k.second, &out));
        return std::make_shared<DeviceTable>(out, schema());
    }
};

struct PhysicalAggregatePlan : PhysicalPlan { // aggregate/mod.rs:29-222
    std::vector<PhysicalExprRef> group_expr;
    std::vector<AggregateOperator> aggr_ops;
    PhysicalPlanRef input;
    static PhysicalPlanRef create(std::vector<PhysicalExprRef> g, std::vector<AggregateOperator> ops, PhysicalPlanRef in) {
        auto p = std::make_shared<PhysicalAggregatePlan>(); p->group_expr = std::move(g); p->aggr_ops = std::move(ops); p->input = std::move(in); return p;
    }
    std::vector<std::string> schema() const override { return input->schema(); } // the INPUT's schema (mod.rs:44,105-107)
    std::vector<PhysicalPlanRef> children() const override { return {input}; }
    DeviceTableRef execute_device() const override {
        Context &ctx = Context::instance();
        std::vector<nqe_agg> aggs;
        std::vector<std::string> out_names;
        nqe_table *out = nullptr;
        auto join = dynamic_cast<const HashJoin *>(input.get());
        auto gcol = group_expr.empty() ? nullptr : dynamic_cast<const ColumnExpr *>(group_expr[0].get());
        if (join && gcol && !join->on.empty()) { // fused join -> aggregate
            auto l = join->left->execute_device(), r = join->right->execute_device();
            auto k = join->keys(*l, *r);
            const auto names = join->schema();
            // Utf8 join / group keys go through the dictionary path of nqe_hash_join / nqe_hash_aggregate, unfused
            auto dtype_at = [&](int c) {
                nqe_column_desc d;
                const int nl = nqe_table_num_columns(l->get());
                nqe_table_column(c < nl ? l->get() : r->get(), c < nl ? c : c - nl, &d);
                return d.dtype;
            };
            if (dtype_at(k.first) != NQE_UTF8 && dtype_at(gcol->resolve(names)) != NQE_UTF8) {
                for (auto &a : aggr_ops) { aggs.push_back(nqe_agg{a.op, a.col_expr->resolve(names)}); out_names.push_back(a.field_name(names)); }
                ctx.check(nqe_join_aggregate(ctx.get(), l->get(), r->get(), k.first, k.second, gcol->resolve(names), aggs.data(),
                                             (int32_t)aggs.size(), &out));
                return std::make_shared<DeviceTable>(out, out_names);
            }
        }
        auto t = input->execute_device();
        for (auto &a : aggr_ops) { aggs.push_back(nqe_agg{a.op, a.col_expr->resolve(t->names())}); out_names.push_back(a.field_name(t->names())); }
        std::vector<nqe_expr_node> gn;
        nqe_expr ge{};
        if (!group_expr.empty()) { // only group_expr[0] is used (mod.rs:141-146)
            group_expr[0]->lower(t->names(), gn);
            ge = nqe_expr{gn.data(), (int32_t)gn.size(), 0};
        }
        ctx.check(nqe_hash_aggregate(ctx.get(), t->get(), group_expr.empty() ? nullptr : &ge, aggs.data(), (int32_t)aggs.size(), &out));
        return std::make_shared<DeviceTable>(out, out_names);
    }
};

struct PhysicalLimitPlan : PhysicalPlan { // limit.rs:17-66
    PhysicalPlanRef input;
    int64_t n;
    static PhysicalPlanRef create(PhysicalPlanRef in, int64_t n) { auto p = std::make_shared<PhysicalLimitPlan>(); p->input = std::move(in); p->n = n; return p; }
    std::vector<std::string> schema() const override { return input->schema(); }
    std::vector<PhysicalPlanRef> children() const override { return {input}; }
    DeviceTableRef execute_device() const override {
        Context &ctx = Context::instance();
        auto t = input->execute_device();
        nqe_table *out = nullptr;
        ctx.check(nqe_table_slice(ctx.get(), t->get(), 0, n < t->num_rows() ? n : t->num_rows(), &out));
        return std::make_shared<DeviceTable>(out, t->names());
    }
};

struct PhysicalOffsetPlan : PhysicalPlan { // offset.rs:17-68
    PhysicalPlanRef input;
    int64_t n;
    static PhysicalPlanRef create(PhysicalPlanRef in, int64_t n) { auto p = std::make_shared<PhysicalOffsetPlan>(); p->input = std::move(in); p->n = n; return p; }
    std::vector<std::string> schema() const override { return input->schema(); }
    std::vector<PhysicalPlanRef> children() const override { return {input}; }
    DeviceTableRef execute_device() const override {
        Context &ctx = Context::instance();
        auto t = input->execute_device();
        const int64_t off = n < t->num_rows() ? n : t->num_rows();
        nqe_table *out = nullptr;
        ctx.check(nqe_table_slice(ctx.get(), t->get(), off, t->num_rows() - off, &out));
        return std::make_shared<DeviceTable>(out, t->names());
    }
};

} // namespace nqe

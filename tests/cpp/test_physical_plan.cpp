// C++ parity tests written like the reference's inline #[test] functions, driving the CUDA
// path through the C++ host mirror (naive-query-engine_b200/host/physical_plan.hpp).
// Expected vectors: projection.rs:88-121, selection.rs:126-178, limit.rs:67-90,
// offset.rs:69-92, README.md:70-111, BASELINE configs[0].
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>

#include "physical_plan.hpp"

using namespace nqe;

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static RecordBatch t1() { // data/test_data.csv
    return {Column::int64("id", {1, 2, 4, 5, 6, 7, 8, 9}),
            Column::utf8("name", {"veeupup", "alex", "lynne", "alice", "bob", "jack", "cock", "primer"}),
            Column::int64("age", {23, 20, 18, 19, 20, 21, 22, 23}),
            Column::float64("score", {60.0, 90.1, 99.99, 81.1, 82.2, 83.3, 84.4, 85.5})};
}
static PhysicalExprRef lit(int64_t v) { return PhysicalLiteralExpr::create(ScalarValue::Int64(v)); }
static PhysicalExprRef col(int i) { return ColumnExpr::by_idx(i); }
static PhysicalExprRef coln(const std::string &n) { return ColumnExpr::by_name(n); }

static int test_projection() { // projection.rs:88-121
    auto scan = ScanPlan::create(MemTable::try_create(t1()));
    auto add = PhysicalBinaryExpr::create(coln("id"), Operator::Plus, lit(1));
    auto plan = ProjectionPlan::create(scan, {"id", "name"}, {add, coln("name")});
    auto res = plan->execute();
    CHECK(res.size() == 1);
    CHECK((res[0][0].as<int64_t>() == std::vector<int64_t>{2, 3, 5, 6, 7, 8, 9, 10}));
    CHECK((res[0][1].strings() == std::vector<std::string>{"veeupup", "alex", "lynne", "alice", "bob", "jack", "cock", "primer"}));
    return 0;
}

static int test_selection() { // selection.rs:126-178
    auto scan = ScanPlan::create(MemTable::try_create(t1()));
    auto proj = ProjectionPlan::create(scan, {"id", "name", "age"}, {col(0), coln("name"), col(2)});
    auto pred = PhysicalBinaryExpr::create(PhysicalBinaryExpr::create(coln("id"), Operator::Plus, lit(1)), Operator::Gt, lit(5));
    auto res = SelectionPlan::create(proj, pred)->execute();
    CHECK(res.size() == 1);
    CHECK((res[0][0].as<int64_t>() == std::vector<int64_t>{5, 6, 7, 8, 9}));
    CHECK((res[0][1].strings() == std::vector<std::string>{"alice", "bob", "jack", "cock", "primer"}));
    return 0;
}

static int test_limit_offset() { // limit.rs:67-90, offset.rs:69-92, README.md:70-76
    auto scan = ScanPlan::create(MemTable::try_create(t1()));
    CHECK((PhysicalLimitPlan::create(scan, 2)->execute()[0][0].as<int64_t>() == std::vector<int64_t>{1, 2}));
    CHECK((PhysicalOffsetPlan::create(scan, 5)->execute()[0][0].as<int64_t>() == std::vector<int64_t>{7, 8, 9}));
    auto sel = SelectionPlan::create(scan, PhysicalBinaryExpr::create(col(0), Operator::Lt, lit(9)));
    auto proj = ProjectionPlan::create(sel, {"id", "name", "age + 100"},
                                       {col(0), col(1), PhysicalBinaryExpr::create(col(2), Operator::Plus, lit(100))});
    auto res = PhysicalLimitPlan::create(PhysicalOffsetPlan::create(proj, 2), 3)->execute();
    CHECK((res[0][0].as<int64_t>() == std::vector<int64_t>{4, 5, 6}));
    CHECK((res[0][1].strings() == std::vector<std::string>{"lynne", "alice", "bob"}));
    CHECK((res[0][2].as<int64_t>() == std::vector<int64_t>{118, 119, 120}));
    return 0;
}

static int test_config1() { // select id, age+100 from t1 where id < 9
    auto scan = ScanPlan::create(MemTable::try_create(t1()));
    auto sel = SelectionPlan::create(scan, PhysicalBinaryExpr::create(col(0), Operator::Lt, lit(9)));
    auto res = ProjectionPlan::create(sel, {"id", "age + 100"}, {col(0), PhysicalBinaryExpr::create(col(2), Operator::Plus, lit(100))})->execute();
    CHECK((res[0][0].as<int64_t>() == std::vector<int64_t>{1, 2, 4, 5, 6, 7, 8}));
    CHECK((res[0][1].as<int64_t>() == std::vector<int64_t>{123, 120, 118, 119, 120, 121, 122}));
    return 0;
}

static int test_readme_join() { // README.md:77-85, printed row order
    RecordBatch emp = {Column::int64("id", {1, 2, 3, 4, 5}), Column::utf8("name", {"vee", "lynne", "Alex", "jack", "mike"}),
                       Column::int64("department_id", {1, 1, 2, 2, 3}), Column::int64("rank", {1, 0, 0, 1, 2})};
    RecordBatch rank = {Column::int64("id", {0, 1, 2}), Column::utf8("rank_name", {"master", "diamond", "grandmaster"})};
    RecordBatch dept = {Column::int64("id", {1, 2, 3}), Column::utf8("department_name", {"IT", "Marketing", "Human Resource"})};
    auto j1 = HashJoin::create(ScanPlan::create(MemTable::try_create(emp)), ScanPlan::create(MemTable::try_create(rank)), {{"rank", "id"}});
    auto j2 = HashJoin::create(j1, ScanPlan::create(MemTable::try_create(dept)), {{"department_id", "id"}});
    auto plan = ProjectionPlan::create(j2, {"id", "name", "rank_name", "department_name"},
                                       {col(0), col(1), coln("rank_name"), coln("department_name")});
    auto res = plan->execute();
    CHECK((res[0][0].as<int64_t>() == std::vector<int64_t>{2, 1, 3, 4, 5}));
    CHECK((res[0][1].strings() == std::vector<std::string>{"lynne", "vee", "Alex", "jack", "mike"}));
    CHECK((res[0][2].strings() == std::vector<std::string>{"master", "diamond", "master", "diamond", "grandmaster"}));
    CHECK((res[0][3].strings() == std::vector<std::string>{"IT", "IT", "Marketing", "Marketing", "Human Resource"}));
    return 0;
}

static int test_readme_group_by() { // README.md:105-111 (row order unspecified)
    auto scan = ScanPlan::create(MemTable::try_create(t1()));
    auto key = PhysicalBinaryExpr::create(col(0), Operator::Modulos, lit(3));
    auto plan = PhysicalAggregatePlan::create({key}, {Count::create(ColumnExpr::by_idx(0)), Sum::create(ColumnExpr::by_idx(2)),
                                                      Sum::create(ColumnExpr::by_idx(3)), Avg::create(ColumnExpr::by_idx(3)),
                                                      Max::create(ColumnExpr::by_idx(3)), Min::create(ColumnExpr::by_idx(3))}, scan);
    auto res = ProjectionPlan::create(plan, {}, {})->execute(); // empty projection schema passes the aggregate through
    CHECK(res[0].size() == 6 && res[0][0].length == 3);
    CHECK(res[0][0].name == "count(id)" && res[0][1].name == "sum(age)" && res[0][5].name == "min(score)");
    const double want[3][6] = {{3, 61, 255.6, 85.2, 90.1, 81.1}, {3, 62, 243.29000000000002, 81.09666666666668, 99.99, 60}, {2, 43, 167.7, 83.85, 85.5, 82.2}};
    auto cnt = res[0][0].as<uint64_t>();
    for (int g = 0; g < 3; g++) {
        bool found = false;
        for (int w = 0; w < 3 && !found; w++) {
            bool ok = cnt[g] == (uint64_t)want[w][0];
            for (int c = 1; c < 6 && ok; c++) {
                const double v = res[0][c].as<double>()[g];
                ok = std::fabs(v - want[w][c]) <= 1e-9 * std::fabs(want[w][c]);
            }
            found = ok;
        }
        CHECK(found);
    }
    return 0;
}

static int test_errors() {
    auto scan = ScanPlan::create(MemTable::try_create(t1()));
    try { // binary.rs:114-119
        auto bad = PhysicalBinaryExpr::create(col(0), Operator::Lt, PhysicalLiteralExpr::create(ScalarValue::Float64(9.5)));
        SelectionPlan::create(scan, bad)->execute();
        CHECK(false);
    } catch (const ErrorCode &e) {
        CHECK(e.code == NQE_ERR_INTERVAL);
        CHECK(std::string(e.what()) == "Cannot evaluate binary expression Lt with types Int64 and Float64");
    }
    try { // hash_join.rs:125-129
        HashJoin::create(scan, scan, {})->execute();
        CHECK(false);
    } catch (const ErrorCode &e) {
        CHECK(e.code == NQE_ERR_PLAN);
    }
    try { // column.rs:24-28
        ColumnExpr::try_create(nullptr, nullptr);
        CHECK(false);
    } catch (const ErrorCode &e) {
        CHECK(e.code == NQE_ERR_LOGICAL);
    }
    return 0;
}

int main() {
    int rc = 0;
    try {
        rc |= test_projection();
        rc |= test_selection();
        rc |= test_limit_offset();
        rc |= test_config1();
        rc |= test_readme_join();
        rc |= test_readme_group_by();
        rc |= test_errors();
    } catch (const ErrorCode &e) {
        std::fprintf(stderr, "ErrorCode %s: %s\n", e.kind(), e.what());
        return 2;
    }
    if (rc == 0) std::puts("ALL OK");
    return rc;
}

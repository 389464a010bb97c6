// Generates and NVRTC-compiles (sm_100a, no GPU needed) the shape-specialised filter/project kernel for the query shapes
// named on the command line and leaves <prefix>.cu / <prefix>.cubin behind (NQE_JIT_DUMP); tests/test_jit_codegen.py
// checks the resource usage with cuobjdump.  Not part of the product: it reaches below the C ABI on purpose.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "nqe_internal.cuh"
int32_t nqe_jit_filter_project(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                               int32_t n_projs, void *const *out_values, uint8_t *const *out_valid,
                               unsigned long long *tile_state, unsigned int *ticket, unsigned long long *out_count,
                               uint32_t *status, bool *used, std::string *source_out);
int main(int argc, char **argv) {
    const std::string shape = argc > 1 ? argv[1] : "plain";
    nqe_ctx ctx;
    nqe_table t;
    t.ctx = &ctx;
    t.nrows = 100000000;
    void *outs[4] = {nullptr, nullptr, nullptr, nullptr};
    bool used = false;
    std::string src;
    int rc;
    if (shape == "gather") { // the partitioned join's fused last pass: [k, a (gathered), fk, b, pos, match (gathered)]
        t.cols.resize(6);
        for (int i = 0; i < 6; i++) { t.cols[i].dtype = NQE_INT64; t.cols[i].values = (void *)0x10000; t.cols[i].length = t.nrows; }
        t.cols[3].dtype = NQE_FLOAT64;
        t.cols[4].dtype = NQE_POS32;
        t.cols[5].dtype = NQE_UINT64;
        t.cols[1].via = 4;
        t.cols[5].via = 4;
        nqe_expr_node gp[3] = {{NQE_NODE_COLUMN, 0, 5, 0, 0, 0, {0}}, {NQE_NODE_LITERAL, 0, 0, NQE_UINT64, 0, 0, {-1}}, {NQE_NODE_BINARY, NQE_OP_NOT_EQ, 0, 0, 0, 0, {0}}};
        nqe_expr gpred{gp, 3, 0};
        nqe_expr_node c[4] = {{NQE_NODE_COLUMN, 0, 2, 0, 0, 0, {0}}, {NQE_NODE_COLUMN, 0, 1, 0, 0, 0, {0}}, {NQE_NODE_COLUMN, 0, 2, 0, 0, 0, {0}}, {NQE_NODE_COLUMN, 0, 3, 0, 0, 0, {0}}};
        nqe_expr gprojs[4] = {{&c[0], 1, 0}, {&c[1], 1, 0}, {&c[2], 1, 0}, {&c[3], 1, 0}};
        rc = nqe_jit_filter_project(&ctx, &t, &gpred, gprojs, 4, outs, nullptr, nullptr, nullptr, nullptr, nullptr, &used, &src);
    } else { // BASELINE configs[1]: select id, age + 100 from t where id < 500  ("nulls": id and age nullable)
        t.cols.resize(3);
        for (int i = 0; i < 3; i++) { t.cols[i].dtype = i == 2 ? NQE_FLOAT64 : NQE_INT64; t.cols[i].values = (void *)0x10000; t.cols[i].length = t.nrows; }
        if (shape == "nulls") for (int i = 0; i < 2; i++) { t.cols[i].validity = (uint8_t *)0x20000; t.cols[i].null_count = 5; }
        nqe_expr_node pn[3] = {{NQE_NODE_COLUMN, 0, 0, 0, 0, 0, {0}}, {NQE_NODE_LITERAL, 0, 0, NQE_INT64, 0, 0, {500}}, {NQE_NODE_BINARY, NQE_OP_LT, 0, 0, 0, 0, {0}}};
        nqe_expr pred{pn, 3, 0};
        nqe_expr_node a[1] = {{NQE_NODE_COLUMN, 0, 0, 0, 0, 0, {0}}};
        nqe_expr_node b[3] = {{NQE_NODE_COLUMN, 0, 1, 0, 0, 0, {0}}, {NQE_NODE_LITERAL, 0, 0, NQE_INT64, 0, 0, {100}}, {NQE_NODE_BINARY, NQE_OP_PLUS, 0, 0, 0, 0, {0}}};
        nqe_expr projs[2] = {{a, 1, 0}, {b, 3, 0}};
        rc = nqe_jit_filter_project(&ctx, &t, &pred, projs, 2, outs, nullptr, nullptr, nullptr, nullptr, nullptr, &used, &src);
    }
    printf("rc=%d source_bytes=%zu err=%s\n", rc, src.size(), ctx.last_error.c_str());
    return rc == 0 && !src.empty() ? 0 : 1;
}

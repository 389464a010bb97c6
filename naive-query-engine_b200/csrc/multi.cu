// multi.cu -- several GPUs of one node behind the C ABI, driven from ONE process.
//
// The reference is a single-process engine: what its planner can bind is a call that takes the shards of a table and
// returns one result, not a rank of a torch.distributed job (naive-query-engine_b200/distributed.py stays the
// one-process-per-GPU path that bench.py measures).  A nqe_multi owns one nqe_ctx per member GPU; every member's share
// of a plan runs on its own host thread (the operators synchronise their stream to read back row counts), tables move
// between members as peer copies over NVLink (cudaMemcpyPeerAsync; peer access is enabled where the topology allows it),
// and nothing larger than the build side or the partial aggregate states ever moves:
//
//   nqe_multi_join_aggregate   the broadcast-build plan: the build side is copied to every member, each member runs the
//                              fused join -> PARTIAL aggregate over its own probe shard, the partial states (one row per
//                              group and member) are gathered on member 0 and merged there.
//   nqe_multi_hash_aggregate   group-by over sharded input: local pre-aggregate, gather, merge.
//
// Partial states: COUNT -> count, SUM -> sum, AVG -> sum and count, MIN -> min, MAX -> max, plus the group key
// (extension op NQE_AGG_GROUP_KEY: the reference's aggregate output has no key column, aggregate/mod.rs:186-221).  The
// merge is one more nqe_hash_aggregate over the gathered partial rows (counts and sums are added, min / max folded),
// then a small kernel turns the merged states into the requested columns (count back to UInt64, avg = sum / count).
#include <chrono>
#include <cstring>
#include <thread>

#include "agg_device.cuh"
#include "nqe_internal.cuh"

struct nqe_multi {
    std::vector<nqe_ctx *> ctx;
    std::string last_error;
};

namespace {

int32_t multi_fail(nqe_multi *m, int32_t code, const std::string &msg) {
    m->last_error = msg;
    return code;
}

// out[i] = (u64) in[i] (a count that was summed as f64: exact below 2^53) | in[i] | a[i] / b[i]
__global__ void multi_finalize_kernel(int kind, const unsigned long long *a, const unsigned long long *b, unsigned long long *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (kind == 0) out[i] = a[i];
    else if (kind == 1) out[i] = (unsigned long long)__longlong_as_double((long long)a[i]);
    else out[i] = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)a[i]) / __longlong_as_double((long long)b[i]));
}

struct PartialPlan {
    std::vector<nqe_agg> partial; // per-member aggregate list; the group key is its last column
    std::vector<nqe_agg> merge;   // over the gathered partial rows (column j = partial j)
    // requested aggregate -> (kind, merged column a, merged column b)
    struct Out { int kind, a, b, dtype; };
    std::vector<Out> outs;
    int key_col = 0; // column of the group key in the partial table and, if present, in the merged one
    bool want_key = false;
};

int32_t make_plan(nqe_multi *m, const nqe_agg *aggs, int32_t n_aggs, PartialPlan *pl) {
    auto find = [&](int32_t op, int32_t col) {
        for (size_t j = 0; j < pl->partial.size(); j++)
            if (pl->partial[j].op == op && pl->partial[j].column == col) return (int)j;
        pl->partial.push_back(nqe_agg{op, col});
        return (int)pl->partial.size() - 1;
    };
    std::vector<int> need_a(n_aggs, -1), need_b(n_aggs, -1);
    for (int a = 0; a < n_aggs; a++) {
        switch (aggs[a].op) {
        case NQE_AGG_COUNT: need_a[a] = find(NQE_AGG_COUNT, aggs[a].column); break;
        case NQE_AGG_SUM: need_a[a] = find(NQE_AGG_SUM, aggs[a].column); break;
        case NQE_AGG_AVG:
            need_a[a] = find(NQE_AGG_SUM, aggs[a].column);
            need_b[a] = find(NQE_AGG_COUNT, aggs[a].column);
            break;
        case NQE_AGG_MIN: need_a[a] = find(NQE_AGG_MIN, aggs[a].column); break;
        case NQE_AGG_MAX: need_a[a] = find(NQE_AGG_MAX, aggs[a].column); break;
        case NQE_AGG_GROUP_KEY: pl->want_key = true; break;
        default: return multi_fail(m, NQE_ERR_INVALID_ARG, "unknown aggregate op");
        }
    }
    const int np = (int)pl->partial.size();
    if (np + 1 > AG_MAX) return multi_fail(m, NQE_ERR_NOT_SUPPORTED, "too many partial aggregate states for one plan");
    for (int j = 0; j < np; j++) {
        const int32_t op = pl->partial[j].op;
        pl->merge.push_back(nqe_agg{op == NQE_AGG_COUNT ? NQE_AGG_SUM : op, j}); // counts and sums add, min / max fold
    }
    pl->key_col = np;
    pl->partial.push_back(nqe_agg{NQE_AGG_GROUP_KEY, 0});
    if (pl->want_key) pl->merge.push_back(nqe_agg{NQE_AGG_GROUP_KEY, 0});
    for (int a = 0; a < n_aggs; a++) {
        switch (aggs[a].op) {
        case NQE_AGG_COUNT: pl->outs.push_back({1, need_a[a], -1, NQE_UINT64}); break;
        case NQE_AGG_AVG: pl->outs.push_back({2, need_a[a], need_b[a], NQE_FLOAT64}); break;
        case NQE_AGG_GROUP_KEY: pl->outs.push_back({0, np, -1, NQE_INT64}); break;
        default: pl->outs.push_back({0, need_a[a], -1, NQE_FLOAT64}); break;
        }
    }
    return NQE_OK;
}

// gather the members' partial tables on member 0, merge, finalise
int32_t merge_partials(nqe_multi *m, const PartialPlan &pl, std::vector<nqe_table *> &part, nqe_table **out) {
    nqe_ctx *c0 = m->ctx[0];
    cudaSetDevice(c0->device);
    std::vector<nqe_table *> on0;
    int32_t rc = NQE_OK;
    { // the partial tables are small (one row per group): all peer copies at once, one host thread each
        std::vector<nqe_table *> moved(part.size(), nullptr);
        std::vector<int32_t> crc(part.size(), NQE_OK);
        std::vector<std::thread> th;
        for (size_t i = 0; i < part.size(); i++)
            if (part[i] && part[i]->ctx != c0) th.emplace_back([&, i] { crc[i] = nqe_multi_table_copy(m, part[i], 0, &moved[i]); });
        for (auto &t : th) t.join();
        for (size_t i = 0; i < part.size(); i++) {
            if (!part[i]) continue;
            if (part[i]->ctx == c0) {
                on0.push_back(part[i]);
                part[i] = nullptr;
            } else if (crc[i] != NQE_OK) {
                rc = crc[i];
            } else {
                on0.push_back(moved[i]);
            }
        }
    }
    nqe_table *all = nullptr, *merged = nullptr;
    if (rc == NQE_OK && on0.empty()) rc = multi_fail(m, NQE_ERR_INVALID_ARG, "no member had any input");
    if (rc == NQE_OK) {
        if (on0.size() == 1) {
            all = on0[0];
            on0.clear();
        } else {
            rc = nqe_table_concat(c0, on0.data(), (int32_t)on0.size(), &all);
        }
    }
    if (rc == NQE_OK) {
        nqe_expr_node kn{NQE_NODE_COLUMN, 0, pl.key_col, 0, 0, 0, {0}};
        const nqe_expr ke{&kn, 1, 0};
        rc = nqe_hash_aggregate(c0, all, &ke, pl.merge.data(), (int32_t)pl.merge.size(), &merged);
    }
    nqe_table *t = nullptr;
    if (rc == NQE_OK) {
        const int64_t g = merged->nrows;
        nqe_table_new(c0, g, &t);
        t->cols.resize(pl.outs.size());
        for (size_t o = 0; o < pl.outs.size() && rc == NQE_OK; o++) {
            const PartialPlan::Out &spec = pl.outs[o];
            rc = nqe_column_alloc(c0, spec.dtype, g, false, &t->cols[o]);
            if (rc != NQE_OK || g == 0) continue;
            multi_finalize_kernel<<<(unsigned)((g + 255) / 256), 256, 0, c0->stream>>>(
                spec.kind, (const unsigned long long *)merged->cols[spec.a].values,
                spec.b >= 0 ? (const unsigned long long *)merged->cols[spec.b].values : nullptr,
                (unsigned long long *)t->cols[o].values, g);
            c0->launches++;
        }
        if (rc == NQE_OK && cudaStreamSynchronize(c0->stream) != cudaSuccess) rc = nqe_fail(c0, NQE_ERR_CUDA, "merge failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (rc != NQE_OK && m->last_error.empty()) m->last_error = c0->last_error;
    for (nqe_table *p : on0) nqe_table_free(p);
    if (all) nqe_table_free(all);
    if (merged) nqe_table_free(merged);
    if (rc != NQE_OK) {
        if (t) nqe_table_free(t);
        return rc;
    }
    *out = t;
    return NQE_OK;
}

// run fn(member) on one host thread per member; first error wins
template <typename F>
int32_t for_each_member(nqe_multi *m, F fn) {
    const int n = (int)m->ctx.size();
    std::vector<int32_t> rc(n, NQE_OK);
    std::vector<std::thread> th;
    for (int i = 1; i < n; i++) th.emplace_back([&, i] { rc[i] = fn(i); });
    rc[0] = fn(0);
    for (auto &t : th) t.join();
    for (int i = 0; i < n; i++)
        if (rc[i] != NQE_OK) {
            m->last_error = "member " + std::to_string(i) + ": " + m->ctx[i]->last_error;
            return rc[i];
        }
    return NQE_OK;
}

} // namespace

extern "C" int32_t nqe_multi_create(const int32_t *devices, int32_t n, nqe_multi **out) {
    if (!devices || n < 1 || !out) return NQE_ERR_INVALID_ARG;
    *out = nullptr;
    nqe_multi *m = new nqe_multi();
    for (int i = 0; i < n; i++) {
        nqe_ctx *c = nullptr;
        const int32_t rc = nqe_ctx_create(devices[i], &c);
        if (rc != NQE_OK) {
            for (nqe_ctx *p : m->ctx) nqe_ctx_destroy(p);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    // direct NVLink loads / stores for the peer copies where the topology allows it (cudaMemcpyPeer works either way)
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                cudaSetDevice(devices[i]);
                cudaDeviceEnablePeerAccess(devices[j], 0);
                cudaGetLastError(); // already enabled is fine
                // tables live in the stream-ordered allocator's default pool, which cudaDeviceEnablePeerAccess does not
                // cover: without this a peer copy is staged through host memory (measured 35 GB/s instead of NVLink)
                cudaMemPool_t pool;
                if (cudaDeviceGetDefaultMemPool(&pool, devices[j]) == cudaSuccess) {
                    cudaMemAccessDesc desc;
                    memset(&desc, 0, sizeof desc);
                    desc.location.type = cudaMemLocationTypeDevice;
                    desc.location.id = devices[i];
                    desc.flags = cudaMemAccessFlagsProtReadWrite;
                    cudaMemPoolSetAccess(pool, &desc, 1);
                    cudaGetLastError();
                }
            }
        }
    *out = m;
    return NQE_OK;
}

extern "C" void nqe_multi_destroy(nqe_multi *m) {
    if (!m) return;
    for (nqe_ctx *c : m->ctx) nqe_ctx_destroy(c);
    delete m;
}

extern "C" int32_t nqe_multi_size(const nqe_multi *m) { return m ? (int32_t)m->ctx.size() : 0; }

extern "C" nqe_ctx *nqe_multi_ctx(nqe_multi *m, int32_t member) {
    return m && member >= 0 && member < (int32_t)m->ctx.size() ? m->ctx[member] : nullptr;
}

extern "C" const char *nqe_multi_last_error(const nqe_multi *m) { return m ? m->last_error.c_str() : ""; }

extern "C" int32_t nqe_multi_table_copy(nqe_multi *m, const nqe_table *src, int32_t dst_member, nqe_table **out) {
    if (!m || !src || !out || dst_member < 0 || dst_member >= (int32_t)m->ctx.size()) return NQE_ERR_INVALID_ARG;
    *out = nullptr;
    nqe_ctx *sc = src->ctx, *dc = m->ctx[dst_member];
    for (const DevColumn &c : src->cols)
        if (c.via >= 0 || c.dtype == NQE_POS32) return multi_fail(m, NQE_ERR_INVALID_ARG, "library-internal columns cannot be copied");
    cudaSetDevice(sc->device);
    if (cudaStreamSynchronize(sc->stream) != cudaSuccess) return multi_fail(m, NQE_ERR_CUDA, "source stream failed"); // the source is complete
    cudaSetDevice(dc->device);
    nqe_table *t = nullptr;
    nqe_table_new(dc, src->nrows, &t);
    t->cols.resize(src->cols.size());
    const int64_t n = src->nrows;
    int32_t rc = NQE_OK;
    auto copy = [&](void *dst, const void *from, size_t bytes) {
        if (rc != NQE_OK || bytes == 0) return;
        const cudaError_t e = sc->device == dc->device ? cudaMemcpyAsync(dst, from, bytes, cudaMemcpyDeviceToDevice, dc->stream)
                                                       : cudaMemcpyPeerAsync(dst, dc->device, from, sc->device, bytes, dc->stream);
        if (e != cudaSuccess) rc = nqe_fail(dc, NQE_ERR_CUDA, "peer copy failed: %s", cudaGetErrorString(e));
    };
    for (size_t i = 0; i < src->cols.size() && rc == NQE_OK; i++) {
        const DevColumn &s = src->cols[i];
        DevColumn &d = t->cols[i];
        rc = nqe_column_alloc(dc, s.dtype, n, s.validity != nullptr, &d);
        if (rc != NQE_OK) break;
        d.null_count = s.null_count;
        const size_t vbytes = s.dtype == NQE_BOOL ? (size_t)((n + 7) / 8) : s.dtype == NQE_UTF8 ? (size_t)(n + 1) * 4 : (size_t)n * 8;
        if (s.dtype == NQE_BOOL) cudaMemsetAsync(d.values, 0, nqe_bitmap_bytes(n), dc->stream); // kernels read whole words
        copy(d.values, s.values, vbytes);
        if (s.validity) {
            cudaMemsetAsync(d.validity, 0, nqe_bitmap_bytes(n), dc->stream);
            copy(d.validity, s.validity, (size_t)((n + 7) / 8));
        }
        if (s.dtype == NQE_UTF8) {
            d.data_bytes = s.data_bytes;
            rc = nqe_dev_alloc(dc, (void **)&d.data, (size_t)s.data_bytes + 64);
            copy(d.data, s.data, (size_t)s.data_bytes);
        }
    }
    // the caller may free the source right after this call
    if (rc == NQE_OK && cudaStreamSynchronize(dc->stream) != cudaSuccess) rc = nqe_fail(dc, NQE_ERR_CUDA, "peer copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (rc != NQE_OK) {
        m->last_error = dc->last_error;
        nqe_table_free(t);
        return rc;
    }
    *out = t;
    return NQE_OK;
}

extern "C" int32_t nqe_multi_join_aggregate(nqe_multi *m, const nqe_table *left, const nqe_table *const *right, int32_t left_key,
                                            int32_t right_key, int32_t group_column, const nqe_agg *aggs, int32_t n_aggs,
                                            nqe_table **out) {
    if (!m || !left || !right || !out || !aggs || n_aggs < 1) return NQE_ERR_INVALID_ARG;
    *out = nullptr;
    m->last_error.clear();
    const int n = (int)m->ctx.size();
    for (int i = 0; i < n; i++)
        if (right[i] && right[i]->ctx != m->ctx[i]) return multi_fail(m, NQE_ERR_INVALID_ARG, "right[i] must live on member i");
    PartialPlan pl;
    NQE_TRY(make_plan(m, aggs, n_aggs, &pl));
    std::vector<nqe_table *> part(n, nullptr);
    static const bool prof = getenv("NQE_MULTI_PROF") != nullptr; // per-member phase times of one call on stderr
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    // ---- the build side travels (the probe side never does): a doubling broadcast -- every member that already holds a
    // copy feeds one that does not, so n members are served in ceil(log2 n) rounds of concurrent peer copies instead of
    // n - 1 copies queued on the owner's links (NVSwitch: any pair at full bandwidth)
    std::vector<const nqe_table *> L(n, nullptr);
    std::vector<nqe_table *> copies(n, nullptr);
    std::vector<const nqe_table *> have;
    std::vector<int> need;
    have.push_back(left);
    for (int i = 0; i < n; i++) {
        if (!right[i]) continue;
        if (left->ctx == m->ctx[i]) L[i] = left;
        else need.push_back(i);
    }
    int32_t rc = NQE_OK;
    for (size_t next = 0; next < need.size() && rc == NQE_OK;) {
        const size_t sources = have.size(), round = need.size() - next < sources ? need.size() - next : sources;
        std::vector<int32_t> crc(round, NQE_OK);
        std::vector<std::thread> th;
        for (size_t k = 0; k < round; k++)
            th.emplace_back([&, k] { crc[k] = nqe_multi_table_copy(m, have[k], need[next + k], &copies[need[next + k]]); });
        for (auto &t : th) t.join();
        for (size_t k = 0; k < round; k++) {
            if (crc[k] != NQE_OK) rc = crc[k];
            L[need[next + k]] = copies[need[next + k]];
            have.push_back(copies[need[next + k]]);
        }
        next += round;
    }
    const double t_bcast = now();
    if (rc == NQE_OK)
        rc = for_each_member(m, [&](int i) -> int32_t {
            nqe_ctx *c = m->ctx[i];
            const double t0 = now();
            cudaSetDevice(c->device);
            if (!right[i]) return NQE_OK; // a member without a shard
            const int32_t jrc = nqe_join_aggregate(c, L[i], right[i], left_key, right_key, group_column, pl.partial.data(),
                                                   (int32_t)pl.partial.size(), &part[i]);
            if (prof) fprintf(stderr, "nqe_multi member %d: thread start %.3f ms, join -> partial aggregate %.3f (kernels %.3f)\n", i,
                              t0 - t_bcast, now() - t0, c->last_op_ms);
            return jrc;
        });
    for (nqe_table *t : copies)
        if (t) nqe_table_free(t);
    const double t_joined = now();
    if (prof) fprintf(stderr, "nqe_multi: build-side broadcast %.3f ms\n", t_bcast - t_start);
    if (rc == NQE_OK) rc = merge_partials(m, pl, part, out);
    if (prof) fprintf(stderr, "nqe_multi: members done after %.3f ms, gather + merge %.3f ms\n", t_joined - t_start, now() - t_joined);
    for (nqe_table *p : part)
        if (p) nqe_table_free(p);
    return rc;
}

extern "C" int32_t nqe_multi_hash_aggregate(nqe_multi *m, const nqe_table *const *in, const nqe_expr *group_expr, const nqe_agg *aggs,
                                            int32_t n_aggs, nqe_table **out) {
    if (!m || !in || !out || !aggs || n_aggs < 1) return NQE_ERR_INVALID_ARG;
    if (!group_expr) return multi_fail(m, NQE_ERR_NOT_SUPPORTED, "the multi-GPU aggregate needs a group expression");
    *out = nullptr;
    m->last_error.clear();
    const int n = (int)m->ctx.size();
    for (int i = 0; i < n; i++)
        if (in[i] && in[i]->ctx != m->ctx[i]) return multi_fail(m, NQE_ERR_INVALID_ARG, "in[i] must live on member i");
    PartialPlan pl;
    NQE_TRY(make_plan(m, aggs, n_aggs, &pl));
    std::vector<nqe_table *> part(n, nullptr);
    int32_t rc = for_each_member(m, [&](int i) -> int32_t {
        cudaSetDevice(m->ctx[i]->device);
        if (!in[i]) return NQE_OK;
        return nqe_hash_aggregate(m->ctx[i], in[i], group_expr, pl.partial.data(), (int32_t)pl.partial.size(), &part[i]);
    });
    if (rc == NQE_OK) rc = merge_partials(m, pl, part, out);
    for (nqe_table *p : part)
        if (p) nqe_table_free(p);
    return rc;
}

// jit.cu -- query-shape specialisation of the filter/project kernel.
//
// The expression trees of one SelectionPlan + ProjectionPlan pair are turned into
// straight-line CUDA C++ (one typed expression per output, literals and column
// pointers stay kernel parameters, so the compiled kernel depends only on the
// SHAPE of the query), compiled for sm_100a with NVRTC and cached per context
// process.  The hand-written skeleton below is the same algorithm as the
// interpreter kernels in filter_project.cu (tile ticket, ballot ranks, decoupled
// look-back, contiguous warp stores); only the per-row arithmetic is generated.
//
// Applies when no referenced column has a validity bitmap and no literal is NULL;
// everything else runs on the pre-compiled interpreter kernels.  If libnvrtc is
// not present the interpreter kernels are used as well (NQE_JIT=0 forces that).
#include <atomic>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "nqe_internal.cuh"

namespace {

const char *kLookbackSrc =
#include "build/lookback_body.inc.h"
    ;

const char *kSkeletonHead = R"SRC(
typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;
typedef unsigned char u8;
struct JP {
    const u64 *col[16];
    void *out[16];
    u64 lit[32];
    i64 n_rows;
    u64 *tile_state;
    u32 *ticket;
    u64 *out_count;
    u32 *status;
    int num_tiles;
    int pad;
    u64 *prof;
    const u32 *valid[16];  // validity bitmap of column slot s, or null
    u8 *out_valid[16];     // validity byte per output row, or null (output cannot be NULL)
};
#define WARPS 8
#if NQE_PROF
#define PROF_T0 long long _t0 = clock64();
#define PROF_ADD(i) { const long long _t1 = clock64(); _prof[i] += (u64)(_t1 - _t0); _t0 = _t1; }
#else
#define PROF_T0
#define PROF_ADD(i)
#endif
#define LB_AGG (1ull << 62)
#define LB_PREFIX (2ull << 62)
#define LB_MASK ((1ull << 62) - 1)
#define I64_MIN (-9223372036854775807ll - 1)

__device__ __forceinline__ u64 ldg_stream(const u64 *p) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u64 ldg_keep(const u64 *p, u64 pol) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ u64 ld_vol(const u64 *p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_vol(u64 *p, u64 v) { asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v)); }

// arrow 13 divide / modulus semantics (see expr_eval.cuh)
__device__ __forceinline__ i64 nqe_div_i64(i64 x, i64 y, bool live, u32 *st) {
    if (y == 0) { if (live) atomicOr(st, 1u); return 0; }
    if (y == -1) { if (x == I64_MIN) { if (live) atomicOr(st, 2u); return 0; } return -x; }
    return x / y;
}
__device__ __forceinline__ i64 nqe_mod_i64(i64 x, i64 y, bool live, u32 *st) {
    if (y == 0) { if (live) atomicOr(st, 1u); return 0; }
    if (y == -1) { if (x == I64_MIN && live) atomicOr(st, 2u); return 0; }
    return x % y;
}
__device__ __forceinline__ u64 nqe_div_u64(u64 x, u64 y, bool live, u32 *st) {
    if (y == 0) { if (live) atomicOr(st, 1u); return 0; }
    return x / y;
}
__device__ __forceinline__ u64 nqe_mod_u64(u64 x, u64 y, bool live, u32 *st) {
    if (y == 0) { if (live) atomicOr(st, 1u); return 0; }
    return x % y;
}
__device__ __forceinline__ double nqe_div_f64(double x, double y, bool live, u32 *st) {
    if (y == 0.0 && live) atomicOr(st, 1u);
    return x / y;
}
__device__ __forceinline__ double nqe_mod_f64(double x, double y, bool live, u32 *st) {
    if (y == 0.0 && live) atomicOr(st, 1u);
    return fmod(x, y);
}

)SRC";

// Count-ahead skeleton.  Every CTA keeps D claimed tiles "counted but not yet written":
//   count(t) : load the predicate's columns, evaluate it, publish the tile aggregate
//   write(t) : D iterations later -- re-load the columns (the predicate's now come from
//              L2), re-evaluate, rank, resolve the exclusive prefix by look-back (all
//              predecessors were counted long ago, so the walk does not wait), project, store
// The HBM loads of count(next) are issued before write(t) and consumed after it, so
// their latency overlaps the write phase.
const char *kSkeletonKernel = R"SRC(
#if HAS_PRED
extern "C" __global__ void __launch_bounds__(THREADS) nqe_fp_jit(const __grid_constant__ JP p) {
    __shared__ u32 s_cnt[K * WARPS];
    __shared__ u32 s_wsum[WARPS];
    __shared__ u64 s_tile_excl;
    __shared__ int s_ring_tile[D];
    __shared__ u32 s_ring_total[D];
    __shared__ int s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 ltmask = (1u << lane) - 1u;
    const int TILE = K * THREADS;
    u64 pol_keep, pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

    // ---- prologue: claim and count D tiles
    for (int d = 0; d < D; d++) {
        if (tid == 0) s_ring_tile[d] = (int)atomicAdd(p.ticket, 1u);
        __syncthreads();
        const int tile = s_ring_tile[d];
        if (tile < p.num_tiles) {
            const i64 e0 = (i64)tile * TILE + tid;
            const bool full = (i64)(tile + 1) * TILE <= p.n_rows;
            LOAD_PRED_COLUMNS
            u32 wsum = 0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const bool inr = full || (e0 + (i64)j * THREADS < p.n_rows);
                const bool keep = inr && (PRED_EXPR);
                wsum += __popc(__ballot_sync(0xffffffffu, keep));
            }
            if (lane == 0) s_wsum[warp] = wsum;
            __syncthreads();
            if (tid == 0) {
                u32 total = 0;
#pragma unroll
                for (int w = 0; w < WARPS; w++) total += s_wsum[w];
                s_ring_total[d] = total;
                nqe_lb_publish(p.tile_state, tile, total);
            }
        }
        __syncthreads();
    }

    for (int head = 0;; head = head + 1 == D ? 0 : head + 1) {
        const int tile = s_ring_tile[head];
        if (tile >= p.num_tiles) break;
        const u32 my_total = s_ring_total[head];
        if (tid == 0) s_next = (int)atomicAdd(p.ticket, 1u);
        __syncthreads();
        const int next = s_next;
        // ---- start the HBM loads of count(next)
        const bool has_next = next < p.num_tiles;
        const i64 n0 = (i64)next * TILE + tid;
        const bool nfull = (i64)(next + 1) * TILE <= p.n_rows;
        DECL_NEXT_COLUMNS
        if (has_next) {
            LOAD_NEXT_COLUMNS
        }
        // ---- write(tile)
        {
            const i64 e0 = (i64)tile * TILE + tid;
            const bool full = (i64)(tile + 1) * TILE <= p.n_rows;
            LOAD_COLUMNS
            bool keep[K];
            u32 bal[K];
#pragma unroll
            for (int j = 0; j < K; j++) {
                const bool inr = full || (e0 + (i64)j * THREADS < p.n_rows);
                keep[j] = inr && (PRED_EXPR);
                bal[j] = __ballot_sync(0xffffffffu, keep[j]);
                if (lane == 0) s_cnt[j * WARPS + warp] = __popc(bal[j]);
            }
            __syncthreads();
            if (warp == 0) {
                const int N = K * WARPS, PER = (N + 31) / 32;
                u32 c[PER], sum = 0;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    const int i = lane * PER + q;
                    c[q] = i < N ? s_cnt[i] : 0;
                    sum += c[q];
                }
                u32 incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                u32 run = incl - sum;
#pragma unroll
                for (int q = 0; q < PER; q++) {
                    const int i = lane * PER + q;
                    if (i < N) s_cnt[i] = run;
                    run += c[q];
                }
                const u64 excl = nqe_lb_walk(p.tile_state, tile, my_total, lane);
                if (lane == 0) {
                    s_tile_excl = excl;
                    if (tile == p.num_tiles - 1) *p.out_count = excl + my_total;
                }
            }
            __syncthreads();
            const u64 tile_excl = s_tile_excl;
            u32 idx[K];
#pragma unroll
            for (int j = 0; j < K; j++) idx[j] = s_cnt[j * WARPS + warp] + __popc(bal[j] & ltmask);
            STORE_OUTPUTS
        }
        // ---- finish count(next): its loads have been in flight during the write phase
        if (has_next) {
            u32 wsum = 0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const bool inr = nfull || (n0 + (i64)j * THREADS < p.n_rows);
                const bool keep = inr && (NEXT_PRED_EXPR);
                wsum += __popc(__ballot_sync(0xffffffffu, keep));
            }
            if (lane == 0) s_wsum[warp] = wsum;
        }
        __syncthreads();
        if (tid == 0) {
            s_ring_tile[head] = next;
            if (has_next) {
                u32 total = 0;
#pragma unroll
                for (int w = 0; w < WARPS; w++) total += s_wsum[w];
                s_ring_total[head] = total;
                nqe_lb_publish(p.tile_state, next, total);
            }
        }
        __syncthreads();
    }
}
#else
extern "C" __global__ void __launch_bounds__(THREADS) nqe_fp_jit(const __grid_constant__ JP p) {
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const i64 e0 = (i64)tile * (K * THREADS) + tid;
    const bool full = (i64)(tile + 1) * (K * THREADS) <= p.n_rows;
    LOAD_COLUMNS
    bool keep[K];
    u32 idx[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        keep[j] = full || (e0 + (i64)j * THREADS < p.n_rows);
        idx[j] = j * THREADS + tid;
    }
    const u64 tile_excl = (u64)tile * (K * THREADS);
    STORE_OUTPUTS
}
#endif
)SRC";

// TMA-ring skeleton (the default when every referenced buffer is 16-byte aligned): see
// jit_tma_skeleton.inc for the design.
const char *kSkeletonTma =
#include "build/jit_tma_skeleton.inc.h"
    ;

struct TNode {
    int kind, op, dtype, col, lit; // col: column slot; lit: literal slot
    int left = -1, right = -1;
    bool is_null = false; // NULL literal
};

struct Gen {
    const nqe_table *in;
    std::vector<int> col_of_slot; // table column per slot
    std::vector<uint64_t> lits;
    std::vector<TNode> nodes;
    std::vector<char> in_proj; // slot is read by some projection
    std::vector<char> has_valid; // slot has a validity bitmap
    std::vector<int> via_slot;   // gathered column (DevColumn::via): slot of its position column, else -1
    bool any_via = false;
    bool parsing_proj = false;
    bool any_nulls = false; // some referenced column is nullable or a literal is NULL

    int slot_of(int col) {
        size_t s = 0;
        for (; s < col_of_slot.size(); s++)
            if (col_of_slot[s] == col) break;
        if (s == col_of_slot.size()) {
            col_of_slot.push_back(col);
            in_proj.push_back(0);
            has_valid.push_back(in->cols[col].validity ? 1 : 0);
            via_slot.push_back(-1);
        }
        if (parsing_proj) in_proj[s] = 1;
        if (in->cols[col].via >= 0) { // the position column is staged wherever the gathered column is read
            const int x = slot_of(in->cols[col].via);
            via_slot[s] = x;
            any_via = true;
        }
        return (int)s;
    }

    // parse one postfix program; returns root index or -1 when the JIT does not apply
    int parse(const nqe_expr *ex) {
        std::vector<int> st;
        for (int i = 0; i < ex->n_nodes; i++) {
            const nqe_expr_node &s = ex->nodes[i];
            TNode n{};
            n.kind = s.kind; n.op = s.op;
            if (s.kind == NQE_NODE_COLUMN) {
                const DevColumn &c = in->cols[s.column];
                if (c.dtype == NQE_UTF8 || c.dtype == NQE_POS32) return -1;
                if (c.via >= 0 && (c.validity || c.dtype == NQE_BOOL || in->cols[c.via].dtype != NQE_POS32 || in->cols[c.via].via >= 0)) return -1;
                if (c.validity) any_nulls = true;
                n.dtype = c.dtype;
                n.col = slot_of(s.column);
                if (col_of_slot.size() > 16) return -1;
            } else if (s.kind == NQE_NODE_LITERAL) {
                n.dtype = s.dtype;
                n.is_null = s.is_null != 0;
                if (n.is_null) any_nulls = true;
                n.lit = (int)lits.size();
                lits.push_back(n.is_null ? 0 : s.dtype == NQE_BOOL ? (s.value.u64 ? 1 : 0) : s.value.u64);
                if (lits.size() > 32) return -1;
            } else if (s.kind == NQE_NODE_BINARY) {
                n.right = st.back(); st.pop_back();
                n.left = st.back(); st.pop_back();
                const int lt = nodes[n.left].dtype;
                n.dtype = (s.op <= NQE_OP_GT_EQ || s.op == NQE_OP_AND || s.op == NQE_OP_OR) ? NQE_BOOL : lt;
            } else {
                n.left = st.back(); st.pop_back();
                n.dtype = NQE_FLOAT64;
            }
            nodes.push_back(n);
            st.push_back((int)nodes.size() - 1);
        }
        return st.back();
    }

    static const char *ctype(int dt) {
        return dt == NQE_INT64 ? "i64" : dt == NQE_UINT64 ? "u64" : dt == NQE_FLOAT64 ? "double" : "bool";
    }

    // NULL-aware form: one (value, valid) pair of temporaries per node, `t<n>v` / `t<n>k`, for row j of this
    // thread.  live = "errors of this row count", rn = "the row's column inputs are forced NULL"
    // (load_operand / apply_binary in expr_eval.cuh are the interpreter's statement of the same rules).
    void emit_stmts(int n, std::ostringstream &o, const std::string &live, const std::string &rn) {
        const TNode &t = nodes[n];
        const std::string T = ctype(t.dtype), v = "t" + std::to_string(n) + "v", k = "t" + std::to_string(n) + "k";
        if (t.kind == NQE_NODE_COLUMN || t.kind == NQE_NODE_LITERAL) {
            o << "const " << T << " " << v << " = " << emit(n, "false") << ";\n";
            if (t.kind == NQE_NODE_LITERAL) o << "const bool " << k << " = " << (t.is_null ? "false" : "true") << ";\n";
            else o << "const bool " << k << " = " << (has_valid[t.col] ? "((v" + std::to_string(t.col) + "m >> j) & 1u) != 0 && " : std::string()) << "!(" << rn << ");\n";
            return;
        }
        if (t.kind == NQE_NODE_UNARY) {
            emit_stmts(t.left, o, live, rn);
            const std::string a = "t" + std::to_string(t.left) + "v";
            o << "const double " << v << " = " << (t.op == NQE_FN_ABS ? "fabs(" : t.op == NQE_FN_SIN ? "sin(" : "cos(") << a << ");\n";
            o << "const bool " << k << " = t" << t.left << "k;\n";
            return;
        }
        emit_stmts(t.left, o, live, rn);
        emit_stmts(t.right, o, live, rn);
        const std::string a = "t" + std::to_string(t.left) + "v", b = "t" + std::to_string(t.right) + "v";
        const std::string ak = "t" + std::to_string(t.left) + "k", bk = "t" + std::to_string(t.right) + "k";
        const int lt = nodes[t.left].dtype;
        static const char *cmp[] = {"==", "!=", "<", "<=", ">", ">="};
        if (t.op <= NQE_OP_GT_EQ) {
            o << "const bool " << v << " = " << a << " " << cmp[t.op] << " " << b << ";\nconst bool " << k << " = " << ak << " && " << bk << ";\n";
        } else if (t.op == NQE_OP_AND || t.op == NQE_OP_OR) { // Kleene
            o << "const bool " << v << "a = " << a << " && " << ak << ", " << v << "b = " << b << " && " << bk << ";\n";
            if (t.op == NQE_OP_AND)
                o << "const bool " << v << " = " << v << "a && " << v << "b;\nconst bool " << k << " = (" << ak << " && " << bk << ") || (" << ak
                  << " && !" << v << "a) || (" << bk << " && !" << v << "b);\n";
            else
                o << "const bool " << v << " = " << v << "a || " << v << "b;\nconst bool " << k << " = (" << ak << " && " << bk << ") || " << v
                  << "a || " << v << "b;\n";
        } else {
            o << "const bool " << k << " = " << ak << " && " << bk << ";\n";
            const char *sfx = lt == NQE_INT64 ? "i64" : lt == NQE_UINT64 ? "u64" : "f64";
            const std::string lv = "(" + live + ") && " + k;
            if (t.op == NQE_OP_DIVIDE) o << "const " << T << " " << v << " = nqe_div_" << sfx << "(" << a << ", " << b << ", " << lv << ", p.status);\n";
            else if (t.op == NQE_OP_MODULOS) o << "const " << T << " " << v << " = nqe_mod_" << sfx << "(" << a << ", " << b << ", " << lv << ", p.status);\n";
            else if (lt == NQE_FLOAT64)
                o << "const double " << v << " = " << (t.op == NQE_OP_PLUS ? "__dadd_rn" : t.op == NQE_OP_MINUS ? "__dsub_rn" : "__dmul_rn") << "(" << a << ", " << b << ");\n";
            else {
                const char *sym = t.op == NQE_OP_PLUS ? "+" : t.op == NQE_OP_MINUS ? "-" : "*";
                o << "const " << T << " " << v << " = (" << T << ")((u64)" << a << " " << sym << " (u64)" << b << ");\n";
            }
        }
    }

    // typed C expression for row j of this thread
    std::string emit(int n, const char *live, const char *v = "c") {
        const TNode &t = nodes[n];
        std::ostringstream o;
        if (t.kind == NQE_NODE_COLUMN) {
            if (t.dtype == NQE_INT64) o << "((i64)" << v << t.col << "_[j])";
            else if (t.dtype == NQE_UINT64) o << "(" << v << t.col << "_[j])";
            else if (t.dtype == NQE_FLOAT64) o << "__longlong_as_double((i64)" << v << t.col << "_[j])";
            else o << "(" << v << t.col << "_[j] != 0)";
            return o.str();
        }
        if (t.kind == NQE_NODE_LITERAL) {
            if (t.dtype == NQE_INT64) o << "((i64)p.lit[" << t.lit << "])";
            else if (t.dtype == NQE_UINT64) o << "(p.lit[" << t.lit << "])";
            else if (t.dtype == NQE_FLOAT64) o << "__longlong_as_double((i64)p.lit[" << t.lit << "])";
            else o << "(p.lit[" << t.lit << "] != 0)";
            return o.str();
        }
        if (t.kind == NQE_NODE_UNARY) {
            const std::string a = emit(t.left, live, v);
            if (t.op == NQE_FN_ABS) return "fabs(" + a + ")";
            if (t.op == NQE_FN_SIN) return "sin(" + a + ")";
            return "cos(" + a + ")"; // Cos and Tan (unary.rs:96)
        }
        const std::string a = emit(t.left, live, v), b = emit(t.right, live, v);
        const int lt = nodes[t.left].dtype;
        static const char *cmp[] = {"==", "!=", "<", "<=", ">", ">="};
        if (t.op <= NQE_OP_GT_EQ) return "(" + a + " " + cmp[t.op] + " " + b + ")";
        if (t.op == NQE_OP_AND) return "(" + a + " & " + b + ")"; // both sides always evaluated
        if (t.op == NQE_OP_OR) return "(" + a + " | " + b + ")";
        const char *sfx = lt == NQE_INT64 ? "i64" : lt == NQE_UINT64 ? "u64" : "f64";
        if (t.op == NQE_OP_DIVIDE) return std::string("nqe_div_") + sfx + "(" + a + ", " + b + ", " + live + ", p.status)";
        if (t.op == NQE_OP_MODULOS) return std::string("nqe_mod_") + sfx + "(" + a + ", " + b + ", " + live + ", p.status)";
        if (lt == NQE_FLOAT64) {
            const char *fn = t.op == NQE_OP_PLUS ? "__dadd_rn" : t.op == NQE_OP_MINUS ? "__dsub_rn" : "__dmul_rn";
            return std::string(fn) + "(" + a + ", " + b + ")";
        }
        const char *sym = t.op == NQE_OP_PLUS ? "+" : t.op == NQE_OP_MINUS ? "-" : "*";
        const std::string r = "((u64)" + a + " " + sym + " (u64)" + b + ")"; // wrapping
        return lt == NQE_INT64 ? "((i64)" + r + ")" : r;
    }
};

struct Nvrtc {
    void *h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *);
    nvrtcResult (*DestroyProgram)(nvrtcProgram *);
    bool ok = false;
};

Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *e = getenv("NQE_JIT");
        if (e && !strcmp(e, "0")) return;
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
        for (const char *nm : names)
            if ((n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL))) break;
        if (!n.h) return;
#define NQE_SYM(f) *(void **)(&n.f) = dlsym(n.h, "nvrtc" #f); if (!n.f) return;
        NQE_SYM(CreateProgram) NQE_SYM(CompileProgram) NQE_SYM(GetCUBINSize) NQE_SYM(GetCUBIN)
        NQE_SYM(GetProgramLogSize) NQE_SYM(GetProgramLog) NQE_SYM(DestroyProgram)
#undef NQE_SYM
        n.ok = true;
    });
    return n;
}

struct JitParams {
    const void *col[16];
    void *out[16];
    uint64_t lit[32];
    int64_t n_rows;
    unsigned long long *tile_state;
    unsigned int *ticket;
    unsigned long long *out_count;
    uint32_t *status;
    int32_t num_tiles;
    int32_t pad;
    unsigned long long *prof;
    const uint32_t *valid[16];
    uint8_t *out_valid[16];
};

struct CachedKernel {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    bool failed = false;
    int occ = 0; // CTAs per SM at this kernel's block size / dynamic shared memory (0: not queried yet)
};

std::map<std::string, CachedKernel> &cache() {
    static std::map<std::string, CachedKernel> c;
    return c;
}
std::map<std::string, CachedKernel> &shape_cache() { // query shape -> kernel, in front of the source-text cache
    static std::map<std::string, CachedKernel> c;
    return c;
}
std::mutex &cache_mutex() {
    static std::mutex m;
    return m;
}

} // namespace

// Returns NQE_OK and sets *used = true when the specialised kernel was launched;
// *used = false means "not applicable, use the interpreter kernels".
static int32_t jit_filter_project_impl(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                                       int32_t n_projs, void *const *out_values, uint8_t *const *out_valid,
                                       unsigned long long *tile_state,
                                       unsigned int *ticket, unsigned long long *out_count, uint32_t *status, bool *used,
                                       std::string *source_out, bool allow_tma) {
    *used = false;
    Nvrtc &rt = nvrtc();
    if (!rt.ok && !source_out) return NQE_OK;
    static int CK = 0;
    if (!CK) {
        const char *e = getenv("NQE_JIT_K");
        CK = e ? atoi(e) : 4;
        if (CK != 1 && CK != 2 && CK != 4 && CK != 8) CK = 4;
    }
    int K = CK;
    Gen g;
    g.in = in;
    int pred_root = -1;
    if (predicate) {
        pred_root = g.parse(predicate);
        if (pred_root < 0) return NQE_OK;
    }
    const size_t n_pred_cols = g.col_of_slot.size();
    g.parsing_proj = true;
    std::vector<int> roots;
    for (int i = 0; i < n_projs; i++) {
        const int r = g.parse(&projs[i]);
        if (r < 0) return NQE_OK;
        roots.push_back(r);
    }
    // ---- generate the source (depends on the query shape only)
    std::ostringstream src;
    static int D = 0;
    if (!D) {
        const char *e = getenv("NQE_JIT_D");
        D = e ? atoi(e) : 4;
        if (D < 1 || D > 8) D = 4;
    }
    // TMA-ring variant: knobs NQE_JIT_IMPL=tma|ca, NQE_JIT_TMA_K / _SP / _SW / _LAG / _WALKERS / _LBW / _OCC
    static std::atomic<int> impl_guard{-1}; // published last: operators may be called from several host threads (multi.cu)
    static int TK = 8, SP = 2, SW = 2, LAG = 4, WALKERS = 2, lbw = 1, prof = 0;
    if (impl_guard.load(std::memory_order_acquire) < 0) {
        auto knob = [](const char *name, int dflt, int lo, int hi) {
            const char *e = getenv(name);
            const int v = e ? atoi(e) : dflt;
            return v < lo || v > hi ? dflt : v;
        };
        // defaults = best of the B200 sweeps in profiles/filter_project_sweeps_r01.md
        TK = knob("NQE_JIT_TMA_K", 8, 1, 16);
        if (TK & (TK - 1)) TK = 8;
        SP = knob("NQE_JIT_TMA_SP", 2, 2, 16);
        SW = knob("NQE_JIT_TMA_SW", 2, 2, 16);
        LAG = knob("NQE_JIT_TMA_LAG", 4, 0, 48);
        WALKERS = knob("NQE_JIT_TMA_WALKERS", 2, 1, 4);
        lbw = knob("NQE_JIT_TMA_LBW", 1, 1, 16); // look-back window: 32*lbw tiles per L2 round trip
        prof = getenv("NQE_JIT_PROF") ? 1 : 0;
        const char *e = getenv("NQE_JIT_IMPL");
        impl_guard.store((e && !strcmp(e, "ca")) ? 0 : 1, std::memory_order_release);
    }
    const int impl = impl_guard.load(std::memory_order_relaxed);
    bool tma = allow_tma && impl == 1 && predicate && n_pred_cols > 0;
    const size_t n_slots = g.col_of_slot.size();
    uint32_t pstage = 0, ptx = 0, wstage = 0, wtx = 0;
    std::vector<uint32_t> poff(n_slots), woff(n_slots), col_bytes(n_slots), pvoff(n_slots), wvoff(n_slots);
    int sp = SP, sw = SW, tk = TK;
    const bool nulls = g.any_nulls; // NULL-aware code: only the TMA-ring skeleton has it
    if ((nulls || g.any_via) && !tma) return NQE_OK; // gathered columns too
    if (g.any_via && nulls) return NQE_OK;
    if (tma) {
        bool any_proj = false;
        for (size_t s = 0; s < n_slots; s++) {
            if (g.via_slot[s] < 0 && ((uintptr_t)in->cols[g.col_of_slot[s]].values & 15)) tma = false;
            if ((uintptr_t)in->cols[g.col_of_slot[s]].validity & 15) tma = false;
            if (g.in_proj[s]) any_proj = true;
        }
        if (!any_proj) tma = false; // projections of literals only: nothing to stage
        // rows per thread: the largest K <= the knob whose two rings leave room for two CTAs per SM
        for (;; tk >>= 1) {
            const uint32_t tile = (uint32_t)tk * 256;
            pstage = ptx = wstage = wtx = 0;
            for (size_t s = 0; s < n_slots; s++) {
                const int sdt = in->cols[g.col_of_slot[s]].dtype;
                col_bytes[s] = g.via_slot[s] >= 0 ? 0 : sdt == NQE_BOOL ? tile / 8 : sdt == NQE_POS32 ? tile * 4 : tile * 8;
                const uint32_t padded = (col_bytes[s] + 127) & ~127u;
                if (s < n_pred_cols) { poff[s] = pstage; pstage += padded; ptx += col_bytes[s]; }
                if (g.in_proj[s]) { woff[s] = wstage; wstage += padded; wtx += col_bytes[s]; }
                if (g.has_valid[s]) { // validity bitmap of the tile, staged next to the values
                    const uint32_t vb = tile / 8, vpad = (vb + 127) & ~127u;
                    if (s < n_pred_cols) { pvoff[s] = pstage; pstage += vpad; ptx += vb; }
                    if (g.in_proj[s]) { wvoff[s] = wstage; wstage += vpad; wtx += vb; }
                }
            }
            if ((size_t)sp * pstage + (size_t)sw * wstage <= 104 * 1024 || tk == 1) break;
        }
        if ((size_t)sp * pstage + (size_t)sw * wstage > 200 * 1024) tma = false;
    }
    if ((nulls || g.any_via) && !tma) return NQE_OK;
    const int nr = LAG + sw + WALKERS + 2; // ring slots of per-tile state: covers claim .. write of a tile
    src << "#define NQE_PROF " << (tma ? prof : 0) << "\n#define NULLS " << (nulls ? 1 : 0) << "\n#define GATHERS " << (g.any_via ? 1 : 0) << "\n";
    if (tma) {
        K = tk;
        src << "#define K " << tk << "\n#define SP " << sp << "\n#define SW " << sw << "\n#define LAG " << LAG << "\n#define NR " << nr
            << "\n#define WALKERS " << WALKERS << "\n#define HAS_PRED 1\n#define NQE_LB_WIDE " << lbw << "\n#define NQE_LB_BACKOFF 40\n"
            << "#define PSTAGE_BYTES " << pstage << "u\n#define PRED_TX_BYTES " << ptx << "u\n#define WSTAGE_BYTES " << wstage
            << "u\n#define WRITE_TX_BYTES " << wtx << "u\n#define THREADS 256\n";
    } else {
        K = CK;
        src << "#define K " << K << "\n#define D " << D << "\n#define HAS_PRED " << (predicate ? 1 : 0) << "\n#define THREADS 256\n";
    }
    // column slots used by the predicate are parsed first, so they are slots [0, n_pred_cols)
    static int HINTS = -1;
    if (HINTS < 0) {
        const char *e = getenv("NQE_JIT_L2_HINTS");
        HINTS = e ? atoi(e) : 0; // measured: no gain on B200 (profiles/README_r01.md)
    }
    // ---- shape key: everything the generated source depends on.  A hit skips source generation altogether
    // (~40 us of string work per call, which would otherwise sit in front of every launch).
    std::string shape_key;
    {
        std::ostringstream k;
        k << tma << ',' << tk << ',' << sp << ',' << sw << ',' << LAG << ',' << WALKERS << ',' << lbw << ',' << prof << ',' << CK << ','
          << D << ',' << HINTS << ',' << (predicate ? 1 : 0) << ',' << n_pred_cols << ',' << pred_root << ',' << nulls << ';';
        for (size_t sl = 0; sl < n_slots; sl++)
            k << in->cols[g.col_of_slot[sl]].dtype << (g.in_proj[sl] ? 'p' : '-') << (g.has_valid[sl] ? 'v' : '-') << g.via_slot[sl] << ' ';
        k << ';';
        for (const TNode &t : g.nodes) k << t.kind << '.' << t.op << '.' << t.dtype << '.' << t.col << '.' << t.lit << '.' << t.left << '.' << t.right << '.' << t.is_null << ' ';
        k << ';';
        for (int r : roots) k << r << ' ';
        shape_key = k.str();
    }
    CachedKernel ck;
    bool have = false;
    if (!source_out) {
        std::lock_guard<std::mutex> lock(cache_mutex());
        auto it = shape_cache().find(shape_key);
        if (it != shape_cache().end()) {
            ck = it->second;
            have = true;
        }
    }
    if (!have) {
    std::vector<size_t> all_slots, pred_slots, proj_slots;
    for (size_t s = 0; s < n_slots; s++) {
        all_slots.push_back(s);
        if (s < n_pred_cols) pred_slots.push_back(s);
        if (g.in_proj[s]) proj_slots.push_back(s);
    }
    auto emit_decls = [&](std::ostringstream &o, const std::vector<size_t> &slots, const char *v) {
        for (size_t s : slots) {
            o << "u64 " << v << s << "_[K];\n";
            if (nulls && g.has_valid[s]) o << "u32 v" << s << "m = 0;\n"; // bit j = row j of this thread is valid
        }
    };
    // gathered columns (DevColumn::via) are loaded after the staged ones: value = base[position of this row]
    auto emit_gathers = [&](std::ostringstream &o, const std::vector<size_t> &slots, const char *v) {
        for (size_t s : slots)
            if (g.via_slot[s] >= 0)
                o << "_Pragma(\"unroll\") for (int j = 0; j < K; j++) " << v << s << "_[j] = __ldg(p.col[" << s << "] + " << v
                  << g.via_slot[s] << "_[j]);\n";
    };
    auto emit_loads = [&](std::ostringstream &o, const std::vector<size_t> &slots, const char *v, const char *e0, const char *full,
                          bool decl, const char *pol = nullptr) {
        for (size_t s : slots) {
            const int dt = in->cols[g.col_of_slot[s]].dtype;
            if (decl) o << "u64 " << v << s << "_[K];\n";
            if (g.via_slot[s] >= 0) continue;
            o << "_Pragma(\"unroll\") for (int j = 0; j < K; j++) { const i64 e = " << e0 << " + (i64)j * THREADS; ";
            if (dt == NQE_POS32)
                o << v << s << "_[j] = (" << full << " || e < p.n_rows) ? (u64)((const u32 *)p.col[" << s << "])[e] : 0ull; }\n";
            else if (dt == NQE_BOOL)
                o << v << s << "_[j] = (" << full << " || e < p.n_rows) ? ((((const u32 *)p.col[" << s << "])[e >> 5] >> (e & 31)) & 1u) : 0ull; }\n";
            else
                if (pol && HINTS && predicate)
                    o << v << s << "_[j] = (" << full << " || e < p.n_rows) ? ldg_keep(p.col[" << s << "] + e, " << pol << ") : 0ull; }\n";
                else
                    o << v << s << "_[j] = (" << full << " || e < p.n_rows) ? ldg_stream(p.col[" << s << "] + e) : 0ull; }\n";
            if (nulls && g.has_valid[s])
                o << "_Pragma(\"unroll\") for (int j = 0; j < K; j++) { const i64 e = " << e0 << " + (i64)j * THREADS; v" << s
                  << "m |= ((" << full << " || e < p.n_rows) ? ((p.valid[" << s << "][e >> 5] >> (e & 31)) & 1u) : 0u) << j; }\n";
        }
        emit_gathers(o, slots, v);
    };
    // operands of a staged (full) tile come from shared memory: row j*256+tid of the stage's column block
    auto emit_smem_loads = [&](std::ostringstream &o, const std::vector<size_t> &slots, const std::vector<uint32_t> &off,
                               const std::vector<uint32_t> &voff) {
        for (size_t s : slots) {
            const int dt = in->cols[g.col_of_slot[s]].dtype;
            if (g.via_slot[s] >= 0) continue;
            o << "_Pragma(\"unroll\") for (int j = 0; j < K; j++) ";
            if (dt == NQE_POS32)
                o << "c" << s << "_[j] = ((const u32 *)(stg + " << off[s] << "))[j * 256 + tid];\n";
            else if (dt == NQE_BOOL)
                o << "c" << s << "_[j] = (((const u32 *)(stg + " << off[s] << "))[j * 8 + warp] >> lane) & 1u;\n";
            else
                o << "c" << s << "_[j] = ((const u64 *)(stg + " << off[s] << "))[j * 256 + tid];\n";
            if (nulls && g.has_valid[s])
                o << "_Pragma(\"unroll\") for (int j = 0; j < K; j++) v" << s << "m |= ((((const u32 *)(stg + " << voff[s]
                  << "))[j * 8 + warp] >> lane) & 1u) << j;\n";
        }
        emit_gathers(o, slots, "c");
    };
    auto emit_copies = [&](std::ostringstream &o, const std::vector<size_t> &slots, const std::vector<uint32_t> &off,
                           const std::vector<uint32_t> &voff, bool pred_ring) {
        for (size_t s : slots) { // a predicate column that a projection reads again is kept in L2 for the write pass
            if (g.via_slot[s] >= 0) continue; // gathered: nothing contiguous to stage
            const char *pol = pred_ring && g.in_proj[s] ? "pol_keep" : "pol_stream";
            o << "bulk_g2s(dst + " << off[s] << ", (const u8 *)p.col[" << s << "] + (size_t)tile * " << col_bytes[s] << "u, " << col_bytes[s]
              << "u, bar, " << pol << ");\n";
            if (nulls && g.has_valid[s])
                o << "bulk_g2s(dst + " << voff[s] << ", (const u8 *)p.valid[" << s << "] + (size_t)tile * (K * 32), K * 32, bar, " << pol
                  << ");\n";
        }
    };
    std::ostringstream loads, pred_loads, next_decl, next_loads;
    emit_loads(loads, tma ? proj_slots : all_slots, "c", "e0", "full", !tma, "pol_stream");
    emit_loads(pred_loads, pred_slots, "c", "e0", "full", !tma, "pol_keep");
    for (size_t s = 0; s < n_pred_cols; s++) next_decl << "u64 n" << s << "_[K];\n";
    emit_loads(next_loads, pred_slots, "n", "n0", "nfull", false, "pol_keep");
    std::ostringstream decls, pred_decls, smem_loads, smem_pred_loads, pred_copies, write_copies;
    if (tma) {
        emit_decls(decls, proj_slots, "c");
        emit_decls(pred_decls, pred_slots, "c");
        emit_smem_loads(smem_loads, proj_slots, woff, wvoff);
        emit_smem_loads(smem_pred_loads, pred_slots, poff, pvoff);
        emit_copies(pred_copies, pred_slots, poff, pvoff, true);
        emit_copies(write_copies, proj_slots, woff, wvoff, false);
    }
    std::ostringstream stores, qstmts;
    std::string qvalid = "true", qvalue = "true";
    if (nulls) {
        g.emit_stmts(pred_root, qstmts, "inr", "false");
        qvalid = "t" + std::to_string(pred_root) + "k";
        qvalue = "t" + std::to_string(pred_root) + "v";
    }
    const bool rows_outer = nulls || g.any_via;
    if (rows_outer) {
        // write pass with the ROWS outermost: a row's operands (and their valid flags / gathered values) are loaded, used
        // by every output and dead again.  The NULL-aware and the gathered variants need this to stay within the 80
        // registers that two CTAs per SM allow.
        auto row_loads = [&](std::ostringstream &o, bool smem) {
            if (!smem) o << "const i64 e = e0 + (i64)j * THREADS; const bool inr = e < p.n_rows;\n";
            for (int pass = 0; pass < 2; pass++) // staged columns first, then the columns gathered through them
                for (size_t s : proj_slots) {
                    const int dt = in->cols[g.col_of_slot[s]].dtype;
                    if ((g.via_slot[s] >= 0) != (pass == 1)) continue;
                    if (pass == 1) {
                        o << "c" << s << "_[j] = __ldg(p.col[" << s << "] + c" << g.via_slot[s] << "_[j]);\n";
                    } else if (smem) {
                        if (dt == NQE_BOOL) o << "c" << s << "_[j] = (((const u32 *)(stg + " << woff[s] << "))[j * 8 + warp] >> lane) & 1u;\n";
                        else if (dt == NQE_POS32) o << "c" << s << "_[j] = ((const u32 *)(stg + " << woff[s] << "))[j * 256 + tid];\n";
                        else o << "c" << s << "_[j] = ((const u64 *)(stg + " << woff[s] << "))[j * 256 + tid];\n";
                        if (nulls && g.has_valid[s]) o << "v" << s << "m |= ((((const u32 *)(stg + " << wvoff[s] << "))[j * 8 + warp] >> lane) & 1u) << j;\n";
                    } else {
                        if (dt == NQE_BOOL) o << "c" << s << "_[j] = inr ? ((((const u32 *)p.col[" << s << "])[e >> 5] >> (e & 31)) & 1u) : 0ull;\n";
                        else if (dt == NQE_POS32) o << "c" << s << "_[j] = inr ? (u64)((const u32 *)p.col[" << s << "])[e] : 0ull;\n";
                        else o << "c" << s << "_[j] = inr ? ldg_stream(p.col[" << s << "] + e) : 0ull;\n";
                        if (nulls && g.has_valid[s]) o << "v" << s << "m |= (inr ? ((p.valid[" << s << "][e >> 5] >> (e & 31)) & 1u) : 0u) << j;\n";
                    }
                }
        };
        std::ostringstream body;
        for (int o = 0; o < n_projs; o++) {
            const TNode &r = g.nodes[roots[o]];
            if (nulls) stores << "u8 *const ov" << o << " = p.out_valid[" << o << "] ? p.out_valid[" << o << "] + tile_excl : (u8 *)0;\n";
            if (r.dtype == NQE_BOOL) stores << "u8 *const out" << o << " = (u8 *)p.out[" << o << "] + tile_excl;\n";
            else stores << "u64 *const out" << o << " = (u64 *)p.out[" << o << "] + tile_excl;\n";
            body << "{\n";
            if (nulls) {
                const std::string rv = "t" + std::to_string(roots[o]) + "v", rk = "t" + std::to_string(roots[o]) + "k";
                g.emit_stmts(roots[o], body, "KEEPJ && !RNJ", "RNJ");
                body << "if (KEEPJ) { out" << o << "[idx[j]] = " << rk << " ? ";
                if (r.dtype == NQE_FLOAT64) body << "(u64)__double_as_longlong(" << rv << ")";
                else if (r.dtype == NQE_BOOL) body << "(u8)" << rv;
                else body << "(u64)" << rv;
                body << " : 0; if (ov" << o << ") ov" << o << "[idx[j]] = (u8)" << rk << "; } }\n";
            } else {
                body << "const " << Gen::ctype(r.dtype) << " v = " << g.emit(roots[o], "keep[j]") << "; if (keep[j]) out" << o << "[idx[j]] = ";
                if (r.dtype == NQE_FLOAT64) body << "(u64)__double_as_longlong(v)";
                else if (r.dtype == NQE_BOOL) body << "(u8)v";
                else body << "(u64)v";
                body << "; }\n";
            }
        }
        stores << "if (full) {\n_Pragma(\"unroll\") for (int j = 0; j < K; j++) {\n";
        row_loads(stores, true);
        stores << body.str() << "} } else {\n_Pragma(\"unroll\") for (int j = 0; j < K; j++) {\n";
        row_loads(stores, false);
        stores << body.str() << "} }\n";
        smem_loads.str("");
        loads.str("");
    }
    for (int o = 0; !rows_outer && o < n_projs; o++) {
        const TNode &r = g.nodes[roots[o]];
        stores << "{ ";
        if (r.dtype == NQE_BOOL) stores << "u8 *out = (u8 *)p.out[" << o << "] + tile_excl;\n";
        else stores << "u64 *out = (u64 *)p.out[" << o << "] + tile_excl;\n";
        stores << "_Pragma(\"unroll\") for (int j = 0; j < K; j++) { const " << Gen::ctype(r.dtype) << " v = "
               << g.emit(roots[o], "keep[j]") << "; if (keep[j]) out[idx[j]] = ";
        if (r.dtype == NQE_FLOAT64) stores << "(u64)__double_as_longlong(v)";
        else if (r.dtype == NQE_BOOL) stores << "(u8)v";
        else stores << "(u64)v";
        stores << "; } }\n";
    }
    std::string pred_expr = predicate ? g.emit(pred_root, "inr") : "true";
    std::string next_pred_expr = predicate ? g.emit(pred_root, "inr", "n") : "true";
    auto replace_all = [](std::string s, const std::string &a, const std::string &b) {
        size_t pos = 0;
        while ((pos = s.find(a, pos)) != std::string::npos) {
            s.replace(pos, a.size(), b);
            pos += b.size();
        }
        return s;
    };
    std::string kernel = tma ? kSkeletonTma : kSkeletonKernel;
    if (tma) {
        kernel = replace_all(kernel, "QSTMTS", qstmts.str());
        kernel = replace_all(kernel, "QVALID", qvalid);
        kernel = replace_all(kernel, "QVALUE", qvalue);
        kernel = replace_all(kernel, "ISSUE_PRED_COPIES", pred_copies.str());
        kernel = replace_all(kernel, "ISSUE_WRITE_COPIES", write_copies.str());
        kernel = replace_all(kernel, "DECL_PRED_COLUMNS", pred_decls.str());
        kernel = replace_all(kernel, "DECL_COLUMNS", decls.str());
        kernel = replace_all(kernel, "SMEM_LOAD_PRED_COLUMNS", smem_pred_loads.str());
        kernel = replace_all(kernel, "SMEM_LOAD_COLUMNS", smem_loads.str());
    }
    kernel = replace_all(kernel, "LOAD_PRED_COLUMNS", pred_loads.str());
    kernel = replace_all(kernel, "DECL_NEXT_COLUMNS", next_decl.str());
    kernel = replace_all(kernel, "LOAD_NEXT_COLUMNS", next_loads.str());
    kernel = replace_all(kernel, "LOAD_COLUMNS", loads.str());
    kernel = replace_all(kernel, "NEXT_PRED_EXPR", next_pred_expr);
    kernel = replace_all(kernel, "PRED_EXPR", pred_expr);
    kernel = replace_all(kernel, "STORE_OUTPUTS", stores.str());
    const std::string source = src.str() + kSkeletonHead + kLookbackSrc + kernel;
    if (source_out) {
        *source_out = source;
        if (!rt.ok) return NQE_OK;
    }

    // ---- compile (cached per source text)
    {
        std::lock_guard<std::mutex> lock(cache_mutex());
        auto it = cache().find(source);
        if (it != cache().end()) ck = it->second;
        else {
            nvrtcProgram prog;
            if (rt.CreateProgram(&prog, source.c_str(), "nqe_fp_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) ck.failed = true;
            if (!ck.failed) {
                const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=false"};
                const nvrtcResult r = rt.CompileProgram(prog, 4, opts);
                if (r != NVRTC_SUCCESS) {
                    size_t n = 0;
                    rt.GetProgramLogSize(prog, &n);
                    std::string log(n, 0);
                    rt.GetProgramLog(prog, &log[0]);
                    nqe_fail(ctx, NQE_ERR_CUDA, "NVRTC compile failed: %s", log.c_str());
                    if (getenv("NQE_JIT_DEBUG")) fprintf(stderr, "NVRTC: %s\n%s\n", log.c_str(), source.c_str());
                    ck.failed = true;
                } else {
                    size_t n = 0;
                    rt.GetCUBINSize(prog, &n);
                    std::vector<char> cubin(n);
                    rt.GetCUBIN(prog, cubin.data());
                    if (const char *dump = getenv("NQE_JIT_DUMP")) { // <dump>.cu / <dump>.cubin for cuobjdump -sass
                        if (FILE *f = fopen((std::string(dump) + ".cu").c_str(), "w")) { fputs(source.c_str(), f); fclose(f); }
                        if (FILE *f = fopen((std::string(dump) + ".cubin").c_str(), "wb")) { fwrite(cubin.data(), 1, n, f); fclose(f); }
                    }
                    if (cudaLibraryLoadData(&ck.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
                        cudaLibraryGetKernel(&ck.kernel, ck.lib, "nqe_fp_jit") != cudaSuccess) {
                        cudaGetLastError();
                        ck.failed = true;
                    }
                }
                rt.DestroyProgram(&prog);
            }
            cache()[source] = ck;
        }
    }
    {
        std::lock_guard<std::mutex> lock(cache_mutex());
        shape_cache()[shape_key] = ck;
    }
    }
    if (ck.failed) return NQE_OK; // interpreter kernels take over

    JitParams jp;
    memset(&jp, 0, sizeof jp);
    for (size_t s = 0; s < g.col_of_slot.size(); s++) jp.col[s] = in->cols[g.col_of_slot[s]].values;
    for (int o = 0; o < n_projs; o++) jp.out[o] = out_values[o];
    for (size_t s = 0; s < g.col_of_slot.size(); s++) jp.valid[s] = (const uint32_t *)in->cols[g.col_of_slot[s]].validity;
    for (int o = 0; o < n_projs; o++) jp.out_valid[o] = out_valid ? out_valid[o] : nullptr;
    for (size_t i = 0; i < g.lits.size(); i++) jp.lit[i] = g.lits[i];
    jp.n_rows = in->nrows;
    jp.tile_state = tile_state;
    jp.ticket = ticket;
    jp.out_count = out_count;
    jp.status = status;
    const int64_t tile = (int64_t)K * 256;
    jp.num_tiles = (int32_t)((in->nrows + tile - 1) / tile);
    if (jp.num_tiles == 0) { *used = true; return NQE_OK; }
    int grid = jp.num_tiles;
    int block = 256;
    size_t dyn = 0;
    if (tma) {
        block = 320 + 32 * WALKERS; // 8 worker warps + loader + publisher + the walker warps
        dyn = (size_t)sp * pstage + (size_t)sw * wstage;
        int occ = ck.occ;
        if (occ == 0) { // first launch of this kernel: opt in to the shared memory, query the occupancy, remember both
            if (cudaFuncSetAttribute((const void *)ck.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess ||
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)ck.kernel, block, dyn) != cudaSuccess || occ < 1) {
                cudaGetLastError();
                return NQE_OK; // the caller retries with the register-staged variant
            }
            std::lock_guard<std::mutex> lock(cache_mutex());
            shape_cache()[shape_key].occ = occ;
        }
        static int max_occ = -1; // knob: cap CTAs per SM (NQE_JIT_TMA_OCC)
        if (max_occ < 0) {
            const char *e = getenv("NQE_JIT_TMA_OCC");
            max_occ = e ? atoi(e) : 0;
        }
        if (max_occ > 0 && occ > max_occ) occ = max_occ;
        grid = ctx->sm_count * occ;
        if (grid > jp.num_tiles) grid = jp.num_tiles;
    } else if (predicate) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)ck.kernel, 256, 0) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            occ = 4;
        }
        grid = ctx->sm_count * occ;
        if (grid > jp.num_tiles) grid = jp.num_tiles;
    }
    unsigned long long *d_prof = nullptr;
    if (tma && prof) {
        cudaMalloc(&d_prof, 32 * 8);
        cudaMemsetAsync(d_prof, 0, 32 * 8, ctx->stream);
        jp.prof = d_prof;
    }
    void *args[] = {&jp};
    // the operator timer brackets the kernel, not the host-side preparation in front of it
    if (ctx->timer_depth == 1) cudaEventRecord(ctx->ev0, ctx->stream);
    NQE_CUDA(ctx, cudaLaunchKernel((const void *)ck.kernel, dim3(grid), dim3(block), args, dyn, ctx->stream));
    ctx->launches++;
    if (d_prof) { // per-phase cycle averages of the control warp / worker warp 0 (tuning aid)
        unsigned long long h[32];
        cudaMemcpyAsync(h, d_prof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(d_prof);
        const double it = h[5] ? (double)h[5] : 1.0;
        fprintf(stderr, "[nqe jit prof] grid %d iters/cta %.1f | publisher: wait_cnt %.0f scan+publish %.0f; walkers: wait_pub %.0f walk %.0f; loader: wait_free %.0f issue %.0f"
                        " | worker: idle %.0f count %.0f blocked %.0f write %.0f (cycles per tile)\n",
                grid, it / grid, h[0] / it, h[1] / it, h[6] / it, h[2] / it, h[3] / it, h[4] / it, h[8] / it, h[9] / it, h[10] / it, h[11] / it);
        fprintf(stderr, "[nqe jit prof] tile life, avg (max) cycles: claim->count %.0f (%llu) count->publish %.0f (%llu) publish->prefix %.0f (%llu)"
                        " prefix->write %.0f (%llu) write %.0f (%llu)\n",
                h[16] / it, h[24], h[17] / it, h[25], h[18] / it, h[26], h[19] / it, h[27], h[20] / it, h[28]);
    }
    *used = true;
    return NQE_OK;
}

int32_t nqe_jit_filter_project(nqe_ctx *ctx, const nqe_table *in, const nqe_expr *predicate, const nqe_expr *projs,
                               int32_t n_projs, void *const *out_values, uint8_t *const *out_valid,
                               unsigned long long *tile_state,
                               unsigned int *ticket, unsigned long long *out_count, uint32_t *status, bool *used,
                               std::string *source_out) {
    int32_t rc = jit_filter_project_impl(ctx, in, predicate, projs, n_projs, out_values, out_valid, tile_state, ticket, out_count,
                                         status, used, source_out, true);
    if (rc == NQE_OK && !*used && !source_out)
        rc = jit_filter_project_impl(ctx, in, predicate, projs, n_projs, out_values, out_valid, tile_state, ticket, out_count, status,
                                     used, source_out, false);
    return rc;
}

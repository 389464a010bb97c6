// hash_join.cu -- HashJoin::build / probe (hash_join.rs:124-254) and the fused
// join -> PhysicalAggregatePlan path.
//
// Build (hash_join.rs:58-78): the build side's raw key words (validity is ignored,
// :67) go into one open-addressing multimap; a slot is claimed with a CAS on its row
// word.  Slots are 16 bytes {key, build row} -- or {key, payload} when the plan needs exactly one
// build-side value per match (JoinTable::rowpay) -- and are inserted with one 128-bit CAS, which
// also tells the build whether any key repeats, so unique-key probes stop at the first match.
//
// Probe (hash_join.rs:80-103, 236-246): one pass over the probe side.  Every tile looks
// its rows up (the first table probe of a thread's K rows is issued together, collisions
// walk sequentially), ranks the matches with warp shuffles + a decoupled look-back across
// tiles (so the output is probe-row-major like the reference's), and writes the joined row
// straight to its final position: no (outer_pos, inner_pos) index arrays and no separate
// `take` pass.  Within one probe row, matches are emitted in ascending build-row order.
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "agg_device.cuh"
#include "hash_common.cuh"
#include "nqe_internal.cuh"
#include "paged_split.cuh"

int32_t nqe_pack_bytes(nqe_ctx *ctx, const uint8_t *bytes, int64_t n, uint32_t *words, unsigned long long *zeros);

int32_t nqe_hash_join_strings(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right, int32_t left_key,
                              int32_t right_key, nqe_table **out);

namespace {

constexpr int HJ_THREADS = 256;
constexpr int HJ_WARPS = HJ_THREADS / 32;
constexpr int HJ_K = 4;
constexpr int HJ_MAX_COLS = 16;
constexpr unsigned long long EMPTY_ROW = ~0ull;

struct Slot {
    unsigned long long key;
    unsigned long long row;
};

struct JoinTable {
    unsigned long long *words; // slot s = words[2 s], words[2 s + 1]
    uint64_t cap;     // number of slots (any size: slots are chosen by multiply-high)
    int32_t has_dups;
    int32_t key_col;  // build-side key column
    // "payload in the row word": when the plan needs exactly ONE build-side value per match (the only non-key
    // build column of a join, or the group key of a fused join -> group-by) the slot's second word holds that
    // value instead of the build row number, so a match costs one random access, not two (slot, then column[row]).
    // A payload equal to EMPTY_ROW cannot be stored (it marks a free slot): the build reports it and the host
    // rebuilds with row numbers.
    const unsigned long long *rowpay;
    int32_t rowpay_col;
    // "direct" table (dense unique build keys, e.g. a primary key): no hashing and no key compares -- slot d = key - lo
    // is ONE 8-byte word, the row word of the build row with that key (EMPTY_ROW: no such key), cap = hi - lo + 1.  A
    // probe is exactly one random 8-byte read and the table is 8 bytes per key of the range instead of 32 per build
    // row: 1e7 keys = 80 MB, most of which stays in the 126 MB L2 (scratch/gather_bw.cu: 1e8 random 8-byte reads cost
    // 0.50 ms out of <= 60 MB, 0.74 ms out of 80 MB, 1.33 ms out of 160 MB, 2.0 ms out of 320 MB), so the probe side
    // needs no split by slot range.  A duplicate key found while building sends the host back to the hashed table.
    int32_t direct;
    long long lo;
    // narrow direct table: 4-byte slots.  Row numbers always fit (build sides stay below 2^32 - 1 rows); a payload fits
    // when its column spans less than 2^32 - 1 values and is then stored as value - pay_lo (frame of reference).
    // 0xffffffff marks a free slot.  Half the bytes = twice the keys that stay L2-resident: 1e7 keys are 40 MB.
    int32_t narrow, pad;
    long long pay_lo;
};

struct ColSrc {
    const void *values;
    const uint32_t *validity;
    int32_t dtype;
    int32_t pad;
};

__device__ __forceinline__ unsigned long long *slot_ptr(const JoinTable &jt, uint64_t s) { return jt.words + (s << 1); }
// Random accesses.  Measured on B200 (scratch/l2gran.cu, profiles/join_groupby_r01.md): an L2 sector miss
// reads 128 bytes from DRAM (the whole line is installed) whatever the load flavour (.nc, .cg,
// .L1::no_allocate) and whatever cudaLimitMaxL2FetchGranularity says; the .L2::64B prefetch-size
// qualifier halves the bytes but not the time -- random probes are bound by the DRAM request rate
// (~1.7e10 /s), not by bytes -- so the plain read-only path is kept.
__device__ __forceinline__ ulonglong2 ld_cg_v2(const unsigned long long *p) { return __ldg((const ulonglong2 *)p); }
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p) { return __ldg(p); }
// Home slot of a key: multiply-high range reduction, so the capacity need not be a power of two.  Home slots are EVEN
// (the capacity is even): a probe sequence starts on a 32-byte sector boundary, so its second step reads the other half
// of the sector the first one just brought into L1 -- every second step of a sequence is an L1 hit instead of an L2
// round trip.  (Reading both slots with one 32-byte load was tried: 32 more registers for K = 4 rows, spills under the
// 64-register cap of the fused kernels.)
__device__ __forceinline__ uint64_t join_slot_of(const JoinTable &jt, unsigned long long key) {
    return __umul64hi(nqe_mix64(key), jt.cap >> 1) << 1;
}
__device__ __forceinline__ Slot ld_slot(const JoinTable &jt, uint64_t s) {
    const ulonglong2 v = ld_cg_v2(slot_ptr(jt, s));
    return Slot{v.x, v.y};
}

__global__ void join_clear_kernel(JoinTable jt, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned long long *p = slot_ptr(jt, i);
        p[0] = 0;
        p[1] = EMPTY_ROW;
    }
}

// 128-bit compare-and-swap of a whole {key, row} slot (ATOMG.CAS.128): the slot becomes visible with both
// words at once, so an insert that walks over an occupied slot can compare keys reliably -- duplicate
// build keys are detected while building, without a second pass over the build side.
__device__ __forceinline__ ulonglong2 cas128(unsigned long long *p, ulonglong2 cmp, ulonglong2 val) {
    ulonglong2 old;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.global.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old.x), "=l"(old.y) : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(p) : "memory");
    return old;
}

__global__ void join_build_kernel(JoinTable jt, const unsigned long long *__restrict__ keys, int64_t n, uint32_t *status,
                                  uint32_t *dupflag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = keys[i];
    const unsigned long long rowword = jt.rowpay ? jt.rowpay[i] : (unsigned long long)i;
    if (rowword == EMPTY_ROW) { // only possible for a payload
        dupflag[2] = 1u;
        return;
    }
    uint64_t s = join_slot_of(jt, key);
    bool dup = false;
    for (uint64_t probe = 0; probe < jt.cap; probe++) {
        unsigned long long *p = slot_ptr(jt, s);
        const ulonglong2 old = cas128(p, make_ulonglong2(0ull, EMPTY_ROW), make_ulonglong2(key, rowword));
        if (old.y == EMPTY_ROW) {
            if (dup) *dupflag = 1u;
            return;
        }
        if (old.x == key) dup = true; // an equal key was inserted earlier on this probe sequence
        s = s + 1 == jt.cap ? 0 : s + 1;
    }
    atomicOr(status, DEV_ERR_TABLE_FULL);
}

__global__ void join_direct_build_kernel(JoinTable jt, const unsigned long long *__restrict__ keys, int64_t n, uint32_t *dupflag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long rowword = jt.rowpay ? jt.rowpay[i] : (unsigned long long)i;
    if (rowword == EMPTY_ROW) { // only possible for a payload; EMPTY_ROW is what a probe returns for "no match"
        dupflag[2] = 1u;
        return;
    }
    const unsigned long long d = keys[i] - (unsigned long long)jt.lo; // < cap: lo and cap come from the keys' own min / max
    if (jt.narrow) {
        const unsigned int w = (unsigned int)(rowword - (unsigned long long)jt.pay_lo); // < 0xffffffff by the host's range check
        if (atomicCAS((unsigned int *)jt.words + d, 0xffffffffu, w) != 0xffffffffu) *dupflag = 1u;
    } else if (atomicCAS(jt.words + d, EMPTY_ROW, rowword) != EMPTY_ROW) {
        *dupflag = 1u;
    }
}

// streaming accesses of the partitioned probe carry an L2 evict_first policy, so that the slot range being
// probed (default policy) is what stays resident
__device__ __forceinline__ unsigned long long pj_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long ld_ef(const unsigned long long *p, unsigned long long pol) {
    unsigned long long v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_ef(unsigned long long *p, unsigned long long v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}

// table reads of the direct probe: evict_last, so that the 8-byte-per-key table outlives the streams flowing past it
__device__ __forceinline__ unsigned long long jt_policy_keep() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long ld_keep(const unsigned long long *p, unsigned long long pol) {
    unsigned long long v;
    asm volatile("ld.global.nc.L2::cache_hint.b64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}

// Direct table: one 8-byte read per row, nothing to compare and nothing to walk.
template <int K, bool KEEP = false>
__device__ __forceinline__ void probe_direct(const JoinTable &jt, const unsigned long long (&key)[K], uint32_t want,
                                             unsigned long long (&brow)[K], uint64_t (&slot)[K]) {
    const unsigned long long pol = KEEP ? jt_policy_keep() : 0ull;
#pragma unroll
    for (int j = 0; j < K; j++) {
        slot[j] = key[j] - (unsigned long long)jt.lo;
        brow[j] = EMPTY_ROW;
    }
    if (jt.narrow) {
        unsigned int w[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            w[j] = 0xffffffffu;
            if (((want >> j) & 1u) && slot[j] < jt.cap) w[j] = __ldg((const unsigned int *)jt.words + slot[j]);
        }
#pragma unroll
        for (int j = 0; j < K; j++)
            if (w[j] != 0xffffffffu) brow[j] = (unsigned long long)w[j] + (unsigned long long)jt.pay_lo;
        return;
    }
#pragma unroll
    for (int j = 0; j < K; j++)
        if (((want >> j) & 1u) && slot[j] < jt.cap) brow[j] = KEEP ? ld_keep(jt.words + slot[j], pol) : ld_cg_u64(jt.words + slot[j]);
}

// First slot (in probe order) holding `key`, for K probe rows.  The first table probe of all
// K rows is issued together (K independent loads in flight; at load factor <= 0.6 most rows
// resolve there), collisions are then walked one row at a time.  (Advancing the K sequences together -- one
// table round trip per pass over all rows instead of per row and step -- was measured and is SLOWER: fused
// join -> group-by 5.17 -> 5.66 ms, paged variant 4.78 -> 5.18 ms; the per-row loop is tighter and most lanes
// never enter it.)
template <int K>
__device__ __forceinline__ void probe_first(const JoinTable &jt, const unsigned long long (&key)[K], uint32_t want,
                                            unsigned long long (&brow)[K], uint64_t (&slot)[K]) {
    if (jt.direct) {
        probe_direct<K>(jt, key, want, brow, slot);
        return;
    }
    Slot first[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        slot[j] = join_slot_of(jt, key[j]);
        if ((want >> j) & 1u) first[j] = ld_slot(jt, slot[j]);
    }
#pragma unroll
    for (int j = 0; j < K; j++) {
        brow[j] = EMPTY_ROW;
        if (!((want >> j) & 1u)) continue;
        Slot sl = first[j];
        while (sl.row != EMPTY_ROW) {
            if (sl.key == key[j]) { brow[j] = sl.row; break; }
            slot[j] = slot[j] + 1 == jt.cap ? 0 : slot[j] + 1;
            sl = ld_slot(jt, slot[j]); // odd slot: the other half of the sector just read, an L1 hit
        }
    }
}

// number of matches of `key`; the smallest matching build row and its slot
__device__ __forceinline__ unsigned int probe_count(const JoinTable &jt, unsigned long long key, unsigned long long *first,
                                                    uint64_t *first_slot) {
    uint64_t s = join_slot_of(jt, key);
    unsigned int c = 0;
    unsigned long long best = EMPTY_ROW;
    while (true) {
        const Slot sl = ld_slot(jt, s);
        if (sl.row == EMPTY_ROW) break;
        if (sl.key == key) {
            c++;
            if (sl.row < best) { best = sl.row; *first_slot = s; }
        }
        s = s + 1 == jt.cap ? 0 : s + 1;
    }
    *first = best;
    return c;
}

// smallest matching build row strictly greater than `after` (and its slot)
__device__ __forceinline__ unsigned long long probe_next(const JoinTable &jt, unsigned long long key, unsigned long long after,
                                                         uint64_t *slot_out) {
    uint64_t s = join_slot_of(jt, key);
    unsigned long long best = EMPTY_ROW;
    while (true) {
        const Slot sl = ld_slot(jt, s);
        if (sl.row == EMPTY_ROW) break;
        if (sl.key == key && sl.row > after && sl.row < best) { best = sl.row; *slot_out = s; }
        s = s + 1 == jt.cap ? 0 : s + 1;
    }
    return best;
}

// ---------------------------------------------------------------------------
// Partitioned probe (unique build keys, table larger than the L2).
//
// A probe row costs one random table access; for a table that does not fit in the 126 MB L2
// that is ~128 bytes of DRAM traffic per row (measured) and the probe runs at DRAM
// random-access speed.  Here the probe KEYS are first split -- stably -- into 2^log2p
// partitions by the top bits of their hash, which are the top bits of their table slot, so
// partition p only touches slot range p of the (unchanged) table and that range stays in L2
// while its keys are probed.  The per-row results land in partition order in sequential
// streams; the gather pass brings them back into the probe side's ORIGINAL row order (every row
// remembers its position in the partitioned order; reads are sequential per partition stream),
// leaving a match bitmap, and the joined rows are compacted -- probe-row-major exactly like the
// direct path -- by the filter/project kernel with that bitmap as the predicate.  Everything
// except the L2-resident table is streamed.
// ---------------------------------------------------------------------------
constexpr int PJ_MAX_PARTS = 32; // one lane per partition in the tile histograms

struct PartJoin {
    const unsigned long long *keys; // probe keys, original order
    int64_t n;
    int32_t log2p;
    int32_t num_tiles;              // tiles of PJ_K * HJ_THREADS rows
    unsigned int *tile_cnt;         // [P][num_tiles]: rows of the tile per partition
    unsigned int *tile_off;         // [P][num_tiles]: rows of that partition in earlier tiles
    unsigned long long *part_base;  // [P + 1]: first position of the partition in the partitioned order
    unsigned long long *pkeys;      // keys in partitioned order
    unsigned int *ppos32;           // per probe row (original order): its position in the partitioned order
    unsigned long long *res0;       // per position: build row or payload (EMPTY_ROW: no match)
};

__device__ __forceinline__ int pj_part(unsigned long long key, int log2p) {
    return log2p ? (int)(nqe_mix64(key) >> (64 - log2p)) : 0;
}

constexpr int PJ_K = 8; // rows per thread in the split passes: 2048-row tiles

// Warp-level partition histogram of K row groups (32 rows each) from log2p ballot bit-planes:
//   rank[j]   = number of lower lanes of this row group in the same partition as this lane's row
//   cnt[j]    = rows of this row group that fall into partition `lane`   (lane < 2^log2p)
// No shared memory and no atomics; row order inside a tile is j major, warp minor, lane.
template <int K>
__device__ __forceinline__ void pj_warp_hist(const unsigned long long (&key)[K], uint32_t inrange, int log2p, int (&pid)[K],
                                             unsigned (&rank)[K], unsigned (&cnt)[K]) {
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const bool live = (inrange >> j) & 1u;
        pid[j] = pj_part(key[j], log2p);
        unsigned own = __ballot_sync(0xffffffffu, live), bucket = own;
        for (int b = 0; b < log2p; b++) {
            const unsigned bb = __ballot_sync(0xffffffffu, (pid[j] >> b) & 1);
            own &= ((pid[j] >> b) & 1) ? bb : ~bb;
            bucket &= ((lane >> b) & 1) ? bb : ~bb;
        }
        rank[j] = __popc(own & ltmask);
        cnt[j] = __popc(bucket);
    }
}

// pass 2: one CTA per partition: exclusive scan of its per-tile counts (contiguous); partition bases
__global__ void __launch_bounds__(1024) pj_scan_kernel(PartJoin pj, const unsigned long long *totals) {
    __shared__ unsigned int s_w[32];
    __shared__ unsigned int s_carry;
    const int p = blockIdx.x, P = 1 << pj.log2p, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int *cnt = pj.tile_cnt + (size_t)p * pj.num_tiles;
    unsigned int *off = pj.tile_off + (size_t)p * pj.num_tiles;
    if (tid == 0) {
        unsigned long long b = 0;
        for (int q = 0; q < p; q++) b += totals[q];
        pj.part_base[p] = b;
        if (p == P - 1) pj.part_base[P] = b + totals[p];
        s_carry = 0;
    }
    __syncthreads();
    for (int t0 = 0; t0 < pj.num_tiles; t0 += 1024) {
        const int t = t0 + tid;
        const unsigned v = t < pj.num_tiles ? cnt[t] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned w = s_w[lane];
            unsigned wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned x = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += x;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const unsigned carry = s_carry;
        if (t < pj.num_tiles) off[t] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_w[warp] + incl;
        __syncthreads();
    }
}

// pass 3: stable split of the probe keys; every row also remembers its position in the partitioned order
__global__ void __launch_bounds__(HJ_THREADS) pj_scatter_kernel(PartJoin pj) {
    constexpr int K = PJ_K, TILE = K * HJ_THREADS;
    __shared__ unsigned short s_c[K * HJ_WARPS][PJ_MAX_PARTS];
    __shared__ unsigned long long s_base[PJ_MAX_PARTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, P = 1 << pj.log2p;
    const unsigned long long pol = pj_policy();
    for (int tile = blockIdx.x; tile < pj.num_tiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TILE + tid;
        unsigned long long key[K];
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            key[j] = e < pj.n ? ld_ef(pj.keys + e, pol) : 0ull;
            if (e < pj.n) inrange |= 1u << j;
        }
        int pid[K];
        unsigned rank[K], cnt[K];
        pj_warp_hist<K>(key, inrange, pj.log2p, pid, rank, cnt);
#pragma unroll
        for (int j = 0; j < K; j++) s_c[j * HJ_WARPS + warp][lane] = (unsigned short)cnt[j];
        if (tid < P) s_base[tid] = pj.part_base[tid] + pj.tile_off[(size_t)tid * pj.num_tiles + tile];
        __syncthreads();
        if (tid < P) { // exclusive prefix over the (j, warp) pairs, per partition
            unsigned run = 0;
#pragma unroll
            for (int i = 0; i < K * HJ_WARPS; i++) {
                const unsigned c = s_c[i][tid];
                s_c[i][tid] = (unsigned short)run;
                run += c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < K; j++)
            if ((inrange >> j) & 1u) {
                const unsigned long long ppos = s_base[pid[j]] + s_c[j * HJ_WARPS + warp][pid[j]] + rank[j];
                st_ef(pj.pkeys + ppos, key[j], pol);
                pj.ppos32[e0 + (int64_t)j * HJ_THREADS] = (unsigned int)ppos;
            }
        __syncthreads(); // s_c / s_base are reused by the next tile
    }
}

// ---- split, second generation (knob NQE_JOIN_SPLIT=2; 3 = atomic count + the stable scatter above).  Same three passes and the same
// tile-major order of every partition's stream (an order that follows the probe rows keeps the gather pass's reads
// nearly sequential -- claiming runs with a global atomic instead of scanning tile counts was measured: the
// scatter got faster but the gather's DRAM reads went from 1.3 to 3.5 GB), but
//   * histograms and ranks come from shared-memory atomics (one ATOMS per row, the returned value is the row's rank
//     inside its tile's run) instead of log2p ballot bit-planes per row group,
//   * the keys are staged in shared memory in partition-major order and leave as contiguous runs: coalesced stores
//     instead of up to 2^log2p distinct lines per warp store.
// Inside a tile's run the order is arrival order, which nothing depends on: every row remembers its position.
__global__ void __launch_bounds__(HJ_THREADS) pj2_count_kernel(PartJoin pj, unsigned long long *totals) {
    constexpr int K = PJ_K, TILE = K * HJ_THREADS;
    __shared__ unsigned int s_hist[PJ_MAX_PARTS];
    const int tid = threadIdx.x, P = 1 << pj.log2p;
    unsigned long long mine = 0;
    if (tid < PJ_MAX_PARTS) s_hist[tid] = 0;
    __syncthreads();
    for (int tile = blockIdx.x; tile < pj.num_tiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TILE + tid;
        unsigned long long key[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            key[j] = e < pj.n ? ld_stream_u64(pj.keys + e) : 0ull;
        }
#pragma unroll
        for (int j = 0; j < K; j++)
            if (e0 + (int64_t)j * HJ_THREADS < pj.n) atomicAdd(&s_hist[pj_part(key[j], pj.log2p)], 1u);
        __syncthreads();
        if (tid < P) {
            const unsigned c = s_hist[tid];
            pj.tile_cnt[(size_t)tid * pj.num_tiles + tile] = c;
            mine += c;
            s_hist[tid] = 0;
        }
        __syncthreads();
    }
    if (tid < P && mine) atomicAdd(totals + tid, mine);
}

__global__ void __launch_bounds__(HJ_THREADS) pj2_scatter_kernel(PartJoin pj) {
    constexpr int K = PJ_K, TILE = K * HJ_THREADS;
    __shared__ unsigned long long s_keys[TILE];
    __shared__ unsigned char s_pid[TILE];
    __shared__ unsigned int s_cnt[PJ_MAX_PARTS], s_start[PJ_MAX_PARTS + 1];
    __shared__ unsigned long long s_gbase[PJ_MAX_PARTS];
    const int tid = threadIdx.x, lane = tid & 31, P = 1 << pj.log2p;
    const unsigned long long pol = pj_policy();
    if (tid < 32) s_cnt[tid] = 0;
    __syncthreads();
    for (int tile = blockIdx.x; tile < pj.num_tiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TILE + tid;
        unsigned long long key[K];
        int pid[K];
        unsigned rank[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            key[j] = e < pj.n ? ld_ef(pj.keys + e, pol) : 0ull;
        }
        if (tid < P) s_gbase[tid] = pj.part_base[tid] + pj.tile_off[(size_t)tid * pj.num_tiles + tile];
#pragma unroll
        for (int j = 0; j < K; j++) {
            pid[j] = pj_part(key[j], pj.log2p);
            rank[j] = 0;
            if (e0 + (int64_t)j * HJ_THREADS < pj.n) rank[j] = atomicAdd(&s_cnt[pid[j]], 1u);
        }
        __syncthreads();
        if (tid < 32) { // local starts of the partitions' runs in the staging order
            const unsigned c = tid < P ? s_cnt[tid] : 0u;
            unsigned incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            s_start[tid] = incl - c;
            if (tid == 31) s_start[32] = incl;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            if (e < pj.n) {
                const unsigned local = s_start[pid[j]] + rank[j];
                s_keys[local] = key[j];
                s_pid[local] = (unsigned char)pid[j];
                pj.ppos32[e] = (unsigned int)(s_gbase[pid[j]] + rank[j]);
            }
        }
        __syncthreads();
        const unsigned total = s_start[32];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const unsigned idx = tid + j * HJ_THREADS;
            if (idx < total) {
                const int p = s_pid[idx];
                st_ef(pj.pkeys + s_gbase[p] + (idx - s_start[p]), s_keys[idx], pol);
            }
        }
        if (tid < 32) s_cnt[tid] = 0;
        __syncthreads(); // staging buffers and counters are reused by the next tile
    }
}

// pass 4: probe in partitioned order (partition after partition, so one slot range of the table is hot in L2)
template <int K>
__global__ void __launch_bounds__(HJ_THREADS) pj_probe_kernel(PartJoin pj, JoinTable jt) {
    constexpr int TILE = K * HJ_THREADS;
    const int tid = threadIdx.x;
    const unsigned long long pol = pj_policy();
    const int64_t num_tiles = (pj.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = (int64_t)tile * TILE + tid;
        unsigned long long key[K], brow[K];
        uint64_t slot[K];
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            key[j] = e < pj.n ? ld_ef(pj.pkeys + e, pol) : 0ull;
            if (e < pj.n) inrange |= 1u << j;
        }
        probe_first<K>(jt, key, inrange, brow, slot);
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            if ((inrange >> j) & 1u) st_ef(pj.res0 + e, brow[j], pol);
        }
    }
}

// pass 5: results back into probe-row order.  Row i fetches the result at its position in the partitioned
// order (sequential within each partition's stream), gathers the build-side columns and leaves a match bitmap;
// the joined rows are then compacted by the filter/project kernel (predicate = the match bitmap).
struct GatherParams {
    PartJoin pj;
    int32_t n_gather;
    int32_t row_is_payload;      // the probe results ARE the one gathered column's values (JoinTable::rowpay)
    const unsigned long long *src[HJ_MAX_COLS]; // build column to gather by build row
    unsigned long long *dst[HJ_MAX_COLS];       // gathered column
    unsigned int *match;                        // match bitmap in probe-row order
};
// Tile-aligned variant for the second-generation split: inside a tile's run the positions are in arrival order, so the
// four results that share a 32-byte sector belong to four arbitrary rows of the SAME 2048-row tile.  One CTA therefore
// handles a whole tile and reads the results through L1 (plain read-only loads, no evict_first): the sector is fetched once.
__global__ void __launch_bounds__(256) pj_gather_tile_kernel(const __grid_constant__ GatherParams gp) {
    constexpr int K = PJ_K, TILE = K * 256;
    const int lane = threadIdx.x & 31;
    const unsigned long long pol = pj_policy();
    const int64_t num_tiles = (gp.pj.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        unsigned int pos[K];
        unsigned long long r[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t i = tile * TILE + j * 256 + threadIdx.x;
            pos[j] = i < gp.pj.n ? __ldg(gp.pj.ppos32 + i) : 0u;
        }
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t i = tile * TILE + j * 256 + threadIdx.x;
            r[j] = i < gp.pj.n ? __ldg(gp.pj.res0 + pos[j]) : EMPTY_ROW;
        }
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t i = tile * TILE + j * 256 + threadIdx.x;
            const bool m = r[j] != EMPTY_ROW;
            if (i < gp.pj.n) st_ef(gp.dst[0] + i, m ? r[j] : 0ull, pol);
            const unsigned b = __ballot_sync(0xffffffffu, m);
            if (lane == 0 && (i - lane) < gp.pj.n) gp.match[i >> 5] = b;
        }
    }
}

__global__ void __launch_bounds__(256) pj_gather_kernel(const __grid_constant__ GatherParams gp) {
    const int lane = threadIdx.x & 31;
    const unsigned long long pol = pj_policy();
    const int64_t n32 = (gp.pj.n + 31) / 32 * 32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n32; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < gp.pj.n;
        bool m = false;
        if (live) {
            const unsigned int p = gp.pj.ppos32[i];
            const unsigned long long brow = ld_ef(gp.pj.res0 + p, pol);
            m = brow != EMPTY_ROW;
            if (gp.row_is_payload) st_ef(gp.dst[0] + i, m ? brow : 0ull, pol);
            else for (int c = 0; c < gp.n_gather; c++) {
                unsigned long long v = 0;
                if (m) v = ld_cg_u64(gp.src[c] + brow);
                st_ef(gp.dst[c] + i, v, pol);
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, m);
        if (lane == 0) gp.match[i >> 5] = b;
    }
}

struct ProbeParams {
    JoinTable jt;
    PartJoin pj;            // scratch of the partitioned probe (host side only)
    const unsigned long long *probe_keys;
    int64_t n_probe;
    int32_t n_left, n_right;
    int32_t key_from_probe; // the build key column carries no validity: its output value is the probe key
    int32_t pad;
    ColSrc left[HJ_MAX_COLS], right[HJ_MAX_COLS];
    void *out_values[2 * HJ_MAX_COLS];   // 8-byte values or one byte per row (Boolean)
    uint8_t *out_valid[2 * HJ_MAX_COLS]; // one byte per row or nullptr
    int64_t out_cap;
    unsigned long long *tile_state;
    unsigned int *ticket;
    unsigned long long *out_count;
    int32_t num_tiles;
};

#ifndef NQE_LB_WIDE
#define NQE_LB_WIDE 1
#endif
#include "lookback_body.inc"

// one output cell: column c of the joined row (build row | probe row) -> position pos
__device__ __forceinline__ void emit_value(const ColSrc &c, int64_t src_row, void *out_values, uint8_t *out_valid, int64_t pos) {
    bool valid = true;
    if (c.validity) valid = (__ldg(c.validity + (src_row >> 5)) >> (src_row & 31)) & 1u;
    if (c.dtype == NQE_BOOL) {
        const uint32_t w = __ldg((const uint32_t *)c.values + (src_row >> 5));
        ((uint8_t *)out_values)[pos] = valid ? (uint8_t)((w >> (src_row & 31)) & 1u) : 0;
    } else {
        ((unsigned long long *)out_values)[pos] = valid ? ld_cg_u64((const unsigned long long *)c.values + src_row) : 0ull;
    }
    if (out_valid) out_valid[pos] = (uint8_t)valid;
}

__device__ __forceinline__ void emit_row(const ProbeParams &pp, int64_t brow, int64_t prow, int64_t pos) {
    const int ncols = pp.n_left + pp.n_right;
    for (int c = 0; c < ncols; c++) {
        const bool is_left = c < pp.n_left;
        emit_value(is_left ? pp.left[c] : pp.right[c - pp.n_left], is_left ? brow : prow, pp.out_values[c], pp.out_valid[c], pos);
    }
}

// DIRECT: the table is a direct one (unique keys: no duplicate walks are compiled in); CM (DIRECT only): bit 0 = the
// streams -- probe columns in, joined columns out -- carry an L2 evict_first policy, bit 1 = the table reads carry
// evict_last, so that the 8-byte-per-key table stays in the L2 while ~5 GB of rows flow past it.
template <bool DIRECT, int CM>
__global__ void __launch_bounds__(HJ_THREADS)
join_probe_kernel(const __grid_constant__ ProbeParams pp) {
    constexpr int K = HJ_K, TILE = K * HJ_THREADS;
    __shared__ unsigned long long s_cnt[K * HJ_WARPS];
    __shared__ unsigned long long s_tile_excl;
    __shared__ int s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long ef = (DIRECT && (CM & 1)) ? pj_policy() : 0ull;
    auto ld_in = [&](const unsigned long long *p) { return (DIRECT && (CM & 1)) ? ld_ef(p, ef) : (unsigned long long)ld_stream_u64(p); };
    auto st_out = [&](unsigned long long *p, unsigned long long v) {
        if (DIRECT && (CM & 1)) st_ef(p, v, ef);
        else *p = v;
    };
    while (true) {
        if (tid == 0) s_tile = (int)atomicAdd(pp.ticket, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= pp.num_tiles) break;
        const int64_t e0 = (int64_t)tile * TILE + tid;
        unsigned long long key[K], first[K];
        uint64_t slot[K];
        unsigned int cnt[K];
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            key[j] = e < pp.n_probe ? ld_in(pp.probe_keys + e) : 0ull;
            if (e < pp.n_probe) inrange |= 1u << j;
        }
        if (DIRECT) {
            probe_direct<K, (CM & 2) != 0>(pp.jt, key, inrange, first, slot);
#pragma unroll
            for (int j = 0; j < K; j++) cnt[j] = first[j] != EMPTY_ROW;
        } else if (!pp.jt.has_dups) {
            probe_first<K>(pp.jt, key, inrange, first, slot);
#pragma unroll
            for (int j = 0; j < K; j++) cnt[j] = first[j] != EMPTY_ROW;
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) {
                slot[j] = 0;
                first[j] = EMPTY_ROW;
                cnt[j] = ((inrange >> j) & 1u) ? probe_count(pp.jt, key[j], &first[j], &slot[j]) : 0u;
            }
        }
        // ranks: warp inclusive scan of counts per j, then scan of the K*WARPS warp totals
        unsigned long long excl_in_warp[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            unsigned long long incl = cnt[j];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            excl_in_warp[j] = incl - cnt[j];
            if (lane == 31) s_cnt[j * HJ_WARPS + warp] = incl;
        }
        __syncthreads();
        if (warp == 0) {
            static_assert(K * HJ_WARPS == 32, "one entry per lane");
            const unsigned long long mine = s_cnt[lane];
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s_cnt[lane] = incl - mine;
            const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) nqe_lb_publish(pp.tile_state, tile, total);
            const unsigned long long excl = nqe_lb_walk(pp.tile_state, tile, total, lane);
            if (lane == 0) {
                s_tile_excl = excl;
                if (tile == pp.num_tiles - 1) *pp.out_count = excl + total;
            }
        }
        __syncthreads();
        const unsigned long long tile_excl = s_tile_excl;
        unsigned long long pos[K];
        uint32_t emit = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            pos[j] = tile_excl + s_cnt[j * HJ_WARPS + warp] + excl_in_warp[j];
            if (cnt[j] && (int64_t)pos[j] < pp.out_cap) emit |= 1u << j;
        }
        // ---- first match of every row, column by column (the loads of the K rows are independent)
        for (int c = 0; c < pp.n_left; c++) {
            unsigned long long *out = (unsigned long long *)pp.out_values[c];
            if (c == pp.jt.key_col && pp.key_from_probe) {
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((emit >> j) & 1u) st_out(out + pos[j], key[j]);
            } else if (c == pp.jt.rowpay_col) { // the slot's row word IS this column's value
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((emit >> j) & 1u) st_out(out + pos[j], first[j]);
            } else {
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((emit >> j) & 1u) emit_value(pp.left[c], (int64_t)first[j], pp.out_values[c], pp.out_valid[c], (int64_t)pos[j]);
            }
        }
        for (int c = 0; c < pp.n_right; c++) {
            const ColSrc &src = pp.right[c];
            if (src.dtype != NQE_BOOL && !src.validity) {
                unsigned long long v[K];
#pragma unroll
                for (int j = 0; j < K; j++) {
                    v[j] = 0;
                    if ((emit >> j) & 1u) v[j] = ld_in((const unsigned long long *)src.values + e0 + (int64_t)j * HJ_THREADS);
                }
                unsigned long long *out = (unsigned long long *)pp.out_values[pp.n_left + c];
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((emit >> j) & 1u) st_out(out + pos[j], v[j]);
            } else {
#pragma unroll
                for (int j = 0; j < K; j++)
                    if ((emit >> j) & 1u)
                        emit_value(src, e0 + (int64_t)j * HJ_THREADS, pp.out_values[pp.n_left + c], pp.out_valid[pp.n_left + c], (int64_t)pos[j]);
            }
        }
        // ---- duplicate build keys: the remaining matches of a row, ascending build row
        if (!DIRECT && pp.jt.has_dups) {
#pragma unroll
            for (int j = 0; j < K; j++) {
                unsigned long long brow = first[j];
                uint64_t s = slot[j];
                for (unsigned int m = 1; m < cnt[j]; m++) {
                    brow = probe_next(pp.jt, key[j], brow, &s);
                    const int64_t p = (int64_t)pos[j] + m;
                    if (p < pp.out_cap) emit_row(pp, (int64_t)brow, e0 + (int64_t)j * HJ_THREADS, p);
                }
            }
        }
        __syncthreads();
    }
}

#include "hash_join_direct.cuh" // direct table: probe pass + staged emit pass

// ---- fused join -> group-by aggregate --------------------------------------
struct JoinAggParams {
    JoinTable jt;
    const unsigned long long *probe_keys;
    int64_t n_probe;
    ColSrc group;          // group key column
    int32_t group_left;    // 1: taken from the build row, 0: from the probe row
    int32_t group_in_slot; // the slot's row word IS the group key (JoinTable::rowpay)
    ColSrc val[AG_MAX];
    int32_t val_left[AG_MAX];
};

__device__ __forceinline__ bool col_valid(const ColSrc &c, int64_t r) {
    return !c.validity || ((__ldg(c.validity + (r >> 5)) >> (r & 31)) & 1u);
}

struct JoinRowSource {
    const JoinAggParams &jp;
    int64_t brow, prow;
    __device__ __forceinline__ bool operator()(int id, int *dtype, uint64_t *bits) const {
        const ColSrc &c = jp.val[id];
        *dtype = c.dtype;
        const int64_t r = jp.val_left[id] ? brow : prow;
        if (!col_valid(c, r)) return false;
        *bits = (c.dtype == NQE_BOOL || c.dtype == NQE_UTF8) ? 0ull : ld_cg_u64((const unsigned long long *)c.values + r);
        return true;
    }
};

__device__ __forceinline__ void join_agg_one(const JoinAggParams &jp, const AggParams &ap, int64_t brow, int64_t prow) {
    uint64_t gkey;
    if (jp.group_in_slot) gkey = (uint64_t)brow;
    else {
        const int64_t grow = jp.group_left ? brow : prow;
        if (!col_valid(jp.group, grow)) return; // NULL group keys are dropped (aggregate/mod.rs:63-71)
        gkey = ld_cg_u64((const unsigned long long *)jp.group.values + grow);
    }
    Sector0 s0;
    unsigned long long *r = find_slot(ap, gkey, &s0);
    if (r) update_states(ap, r, s0, JoinRowSource{jp, brow, prow});
}

__global__ void __launch_bounds__(HJ_THREADS)
join_aggregate_kernel(const __grid_constant__ JoinAggParams jp, const __grid_constant__ AggParams ap) {
    constexpr int K = HJ_K, TILE = K * HJ_THREADS;
    const int64_t num_tiles = (jp.n_probe + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE + threadIdx.x;
        unsigned long long key[K], brow[K];
        uint64_t slot[K];
        uint32_t inrange = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int64_t e = e0 + (int64_t)j * HJ_THREADS;
            key[j] = e < jp.n_probe ? ld_stream_u64(jp.probe_keys + e) : 0ull;
            if (e < jp.n_probe) inrange |= 1u << j;
        }
        probe_first<K>(jp.jt, key, inrange, brow, slot);
        // group keys of the (first) matches: K independent loads
        uint64_t gkey[K];
        uint32_t have = 0;
#pragma unroll
        for (int j = 0; j < K; j++) {
            gkey[j] = 0;
            if (brow[j] == EMPTY_ROW) continue;
            if (jp.group_in_slot) {
                gkey[j] = brow[j];
            } else {
                const int64_t grow = jp.group_left ? (int64_t)brow[j] : e0 + (int64_t)j * HJ_THREADS;
                if (!col_valid(jp.group, grow)) continue; // NULL group keys are dropped
                gkey[j] = ld_cg_u64((const unsigned long long *)jp.group.values + grow);
            }
            have |= 1u << j;
        }
        unsigned long long *rec[K];
        Sector0 s0[K];
        find_slots<K>(ap, gkey, have, rec, s0);
#pragma unroll
        for (int j = 0; j < K; j++)
            if (rec[j]) update_states(ap, rec[j], s0[j], JoinRowSource{jp, (int64_t)brow[j], e0 + (int64_t)j * HJ_THREADS});
        if (jp.jt.has_dups) {
            // duplicate build keys: walk on from the first match for the remaining ones
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (brow[j] == EMPTY_ROW) continue;
                uint64_t s = slot[j] + 1 == jp.jt.cap ? 0 : slot[j] + 1;
                while (true) {
                    const Slot sl = ld_slot(jp.jt, s);
                    if (sl.row == EMPTY_ROW) break;
                    if (sl.key == key[j]) join_agg_one(jp, ap, (int64_t)sl.row, e0 + (int64_t)j * HJ_THREADS);
                    s = s + 1 == jp.jt.cap ? 0 : s + 1;
                }
            }
        }
    }
}

#include "hash_join_paged.cuh" // fused join -> group-by over paged streams

// smallest probe side that takes the partitioned / paged paths (knob NQE_JOIN_PART_MIN_ROWS; tests and the
// compute-sanitizer runs lower it so that those kernels run on small inputs)
int64_t join_part_min_rows() {
    static int64_t v = -1;
    if (v < 0) {
        const char *e = getenv("NQE_JOIN_PART_MIN_ROWS");
        v = e ? atoll(e) : ((int64_t)1 << 22);
    }
    return v;
}

int32_t check_join_keys(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right, int32_t lk, int32_t rk) {
    if (lk < 0 || lk >= (int)left->cols.size() || rk < 0 || rk >= (int)right->cols.size())
        return nqe_fail(ctx, NQE_ERR_LOGICAL, "ColumnExpr must has name or idx"); // key column not found (column.rs:53-55)
    const int ld = left->cols[lk].dtype, rd = right->cols[rk].dtype;
    // hash_join.rs:139-162 / :185-226
    if (ld == NQE_UTF8 && rd == NQE_UTF8) return NQE_OK; // dictionary ids, utf8.cu
    if (ld == NQE_UTF8 || rd == NQE_UTF8) return nqe_fail(ctx, NQE_ERR_PANIC, "join key dtypes differ (downcast_ref unwrap on None)");
    if ((ld != NQE_INT64 && ld != NQE_UINT64) || (rd != NQE_INT64 && rd != NQE_UINT64))
        return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "join key dtype must be Int64, UInt64 or Utf8");
    if (ld != rd) return nqe_fail(ctx, NQE_ERR_PANIC, "join key dtypes differ (downcast_ref unwrap on None)");
    return NQE_OK;
}

// build the multimap over left->cols[lk]
// rowpay_col >= 0: store that column's value in the slots' row word (see JoinTable); falls back to row numbers when a
// value collides with the free-slot marker or (dups_ok == false) the build keys are not unique
int32_t build_table(nqe_ctx *ctx, const nqe_table *left, int32_t lk, JoinTable *jt, int rowpay_col = -1, bool dups_ok = true) {
    const int64_t nl = left->nrows;
    memset(jt, 0, sizeof *jt);
    jt->key_col = lk;
    jt->rowpay_col = -1;
    if (rowpay_col >= 0) {
        const DevColumn &pc = left->cols[rowpay_col];
        if (!pc.validity && (pc.dtype == NQE_INT64 || pc.dtype == NQE_UINT64 || pc.dtype == NQE_FLOAT64)) {
            jt->rowpay = (const unsigned long long *)pc.values;
            jt->rowpay_col = rowpay_col;
        }
    }
    const uint64_t cap = ((uint64_t)((double)nl / 0.5) + 16) & ~(uint64_t)1; // load factor 0.5; even: probes read slot pairs
    void *slots = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &slots, cap * 16));
    jt->words = (unsigned long long *)slots;
    jt->cap = cap;
    uint32_t *status = (uint32_t *)(ctx->d_scratch + 1);
    uint32_t *dupflag = (uint32_t *)(ctx->d_scratch + 3);
    join_clear_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, ctx->stream>>>(*jt, cap);
    ctx->launches++;
    if (nl > 0) {
        const unsigned long long *keys = (const unsigned long long *)left->cols[lk].values;
        join_build_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, ctx->stream>>>(*jt, keys, nl, status, dupflag);
        ctx->launches++;
    }
    cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
        return nqe_fail(ctx, NQE_ERR_CUDA, "join build failed: %s", cudaGetErrorString(cudaGetLastError()));
    if ((uint32_t)ctx->h_scratch[1] & DEV_ERR_TABLE_FULL) return nqe_fail(ctx, NQE_ERR_CUDA, "join table overflow");
    jt->has_dups = (uint32_t)ctx->h_scratch[3] ? 1 : 0;
    if (jt->rowpay && ((uint32_t)ctx->h_scratch[4] || (jt->has_dups && !dups_ok))) {
        nqe_dev_free(ctx, jt->words);
        cudaMemsetAsync(ctx->d_scratch + 1, 0, 4 * sizeof(uint64_t), ctx->stream); // status, ticket, dup flag, payload flag
        return build_table(ctx, left, lk, jt, -1, true);
    }
    return NQE_OK;
}

// Direct table over dense unique build keys (JoinTable::direct).  *done = false (and nothing allocated) when the keys are
// not dense enough, not unique, or a payload collides with the free-slot marker: the caller then builds the hashed table.
// Knobs: NQE_JOIN_DIRECT=0 switches the direct table off, NQE_JOIN_DIRECT_MIN_ROWS (default 16384) is the smallest build
// side worth the extra min/max pass and its synchronisation.
int32_t build_table_direct(nqe_ctx *ctx, const nqe_table *left, int32_t lk, JoinTable *jt, int rowpay_col, bool *done,
                           const long long *pay_minmax = nullptr) {
    *done = false;
    // (knob blocks with several variables publish their guard LAST, with release / acquire order: the multi-GPU entry
    // points call the operators from one host thread per member)
    static std::atomic<int> allow{-1};
    static int allow_narrow = 1;
    static int64_t min_rows = 0;
    if (allow.load(std::memory_order_acquire) < 0) {
        const char *e = getenv("NQE_JOIN_DIRECT_MIN_ROWS");
        min_rows = e ? atoll(e) : 16384;
        e = getenv("NQE_JOIN_DIRECT_NARROW"); // 0: always 8-byte slots
        allow_narrow = e ? atoi(e) : 1;
        e = getenv("NQE_JOIN_DIRECT");
        allow.store(e ? atoi(e) != 0 : 1, std::memory_order_release);
    }
    const int64_t nl = left->nrows;
    const DevColumn &kc = left->cols[lk];
    if (!allow || nl < min_rows || nl < 1 || (kc.dtype != NQE_INT64 && kc.dtype != NQE_UINT64)) return NQE_OK;
    // key range and, when a payload may ride in the table, its range (for 4-byte slots): one pass each, one synchronisation
    const DevColumn *pcand = rowpay_col >= 0 ? &left->cols[rowpay_col] : nullptr;
    const bool pay_ok = pcand && !pcand->validity && (pcand->dtype == NQE_INT64 || pcand->dtype == NQE_UINT64 || pcand->dtype == NQE_FLOAT64);
    const bool want_pay_range = pay_ok && allow_narrow && !pay_minmax && pcand->dtype != NQE_FLOAT64;
    long long mm[4];
    NQE_TRY(nqe_minmax2_i64(ctx, (const unsigned long long *)kc.values, want_pay_range ? (const unsigned long long *)pcand->values : nullptr, nl, mm));
    const long long lo = mm[0], hi = mm[1];
    // as two's-complement offsets from the signed minimum the keys of either dtype fall in [0, range)
    const unsigned long long range = (unsigned long long)hi - (unsigned long long)lo + 1ull;
    if (lo > hi || range == 0 || range > 4ull * (unsigned long long)nl + 1024ull) return NQE_OK; // sparse keys: <= 32 bytes per build row
    memset(jt, 0, sizeof *jt);
    jt->key_col = lk;
    jt->rowpay_col = -1;
    jt->direct = 1;
    jt->lo = lo;
    jt->cap = range;
    if (rowpay_col >= 0) {
        const DevColumn &pc = left->cols[rowpay_col];
        if (!pc.validity && (pc.dtype == NQE_INT64 || pc.dtype == NQE_UINT64 || pc.dtype == NQE_FLOAT64)) {
            jt->rowpay = (const unsigned long long *)pc.values;
            jt->rowpay_col = rowpay_col;
        }
    }
    if (allow_narrow) {
        if (!jt->rowpay) {
            jt->narrow = nl < (int64_t)0xffffffffu;
        } else if (left->cols[rowpay_col].dtype != NQE_FLOAT64) {
            const long long plo = pay_minmax ? pay_minmax[0] : mm[2], phi = pay_minmax ? pay_minmax[1] : mm[3];
            if (plo <= phi && (unsigned long long)phi - (unsigned long long)plo < 0xffffffffull) {
                jt->narrow = 1;
                jt->pay_lo = plo;
            }
        }
    }
    const size_t slot_bytes = jt->narrow ? 4 : 8;
    void *slots = nullptr;
    NQE_TRY(nqe_dev_alloc(ctx, &slots, range * slot_bytes + 8));
    jt->words = (unsigned long long *)slots;
    uint32_t *dupflag = (uint32_t *)(ctx->d_scratch + 3);
    cudaMemsetAsync(slots, 0xff, range * slot_bytes, ctx->stream); // free-slot marker everywhere
    join_direct_build_kernel<<<(unsigned)((nl + 255) / 256), 256, 0, ctx->stream>>>(*jt, (const unsigned long long *)kc.values, nl, dupflag);
    ctx->launches++;
    cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 5 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        nqe_dev_free(ctx, slots);
        return nqe_fail(ctx, NQE_ERR_CUDA, "join build failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if ((uint32_t)ctx->h_scratch[3] || (uint32_t)ctx->h_scratch[4]) { // duplicate keys / payload == free-slot marker
        nqe_dev_free(ctx, slots);
        memset(jt, 0, sizeof *jt);
        cudaMemsetAsync(ctx->d_scratch + 1, 0, 4 * sizeof(uint64_t), ctx->stream);
        return NQE_OK;
    }
    *done = true;
    return NQE_OK;
}

void fill_src(ColSrc *s, const DevColumn &c) {
    s->values = c.values;
    s->validity = (const uint32_t *)c.validity;
    s->dtype = c.dtype;
    s->pad = 0;
}

} // namespace

extern "C" int32_t nqe_hash_join(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right, int32_t left_key,
                                 int32_t right_key, nqe_table **out) {
    if (!ctx || !left || !right || !out) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    NQE_TRY(check_join_keys(ctx, left, right, left_key, right_key));
    const int nl = (int)left->cols.size(), nr = (int)right->cols.size();
    if (nl > HJ_MAX_COLS || nr > HJ_MAX_COLS) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "more than %d columns on one join side", HJ_MAX_COLS);
    for (auto *t : {left, right})
        for (auto &c : t->cols)
            if (c.dtype == NQE_UTF8) return nqe_hash_join_strings(ctx, left, right, left_key, right_key, out); // utf8.cu

    OpTimer timer(ctx);
    NQE_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream));
    ProbeParams pp;
    memset(&pp, 0, sizeof pp);
    static std::atomic<int> allow_part{-1};
    static int allow_rowpay = 1;
    static size_t l2_budget = 0, part_min = 0;
    if (allow_part.load(std::memory_order_acquire) < 0) {
        const char *e = getenv("NQE_JOIN_PART_MB"); // slot range of one partition, MiB
        l2_budget = (size_t)(e ? atoi(e) : 48) << 20; // measured 12 / 24 / 48 / 96 MiB: 3.98 / 3.73 / 3.63 / 3.80 ms (1e8 x 1e7)
        e = getenv("NQE_JOIN_PART_MIN_MB"); // tables up to this size are probed directly
        part_min = (size_t)(e ? atoi(e) : 48) << 20;
        e = getenv("NQE_JOIN_ROWPAY");
        allow_rowpay = e ? atoi(e) : 1;
        e = getenv("NQE_JOIN_PART");
        allow_part.store(e ? atoi(e) != 0 : 1, std::memory_order_release);
    }
    pp.n_probe = right->nrows;
    // the partitioned probe (below) will run if the build keys turn out unique; with a single non-key build column its
    // values ride in the slots' row word and the gather pass has nothing left to chase
    const bool part_sizes = allow_part && pp.n_probe >= join_part_min_rows() && pp.n_probe < (int64_t)1 << 32 &&
                            ((size_t)((double)left->nrows / 0.5) + 16) * 16 > part_min && nl + nr + 1 <= 16;
    const int rowpay_col = allow_rowpay && part_sizes && nl == 2 && !left->cols[left_key].validity ? 1 - left_key : -1;
    pp.probe_keys = (const unsigned long long *)right->cols[right_key].values;
    // The split of the probe keys depends only on sizes, not on the table: it runs on the auxiliary stream WHILE the
    // table is built (the build is bound by its random 128-bit CAS traffic, the split by its scattered stores; measured
    // alone: 0.49 ms and 0.91 ms).  If the build then finds duplicate keys the split is simply not used.
    std::vector<void *> pj_bufs;
    int32_t rc = NQE_OK;
    bool split_started = false;
    static int allow_overlap = -1;
    if (allow_overlap < 0) {
        const char *e = getenv("NQE_JOIN_OVERLAP");
        allow_overlap = e ? atoi(e) : 1;
    }
    // dense unique build keys: a direct table, probed in place by the general kernel below (no split, no gather pass);
    // with a single non-key build column its values ride in the table
    bool direct = false;
    {
        const DevColumn *oc = nl == 2 ? &left->cols[1 - left_key] : nullptr;
        const bool pay = allow_rowpay && oc && !oc->validity && (oc->dtype == NQE_INT64 || oc->dtype == NQE_UINT64 || oc->dtype == NQE_FLOAT64);
        rc = build_table_direct(ctx, left, left_key, &pp.jt, pay ? 1 - left_key : -1, &direct);
    }
    if (rc == NQE_OK && part_sizes && !direct) {
        const size_t table_bytes = ((size_t)((double)left->nrows / 0.5) + 16) * 16;
        int log2p = 1;
        while (log2p < 5 && (table_bytes >> log2p) > l2_budget) log2p++;
        PartJoin &pj = pp.pj;
        pj.keys = pp.probe_keys;
        pj.n = pp.n_probe;
        pj.log2p = log2p;
        pj.num_tiles = (int32_t)((pp.n_probe + 2048 - 1) / 2048); // PJ_K * HJ_THREADS rows per tile
        const size_t P = (size_t)1 << log2p, nt = (size_t)pj.num_tiles;
        auto alloc = [&](void **p, size_t bytes) {
            if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, p, bytes);
            if (rc == NQE_OK) pj_bufs.push_back(*p);
        };
        void *totals = nullptr;
        alloc((void **)&pj.tile_cnt, nt * P * 4);
        alloc((void **)&pj.tile_off, nt * P * 4);
        alloc((void **)&pj.part_base, (P + 1) * 8);
        alloc(&totals, P * 8);
        alloc((void **)&pj.pkeys, (size_t)pp.n_probe * 8);
        alloc((void **)&pj.ppos32, (size_t)pp.n_probe * 4);
        alloc((void **)&pj.res0, (size_t)pp.n_probe * 8);
        if (rc == NQE_OK && !ctx->s_aux) {
            if (cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess)
                rc = nqe_fail(ctx, NQE_ERR_CUDA, "auxiliary stream creation failed");
        }
        if (rc == NQE_OK) {
            cudaStream_t ss = allow_overlap ? ctx->s_aux : ctx->stream;
            if (allow_overlap) {
                cudaEventRecord(ctx->ev_fork, ctx->stream); // the buffers above were allocated in ctx->stream's order
                cudaStreamWaitEvent(ss, ctx->ev_fork, 0);
            }
            static int split_ctas = 0; // knob NQE_JOIN_SPLIT_CTAS: CTAs per SM of the split passes (fewer leave room for the build)
            if (!split_ctas) {
                const char *e = getenv("NQE_JOIN_SPLIT_CTAS");
                split_ctas = e && atoi(e) > 0 ? atoi(e) : 8;
            }
            int grid = ctx->sm_count * split_ctas;
            if (grid > pj.num_tiles) grid = pj.num_tiles;
            static int split_gen = 0; // knob NQE_JOIN_SPLIT: 2 = atomic-rank staged split, 3 = stable scatter
            if (!split_gen) {
                const char *e = getenv("NQE_JOIN_SPLIT");
                split_gen = e ? atoi(e) : 0;
                if (split_gen != 2 && split_gen != 3) split_gen = -1; // no knob: chosen per call below
            }
            // default: the staged split when the tile-aligned gather follows (payload-in-row tables), else the stable one
            const int split_mode = split_gen > 0 ? split_gen : rowpay_col >= 0 ? 2 : 3;
            // 2: atomic histograms + staged (arrival-order) scatter; 3: atomic histograms for the count, stable scatter
            // (the stable order keeps the streaming gather's reads sequential)
            cudaMemsetAsync(totals, 0, P * 8, ss);
            pj2_count_kernel<<<grid, HJ_THREADS, 0, ss>>>(pj, (unsigned long long *)totals);
            pj_scan_kernel<<<(unsigned)P, 1024, 0, ss>>>(pj, (const unsigned long long *)totals);
            if (split_mode == 2) pj2_scatter_kernel<<<grid, HJ_THREADS, 0, ss>>>(pj);
            else pj_scatter_kernel<<<grid, HJ_THREADS, 0, ss>>>(pj);
            ctx->launches += 3;
            if (allow_overlap) cudaEventRecord(ctx->ev_join, ss);
            if (cudaGetLastError() != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "probe-side split launch failed");
            split_started = true;
        }
    }
    if (rc == NQE_OK && !direct) rc = build_table(ctx, left, left_key, &pp.jt, rowpay_col, false);
    if (split_started && allow_overlap) cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0); // also orders the frees below
    pp.n_left = nl;
    pp.n_right = nr;
    pp.key_from_probe = left->cols[left_key].validity == nullptr;
    for (int c = 0; c < nl; c++) fill_src(&pp.left[c], left->cols[c]);
    for (int c = 0; c < nr; c++) fill_src(&pp.right[c], right->cols[c]);
    constexpr int TILE = HJ_K * HJ_THREADS;
    pp.num_tiles = (int32_t)((pp.n_probe + TILE - 1) / TILE);
    void *lb = nullptr;
    if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, &lb, (size_t)(pp.num_tiles + 1) * 8);
    pp.tile_state = (unsigned long long *)lb;
    pp.out_count = (unsigned long long *)ctx->d_scratch;
    pp.ticket = (unsigned int *)(ctx->d_scratch + 2);

    // The build key's output VALUES are the probe keys (they matched), but its validity is the build column's own:
    // key validity is ignored by the join (hash_join.rs:67,86) and `take(left_key, outer_pos)` returns the build
    // side's slots.  The view's build-key column aliases the probe key values WITHOUT the probe key's bitmap; when the
    // probe key is NULL-free both key outputs read the one staged probe-key column instead.
    const int key_src = right->cols[right_key].validity ? left_key : nl + right_key;
    // (A direct-table join is a filter/project over the probe side with the build column gathered through the probe key;
    // running it in the filter/project kernel was tried -- key-indexed gathered columns -- and is much slower, 5.8 vs
    // 2.5 ms at 1e8 x 1e7: that kernel's 16 worker warps per SM keep too few random reads in flight.)
    // ---- partitioned probe: unique build keys and a table that does not fit in the L2
    bool part = false;
    if (rc == NQE_OK && split_started) {
        bool emit_fp = true; // emit through the filter/project kernel: NULL-free 8-byte build columns only
        for (int c = 0; c < nl; c++)
            if (left->cols[c].validity || left->cols[c].dtype == NQE_BOOL) emit_fp = false;
        if (!pp.jt.has_dups && emit_fp) {
            PartJoin &pj = pp.pj;
            auto alloc = [&](void **p, size_t bytes) {
                if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, p, bytes);
                if (rc == NQE_OK) pj_bufs.push_back(*p);
            };
            if (rc == NQE_OK) {
                int grid = ctx->sm_count * 8;
                if (grid > pj.num_tiles) grid = pj.num_tiles;
                static int probe_k = 0; // knob NQE_JOIN_PROBE_K: independent probes in flight per thread
                if (!probe_k) {
                    const char *e = getenv("NQE_JOIN_PROBE_K");
                    probe_k = e && atoi(e) == 8 ? 8 : 4;
                }
                if (probe_k == 8) pj_probe_kernel<8><<<grid, HJ_THREADS, 0, ctx->stream>>>(pj, pp.jt);
                else pj_probe_kernel<4><<<grid, HJ_THREADS, 0, ctx->stream>>>(pj, pp.jt);
                ctx->launches++;
                if (cudaGetLastError() != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "partitioned probe launch failed");
                part = true;
            }
        }
    }
    if (rc == NQE_OK && part && pp.jt.rowpay && nl + nr + 2 <= 16) {
        // ---- payload-in-row table: the probe results ARE the build column's values, in partitioned order.  The
        // filter/project kernel reads them through the rows' positions (a gathered column, DevColumn::via) while it
        // compacts the joined rows: no separate gather pass, no intermediate column, no match bitmap.
        static int allow_fuse = -1;
        if (allow_fuse < 0) {
            const char *e = getenv("NQE_JOIN_FUSE");
            allow_fuse = e ? atoi(e) : 0; // measured: 3.80 ms fused vs 3.75 ms with the gather pass (1e8 x 1e7) -- the fused kernel gathers twice
        }
        if (allow_fuse) {
            const int64_t n = pp.n_probe;
            const int pos_col = nl + nr, match_col = nl + nr + 1;
            nqe_table view;
            view.ctx = ctx;
            view.nrows = n;
            view.cols.resize(nl + nr + 2);
            for (int c = 0; c < nl; c++) {
                DevColumn &d = view.cols[c];
                d = left->cols[c];
                d.length = n;
                d.owned = false;
                if (c == left_key) d.values = right->cols[right_key].values;
                else { d.values = pp.pj.res0; d.via = pos_col; }
            }
            for (int c = 0; c < nr; c++) {
                view.cols[nl + c] = right->cols[c];
                view.cols[nl + c].owned = false;
            }
            DevColumn &pc = view.cols[pos_col];
            pc.dtype = NQE_POS32;
            pc.length = n;
            pc.values = pp.pj.ppos32;
            pc.owned = false;
            DevColumn &mc = view.cols[match_col]; // the same results once more, typed UInt64, for the "found a match" test
            mc.dtype = NQE_UINT64;
            mc.length = n;
            mc.values = pp.pj.res0;
            mc.via = pos_col;
            mc.owned = false;
            std::vector<nqe_expr_node> nodes(nl + nr);
            std::vector<nqe_expr> projs(nl + nr);
            for (int c = 0; c < nl + nr; c++) { // the build key's values are the probe keys: one staged column serves both outputs
                nodes[c] = nqe_expr_node{NQE_NODE_COLUMN, 0, c == left_key ? key_src : c, 0, 0, 0, {0}};
                projs[c] = nqe_expr{&nodes[c], 1, 0};
            }
            nqe_expr_node pn[3] = {{NQE_NODE_COLUMN, 0, match_col, 0, 0, 0, {0}},
                                   {NQE_NODE_LITERAL, 0, 0, NQE_UINT64, 0, 0, {0}},
                                   {NQE_NODE_BINARY, NQE_OP_NOT_EQ, 0, 0, 0, 0, {0}}};
            pn[1].value.u64 = EMPTY_ROW;
            const nqe_expr pred{pn, 3, 0};
            nqe_table *joined = nullptr;
            const int32_t frc = nqe_filter_project(ctx, &view, &pred, projs.data(), nl + nr, &joined);
            if (frc != NQE_ERR_NOT_IMPLEMENTED) { // NOT_IMPLEMENTED: no specialised kernel (NVRTC missing) -> gather pass below
                if (frc == NQE_OK) *out = joined;
                timer.stop();
                nqe_dev_free(ctx, lb);
                nqe_dev_free(ctx, pp.jt.words);
                for (void *p : pj_bufs) nqe_dev_free(ctx, p);
                return frc;
            }
        }
    }
    if (rc == NQE_OK && part) {
        // ---- results back into probe-row order, then one fused compaction of the joined rows
        GatherParams gp;
        memset(&gp, 0, sizeof gp);
        gp.pj = pp.pj;
        gp.row_is_payload = pp.jt.rowpay ? 1 : 0;
        const int64_t n = pp.n_probe;
        std::vector<void *> gathered(nl, nullptr); // per build column: its values in probe-row order
        auto alloc = [&](void **p, size_t bytes) {
            if (rc == NQE_OK) rc = nqe_dev_alloc(ctx, p, bytes);
            if (rc == NQE_OK) pj_bufs.push_back(*p);
        };
        alloc((void **)&gp.match, nqe_bitmap_bytes(n));
        {
            for (int c = 0; c < nl; c++) {
                if (c == left_key) continue;
                alloc(&gathered[c], (size_t)n * 8);
                gp.src[gp.n_gather] = (const unsigned long long *)left->cols[c].values;
                gp.dst[gp.n_gather] = (unsigned long long *)gathered[c];
                gp.n_gather++;
            }
        }
        if (rc == NQE_OK) {
            static int gather_gen = 0; // knob NQE_JOIN_GATHER: 2 (default) = tile-aligned gather through L1 for payload-in-row tables, 1 = streaming gather
            if (!gather_gen) {
                const char *e = getenv("NQE_JOIN_GATHER");
                gather_gen = e && atoi(e) == 1 ? 1 : 2;
            }
            if (gather_gen == 2 && gp.row_is_payload) pj_gather_tile_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(gp);
            else pj_gather_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(gp);
            ctx->launches++;
            // joined table before compaction: [build columns (key = probe key) | probe columns | match]
            nqe_table view;
            view.ctx = ctx;
            view.nrows = n;
            view.cols.resize(nl + nr + 1);
            for (int c = 0; c < nl; c++) {
                DevColumn &d = view.cols[c];
                d = left->cols[c];
                d.length = n;
                d.owned = false;
                d.values = c == left_key ? right->cols[right_key].values : gathered[c];
            }
            for (int c = 0; c < nr; c++) {
                view.cols[nl + c] = right->cols[c];
                view.cols[nl + c].owned = false;
            }
            DevColumn &mc = view.cols[nl + nr];
            mc.dtype = NQE_BOOL;
            mc.length = n;
            mc.values = gp.match;
            mc.owned = false;
            std::vector<nqe_expr_node> nodes(nl + nr + 1);
            std::vector<nqe_expr> projs(nl + nr);
            // the build key's output values are the probe keys: both outputs read the one staged probe-key column
            for (int c = 0; c <= nl + nr; c++) nodes[c] = nqe_expr_node{NQE_NODE_COLUMN, 0, c == left_key ? key_src : c, 0, 0, 0, {0}};
            for (int c = 0; c < nl + nr; c++) projs[c] = nqe_expr{&nodes[c], 1, 0};
            const nqe_expr pred{&nodes[nl + nr], 1, 0};
            nqe_table *joined = nullptr;
            rc = nqe_filter_project(ctx, &view, &pred, projs.data(), nl + nr, &joined);
            if (rc == NQE_OK) *out = joined;
        }
        timer.stop();
        nqe_dev_free(ctx, lb);
        nqe_dev_free(ctx, pp.jt.words);
        for (void *p : pj_bufs) nqe_dev_free(ctx, p);
        return rc;
    }

    if (rc == NQE_OK && pp.jt.rowpay && !pp.jt.direct) rc = nqe_fail(ctx, NQE_ERR_CUDA, "internal: payload-in-row table outside the partitioned probe");
    nqe_table *t = nullptr;
    int64_t cap = pp.n_probe > 0 ? pp.n_probe : 1;
    int64_t out_rows = 0;
    std::vector<uint8_t *> bool_bytes, valid_bytes;
    for (int attempt = 0; rc == NQE_OK && attempt < 2; attempt++) {
        nqe_table_new(ctx, 0, &t);
        t->cols.resize(nl + nr);
        bool_bytes.assign(nl + nr, nullptr);
        valid_bytes.assign(nl + nr, nullptr);
        pp.out_cap = cap;
        for (int c = 0; c < nl + nr && rc == NQE_OK; c++) {
            const DevColumn &src = c < nl ? left->cols[c] : right->cols[c - nl];
            rc = nqe_column_alloc(ctx, src.dtype, cap, src.validity != nullptr, &t->cols[c]);
            if (rc != NQE_OK) break;
            if (src.dtype == NQE_BOOL) {
                rc = nqe_dev_alloc(ctx, (void **)&bool_bytes[c], (size_t)cap + 64);
                pp.out_values[c] = bool_bytes[c];
            } else {
                pp.out_values[c] = t->cols[c].values;
            }
            pp.out_valid[c] = nullptr;
            if (rc == NQE_OK && src.validity) {
                rc = nqe_dev_alloc(ctx, (void **)&valid_bytes[c], (size_t)cap + 64);
                pp.out_valid[c] = valid_bytes[c];
            }
        }
        if (rc == NQE_OK) {
            cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
            cudaMemsetAsync(lb, 0, (size_t)(pp.num_tiles + 1) * 8, ctx->stream);
            if (pp.num_tiles > 0) {
                // direct table and nothing but plain 8-byte columns: the staged kernel (knob NQE_JOIN_DIRECT_STAGED=0: the general one)
                static int allow_staged = -1;
                if (allow_staged < 0) {
                    const char *e = getenv("NQE_JOIN_DIRECT_STAGED");
                    allow_staged = e ? atoi(e) : 1;
                }
                // the key and at most the one column that rides in the table on the build side, up to DE_MAX columns on the probe side
                bool staged = allow_staged && pp.jt.direct && (nl == 1 || (nl == 2 && pp.jt.rowpay)) && nr <= DE_MAX;
                for (int c = 0; c < nl + nr && staged; c++) {
                    const DevColumn &src = c < nl ? left->cols[c] : right->cols[c - nl];
                    if (src.validity || (src.dtype != NQE_INT64 && src.dtype != NQE_UINT64 && src.dtype != NQE_FLOAT64)) staged = false;
                    if (c >= nl && ((uintptr_t)src.values & 15)) staged = false; // bulk copies read 16-byte aligned columns
                }
                if (staged) {
                    DirectEmitParams de;
                    memset(&de, 0, sizeof de);
                    de.jt = pp.jt;
                    de.n_probe = pp.n_probe;
                    de.nl = nl;
                    de.nr = nr;
                    de.left_key = left_key;
                    de.right_key = right_key;
                    for (int c = 0; c < nr; c++) de.right[c] = (const unsigned long long *)right->cols[c].values;
                    for (int c = 0; c < nl + nr; c++) de.out[c] = (unsigned long long *)pp.out_values[c];
                    de.out_count = pp.out_count;
                    de.num_tiles = pp.num_tiles;
                    de.chunk_count = pp.tile_state; // num_tiles + 1 words: room for one count per chunk
                    void *rw = nullptr;
                    rc = nqe_dev_alloc(ctx, &rw, (size_t)pp.n_probe * 8 + 16);
                    de.rowwords = (unsigned long long *)rw;
                    if (rc != NQE_OK) break;
                    switch (nl * 8 + nr) {
                    case 8 + 1: rc = join_direct_emit_launch<1, 1>(ctx, de); break;
                    case 8 + 2: rc = join_direct_emit_launch<1, 2>(ctx, de); break;
                    case 8 + 3: rc = join_direct_emit_launch<1, 3>(ctx, de); break;
                    case 8 + 4: rc = join_direct_emit_launch<1, 4>(ctx, de); break;
                    case 16 + 1: rc = join_direct_emit_launch<2, 1>(ctx, de); break;
                    case 16 + 2: rc = join_direct_emit_launch<2, 2>(ctx, de); break;
                    case 16 + 3: rc = join_direct_emit_launch<2, 3>(ctx, de); break;
                    default: rc = join_direct_emit_launch<2, 4>(ctx, de); break;
                    }
                    nqe_dev_free(ctx, rw); // stream-ordered: after the two passes
                } else {
                // over a direct table: table reads evict_last, plain streams (measured best of the four policy combinations)
                auto kern = pp.jt.direct ? join_probe_kernel<true, 2> : join_probe_kernel<false, 0>;
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, HJ_THREADS, 0);
                int grid = ctx->sm_count * (occ > 0 ? occ : 1);
                if (grid > pp.num_tiles) grid = pp.num_tiles;
                kern<<<grid, HJ_THREADS, 0, ctx->stream>>>(pp);
                ctx->launches++;
                }
            }
            cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                rc = nqe_fail(ctx, NQE_ERR_CUDA, "join probe failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        if (rc != NQE_OK) break;
        out_rows = (int64_t)ctx->h_scratch[0];
        if (out_rows <= cap) break;
        // output larger than the first guess (duplicate build keys): exact size is now known
        for (auto p : bool_bytes) nqe_dev_free(ctx, p);
        for (auto p : valid_bytes) nqe_dev_free(ctx, p);
        nqe_table_free(t);
        t = nullptr;
        cap = out_rows;
    }
    if (rc == NQE_OK) {
        bool any = false;
        cudaMemsetAsync(ctx->d_scratch + 8, 0, 32 * sizeof(uint64_t), ctx->stream);
        for (int c = 0; c < nl + nr && rc == NQE_OK; c++) {
            if (bool_bytes[c]) { rc = nqe_pack_bytes(ctx, bool_bytes[c], out_rows, (uint32_t *)t->cols[c].values, nullptr); any = true; }
            if (rc == NQE_OK && valid_bytes[c]) {
                rc = nqe_pack_bytes(ctx, valid_bytes[c], out_rows, (uint32_t *)t->cols[c].validity,
                                    (unsigned long long *)(ctx->d_scratch + 8 + c));
                any = true;
            }
        }
        if (rc == NQE_OK && any) {
            cudaMemcpyAsync(ctx->h_scratch + 8, ctx->d_scratch + 8, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = nqe_fail(ctx, NQE_ERR_CUDA, "pack kernel failed");
        }
    }
    timer.stop();
    for (auto p : bool_bytes) nqe_dev_free(ctx, p);
    for (auto p : valid_bytes) nqe_dev_free(ctx, p);
    nqe_dev_free(ctx, lb);
    nqe_dev_free(ctx, pp.jt.words);
    for (void *p : pj_bufs) nqe_dev_free(ctx, p);
    if (rc != NQE_OK) {
        if (t) nqe_table_free(t);
        return rc;
    }
    t->nrows = out_rows;
    for (int c = 0; c < nl + nr; c++) {
        DevColumn &col = t->cols[c];
        col.length = out_rows;
        if (col.validity) {
            col.null_count = (int64_t)ctx->h_scratch[8 + c];
            if (col.null_count == 0) { // arrow `take`: no nulls selected => no bitmap
                nqe_dev_free(ctx, col.validity);
                col.validity = nullptr;
            }
        }
    }
    *out = t;
    return NQE_OK;
}

extern "C" int32_t nqe_join_aggregate(nqe_ctx *ctx, const nqe_table *left, const nqe_table *right, int32_t left_key,
                                      int32_t right_key, int32_t group_column, const nqe_agg *aggs, int32_t n_aggs,
                                      nqe_table **out) {
    if (!ctx || !left || !right || !out || (!aggs && n_aggs > 0)) return NQE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    *out = nullptr;
    NQE_TRY(check_join_keys(ctx, left, right, left_key, right_key));
    if (left->cols[left_key].dtype == NQE_UTF8)
        return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "fused join+aggregate over Utf8 join keys: run nqe_hash_join, then nqe_hash_aggregate");
    const int nl = (int)left->cols.size(), nr = (int)right->cols.size();
    auto col_at = [&](int c) -> const DevColumn * {
        if (c < 0 || c >= nl + nr) return nullptr;
        return c < nl ? &left->cols[c] : &right->cols[c - nl];
    };
    const DevColumn *g = col_at(group_column);
    if (!g) return nqe_fail(ctx, NQE_ERR_PANIC, "group column index %d out of range", group_column);
    if (g->dtype != NQE_INT64 && g->dtype != NQE_UINT64) {
        if (g->dtype == NQE_UTF8) return nqe_fail(ctx, NQE_ERR_NOT_IMPLEMENTED, "Utf8 group keys are not implemented on the CUDA path yet");
        return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "group by only support by `Int64`, `UInt64`, `String`");
    }
    if (n_aggs > AG_MAX) return nqe_fail(ctx, NQE_ERR_NOT_SUPPORTED, "at most %d aggregates per plan", AG_MAX);
    JoinAggParams jp;
    memset(&jp, 0, sizeof jp);
    AggParams ap;
    memset(&ap, 0, sizeof ap);
    int32_t dts[AG_MAX], src_ids[AG_MAX];
    for (int a = 0; a < n_aggs; a++) {
        if (aggs[a].op == NQE_AGG_GROUP_KEY) { // extension: emits the group key; no argument
            dts[a] = NQE_INT64;
            src_ids[a] = 0;
            continue;
        }
        const DevColumn *c = col_at(aggs[a].column);
        if (!c) return nqe_fail(ctx, NQE_ERR_PANIC, "aggregate column index %d out of range", aggs[a].column);
        dts[a] = c->dtype;
        fill_src(&jp.val[a], *c);
        jp.val_left[a] = aggs[a].column < nl;
        src_ids[a] = a; // jp.val[a]; identical columns share an id
        for (int b2 = 0; b2 < a; b2++)
            if (aggs[b2].op != NQE_AGG_GROUP_KEY && aggs[b2].column == aggs[a].column) { src_ids[a] = src_ids[b2]; break; }
    }
    NQE_TRY(nqe_agg_layout(ctx, aggs, n_aggs, dts, src_ids, true, &ap));
    fill_src(&jp.group, *g);
    jp.group_left = group_column < nl;
    ap.n_rows = right->nrows;
    ap.status = (uint32_t *)(ctx->d_scratch + 1);

    OpTimer timer(ctx);
    NQE_CUDA(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream));
    // only the group key is needed from the build side: it rides in the slots' row word
    bool only_group_from_build = jp.group_left && group_column != left_key;
    for (int a = 0; a < n_aggs; a++)
        if (jp.val_left[a]) only_group_from_build = false;
    static int allow_rowpay = -1;
    if (allow_rowpay < 0) {
        const char *e = getenv("NQE_JOIN_ROWPAY");
        allow_rowpay = e ? atoi(e) : 1;
    }
    jp.probe_keys = (const unsigned long long *)right->cols[right_key].values;
    jp.n_probe = right->nrows;
    int32_t rc = NQE_OK;

    // ---- paged path (see "Fused join -> group-by over paged streams" above): unique build keys, the group key rides
    // in the slots, every aggregate reads ONE NULL-free 8-byte probe-side column.  Everything but "unique build keys,
    // no payload collision" is known before the build, and the first split depends only on the table's SIZE: it is
    // started on the auxiliary stream and runs beside the build (if the build then disqualifies the plan the split is
    // simply dropped).
    static std::atomic<int> allow_paged{-1};
    static size_t l2_budget = 0;
    if (allow_paged.load(std::memory_order_acquire) < 0) {
        const char *e;
        // slot range of one partition of the first split.  Measured (1e8 x 1e7, whole operator, two runs each): 48 MiB
        // (7 partitions) 4.29 / 12 MiB 4.40 / 8 MiB 4.38 / 6 MiB (54 partitions) 4.09 / 4 MiB 4.17 / 3 MiB 4.44 ms: every
        // range is L2-resident; with few partitions the split's shared-memory counters are hot
        e = getenv("NQE_JA_PART_MB");
        l2_budget = (size_t)(e && atoi(e) > 0 ? atoi(e) : 6) << 20;
        e = getenv("NQE_JOINAGG_PAGED");
        allow_paged.store(e ? atoi(e) != 0 : 1, std::memory_order_release);
    }
    const DevColumn *gc = only_group_from_build && allow_rowpay ? &left->cols[group_column] : nullptr;
    bool paged = allow_paged && gc && !gc->validity && jp.n_probe >= join_part_min_rows() && ap.n_states > 0 && left->nrows > 0;
    int need = 0;
    for (int q = 0; q < ap.n_states && paged; q++) {
        if (ap.st_src[q] != ap.st_src[0]) paged = false;
        need |= 1 << ap.st_kind[q];
    }
    const ColSrc &vc = jp.val[ap.n_states > 0 ? ap.st_src[0] : 0];
    if (paged && (vc.validity || (vc.dtype != NQE_INT64 && vc.dtype != NQE_UINT64 && vc.dtype != NQE_FLOAT64))) paged = false;
    int P2 = 0, m2 = 1;
    double est_groups = 0;
    long long dense_lo = 0;
    uint32_t dense_width = 0;
    if (paged) {
        rc = nqe_estimate_distinct_u64(ctx, (const unsigned long long *)gc->values, left->nrows, &est_groups);
        if (rc == NQE_OK && (est_groups < 2048.0 || !nqe_gp2_plan(ctx, est_groups, &P2, &m2))) paged = false;
    }
    static int allow_dense = -1; // knob NQE_AGG_DENSE=0: always hash the group keys
    if (allow_dense < 0) {
        const char *e = getenv("NQE_AGG_DENSE");
        allow_dense = e ? atoi(e) : 1;
    }
    long long group_minmax[2] = {0, -1};
    bool have_group_minmax = false;
    if (rc == NQE_OK && paged && allow_dense) {
        // dense group keys: the build side gives their exact range, so the re-split goes by key range and the groups are
        // aggregated in directly indexed shared-memory tables (no key compares, no inserts)
        long long lo, hi;
        rc = nqe_minmax_i64(ctx, (const unsigned long long *)gc->values, left->nrows, &lo, &hi);
        if (rc == NQE_OK) {
            group_minmax[0] = lo;
            group_minmax[1] = hi;
            have_group_minmax = true;
        }
        if (rc == NQE_OK && lo <= hi) {
            const unsigned long long range = (unsigned long long)hi - (unsigned long long)lo + 1ull;
            const unsigned long long per_part = (range + nqe_dense_parts(ctx, range) - 1) / nqe_dense_parts(ctx, range);
            if (range < (1ull << 32) && (double)range <= 4.0 * est_groups + 1024.0 && per_part <= NQE_GP2_DENSE_MAX_WIDTH &&
                ctx->sm_count <= PS_MAX_PARTS) {
                dense_lo = lo;
                dense_width = (uint32_t)(per_part < 64 ? 64 : per_part);
                P2 = (int)((range + dense_width - 1) / dense_width);
                m2 = ctx->sm_count / P2;
            }
        }
    }
    PagedStreams s1, s2;
    memset(&s1, 0, sizeof s1);
    memset(&s2, 0, sizeof s2);
    bool split_started = false;
    uint32_t *split_status = (uint32_t *)(ctx->d_scratch + 20); // the split's own status word: the build uses words 1..4
    // dense unique build keys: a direct table (JoinTable::direct) -- no first split, the probe rows are read in place
    bool direct = false;
    if (rc == NQE_OK)
        rc = build_table_direct(ctx, left, left_key, &jp.jt, only_group_from_build && allow_rowpay ? group_column : -1, &direct,
                                have_group_minmax ? group_minmax : nullptr);
    if (rc == NQE_OK && paged && !direct) {
        const size_t table_bytes = (((size_t)((double)left->nrows / 0.5) + 16) & ~(size_t)1) * 16; // = build_table's capacity
        int P1 = (int)((table_bytes + l2_budget - 1) / l2_budget);
        if (P1 < 1) P1 = 1;
        if (P1 > ctx->sm_count) P1 = ctx->sm_count;
        if (P1 > PS_MAX_PARTS) P1 = PS_MAX_PARTS;
        rc = nqe_ps_create(ctx, jp.n_probe, P1, &s1);
        if (rc == NQE_OK && !ctx->s_aux) {
            if (cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess)
                rc = nqe_fail(ctx, NQE_ERR_CUDA, "auxiliary stream creation failed");
        }
        if (rc == NQE_OK) {
            s1.status = split_status;
            cudaMemsetAsync(split_status, 0, sizeof(uint64_t), ctx->stream);
            cudaEventRecord(ctx->ev_fork, ctx->stream); // the streams' buffers were allocated in ctx->stream's order
            cudaStreamWaitEvent(ctx->s_aux, ctx->ev_fork, 0);
            const PsSplitArgs sa{jp.probe_keys, (const unsigned long long *)vc.values, jp.n_probe, vc.dtype};
            DevProgramSet none;
            memset(&none, 0, sizeof none);
            cudaStream_t main_stream = ctx->stream;
            ctx->stream = ctx->s_aux; // the split's launches go to the auxiliary stream
            rc = ps_split_launch<false, PartBySlotRange>(ctx, s1, sa, PartBySlotRange{(uint64_t)P1}, none, split_status);
            ctx->stream = main_stream;
            cudaEventRecord(ctx->ev_join, ctx->s_aux);
            split_started = true;
        }
    }
    if (rc == NQE_OK && !direct) rc = build_table(ctx, left, left_key, &jp.jt, only_group_from_build && allow_rowpay ? group_column : -1, true);
    if (split_started) cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0); // also orders the frees below
    jp.group_in_slot = jp.jt.rowpay ? 1 : 0;
    nqe_table *t;
    nqe_table_new(ctx, 0, &t);
    if (paged && (rc != NQE_OK || !jp.jt.rowpay || jp.jt.has_dups)) { // duplicate build keys / a payload equal to the free-slot marker
        paged = false;
        nqe_ps_destroy(ctx, &s1);
    }
    if (rc == NQE_OK && paged) {
        rc = nqe_ps_create(ctx, jp.n_probe, P2, &s2);
        if (rc == NQE_OK) {
            cudaMemsetAsync(ctx->d_scratch, 0, 16 * sizeof(uint64_t), ctx->stream); // not the split's word
            if (direct) {
                const PsSplitArgs sa{jp.probe_keys, (const unsigned long long *)vc.values, jp.n_probe, vc.dtype};
                rc = ja_direct_scatter_launch(ctx, sa, s2, jp.jt, (uint32_t)P2, dense_lo, dense_width);
            } else {
                rc = ja_probe_scatter_launch(ctx, s1, s2, jp.jt, (uint32_t)P2, dense_lo, dense_width);
            }
        }
        uint64_t capacity = nqe_agg_capacity(est_groups);
        for (int attempt = 0; rc == NQE_OK && attempt < 8; attempt++) {
            rc = nqe_agg_table_create(ctx, &ap, capacity);
            if (rc != NQE_OK) break;
            rc = nqe_gp2_aggregate(ctx, s2, ap, m2, need, dense_lo, dense_width);
            if (rc != NQE_OK) break;
            cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 24 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
                rc = nqe_fail(ctx, NQE_ERR_CUDA, "join-aggregate kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
            const uint32_t st = (uint32_t)ctx->h_scratch[1] | ((uint32_t)ctx->h_scratch[20] & DEV_ERR_CAPACITY); // [20]: the first split's word
            if (st & DEV_ERR_CAPACITY) { rc = nqe_fail(ctx, NQE_ERR_CUDA, "internal: paged stream pool exhausted"); break; }
            if (!(st & DEV_ERR_TABLE_FULL)) break;
            nqe_dev_free(ctx, ap.table);
            ap.table = nullptr;
            capacity *= 8;
            cudaMemsetAsync(ctx->d_scratch, 0, 16 * sizeof(uint64_t), ctx->stream);
            if (attempt == 7) rc = nqe_fail(ctx, NQE_ERR_OOM, "group-by table kept overflowing");
        }
        if (rc == NQE_OK) rc = nqe_agg_extract(ctx, ap, false, left->nrows + 1, t);
        timer.stop();
        nqe_ps_destroy(ctx, &s1);
        nqe_ps_destroy(ctx, &s2);
        nqe_dev_free(ctx, ap.table);
        nqe_dev_free(ctx, jp.jt.words);
        if (rc != NQE_OK) {
            nqe_table_free(t);
            return rc;
        }
        *out = t;
        return NQE_OK;
    }

    // ---- direct path: one pass over the probe side, group table in L2
    // groups come from one column of one side: at most that side's row count
    const int64_t side_rows = jp.group_left ? left->nrows : right->nrows;
    uint64_t capacity = nqe_agg_capacity((double)(side_rows < (1 << 18) ? side_rows : (1 << 18)));
    for (int attempt = 0; rc == NQE_OK && attempt < 8; attempt++) {
        cudaMemsetAsync(ctx->d_scratch, 0, 64 * sizeof(uint64_t), ctx->stream);
        rc = nqe_agg_table_create(ctx, &ap, capacity);
        if (rc != NQE_OK) break;
        if (jp.n_probe > 0) {
            const int64_t tiles = (jp.n_probe + HJ_K * HJ_THREADS - 1) / (HJ_K * HJ_THREADS);
            int grid = ctx->sm_count * 8;
            if (grid > tiles) grid = (int)tiles;
            join_aggregate_kernel<<<grid, HJ_THREADS, 0, ctx->stream>>>(jp, ap);
            ctx->launches++;
        }
        cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            rc = nqe_fail(ctx, NQE_ERR_CUDA, "join-aggregate kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if (!((uint32_t)ctx->h_scratch[1] & DEV_ERR_TABLE_FULL)) break;
        nqe_dev_free(ctx, ap.table);
        ap.table = nullptr;
        capacity *= 8;
        if (attempt == 7) rc = nqe_fail(ctx, NQE_ERR_OOM, "group-by table kept overflowing");
    }
    if (rc == NQE_OK) rc = nqe_agg_extract(ctx, ap, false, side_rows + 1, t);
    timer.stop();
    nqe_dev_free(ctx, ap.table);
    nqe_dev_free(ctx, jp.jt.words);
    if (rc != NQE_OK) {
        nqe_table_free(t);
        return rc;
    }
    *out = t;
    return NQE_OK;
}

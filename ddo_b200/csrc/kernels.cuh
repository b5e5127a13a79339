// kernels.cuh -- hand-written sm_100a kernels of the batch DD-compilation engine (MISP device model).
//
// One "layer step" t -> t+1 of every DD of the batch is three launches:
//   k_expand  (flat, G lanes per node)   clean.rs:360-370 + 728-776 + misp/main.rs:77-102,191-193:
//             rough-upper-bound prune, transition / transition_cost of both decisions, 128-bit coalesced row loads and
//             stores, open-addressing dedup (atomicCAS claim + 64-bit atomicMax of (value_top, candidate) per duplicate).
//   k_finish  (one CTA per DD)           clean.rs:350 (misp/main.rs:109-143 next_variable), :779-876 (_restrict/_relax):
//             canonical representatives, stable scan, positional-popcount histogram (warp bit-matrix transposes),
//             MSD radix-select of the width cut with the full (value_top, popcount, lexicographic) key, OR-merge of the
//             overflow, recycled-node lookup, last-exact-layer bookkeeping.
//   k_compact (flat, G lanes per cand)   clean.rs:657-687: scatter survivors to the next layer (ping-pong SoA), parent log,
//             child log, hash-table slot release, last-exact-layer snapshot.
// After the last layer: k_finalize (best nodes, exact-best-path walk, clean.rs:620-655), k_bottomup (local bounds, clean.rs:448-475,
// cutset upper bounds :417-445), k_cutset_{count,offsets,write} (drain_cutset compaction).
#pragma once
#include "engine.hpp"

#include <cooperative_groups.h>

namespace ddo {

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
__device__ __forceinline__ uint64_t word_hash(uint64_t w, int j) { return mix64(w + 0x9E3779B97F4A7C15ULL * (uint64_t)(j + 1)); }
__device__ __forceinline__ unsigned long long pack_key(int32_t value, uint32_t cand) {
    return ((unsigned long long)((uint32_t)value ^ 0x80000000u) << 32) | cand;
}
__device__ __forceinline__ int32_t key_value(unsigned long long key) { return (int32_t)((uint32_t)(key >> 32) ^ 0x80000000u); }
// ranking tie-break word: BitSet::cmp (lexicographic over ascending members) == unsigned order of ~bitreverse, larger = Greater
__device__ __forceinline__ uint64_t lex_word(uint64_t w) { return ~__brevll(w); }
// Low 20 bits of a candidate's rank word: its two smallest members, RB bits each (9 for n <= 512, 10 for n <= 1024), all ones where there
// is none.  Among sets of equal size BitSet::cmp is the lexicographic order of the ascending member lists, so (popcount, m1, m2) is an
// order-preserving prefix of the ranking -- and a dense one: the first members of a late, sparse state say far more than the membership
// of the 20 lowest vertices.
// Programmatic dependent launch (the three kernels of a layer step are launched with programmatic stream serialization, engine.cu): wait
// until the previous kernel of the stream has completed and its writes are visible, then let the NEXT kernel's CTAs become resident behind
// this one so that they start the moment this grid drains (launch latency and CTA rasterisation leave the critical path of narrow batches).
// Both instructions are no-ops in a grid launched without the attribute.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <int S> struct RankBits { static constexpr int RB = S <= 8 ? 9 : 10; };
template <int S> __device__ __forceinline__ void rank_take(uint64_t w, int base, int& got, uint32_t& mk) {
    while (w && got < 2) { mk = (mk << RankBits<S>::RB) | (uint32_t)(base + __ffsll((long long)w) - 1); ++got; w &= w - 1; }
}
template <int S> __device__ __forceinline__ uint32_t rank_pad(int got, uint32_t mk) {
    for (; got < 2; ++got) mk = (mk << RankBits<S>::RB) | ((1u << RankBits<S>::RB) - 1u);
    return mk;
}

__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {  // L2-only load (coherent across SMs inside one kernel)
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {  // streaming read-once row load
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(uint4* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint64_t u4lo(uint4 v) { return (uint64_t)v.x | ((uint64_t)v.y << 32); }
__device__ __forceinline__ uint64_t u4hi(uint4 v) { return (uint64_t)v.z | ((uint64_t)v.w << 32); }
__device__ __forceinline__ uint4 mk_u4(uint64_t lo, uint64_t hi) { return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)); }

// ---- sub-warp group helpers (G = lanes per node, power of two <= 32) ------------------------------------------
template <int G> __device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return FULL_MASK;
    unsigned lane = threadIdx.x & 31;
    return ((G == 32 ? 0u : (1u << G)) - 1u) << (lane & ~(G - 1));
}
template <int G> __device__ __forceinline__ int group_sum(int v, unsigned m) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v += __shfl_xor_sync(m, v, d);
    return v;
}
template <int G> __device__ __forceinline__ uint64_t group_xor64(uint64_t v, unsigned m) {
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) v ^= __shfl_xor_sync(m, v, d);
    return v;
}
template <int G> __device__ __forceinline__ bool group_all(bool p, unsigned m) { return (__ballot_sync(m, p) & m) == m; }
template <int G> __device__ __forceinline__ bool group_any(bool p, unsigned m) { return (__ballot_sync(m, p) & m) != 0; }

// ---- block helpers (blockDim.x == NT, multiple of 32, <= 1024) --------------------------------------------------
template <typename T, typename Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(FULL_MASK, v, d));
    return v;
}
// reduce over the block; result broadcast to all threads. `scratch` has >= 33 T slots.
template <typename T, typename Op>
__device__ T block_reduce(T v, Op op, T identity, T* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_reduce(v, op);
    __syncthreads();  // protect scratch reuse
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    if (w == 0) {
        T x = lane < nw ? scratch[lane] : identity;
        x = warp_reduce(x, op);
        if (lane == 0) scratch[32] = x;
    }
    __syncthreads();
    return scratch[32];
}
// exclusive scan of one int per thread; returns exclusive prefix, *total = sum.  scratch: >= 33 ints.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int* scratch) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int n = __shfl_up_sync(FULL_MASK, inc, d); if (lane >= d) inc += n; }
    __syncthreads();
    if (lane == 31) scratch[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < nw ? scratch[lane] : 0;
        int xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int n = __shfl_up_sync(FULL_MASK, xi, d); if (lane >= d) xi += n; }
        scratch[lane] = xi - x;  // exclusive warp offsets
        if (lane == 31) scratch[32] = xi;
    }
    __syncthreads();
    *total = scratch[32];
    return scratch[w] + inc - v;
}

// 32x32 bit-matrix transpose across the lanes of a warp: afterwards bit r of lane b == bit b of (former) lane r.
// Five butterfly steps of shuffle + rotate + one LOP3: a lane whose index has bit j clear keeps its fields with (index & j) == 0 and takes
// the partner's same fields into the others, a lane with bit j set does the opposite (the bits a rotate wraps around are masked away).
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : (j == 8 ? 0x00FF00FFu : (j == 4 ? 0x0F0F0F0Fu : (j == 2 ? 0x33333333u : 0x55555555u)));
        const bool hi = (lane & j) != 0;
        const uint32_t p = __shfl_xor_sync(FULL_MASK, x, j);
        const uint32_t q = __funnelshift_l(p, p, hi ? 32 - j : j);
        const uint32_t mm = hi ? ~m : m;
        x = (x & mm) | (q & ~mm);
    }
    return x;
}

// state hash of the dedup table: multilinear over the 64-bit words, h = mix64(sum_j word_j * hash_mul(j)) (mod 2^64, odd multipliers)
__device__ __forceinline__ uint64_t hash_mul(int j) { return mix64(0x9E3779B97F4A7C15ULL * (uint64_t)(j + 1)) | 1ull; }

// =================================================================================================================
// k_init: root of every DD becomes the single "candidate" of layer 0 (clean.rs:383-405)
// =================================================================================================================
template <int S>
__global__ void k_init(EV ev, int count, int comp_type, long long best_lb, int dual) {
    const int k = blockIdx.x;
    if (k >= count) {  // dual mode: slot count + j is the (not yet forked) relaxed twin of DD j
        if (dual && k < 2 * count && threadIdx.x == 0) {
            const int p = k - count;
            DDCtl c{};
            c.status = ST_WAITING; c.ncand = 0; c.n_cur = 0; c.var = -1;
            c.width = ev.root_width[p]; c.comp_type = DDO_RELAXED; c.root_depth = ev.root_depth[p]; c.lel = -1;
            c.t_term = -1; c.best_pos = -1; c.best_exact_pos = -1; c.root_value = ev.root_val[p];
            c.best_lb = best_lb; c.primary = p; c.fork_t = -1;
            ev.ctl[k] = c;
        }
        return;
    }
    DDCtl* ctl = ev.ctl + k;
    const size_t cb = (size_t)k * ev.C;
    if (threadIdx.x < S) ev.cand_state[cb * S + threadIdx.x] = ev.root_state[(size_t)k * S + threadIdx.x];
    if (threadIdx.x == 0) {
        DDCtl c{};
        c.status = ST_ACTIVE; c.ncand = 1; c.n_cur = 0; c.var = -1;
        c.width = ev.root_width[k]; c.comp_type = comp_type; c.root_depth = ev.root_depth[k]; c.lel = -1;
        c.t_term = -1; c.best_pos = -1; c.best_exact_pos = -1; c.root_value = ev.root_val[k];
        c.best_lb = best_lb; c.primary = -1; c.fork_t = -1;
        *ctl = c;
        ev.ucount[k] = 1;
        int pc = 0;
        for (int j = 0; j < S; ++j) pc += __popcll(ev.root_state[(size_t)k * S + j]);
        ev.cand_rep[cb] = 0; ev.cand_first[cb] = 0; ev.cand_agg[cb] = pack_key(ev.root_val[k], PLOG_CAND_MASK); ev.cand_inex[cb] = 0;
        {
            int got = 0; uint32_t mk = 0;
            for (int j = 0; j < S; ++j) rank_take<S>(ev.root_state[(size_t)k * S + j], 64 * j, got, mk);
            ev.cand_rank[cb] = ((uint32_t)pc << 20) | rank_pad<S>(got, mk);
        }
        ev.cand_slot[cb] = NONE32; ev.uflag[cb] = 0;
        if (k == 0) { *ev.active = count; ev.fin_cnt[0] = 0; ev.fin_cnt[1] = count; }
        ev.fin_list[ev.K + k] = k;
    }
    // vertex histogram of the (single) root state, consumed by k_finish(0)
    for (int u = threadIdx.x; u < S * 64; u += blockDim.x)
        ev.vhist[(size_t)k * ev.HN + u] = (uint32_t)((ev.root_state[(size_t)k * S + (u >> 6)] >> (u & 63)) & 1ull);
}

// =================================================================================================================
// k_expand: layer t -> candidates of layer t+1
// =================================================================================================================
// tile -> (DD, first item) lookup over the per-layer work plan written by the last CTA of k_finish
__device__ __forceinline__ int plan_find(const int* off, int count, int tile) {
    int lo = 0, hi = count;  // largest k with off[k] <= tile
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= tile) lo = mid; else hi = mid; }
    return lo;
}

// The same lookup by a whole warp in two dependent round trips instead of log2(count): 32 samples of the plan pick a segment, the segment's
// entries pick the DD.  Every lane of the warp must call it with the same arguments.
__device__ __forceinline__ int plan_find_warp(const int* off, int count, int tile) {
    const int lane = threadIdx.x & 31;
    const int str = (count + 31) >> 5;
    const bool le1 = off[min(lane * str, count - 1)] <= tile;  // monotone over the lanes; lane 0 reads off[0] = 0
    const int seg = 31 - __clz(__ballot_sync(0xffffffffu, le1));
    const int base = min(seg * str, count - 1), end = min(base + str, count);
    int best = base;
    for (int j = base; j < end; j += 32) {
        const int i = j + lane;
        const unsigned b = __ballot_sync(0xffffffffu, i < end && off[i] <= tile);
        if (!b) break;
        best = j + 31 - __clz(b);
    }
    return best;
}

#ifndef DDO_EXPAND_MINB
#define DDO_EXPAND_MINB 8
#endif
template <int S>
__global__ void __launch_bounds__(256, DDO_EXPAND_MINB) k_expand(EV ev, int t, int count) {
    pdl_enter();
    constexpr int G = S / 2;          // lanes per node, each owning one 128-bit chunk (two words)
    constexpr int NPB = 256 / G;      // nodes per tile
    constexpr int RP = G >= 8 ? 1 : 8 / G;  // claimed rows per 128 bytes of the staging buffer (bank swizzle below)
    // per-tile counters and the per-DD claim counter live in separate 16-byte slots: thread 0 updates the claim counter while the other
    // threads may still be reading the tile counters (the compiler is free to fetch neighbouring words with one vector load)
    __shared__ __align__(16) unsigned int s_tilectr[4];
    __shared__ __align__(16) unsigned int s_claimbox[4];
    unsigned int& s_exp = s_tilectr[0]; unsigned int& s_tr = s_tilectr[1]; unsigned int& s_rows = s_tilectr[2]; unsigned int& s_claims = s_claimbox[0];
    __shared__ uint4 s_claim[2 * NPB * G];     // distinct states claimed by this tile (zero rows elsewhere): row r, chunk q at [r*G + (q ^ ((r / RP) & (G-1)))]
    __shared__ unsigned int s_hist[64 * S];    // per-vertex occurrence counts of the claimed states, flushed per DD
    const int* off = ev.tile_off_e;  // (L1-resident; staging the plan in shared memory measured slower)
    const int total = off[count];
    const int sub = threadIdx.x % G;
    const unsigned gm = group_mask<G>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t hc0 = hash_mul(2 * sub), hc1 = hash_mul(2 * sub + 1);
    // contiguous tile range per block, so that the histogram of consecutive tiles of one DD is flushed once
    const int tpb = (total + gridDim.x - 1) / gridDim.x;
    const int tile_lo = min((int)blockIdx.x * tpb, total), tile_hi = min(tile_lo + tpb, total);
    for (int i = threadIdx.x; i < 64 * S; i += 256) s_hist[i] = 0;
    if (threadIdx.x == 0) s_claims = 0;
    int hist_k = -1;
    int k = tile_lo < tile_hi ? plan_find(off, count, tile_lo) : 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
    while (off[k + 1] <= tile) ++k;  // tiles of a block are consecutive: the DD index only moves forward
    if (k != hist_k) {
        __syncthreads();
        if (hist_k >= 0) {
            for (int i = threadIdx.x; i < 64 * S; i += 256) { const unsigned v = s_hist[i]; if (v) { atomicAdd(ev.vhist + (size_t)hist_k * ev.HN + i, v); s_hist[i] = 0; } }
            if (threadIdx.x == 0 && s_claims) { atomicAdd(ev.ucount + hist_k, s_claims); s_claims = 0; }
        }
        hist_k = k;
    }
    DDCtl* ctl = ev.ctl + k;
    const int n_cur = ctl->n_cur;
    const int node = (tile - off[k]) * NPB + threadIdx.x / G;
    __syncthreads();
    if (threadIdx.x == 0) { s_exp = 0; s_tr = 0; s_rows = 0; }
    s_claim[threadIdx.x] = make_uint4(0, 0, 0, 0); s_claim[256 + threadIdx.x] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (node < n_cur) {
        const int buf = t & 1;
        const size_t nb = (size_t)k * ev.Wcap + node;
        const uint4 s4 = ld_stream_u4(reinterpret_cast<const uint4*>(ev.cur_state[buf] + nb * S) + sub);
        uint64_t w0 = u4lo(s4), w1 = u4hi(s4);
        const int val = ev.cur_val[buf][nb];
        const uint32_t fl = ev.cur_flag[buf][nb];
        // rough upper bound: sum of the weights of the remaining vertices (misp/main.rs:191-193)
        int rub;
        if (ev.unit_weights) rub = __popcll(w0) + __popcll(w1);
        else {
            rub = 0;
            uint64_t x = w0; const int32_t* wp = ev.weight + (2 * sub) * 64;
            while (x) { int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; }
            x = w1; wp += 64;
            while (x) { int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; }
        }
        rub = group_sum<G>(rub, gm);
        const bool expandable = ((long long)rub + (long long)val) > ctl->best_lb;  // clean.rs:364-365 (no saturation possible in 32+32 bits)
        const int v = ctl->var;
        const int vw = v >> 6;
        const bool owner = (vw >> 1) == sub;
        const uint64_t bit = 1ull << (v & 63);
        const bool has_v = group_any<G>(owner && (((vw & 1) ? w1 : w0) & bit), gm);  // misp/main.rs:96
        const size_t cb = (size_t)k * ev.C;
        const uint32_t c_yes = 2u * node, c_no = 2u * node + 1u;  // for_each_in_domain order: YES then NO (main.rs:95-102)
        if (sub == 0) {
            ev.cur_rub[nb] = rub;
            ev.cand_rep[cb + c_yes] = NONE32; ev.cand_rep[cb + c_no] = NONE32;
            ev.uflag[cb + c_yes] = 0; ev.uflag[cb + c_no] = 0;
            if (expandable) { atomicAdd(&s_exp, 1u); atomicAdd(&s_tr, has_v ? 2u : 1u); }
        }
        if (expandable) {
            if (owner) { if (vw & 1) w1 &= ~bit; else w0 &= ~bit; }  // res.remove(var) main.rs:79
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                // d == 0: YES child (only when v is in the state), d == 1: NO child
                if (d == 0 && !has_v) continue;
                uint64_t a0 = w0, a1 = w1;
                int value = val;
                if (d == 0) {
                    const uint4 nc4 = __ldg(reinterpret_cast<const uint4*>(ev.nc + (size_t)v * S) + sub);  // main.rs:82
                    a0 &= u4lo(nc4); a1 &= u4hi(nc4);
                    value += ev.weight[v];  // main.rs:87-93
                }
                const uint32_t c = d == 0 ? c_yes : c_no;
                uint4* dst = reinterpret_cast<uint4*>(ev.cand_state + (cb + c) * S) + sub;
                st_stream_u4(dst, mk_u4(a0, a1));
                uint64_t hs = a0 * hc0 + a1 * hc1;
#pragma unroll
                for (int dd = G / 2; dd > 0; dd >>= 1) hs += __shfl_xor_sync(gm, hs, dd);
                const uint64_t h = mix64(hs);
                const int pc = group_sum<G>(__popcll(a0) + __popcll(a1), gm);
                uint32_t mk2;
                {   // the two smallest members of the child: every lane lists up to two of its own, the group merges them in lane order
                    int lcnt = 0; uint32_t lmk = 0;
                    rank_take<S>(a0, 128 * sub, lcnt, lmk); rank_take<S>(a1, 128 * sub + 64, lcnt, lmk);
                    int got = 0; uint32_t mk = 0;
#pragma unroll
                    for (int s2 = 0; s2 < G; ++s2) {
                        const int srcl = ((threadIdx.x & 31) & ~(G - 1)) + s2;
                        const int c2 = __shfl_sync(gm, lcnt, srcl); const uint32_t l2 = __shfl_sync(gm, lmk, srcl);
                        if (got < 2 && c2 > 0) {
                            if (c2 == 2) { if (got == 0) { mk = l2; got = 2; } else { mk = (mk << RankBits<S>::RB) | (l2 >> RankBits<S>::RB); got = 2; } }
                            else { mk = (mk << RankBits<S>::RB) | l2; got += 1; }
                        }
                    }
                    mk2 = rank_pad<S>(got, mk);
                }
                if (sub == 0) {
                    ev.cand_rank[cb + c] = ((uint32_t)pc << 20) | mk2;
                    ev.cand_agg[cb + c] = pack_key(value, c);
                    ev.cand_first[cb + c] = c;
                    ev.cand_inex[cb + c] = (uint8_t)(fl & NF_INEXACT);
                }
                __threadfence();
                __syncwarp(gm);
                // open-addressing insert (next_l.entry(), clean.rs:738)
                const uint32_t tag = (uint32_t)(h >> 32);
                const unsigned long long entry = ((unsigned long long)tag << 32) | c;
                uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
                unsigned long long* tab = ev.table + (size_t)k * ev.T;
                for (;;) {
                    unsigned long long old = 0;
                    if (sub == 0) old = atomicCAS(tab + slot, EMPTY64, entry);
                    old = __shfl_sync(gm, old, (threadIdx.x & 31) & ~(G - 1));
                    if (old == EMPTY64) {  // Entry::Vacant, clean.rs:739-765: a new distinct state of the next layer
                        const unsigned row = d * NPB + threadIdx.x / G;
                        if (sub == 0) { ev.cand_rep[cb + c] = c; ev.cand_slot[cb + c] = slot; atomicAdd(&s_rows, 1u); }
                        s_claim[row * G + (sub ^ ((row / RP) & (G - 1)))] = mk_u4(a0, a1);
                        break;
                    }
                    if ((uint32_t)(old >> 32) == tag) {
                        const uint32_t oc = (uint32_t)old;
                        const uint4 o4 = ld_cg_u4(reinterpret_cast<const uint4*>(ev.cand_state + (cb + oc) * S) + sub);
                        const bool eq = group_all<G>(u4lo(o4) == a0 && u4hi(o4) == a1, gm);
                        if (eq) {  // Entry::Occupied, clean.rs:766-774 + append_edge_to! :199-220
                            if (sub == 0) {
                                atomicMax(ev.cand_agg + cb + oc, pack_key(value, c));   // value_top = max, `>=`: last (largest) candidate wins
                                atomicMin(ev.cand_first + cb + oc, c);                    // canonical identity = first candidate
                                if (fl & NF_INEXACT) ev.cand_inex[cb + oc] = 1;           // exact &= parent.exact
                                ev.cand_rep[cb + c] = oc;
                            }
                            break;
                        }
                    }
                    slot = (slot + 1) & (uint32_t)(ev.T - 1);
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_exp) { atomicAdd(&ctl->expanded, (unsigned long long)s_exp); atomicAdd(&ctl->transitions, (unsigned long long)s_tr); }
    // next_variable's histogram (misp/main.rs:131-135), fused: every claimed row is a distinct state of the next layer.  Warp w transposes
    // 32 rows x 32-bit columns at a time (one 128-bit chunk = four column blocks per conflict-free shared-memory load).
    const int nrows = (int)s_rows;
    if (nrows > 0) {
        const int nblk = (2 * NPB) >> 5;
        if (threadIdx.x == 0) s_claims += (unsigned)nrows;
        for (int job = warp; job < nblk * G; job += 8) {
            const int rblk = job / G, q = job % G;
            const int row = rblk * 32 + lane;
            const uint4 x4 = s_claim[row * G + (q ^ ((row / RP) & (G - 1)))];
            const uint32_t xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const unsigned cnt = __popc(warp_transpose32(xs[e]));  // lane b: #rows holding vertex 32*(4q+e) + b
                if (cnt) atomicAdd(&s_hist[32 * (4 * q + e) + lane], cnt);
            }
        }
    }
    }  // tile loop
    __syncthreads();
    if (hist_k >= 0) {
        for (int i = threadIdx.x; i < 64 * S; i += 256) { const unsigned v = s_hist[i]; if (v) atomicAdd(ev.vhist + (size_t)hist_k * ev.HN + i, v); }
        if (threadIdx.x == 0 && s_claims) atomicAdd(ev.ucount + hist_k, s_claims);
    }
}


// =================================================================================================================
// k_expand1: the same layer expansion with ONE THREAD per node (the whole S-word state in registers), for wide batches.  With G lanes
// per node (k_expand) every warp instruction serves 32/G nodes and the group reductions (rough upper bound, hash, popcount, equality)
// are shuffles; here a warp instruction serves 32 nodes and those reductions are plain register arithmetic: ~4x fewer warp instructions
// per node.  A CTA handles up to G consecutive tiles of the work plan (256 nodes) of ONE DD at a time; the claimed rows are staged in
// shared memory, compacted through a bit mask, and bit-transposed into the per-vertex histogram exactly as in k_expand.
// =================================================================================================================
#ifndef DDO_EXPAND1_MINB
#define DDO_EXPAND1_MINB 4
#endif
template <int S>
__global__ void __launch_bounds__(256, DDO_EXPAND1_MINB) k_expand1(EV ev, int t, int count) {
    pdl_enter();
    constexpr int G = S / 2;                 // 128-bit chunks per state
    constexpr int PLAN = 256 / G;            // nodes per tile of the work plan (written by k_finish for k_expand's geometry)
    constexpr int RP = G >= 8 ? 1 : 8 / G;   // bank swizzle of the staged rows, as in k_expand
    extern __shared__ __align__(16) uint4 s_claim[];   // [512 row slots][G]: slot = d * 256 + thread
    __shared__ unsigned int s_hist[64 * S];
    __shared__ unsigned int s_mask[16], s_pref[17];     // claimed row slots (bit mask) and the exclusive prefix of its popcounts
    __shared__ unsigned short s_rowmap[512];            // r-th claimed row -> slot
    __shared__ __align__(16) unsigned int s_tilectr[4];
    __shared__ __align__(16) unsigned int s_claimbox[4];
    unsigned int& s_exp = s_tilectr[0]; unsigned int& s_tr = s_tilectr[1]; unsigned int& s_claims = s_claimbox[0];
    const int* off = ev.tile_off_e;
    const int total = off[count];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tpb = (total + gridDim.x - 1) / gridDim.x;
    const int tile_lo = min((int)blockIdx.x * tpb, total), tile_hi = min(tile_lo + tpb, total);
    for (int i = tid; i < 64 * S; i += 256) s_hist[i] = 0;
    if (tid == 0) s_claims = 0;
    int hist_k = -1;
    int k = tile_lo < tile_hi ? plan_find_warp(off, count, tile_lo) : 0;  // (block-uniform condition: whole warps call it)
    const int buf = t & 1;
    for (int tile = tile_lo; tile < tile_hi;) {
        while (off[k + 1] <= tile) ++k;
        const int ntiles = min(G, min(off[k + 1], tile_hi) - tile);  // consecutive plan tiles of this DD taken together
        if (k != hist_k) {
            __syncthreads();
            if (hist_k >= 0) {
                for (int i = tid; i < 64 * S; i += 256) { const unsigned v = s_hist[i]; if (v) { atomicAdd(ev.vhist + (size_t)hist_k * ev.HN + i, v); s_hist[i] = 0; } }
                if (tid == 0 && s_claims) { atomicAdd(ev.ucount + hist_k, s_claims); s_claims = 0; }
            }
            hist_k = k;
        }
        DDCtl* ctl = ev.ctl + k;
        const int n_cur = ctl->n_cur;
        const int node = (tile - off[k]) * PLAN + tid;
        const bool active = tid < ntiles * PLAN && node < n_cur;
        __syncthreads();
        if (tid < 16) s_mask[tid] = 0;
        if (tid == 0) { s_exp = 0; s_tr = 0; }
        __syncthreads();
        bool claimed0 = false, claimed1 = false;
        {
            const int v = ctl->var;
            const int vw = v >> 6;
            const uint64_t bit = 1ull << (v & 63);
            const size_t cb = (size_t)k * ev.C;
            uint64_t w[S];
            int val = 0; uint32_t fl = 0;
            bool expandable = false, has_v = false;
            if (active) {
                const size_t nb = (size_t)k * ev.Wcap + node;
                const uint4* src = reinterpret_cast<const uint4*>(ev.cur_state[buf] + nb * S);
#pragma unroll
                for (int q = 0; q < G; ++q) { const uint4 x = ld_stream_u4(src + q); w[2 * q] = u4lo(x); w[2 * q + 1] = u4hi(x); }
                val = ev.cur_val[buf][nb];
                fl = ev.cur_flag[buf][nb];
                int rub = 0;
                if (ev.unit_weights) {
#pragma unroll
                    for (int j = 0; j < S; ++j) rub += __popcll(w[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < S; ++j) { uint64_t x = w[j]; const int32_t* wp = ev.weight + j * 64; while (x) { const int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; } }
                }
                expandable = ((long long)rub + (long long)val) > ctl->best_lb;  // clean.rs:364-365
#pragma unroll
                for (int j = 0; j < S; ++j) if (j == vw && (w[j] & bit)) has_v = true;  // misp/main.rs:96
                ev.cur_rub[nb] = rub;
                *reinterpret_cast<uint2*>(ev.cand_rep + cb + 2u * node) = make_uint2(NONE32, NONE32);
                *reinterpret_cast<uchar2*>(ev.uflag + cb + 2u * node) = make_uchar2(0, 0);
            }
            {   // counters, warp-aggregated
                const unsigned me = __ballot_sync(FULL_MASK, expandable), mv = __ballot_sync(FULL_MASK, expandable && has_v);
                if (lane == 0 && me) { atomicAdd(&s_exp, (unsigned)__popc(me)); atomicAdd(&s_tr, (unsigned)(__popc(me) + __popc(mv))); }
            }
            if (expandable) {
                const uint32_t c_yes = 2u * node, c_no = 2u * node + 1u;  // for_each_in_domain order: YES then NO (main.rs:95-102)
#pragma unroll
                for (int j = 0; j < S; ++j) if (j == vw) w[j] &= ~bit;  // res.remove(var), main.rs:79
                const uint4* ncrow = reinterpret_cast<const uint4*>(ev.nc + (size_t)v * S);
                uint64_t hsh[2]; int vals[2];
                // both children are written first, ONE fence publishes their rows, then both are inserted
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    if (d == 0 && !has_v) continue;
                    const uint32_t c = d == 0 ? c_yes : c_no;
                    uint4* dst = reinterpret_cast<uint4*>(ev.cand_state + (cb + c) * S);
                    uint64_t hs = 0; int pc = 0; int got = 0; uint32_t mk = 0;
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        uint64_t a0 = w[2 * q], a1 = w[2 * q + 1];
                        if (d == 0) { const uint4 n4 = __ldg(ncrow + q); a0 &= u4lo(n4); a1 &= u4hi(n4); }  // main.rs:82
                        st_stream_u4(dst + q, mk_u4(a0, a1));
                        hs += a0 * hash_mul(2 * q) + a1 * hash_mul(2 * q + 1);
                        pc += __popcll(a0) + __popcll(a1);
                        rank_take<S>(a0, 128 * q, got, mk); rank_take<S>(a1, 128 * q + 64, got, mk);
                    }
                    const int value = d == 0 ? val + ev.weight[v] : val;  // main.rs:87-93
                    hsh[d] = mix64(hs); vals[d] = value;
                    ev.cand_rank[cb + c] = ((uint32_t)pc << 20) | rank_pad<S>(got, mk);
                    ev.cand_agg[cb + c] = pack_key(value, c);
                    ev.cand_first[cb + c] = c;
                    ev.cand_inex[cb + c] = (uint8_t)(fl & NF_INEXACT);
                }
                __threadfence();
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    if (d == 0 && !has_v) continue;
                    const uint32_t c = d == 0 ? c_yes : c_no;
                    const uint64_t h = hsh[d];
                    const int value = vals[d];
                    const uint32_t tag = (uint32_t)(h >> 32);
                    const unsigned long long entry = ((unsigned long long)tag << 32) | c;
                    uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
                    unsigned long long* tab = ev.table + (size_t)k * ev.T;
                    for (;;) {
                        const unsigned long long old = atomicCAS(tab + slot, EMPTY64, entry);
                        if (old == EMPTY64) {  // Entry::Vacant: a new distinct state of the next layer
                            ev.cand_rep[cb + c] = c; ev.cand_slot[cb + c] = slot;
                            if (d == 0) claimed0 = true; else claimed1 = true;
                            break;
                        }
                        if ((uint32_t)(old >> 32) == tag) {
                            const uint32_t oc = (uint32_t)old;
                            const uint4* orow = reinterpret_cast<const uint4*>(ev.cand_state + (cb + oc) * S);
                            bool eq = true;
#pragma unroll
                            for (int q = 0; q < G; ++q) {
                                uint64_t a0 = w[2 * q], a1 = w[2 * q + 1];
                                if (d == 0) { const uint4 n4 = __ldg(ncrow + q); a0 &= u4lo(n4); a1 &= u4hi(n4); }
                                const uint4 o4 = ld_cg_u4(orow + q);
                                eq = eq && u4lo(o4) == a0 && u4hi(o4) == a1;
                            }
                            if (eq) {  // Entry::Occupied, clean.rs:766-774 + append_edge_to! :199-220
                                atomicMax(ev.cand_agg + cb + oc, pack_key(value, c));
                                atomicMin(ev.cand_first + cb + oc, c);
                                if (fl & NF_INEXACT) ev.cand_inex[cb + oc] = 1;
                                ev.cand_rep[cb + c] = oc;
                                break;
                            }
                        }
                        slot = (slot + 1) & (uint32_t)(ev.T - 1);
                    }
                }
                // stage the claimed rows for the vertex histogram
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    if (!(d == 0 ? claimed0 : claimed1)) continue;
                    const int row = d * 256 + tid;
                    atomicOr(&s_mask[row >> 5], 1u << (row & 31));
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        uint64_t a0 = w[2 * q], a1 = w[2 * q + 1];
                        if (d == 0) { const uint4 n4 = __ldg(ncrow + q); a0 &= u4lo(n4); a1 &= u4hi(n4); }
                        s_claim[row * G + (q ^ ((row / RP) & (G - 1)))] = mk_u4(a0, a1);
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0 && s_exp) { atomicAdd(&ctl->expanded, (unsigned long long)s_exp); atomicAdd(&ctl->transitions, (unsigned long long)s_tr); }
        if (warp == 0) {  // exclusive prefix of the mask popcounts
            const unsigned p = lane < 16 ? (unsigned)__popc(s_mask[lane]) : 0u;
            unsigned inc = p;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned n = __shfl_up_sync(FULL_MASK, inc, d); if (lane >= d) inc += n; }
            if (lane < 16) s_pref[lane] = inc - p;
            if (lane == 15) s_pref[16] = inc;
        }
        __syncthreads();
        const int nrows = (int)s_pref[16];
        if (nrows > 0) {
#pragma unroll
            for (int d = 0; d < 2; ++d) if (d == 0 ? claimed0 : claimed1) {
                const int row = d * 256 + tid;
                const unsigned r = s_pref[row >> 5] + (unsigned)__popc(s_mask[row >> 5] & ((1u << (row & 31)) - 1u));
                s_rowmap[r] = (unsigned short)row;
            }
            if (tid == 0) s_claims += (unsigned)nrows;
            __syncthreads();
            const int nblk = (nrows + 31) >> 5;
            for (int job = warp; job < nblk * G; job += 8) {
                const int rblk = job / G, q = job % G;
                const int ri = rblk * 32 + lane;
                uint4 x4 = make_uint4(0, 0, 0, 0);
                if (ri < nrows) { const int row = s_rowmap[ri]; x4 = s_claim[row * G + (q ^ ((row / RP) & (G - 1)))]; }
                const uint32_t xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const unsigned cnt = __popc(warp_transpose32(xs[e]));
                    if (cnt) atomicAdd(&s_hist[32 * (4 * q + e) + lane], cnt);
                }
            }
        }
        tile += ntiles;
    }
    __syncthreads();
    if (hist_k >= 0) {
        for (int i = tid; i < 64 * S; i += 256) { const unsigned v = s_hist[i]; if (v) atomicAdd(ev.vhist + (size_t)hist_k * ev.HN + i, v); }
        if (tid == 0 && s_claims) atomicAdd(ev.ucount + hist_k, s_claims);
    }
}

// =================================================================================================================
// k_expand2: the thread-per-node expansion of k_expand1 with WARP-autonomous bookkeeping.  The round-1 profile of k_expand1 in the batched
// regime put 39 % of its stall samples on block barriers (per-tile counter reset / flush, claim-mask prefix, histogram hand-over) and 12 %
// on the binary search over the work plan that every thread repeated.  Here a warp owns a contiguous range of 32-node units of the plan:
//   * the claimed (= distinct) rows are staged in a 32-row buffer of the warp's own, bit-transposed when it fills, and counted into
//     per-lane registers (lane b of column j counts vertex 32 j + b); the registers are flushed to the DD's vertex histogram only when
//     the warp moves on to another DD or ends -- no block barrier anywhere in the kernel;
//   * expanded / transition / distinct counters are warp ballots, added to the DD's control block at the same moments;
//   * the plan is searched once per warp and then walked forward.
// Everything written (candidate rows and records, hash table, histogram, counters) is what k_expand1 writes.
// =================================================================================================================
template <int S>
__global__ void __launch_bounds__(256, 3) k_expand2(EV ev, int t, int count) {
    pdl_enter();
    constexpr int G = S / 2;                 // 128-bit chunks per state
    constexpr int PLAN = 256 / G;            // nodes per tile of the work plan (written by k_finish for k_expand's geometry)
    constexpr int UPT = PLAN / 32;           // 32-node units per plan tile
    constexpr int W32 = 2 * S, SROW = W32 + 1;
    __shared__ uint32_t s_stage[8][32 * SROW];
    const int* off = ev.tile_off_e;
    const int total_units = off[count] * UPT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* stage = s_stage[warp];
    const int nwarps = gridDim.x * 8, gw = warp * gridDim.x + blockIdx.x;  // consecutive unit ranges go to different CTAs: a narrow batch still spreads over every SM
    const int upw = (total_units + nwarps - 1) / nwarps;
    const int u_lo = min(gw * upw, total_units), u_hi = min(u_lo + upw, total_units);
    if (u_lo >= u_hi) return;
    const int buf = t & 1;
    int hcnt[W32];
#pragma unroll
    for (int j = 0; j < W32; ++j) hcnt[j] = 0;
    int st_cnt = 0;
    unsigned n_exp = 0, n_tr = 0, n_claim = 0;
    int k = plan_find(off, count, u_lo / UPT);
    int cur_k = k;
    auto stage_flush = [&]() {
        if (st_cnt == 0) return;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < W32; ++j) hcnt[j] += __popc(warp_transpose32(lane < st_cnt ? stage[lane * SROW + j] : 0u));
        __syncwarp();
        st_cnt = 0;
    };
    auto dd_flush = [&](int kk) {  // counters and vertex occurrences of DD kk gathered by this warp
        stage_flush();
#pragma unroll
        for (int j = 0; j < W32; ++j) if (hcnt[j]) { atomicAdd(ev.vhist + (size_t)kk * ev.HN + 32 * j + lane, (unsigned)hcnt[j]); hcnt[j] = 0; }
        if (lane == 0) {
            if (n_exp) { atomicAdd(&ev.ctl[kk].expanded, (unsigned long long)n_exp); atomicAdd(&ev.ctl[kk].transitions, (unsigned long long)n_tr); }
            if (n_claim) atomicAdd(ev.ucount + kk, n_claim);
        }
        n_exp = 0; n_tr = 0; n_claim = 0;
    };
    for (int u = u_lo; u < u_hi; ++u) {
        const int tile = u / UPT;
        while (off[k + 1] <= tile) ++k;
        if (k != cur_k) { dd_flush(cur_k); cur_k = k; }
        const DDCtl* ctl = ev.ctl + k;
        const int n_cur = ctl->n_cur;
        const int node = (tile - off[k]) * PLAN + (u % UPT) * 32 + lane;
        const bool active = node < n_cur;
        const int v = ctl->var;
        const int vw = v >> 6;
        const uint64_t bit = 1ull << (v & 63);
        const size_t cb = (size_t)k * ev.C;
        uint64_t w[S], wy[S];
        int val = 0; uint32_t fl = 0;
        bool expandable = false, has_v = false, claimed0 = false, claimed1 = false;
        if (active) {
            const size_t nb = (size_t)k * ev.Wcap + node;
            const uint4* src = reinterpret_cast<const uint4*>(ev.cur_state[buf] + nb * S);
#pragma unroll
            for (int q = 0; q < G; ++q) { const uint4 x = ld_stream_u4(src + q); w[2 * q] = u4lo(x); w[2 * q + 1] = u4hi(x); }
            val = ev.cur_val[buf][nb];
            fl = ev.cur_flag[buf][nb];
            int rub = 0;
            if (ev.unit_weights) {
#pragma unroll
                for (int j = 0; j < S; ++j) rub += __popcll(w[j]);
            } else {
#pragma unroll
                for (int j = 0; j < S; ++j) { uint64_t x = w[j]; const int32_t* wp = ev.weight + j * 64; while (x) { const int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; } }
            }
            expandable = ((long long)rub + (long long)val) > ctl->best_lb;  // clean.rs:364-365
#pragma unroll
            for (int j = 0; j < S; ++j) if (j == vw && (w[j] & bit)) has_v = true;  // misp/main.rs:96
            ev.cur_rub[nb] = rub;
            *reinterpret_cast<uint2*>(ev.cand_rep + cb + 2u * node) = make_uint2(NONE32, NONE32);
            *reinterpret_cast<uchar2*>(ev.uflag + cb + 2u * node) = make_uchar2(0, 0);
        }
        n_exp += __popc(__ballot_sync(FULL_MASK, expandable));
        n_tr += __popc(__ballot_sync(FULL_MASK, expandable)) + __popc(__ballot_sync(FULL_MASK, expandable && has_v));
        if (expandable) {
            const uint32_t c_yes = 2u * node, c_no = 2u * node + 1u;  // for_each_in_domain order: YES then NO (main.rs:95-102)
#pragma unroll
            for (int j = 0; j < S; ++j) if (j == vw) w[j] &= ~bit;  // res.remove(var), main.rs:79: w is the NO child from here on
            const uint4* ncrow = reinterpret_cast<const uint4*>(ev.nc + (size_t)v * S);
            uint64_t hsh[2]; int vals[2];
            // both children are written first, ONE fence publishes their rows, then both are inserted
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                if (d == 0 && !has_v) continue;
                const uint32_t c = d == 0 ? c_yes : c_no;
                uint4* dst = reinterpret_cast<uint4*>(ev.cand_state + (cb + c) * S);
                uint64_t hs = 0; int pc = 0; int got = 0; uint32_t mk = 0;
#pragma unroll
                for (int q = 0; q < G; ++q) {
                    uint64_t a0 = w[2 * q], a1 = w[2 * q + 1];
                    if (d == 0) { const uint4 n4 = __ldg(ncrow + q); a0 &= u4lo(n4); a1 &= u4hi(n4); wy[2 * q] = a0; wy[2 * q + 1] = a1; }  // main.rs:82
                    st_stream_u4(dst + q, mk_u4(a0, a1));
                    hs += a0 * hash_mul(2 * q) + a1 * hash_mul(2 * q + 1);
                    pc += __popcll(a0) + __popcll(a1);
                    rank_take<S>(a0, 128 * q, got, mk); rank_take<S>(a1, 128 * q + 64, got, mk);
                }
                const int value = d == 0 ? val + ev.weight[v] : val;  // main.rs:87-93
                hsh[d] = mix64(hs); vals[d] = value;
                ev.cand_rank[cb + c] = ((uint32_t)pc << 20) | rank_pad<S>(got, mk);
                ev.cand_agg[cb + c] = pack_key(value, c);
                ev.cand_first[cb + c] = c;
                ev.cand_inex[cb + c] = (uint8_t)(fl & NF_INEXACT);
            }
            __threadfence();
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                if (d == 0 && !has_v) continue;
                const uint32_t c = d == 0 ? c_yes : c_no;
                const uint64_t h = hsh[d];
                const int value = vals[d];
                const uint32_t tag = (uint32_t)(h >> 32);
                const unsigned long long entry = ((unsigned long long)tag << 32) | c;
                uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
                unsigned long long* tab = ev.table + (size_t)k * ev.T;
                for (;;) {
                    const unsigned long long old = atomicCAS(tab + slot, EMPTY64, entry);
                    if (old == EMPTY64) {  // Entry::Vacant: a new distinct state of the next layer
                        ev.cand_rep[cb + c] = c; ev.cand_slot[cb + c] = slot;
                        if (d == 0) claimed0 = true; else claimed1 = true;
                        break;
                    }
                    if ((uint32_t)(old >> 32) == tag) {
                        const uint32_t oc = (uint32_t)old;
                        const uint4* orow = reinterpret_cast<const uint4*>(ev.cand_state + (cb + oc) * S);
                        bool eq = true;
#pragma unroll
                        for (int q = 0; q < G; ++q) {
                            const uint64_t a0 = d == 0 ? wy[2 * q] : w[2 * q], a1 = d == 0 ? wy[2 * q + 1] : w[2 * q + 1];
                            const uint4 o4 = ld_cg_u4(orow + q);
                            eq = eq && u4lo(o4) == a0 && u4hi(o4) == a1;
                        }
                        if (eq) {  // Entry::Occupied, clean.rs:766-774 + append_edge_to! :199-220
                            atomicMax(ev.cand_agg + cb + oc, pack_key(value, c));
                            atomicMin(ev.cand_first + cb + oc, c);
                            if (fl & NF_INEXACT) ev.cand_inex[cb + oc] = 1;
                            ev.cand_rep[cb + c] = oc;
                            break;
                        }
                    }
                    slot = (slot + 1) & (uint32_t)(ev.T - 1);
                }
            }
        }
        // the claimed rows join the vertex histogram of the layer being built (misp/main.rs:131-135)
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const bool has = d == 0 ? claimed0 : claimed1;
            const unsigned m = __ballot_sync(FULL_MASK, has);
            if (!m) continue;
            const int add = __popc(m);
            n_claim += (unsigned)add;
            if (st_cnt + add > 32) stage_flush();
            if (has) {
                const int r = st_cnt + __popc(m & ((1u << lane) - 1u));
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const uint64_t x = d == 0 ? wy[j] : w[j];
                    stage[r * SROW + 2 * j] = (uint32_t)x; stage[r * SROW + 2 * j + 1] = (uint32_t)(x >> 32);
                }
            }
            st_cnt += add;
        }
    }
    dd_flush(cur_k);
}

// =================================================================================================================
// k_finish: one CTA per DD.  Decides everything about layer t (whose candidates were produced by k_expand(t-1)).
// =================================================================================================================
struct FinishSmem {
    int scan[40];
    unsigned long long red64[40];
    unsigned int hist[2048];   // vertex counters (n <= 2048) ; reused as 256-bin digit histogram by the select
    unsigned long long merged[32];
    int misc[16];
};

// compare two unique candidates by the cut order (clean.rs:803-808 + misp/main.rs:205-208): returns true if a is BETTER than b
template <int S>
__device__ bool cand_better(const EV& ev, size_t cb, uint32_t a, uint32_t b) {
    const unsigned long long ka = (ev.cand_agg[cb + a] & 0xFFFFFFFF00000000ull) | ev.cand_rank[cb + a];
    const unsigned long long kb = (ev.cand_agg[cb + b] & 0xFFFFFFFF00000000ull) | ev.cand_rank[cb + b];
    if (ka != kb) return ka > kb;
    for (int j = 0; j < S; ++j) {
        const uint64_t xa = lex_word(ev.cand_state[(cb + a) * S + j]), xb = lex_word(ev.cand_state[(cb + b) * S + j]);
        if (xa != xb) return xa > xb;
    }
    return false;
}

// Tie-break keys of the width cut beyond (value_top, popcount, two smallest members).  BitSet::cmp orders two sets of EQUAL size by their
// ascending member lists: at the first difference the set owning the smaller vertex is Less.  So chunk j >= 1 of the radix-select key packs
// the members of rank MK*(j-1) .. MK*j-1 of the state, BK bits each, most significant first: dense information (a raw 64-bit word of a
// late, sparse state holds a member or two, and a digit over 8 such vertices splits a tied group by ~20 % -- twenty-odd passes; a digit
// over member indices splits it 256-fold).  Sets of equal size pad identically, so the padding value never decides.
template <int S> struct MemberKey { static constexpr int BK = S <= 8 ? 9 : 10, MK = S <= 8 ? 7 : 6, CHUNKS = (64 * S + MK - 1) / MK; };
template <int S>
__device__ unsigned long long member_key(const uint64_t* st, int first_rank) {
    constexpr int BK = MemberKey<S>::BK, MK = MemberKey<S>::MK;
    unsigned long long key = 0; int got = 0, skip = first_rank;
    for (int j = 0; j < S && got < MK; ++j) {
        uint64_t w = st[j];
        const int pc = __popcll(w);
        if (skip >= pc) { skip -= pc; continue; }
        while (skip > 0) { w &= w - 1; --skip; }
        while (w && got < MK) { key = (key << BK) | (unsigned long long)(j * 64 + __ffsll((long long)w) - 1); ++got; w &= w - 1; }
    }
    for (; got < MK; ++got) key = (key << BK) | ((1ull << BK) - 1ull);
    return key;
}

// Keys / status of the distinct candidates live in shared memory when 2*Wcap of them fit (ev.smem_keys), else in global scratch.
// The two size classes of the finish of a large batch: DDs whose coming layer has more than FIN_SMALL_C candidates (k_finish: one DD per
// SM) and the others (k_finish_s: five to six per SM).  The DDs of each class are LISTED by the plan step of the previous layer (a DD
// that is done is in neither list, so the tail of a batch -- a few long DDs among hundreds of finished ones -- costs no empty CTAs).
#ifndef DDO_FIN_SMALL_C
#define DDO_FIN_SMALL_C 2048
#endif
constexpr int FIN_SMALL_C = DDO_FIN_SMALL_C;
template <int S, int NT>
__device__ void finish_body(const EV& ev, int t, FinishSmem& sm, unsigned long long* keys, uint8_t* stat, int k) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DDCtl* ctl = ev.ctl + k;
    const int status = ctl->status;
    if (status == ST_DONE) return;
    if (status == ST_TERMINAL) { __syncthreads(); if (tid == 0) ctl->status = ST_DONE; return; }
    const size_t cb = (size_t)k * ev.C;
    const size_t lb = (size_t)k * ev.Lmax;
    int vh_src = k;  // whose vertex histogram describes my candidates
    if (status == ST_WAITING) {
        // ---- relaxed twin: fork when the primary (restricted) DD is about to take its FIRST cut in this very step ---------------
        // Until then both DDs are identical (clean.rs:779-795 never fired), so the twin starts from a copy of the primary's candidates.
        // Everything read here is stable while the primary's own CTA runs finish(t) concurrently (its ctl->lel may already say t-1).
        const int p = ctl->primary;
        const DDCtl* pc = ev.ctl + p;
        const int plel = pc->lel;
        const bool fork = t >= 1 && ev.ucount[p] > (uint32_t)pc->width && (plel < 0 || plel == t - 1) && ev.nlog[(size_t)p * ev.Lmax + t - 1] > 0;
        if (!fork) return;
        const size_t pb = (size_t)p * ev.C, plb = (size_t)p * ev.Lmax;
        const int n_prev = ev.nlog[plb + t - 1];
        const int nc = 2 * n_prev;
        {   // candidates (state rows + metadata), hash table, parent layer (for the LEL snapshot), per-layer logs
            const uint4* s4 = reinterpret_cast<const uint4*>(ev.cand_state + pb * S); uint4* d4 = reinterpret_cast<uint4*>(ev.cand_state + cb * S);
            for (int i = tid; i < nc * (S / 2); i += NT) d4[i] = s4[i];
            for (int c = tid; c < nc; c += NT) {
                ev.cand_agg[cb + c] = ev.cand_agg[pb + c]; ev.cand_first[cb + c] = ev.cand_first[pb + c]; ev.cand_rep[cb + c] = ev.cand_rep[pb + c];
                ev.cand_inex[cb + c] = ev.cand_inex[pb + c]; ev.cand_rank[cb + c] = ev.cand_rank[pb + c]; ev.cand_slot[cb + c] = ev.cand_slot[pb + c];
                ev.uflag[cb + c] = 0;
            }
            const unsigned long long* st = ev.table + (size_t)p * ev.T; unsigned long long* dt = ev.table + (size_t)k * ev.T;
            for (int i = tid; i < ev.T; i += NT) dt[i] = st[i];
            const int pbuf = (t - 1) & 1;
            const uint4* cs4 = reinterpret_cast<const uint4*>(ev.cur_state[pbuf] + (size_t)p * ev.Wcap * S); uint4* cd4 = reinterpret_cast<uint4*>(ev.cur_state[pbuf] + (size_t)k * ev.Wcap * S);
            for (int i = tid; i < n_prev * (S / 2); i += NT) cd4[i] = cs4[i];
            for (int i = tid; i < n_prev; i += NT) {
                ev.cur_val[pbuf][(size_t)k * ev.Wcap + i] = ev.cur_val[pbuf][(size_t)p * ev.Wcap + i];
                ev.cur_rub[(size_t)k * ev.Wcap + i] = ev.cur_rub[(size_t)p * ev.Wcap + i];
            }
            for (int i = tid; i < t; i += NT) {
                ev.nlog[lb + i] = ev.nlog[plb + i]; ev.vlog[lb + i] = ev.vlog[plb + i];
                ev.rslog[(lb + i) * 2] = ev.rslog[(plb + i) * 2]; ev.rslog[(lb + i) * 2 + 1] = ev.rslog[(plb + i) * 2 + 1];
            }
        }
        __syncthreads();
        if (tid == 0) {
            ctl->status = ST_ACTIVE; ctl->n_cur = n_prev; ctl->fork_t = t;
            ctl->expanded = pc->expanded; ctl->transitions = pc->transitions;  // the shared prefix counts for both DDs
            atomicAdd(ev.active, 1);
        }
        vh_src = p;
        __syncthreads();
    }
    const int ncand = t == 0 ? 1 : 2 * ctl->n_cur;  // candidates produced by k_expand(t-1): two slots per node of layer t-1
    __syncthreads();
    if (tid == 0) { ctl->ncand = ncand; ctl->lel_pending = 0; }  // (the LEL snapshot requested by the previous step has been taken)

    // every thread owns a contiguous chunk of candidates, a multiple of 4 so that the metadata is read with 128-bit loads
    const int per = (((ncand + NT - 1) / NT) + 3) & ~3;
    const int lo = min(tid * per, ncand), hi = min(lo + per, ncand);

    // ---- B (issued first, consumed later). next_variable (misp/main.rs:109-143): vertex occurring in the fewest states, lowest
    //      index on ties; the occurrence counts were accumulated by k_expand / k_init over the distinct states.
    unsigned long long best = ~0ull;
    {
        const uint32_t* vh = ev.vhist + (size_t)vh_src * ev.HN;  // cleared by k_compact (a forking twin reads its primary's)
        for (int i = tid; i < ev.n; i += NT) {
            const unsigned c = __ldcg(vh + i);
            if (c) best = min(best, ((unsigned long long)c << 32) | (unsigned)i);
        }
    }
    // ---- A. canonical representative of every distinct state = its first candidate (rule C1) -----------------
    // claimers have cand_rep[c] == c; cand_first[c] is then the smallest candidate index with that state.  98 % of the time
    // first == c and nothing moves.  The "is canonical" flags live in the (not yet used) status array.
    uint8_t* uniq = stat;
    for (int c0 = lo; c0 < hi; c0 += 4) *reinterpret_cast<uint32_t*>(uniq + c0) = 0u;
    __syncthreads();
    for (int c0 = lo; c0 < hi; c0 += 4) {
        const uint4 r4 = *reinterpret_cast<const uint4*>(ev.cand_rep + cb + c0);
        const uint4 f4 = *reinterpret_cast<const uint4*>(ev.cand_first + cb + c0);
        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w}, ff[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = (uint32_t)(c0 + j);
            if ((int)c < hi && rr[j] == c) {
                const uint32_t f = ff[j];
                uniq[f] = 1;
                if (f != c) { ev.cand_agg[cb + f] = ev.cand_agg[cb + c]; ev.cand_inex[cb + f] = ev.cand_inex[cb + c]; }
            }
        }
    }
    __syncthreads();
    // ---- A'. ordered list of the distinct candidates ---------------------------------------------------------------
    int cnt = 0;
    uint32_t myflags[8];  // unique flags of my chunk, 4 per word (per <= 32 covers 2 x Wcap <= 32768; longer chunks re-read shared memory)
#pragma unroll
    for (int q = 0; q < 8; ++q) myflags[q] = 0;
    for (int c0 = lo, q = 0; c0 < hi; c0 += 4, ++q) {
        uint32_t fl = *reinterpret_cast<const uint32_t*>(uniq + c0);
        if (c0 + 4 > hi) fl &= (1u << (8 * (hi - c0))) - 1u;  // ignore flags beyond my chunk
        if (q < 8) myflags[q] = fl;
        cnt += __popc(fl & 0x01010101u);
    }
    int U;
    const int off0 = block_excl_scan(cnt, &U, sm.scan);  // (its barriers also order the reads of `uniq` before `stat` is reused)
    if (U == 0) {  // every node was pruned: empty layer (clean.rs:667-669) -> no best node
        if (tid == 0) { ctl->status = ST_DONE; ctl->t_term = t; ctl->has_best = 0; ctl->has_best_exact = 0; ev.nlog[lb + t] = 0; atomicSub(ev.active, 1); }
        return;
    }
    best = block_reduce(best, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, ~0ull, sm.red64);
    const bool terminal = best == ~0ull;  // next_variable == None: the layer is the terminal layer (clean.rs:350,608-632)
    const int var = terminal ? -1 : (int)(uint32_t)best;

    // ---- C. width cut ----------------------------------------------------------------------------------------
    const int W = ctl->width, comp = ctl->comp_type;
    bool cut = false; int need = 0;
    if (!terminal) {
        if (comp == DDO_RESTRICTED && U > W) { cut = true; need = W; }               // clean.rs:782-787
        else if (comp == DDO_RELAXED && U > W && t >= 2) { cut = true; need = W - 1; }  // clean.rs:788-793 (layers.len() > 1)
    }
    if (!cut && U > ev.Wcap) {
        if (tid == 0) { ctl->status = ST_DONE; ctl->overflow = 1; ctl->t_term = t; atomicSub(ev.active, 1); }
        return;
    }
    {   // my distinct candidates, in order: list them; when a cut is needed also build the 64-bit keys and reset the status bytes
        int off = off0;
        for (int c0 = lo, q = 0; c0 < hi; c0 += 4, ++q) {
            uint32_t fl;
            if (q < 8) fl = myflags[q];
            else { fl = *reinterpret_cast<const uint32_t*>(uniq + c0); if (c0 + 4 > hi) fl &= (1u << (8 * (hi - c0))) - 1u; }
            if (!fl) continue;
            unsigned long long ag[4] = {0, 0, 0, 0}; uint32_t rk[4] = {0, 0, 0, 0};
            if (cut) {
#pragma unroll
                for (int j = 0; j < 4; ++j) if ((fl >> (8 * j)) & 0xff) { ag[j] = ev.cand_agg[cb + c0 + j]; rk[j] = ev.cand_rank[cb + c0 + j]; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((fl >> (8 * j)) & 0xff) {
                ev.ulist[cb + off] = (uint32_t)(c0 + j);
                if (cut) keys[off] = (ag[j] & 0xFFFFFFFF00000000ull) | rk[j];  // (value_top, popcount, two smallest members)
                ++off;
            }
        }
    }
    // stat[ui]: 0 undecided, 1 keep, 2 drop   (long chunks, q >= 8, still read `uniq` above: keep the barrier before overwriting it)
    __syncthreads();
    if (cut) for (int ui = tid; ui < U; ui += NT) stat[ui] = 0;
    __syncthreads();
    if (cut) {
        int nactive = U;
        bool done = false;
        if (need == 0) { for (int ui = tid; ui < U; ui += NT) stat[ui] = 2; done = true; }
        if (!done) {
            // first pass over the DENSE key (value_top - vmin) * PR + (popcount - pcmin) (see finish_body_cl): one histogram pass places
            // the cut inside one (value_top, popcount) bucket; the radix passes below only order that bucket
            unsigned long long hi = 0, lo = 0;  // (max value, max popcount) and (max -value, max -popcount), biased, packed for two block folds each
            int mm[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};
            for (int ui = tid; ui < U; ui += NT) {
                const unsigned long long x = keys[ui];
                const int v = key_value(x), pc = (int)(((uint32_t)x) >> 20);
                mm[0] = max(mm[0], v); mm[1] = max(mm[1], -v); mm[2] = max(mm[2], pc); mm[3] = max(mm[3], -pc);
            }
            (void)hi; (void)lo;
            unsigned long long m4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) m4[q] = block_reduce((unsigned long long)((uint32_t)mm[q] ^ 0x80000000u), [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
            const int vmax = (int)((uint32_t)m4[0] ^ 0x80000000u), vmin = -(int)((uint32_t)m4[1] ^ 0x80000000u);
            const int pcmax = (int)((uint32_t)m4[2] ^ 0x80000000u), pcmin = -(int)((uint32_t)m4[3] ^ 0x80000000u);
            const long long VR = (long long)vmax - vmin + 1, PR = (long long)pcmax - pcmin + 1;
            if (VR * PR <= 2048) {
                const int nbins = (int)(VR * PR);
                __syncthreads();
                for (int i = tid; i < nbins; i += NT) sm.hist[i] = 0;
                __syncthreads();
                for (int ui = tid; ui < U; ui += NT) {
                    const unsigned long long x = keys[ui];
                    atomicAdd(&sm.hist[(key_value(x) - vmin) * (int)PR + ((int)(((uint32_t)x) >> 20) - pcmin)], 1u);
                }
                __syncthreads();
                {
                    constexpr int PER = 2048 / NT;
                    int c2[PER]; int s2 = 0;
#pragma unroll
                    for (int q = 0; q < PER; ++q) { const int bi = nbins - 1 - (PER * tid + q); c2[q] = bi >= 0 ? (int)sm.hist[bi] : 0; s2 += c2[q]; }
                    int tot;
                    const int before = block_excl_scan(s2, &tot, sm.scan);
                    if (before < need && need <= before + s2) {
                        int acc = before;
#pragma unroll
                        for (int q = 0; q < PER; ++q) {
                            if (acc < need && need <= acc + c2[q]) { sm.misc[0] = nbins - 1 - (PER * tid + q); sm.misc[1] = acc; sm.misc[2] = c2[q]; }
                            acc += c2[q];
                        }
                    }
                    __syncthreads();
                }
                const int b = sm.misc[0], above = sm.misc[1], inb = sm.misc[2];
                need -= above; nactive = inb;
                const bool all_keep = need == inb;
                for (int ui = tid; ui < U; ui += NT) {
                    const unsigned long long x = keys[ui];
                    const int bin = (key_value(x) - vmin) * (int)PR + ((int)(((uint32_t)x) >> 20) - pcmin);
                    if (bin > b) stat[ui] = 1; else if (bin < b) stat[ui] = 2; else if (all_keep) stat[ui] = 1;
                }
                __syncthreads();
                if (all_keep) done = true;
            }
        }
        for (int chunk = 0; chunk <= MemberKey<S>::CHUNKS && !done; ++chunk) {
            // key chunk 0: (value_top, popcount, two smallest members); chunk j >= 1: the next MK members of the state (member_key)
            if (chunk > 0) {
                for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 0) keys[ui] = member_key<S>(ev.cand_state + (cb + ev.ulist[cb + ui]) * S, MemberKey<S>::MK * (chunk - 1));
                __syncthreads();
            }
            auto key_of = [&](int ui) -> unsigned long long { return keys[ui]; };
            unsigned long long kor = 0, kand = ~0ull;
            for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 0) { unsigned long long x = key_of(ui); kor |= x; kand &= x; }
            kor = block_reduce(kor, [](unsigned long long a, unsigned long long b) { return a | b; }, 0ull, sm.red64);
            kand = block_reduce(kand, [](unsigned long long a, unsigned long long b) { return a & b; }, ~0ull, sm.red64);
            const unsigned long long diff = kor ^ kand;
            for (int byte = 7; byte >= 0 && !done; --byte) {
                if (((diff >> (8 * byte)) & 0xff) == 0) continue;  // every undecided key has the same digit here
                for (int i = tid; i < 256; i += NT) sm.hist[i] = 0;
                __syncthreads();
                for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 0) atomicAdd(&sm.hist[(key_of(ui) >> (8 * byte)) & 0xff], 1u);
                __syncthreads();
                if (warp == 0) {  // bucket b with  #(digit > b) < need <= #(digit >= b)
                    int c8[8]; int s8 = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { c8[q] = (int)sm.hist[255 - (lane * 8 + q)]; s8 += c8[q]; }  // lane 0 owns the 8 largest digits
                    int inc = s8;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { int nn = __shfl_up_sync(FULL_MASK, inc, d); if (lane >= d) inc += nn; }
                    int before = inc - s8;  // count of digits strictly greater than this lane's 8 buckets
                    if (before < need && need <= inc) {
                        int acc = before;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (acc < need && need <= acc + c8[q]) { sm.misc[0] = 255 - (lane * 8 + q); sm.misc[1] = acc; sm.misc[2] = c8[q]; }
                            acc += c8[q];
                        }
                    }
                }
                __syncthreads();
                const int b = sm.misc[0], above = sm.misc[1], inb = sm.misc[2];
                need -= above; nactive = inb;
                const bool all_keep = (need == nactive);
                for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 0) {
                    const int d = (int)((key_of(ui) >> (8 * byte)) & 0xff);
                    if (d > b) stat[ui] = 1; else if (d < b) stat[ui] = 2; else if (all_keep) stat[ui] = 1;
                }
                __syncthreads();
                if (all_keep) done = true;
            }
        }
        __syncthreads();
    }

    // ---- D. stable positions of the survivors (rule C3): every thread places the distinct candidates of its own chunk ----
    int nkeep, kp;
    if (cut) {
        int kc = 0;
        for (int i = 0; i < cnt; ++i) kc += (stat[off0 + i] == 1);
        kp = block_excl_scan(kc, &nkeep, sm.scan);
        for (int i = 0; i < cnt; ++i) {
            const uint32_t c = ev.ulist[cb + off0 + i];
            if (stat[off0 + i] == 1) { ev.pos_of[cb + c] = (uint32_t)kp++; ev.uflag[cb + c] = 2; } else ev.pos_of[cb + c] = NONE32;
        }
    } else {  // no cut: every distinct candidate survives at its rank in the ordered list
        nkeep = U; kp = off0;
        for (int c0 = lo, q = 0; c0 < hi; c0 += 4, ++q) {
            uint32_t fl;
            if (q < 8) fl = myflags[q];
            else { fl = *reinterpret_cast<const uint32_t*>(uniq + c0); if (c0 + 4 > hi) fl &= (1u << (8 * (hi - c0))) - 1u; }  // `uniq` is intact: no cut
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((fl >> (8 * j)) & 0xff) { ev.pos_of[cb + c0 + j] = (uint32_t)kp++; ev.uflag[cb + c0 + j] = 2; }
        }
    }
    int n_next = nkeep;
    int s_pos = -1, r_pos = -1;
    __syncthreads();

    // ---- E. relaxation: merge the overflow (clean.rs:826-876; misp/main.rs:172-178 union) ------------------------
    if (cut && comp == DDO_RELAXED) {
        if (tid < 32) sm.merged[tid] = 0;
        __syncthreads();
        uint64_t acc[S];
#pragma unroll
        for (int j = 0; j < S; ++j) acc[j] = 0;
        unsigned long long mkey = 0;
        for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 2) {
            const uint32_t c = ev.ulist[cb + ui];
            const uint4* sp = reinterpret_cast<const uint4*>(ev.cand_state + (cb + c) * S);
#pragma unroll
            for (int j = 0; j < S / 2; ++j) { const uint4 v4 = sp[j]; acc[2 * j] |= u4lo(v4); acc[2 * j + 1] |= u4hi(v4); }
            mkey = max(mkey, ev.cand_agg[cb + c]);
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            uint64_t x = warp_reduce(acc[j], [](uint64_t a, uint64_t b) { return a | b; });
            if (lane == 0 && x) atomicOr(&sm.merged[j], (unsigned long long)x);
        }
        mkey = block_reduce(mkey, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
        __syncthreads();
        // recycled ? (clean.rs:830): a KEPT node whose state equals the merged state
        if (tid == 0) {
            uint64_t h = 0;
            for (int j = 0; j < S; ++j) h += sm.merged[j] * hash_mul(j);
            h = mix64(h);
            const uint32_t tag = (uint32_t)(h >> 32);
            uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
            const unsigned long long* tab = ev.table + (size_t)k * ev.T;
            int recycled = -1;
            for (;;) {
                const unsigned long long e = tab[slot];
                if (e == EMPTY64) break;
                if ((uint32_t)(e >> 32) == tag) {
                    const uint32_t oc = (uint32_t)e;
                    bool eq = true;
                    for (int j = 0; j < S; ++j) eq = eq && ev.cand_state[(cb + oc) * S + j] == sm.merged[j];
                    if (eq) { const uint32_t f = ev.cand_first[cb + oc]; if (ev.pos_of[cb + f] != NONE32) recycled = (int)f; break; }
                }
                slot = (slot + 1) & (uint32_t)(ev.T - 1);
            }
            sm.misc[4] = recycled;
        }
        __syncthreads();
        const int recycled = sm.misc[4];
        const int mpos = (recycled >= 0) ? (int)ev.pos_of[cb + recycled] : nkeep;
        if (recycled >= 0) {
            // clean.rs:868-871: the best merged-away node ("saved") stays in the layer, un-deleted, next to the recycled node
            uint32_t bestc = NONE32;
            for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 2) {
                const uint32_t c = ev.ulist[cb + ui];
                if (bestc == NONE32 || cand_better<S>(ev, cb, c, bestc)) bestc = c;
            }
            __shared__ uint32_t s_best[NT];
            s_best[tid] = bestc;
            __syncthreads();
            for (int d = NT / 2; d > 0; d >>= 1) {
                if (tid < d) {
                    const uint32_t a = s_best[tid], b2 = s_best[tid + d];
                    if (a == NONE32 || (b2 != NONE32 && cand_better<S>(ev, cb, b2, a))) s_best[tid] = b2;
                }
                __syncthreads();
            }
            const uint32_t saved = s_best[0];
            s_pos = nkeep; r_pos = mpos; n_next = nkeep + 1;
            if (tid == 0) {
                // the recycled node receives every relaxed edge: RELAXED flag, value_top = max (`>=`: the appended edges win ties)
                const unsigned long long rk = ev.cand_agg[cb + recycled];
                if (key_value(mkey) >= key_value(rk)) ev.cand_agg[cb + recycled] = mkey;
                ev.cand_inex[cb + recycled] |= (uint8_t)(NF_INEXACT | NF_RELAXED);
            }
            __syncthreads();
            for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 2) {
                const uint32_t c = ev.ulist[cb + ui];
                if (c == saved) { ev.pos_of[cb + c] = (uint32_t)s_pos; stat[ui] = 1; ev.uflag[cb + c] = 2; }
                else ev.pos_of[cb + c] = (uint32_t)r_pos;
            }
        } else {
            n_next = nkeep + 1;
            for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 2) ev.pos_of[cb + ev.ulist[cb + ui]] = (uint32_t)mpos;
            // new merged node (clean.rs:832-849) written straight into the next layer
            const int nbuf = t & 1;
            const size_t nb = (size_t)k * ev.Wcap + mpos;
            if (tid < S) ev.cur_state[nbuf][nb * S + tid] = sm.merged[tid];
            if (tid == 0) {
                ev.cur_val[nbuf][nb] = key_value(mkey);
                ev.cur_flag[nbuf][nb] = (uint8_t)(NF_INEXACT | NF_RELAXED);
                ev.plog[(lb + t) * ev.Wcap + mpos] = ((uint32_t)mkey & PLOG_CAND_MASK) | PLOG_INEXACT | PLOG_RELAXED;
            }
        }
    }

    // ---- F. terminal layer: best nodes (clean.rs:620-632, rule C4: last maximum) ----------------------------------
    if (terminal) {
        unsigned long long b_all = 0, b_ex = 0;  // (biased value, pos + 1)
        for (int ui = tid; ui < U; ui += NT) {
            const uint32_t c = ev.ulist[cb + ui];
            const unsigned long long kk = (ev.cand_agg[cb + c] & 0xFFFFFFFF00000000ull) | (unsigned)(ev.pos_of[cb + c] + 1);
            b_all = max(b_all, kk);
            if (!(ev.cand_inex[cb + c] & (NF_INEXACT | NF_RELAXED))) b_ex = max(b_ex, kk);
        }
        b_all = block_reduce(b_all, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
        b_ex = block_reduce(b_ex, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
        if (tid == 0) {
            ctl->has_best = 1; ctl->best_value = key_value(b_all); ctl->best_pos = (int)(uint32_t)b_all - 1;
            ctl->has_best_exact = b_ex != 0;
            if (b_ex) { ctl->best_exact_value = key_value(b_ex); ctl->best_exact_pos = (int)(uint32_t)b_ex - 1; }
        }
    }

    if (tid == 0) {
        ev.nlog[lb + t] = n_next;
        ev.vlog[lb + t] = var;
        ev.rslog[(lb + t) * 2] = s_pos; ev.rslog[(lb + t) * 2 + 1] = r_pos;
        ctl->n_cur = n_next; ctl->var = var;
        if (cut && ctl->lel < 0) { ctl->lel = t - 1; ctl->lel_pending = 1; }  // _maybe_save_lel, clean.rs:796-800
        if (terminal) { ctl->status = ST_TERMINAL; ctl->t_term = t; atomicSub(ev.active, 1); }
    }
}

// work plan of the two flat kernels that follow a finish: the last CTA to finish scans the per-DD tile counts
template <int S>
__device__ void finish_plan(const EV& ev, FinishSmem& sm, int* s_last, int count) {  // count = DD slots of the batch
    constexpr int G = S / 2, PER_TILE = 256 / G;
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid == 0) { __threadfence(); *s_last = (atomicAdd(ev.finish_counter, 1u) == gridDim.x - 1u); }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    const volatile DDCtl* vc = ev.ctl;
    const int per = (count + blockDim.x - 1) / blockDim.x;
    const int lo = min(tid * per, count), hi = min(lo + per, count);
    int te = 0, tc = 0;
    for (int k = lo; k < hi; ++k) {
        const int st = vc[k].status;
        te += st == ST_ACTIVE ? (vc[k].n_cur + PER_TILE - 1) / PER_TILE : 0;
        tc += (st == ST_ACTIVE || st == ST_TERMINAL) ? (vc[k].ncand + PER_TILE - 1) / PER_TILE : 0;
    }
    int tote, totc;
    int oe = block_excl_scan(te, &tote, sm.scan);
    int oc = block_excl_scan(tc, &totc, sm.scan);
    for (int k = lo; k < hi; ++k) {
        const int st = vc[k].status;
        ev.tile_off_e[k] = oe; ev.tile_off_c[k] = oc;
        oe += st == ST_ACTIVE ? (vc[k].n_cur + PER_TILE - 1) / PER_TILE : 0;
        oc += (st == ST_ACTIVE || st == ST_TERMINAL) ? (vc[k].ncand + PER_TILE - 1) / PER_TILE : 0;
    }
    // the finish of the NEXT layer: its candidates are two per node of the layer decided now (a waiting twin forks with its primary's)
    int nb = 0, ns = 0;
    for (int k = lo; k < hi; ++k) {
        const int st = vc[k].status;
        if (st == ST_DONE) continue;
        const int src = st == ST_WAITING ? vc[k].primary : k;
        if (st == ST_WAITING && (src < 0 || vc[src].status == ST_DONE || vc[src].status == ST_TERMINAL)) continue;  // its primary ended without a cut: the twin never forks
        (st != ST_TERMINAL && 2 * vc[src].n_cur > FIN_SMALL_C) ? ++nb : ++ns;
    }
    int totb, tots;
    int ob = block_excl_scan(nb, &totb, sm.scan);
    int os = block_excl_scan(ns, &tots, sm.scan);
    for (int k = lo; k < hi; ++k) {
        const int st = vc[k].status;
        if (st == ST_DONE) continue;
        const int src = st == ST_WAITING ? vc[k].primary : k;
        if (st == ST_WAITING && (src < 0 || vc[src].status == ST_DONE || vc[src].status == ST_TERMINAL)) continue;
        if (st != ST_TERMINAL && 2 * vc[src].n_cur > FIN_SMALL_C) ev.fin_list[ob++] = k; else ev.fin_list[ev.K + os++] = k;
    }
    if (tid == 0) { ev.tile_off_e[count] = tote; ev.tile_off_c[count] = totc; *ev.finish_counter = 0; ev.fin_cnt[0] = totb; ev.fin_cnt[1] = tots; }
}

// size_class 0: every DD of the batch (blockIdx.x = DD slot), followed by the work plan; 1: the listed wide DDs, no plan (k_finish_s runs
// next and writes it)
template <int S>
__global__ void __launch_bounds__(1024, 1) k_finish(EV ev, int t, int size_class) {
    pdl_enter();
    __shared__ FinishSmem sm;
    __shared__ int s_last;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    {
        unsigned long long* keys; uint8_t* stat;
        if (ev.smem_keys) { keys = reinterpret_cast<unsigned long long*>(dyn_smem); stat = dyn_smem + (size_t)ev.C * 8; }
        else { keys = ev.gkeys + (size_t)blockIdx.x * ev.C; stat = ev.ustat + (size_t)blockIdx.x * ev.C; }
        if (size_class == 0) finish_body<S, 1024>(ev, t, sm, keys, stat, blockIdx.x);
        else {
            const int nb = ev.fin_cnt[0];
            for (int j = blockIdx.x; j < nb; j += gridDim.x) {
                const int k = ev.fin_list[j];
                if (!ev.smem_keys) { keys = ev.gkeys + (size_t)k * ev.C; stat = ev.ustat + (size_t)k * ev.C; }
                finish_body<S, 1024>(ev, t, sm, keys, stat, k);
                __syncthreads();
            }
        }
    }
    if (size_class == 0) finish_plan<S>(ev, sm, &s_last, gridDim.x);
}

// k_finish_s: the finish of the NARROW DDs of a large batch.  k_finish holds the cut keys of up to 2 * Wcap candidates in shared memory
// and runs 1024 threads, so one DD occupies an SM; a batch of a thousand DDs whose layers hold a few hundred nodes then takes seven
// rounds of CTAs per layer step for work a quarter of a CTA could do.  Here a DD with at most FIN_SMALL_C candidates gets 256 threads
// and 18 KB of keys: five to six DDs per SM.
#ifndef DDO_FIN_S_NT
#define DDO_FIN_S_NT 256
#endif
constexpr int FIN_S_NT = DDO_FIN_S_NT;  // threads of a k_finish_s CTA (128 threads, seven CTAs per SM, measured slower: k_finish 246 vs 224 ms per config-2 solve)
template <int S>
__global__ void __launch_bounds__(FIN_S_NT) k_finish_s(EV ev, int t, int slots) {
    pdl_enter();
    __shared__ FinishSmem sm;
    __shared__ int s_last;
    __shared__ __align__(16) unsigned long long s_keys[FIN_SMALL_C];
    __shared__ __align__(16) uint8_t s_stat[FIN_SMALL_C];
    const int ns = ev.fin_cnt[1];
    for (int j = blockIdx.x; j < ns; j += gridDim.x) {
        finish_body<S, FIN_S_NT>(ev, t, sm, s_keys, s_stat, ev.fin_list[ev.K + j]);
        __syncthreads();
    }
    finish_plan<S>(ev, sm, &s_last, slots);
}

// =================================================================================================================
// k_finish_cl: the same decisions as k_finish taken by a thread-block CLUSTER per DD (FCL_CS CTAs of FCL_NT threads), for narrow
// batches, where one CTA per DD leaves the machine idle and its ~45 us chain of block barriers and L2 round trips is the latency of
// a layer step.  Every CTA owns a contiguous slice of the candidates (keys and status bytes of its distinct candidates in its own shared
// memory); counts, offsets, digit histograms, the merged state and maxima are exchanged through distributed shared memory between
// cluster barriers.  All control decisions are cluster-uniform (every CTA executes the same sequence of cluster barriers).
// =================================================================================================================
namespace cg = cooperative_groups;
constexpr int FCL_CS = 8;
constexpr int FCL_NT = 256;

constexpr int FCL_DB = 2048;  // bins of the dense (value_top, popcount) histogram of the width cut

struct FinishClSmem {
    int scan[40];
    unsigned long long red64[40];
    unsigned int hist[256];           // local digit histogram of one radix pass
    unsigned int ghist[2][256];       // cluster-wide digit histogram: every CTA adds its non-zero bins into every peer (remote atomics), double-buffered
    unsigned int dh[FCL_DB];          // local / cluster-wide histogram of the dense (value_top, popcount) key of the first pass
    unsigned int dgh[FCL_DB];
    unsigned long long merged[32];
    unsigned long long xch[2][4][FCL_CS];  // exchange slots [phase][value][source rank]: PUSHED into every peer before a cluster barrier, read locally after it
    int misc[16];
    int red4[FCL_NT / 32][4];
};

// Exchanges between the CTAs of the cluster.  A value is pushed (remote store, fire and forget) into every peer's shared memory before the
// barrier and read locally after it: a pulled value costs one ~215-cycle DSMEM round trip PER PEER, serialised in the fold loop (round 1).
struct ClusterXchg {
    cg::cluster_group cl; FinishClSmem* sm; int phase; int hphase; unsigned rank;
    // fold of up to 4 block-uniform values over the CTAs of the cluster; result uniform in the whole cluster
    template <class Op> __device__ void allreduce(unsigned long long* v, int n, Op op) {
        if ((int)threadIdx.x < n * FCL_CS) { const int i = threadIdx.x / FCL_CS; *cl.map_shared_rank(&sm->xch[phase][i][rank], threadIdx.x % FCL_CS) = v[i]; }
        cl.sync();
        for (int i = 0; i < n; ++i) {
            unsigned long long acc = sm->xch[phase][i][0];
            for (unsigned r = 1; r < FCL_CS; ++r) acc = op(acc, sm->xch[phase][i][r]);
            v[i] = acc;
        }
        phase ^= 1;
    }
    // exclusive prefix over the ranks of one block-uniform count; *total = sum over the cluster
    __device__ int scan(int v, int* total) {
        if (threadIdx.x < FCL_CS) *cl.map_shared_rank(&sm->xch[phase][0][rank], threadIdx.x) = (unsigned long long)(unsigned)v;
        cl.sync();
        int pre = 0, tot = 0;
        for (unsigned r = 0; r < FCL_CS; ++r) { const int x = (int)sm->xch[phase][0][r]; if (r < rank) pre += x; tot += x; }
        phase ^= 1;
        *total = tot;
        return pre;
    }
};

template <int S>
__device__ void finish_body_cl(const EV& ev, int t, FinishClSmem& sm, unsigned long long* keys, uint8_t* stat, int kcap) {
    constexpr int NT = FCL_NT;
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank();
    ClusterXchg X{cl, &sm, 0, 0, rank};
    const int k = blockIdx.x / FCL_CS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 512; i += FCL_NT) (&sm.ghist[0][0])[i] = 0;   // (the peers add into them only after the cluster barriers of step A)
    for (int i = tid; i < FCL_DB; i += FCL_NT) sm.dgh[i] = 0;
    const int gt = (int)rank * NT + tid;  // thread index in the cluster: candidate chunks are contiguous in this order
    DDCtl* ctl = ev.ctl + k;
    const int status = ctl->status;
    if (status == ST_DONE) return;
    if (status == ST_TERMINAL) { cl.sync(); if (rank == 0 && tid == 0) ctl->status = ST_DONE; return; }
    const size_t cb = (size_t)k * ev.C;
    const size_t lb = (size_t)k * ev.Lmax;
    int vh_src = k;
    if (status == ST_WAITING) {  // relaxed twin: fork at the primary's first cut (see finish_body)
        const int p = ctl->primary;
        const DDCtl* pc = ev.ctl + p;
        const int plel = pc->lel;
        const bool fork = t >= 1 && ev.ucount[p] > (uint32_t)pc->width && (plel < 0 || plel == t - 1) && ev.nlog[(size_t)p * ev.Lmax + t - 1] > 0;
        if (!fork) return;
        const size_t pb = (size_t)p * ev.C, plb = (size_t)p * ev.Lmax;
        const int n_prev = ev.nlog[plb + t - 1];
        const int nc = 2 * n_prev;
        constexpr int GN = FCL_CS * NT;
        {
            const uint4* s4 = reinterpret_cast<const uint4*>(ev.cand_state + pb * S); uint4* d4 = reinterpret_cast<uint4*>(ev.cand_state + cb * S);
            for (int i = gt; i < nc * (S / 2); i += GN) d4[i] = s4[i];
            for (int c = gt; c < nc; c += GN) {
                ev.cand_agg[cb + c] = ev.cand_agg[pb + c]; ev.cand_first[cb + c] = ev.cand_first[pb + c]; ev.cand_rep[cb + c] = ev.cand_rep[pb + c];
                ev.cand_inex[cb + c] = ev.cand_inex[pb + c]; ev.cand_rank[cb + c] = ev.cand_rank[pb + c]; ev.cand_slot[cb + c] = ev.cand_slot[pb + c];
                ev.uflag[cb + c] = 0;
            }
            const unsigned long long* st = ev.table + (size_t)p * ev.T; unsigned long long* dt = ev.table + (size_t)k * ev.T;
            for (int i = gt; i < ev.T; i += GN) dt[i] = st[i];
            const int pbuf = (t - 1) & 1;
            const uint4* cs4 = reinterpret_cast<const uint4*>(ev.cur_state[pbuf] + (size_t)p * ev.Wcap * S); uint4* cd4 = reinterpret_cast<uint4*>(ev.cur_state[pbuf] + (size_t)k * ev.Wcap * S);
            for (int i = gt; i < n_prev * (S / 2); i += GN) cd4[i] = cs4[i];
            for (int i = gt; i < n_prev; i += GN) {
                ev.cur_val[pbuf][(size_t)k * ev.Wcap + i] = ev.cur_val[pbuf][(size_t)p * ev.Wcap + i];
                ev.cur_rub[(size_t)k * ev.Wcap + i] = ev.cur_rub[(size_t)p * ev.Wcap + i];
            }
            for (int i = gt; i < t; i += GN) {
                ev.nlog[lb + i] = ev.nlog[plb + i]; ev.vlog[lb + i] = ev.vlog[plb + i];
                ev.rslog[(lb + i) * 2] = ev.rslog[(plb + i) * 2]; ev.rslog[(lb + i) * 2 + 1] = ev.rslog[(plb + i) * 2 + 1];
            }
        }
        __threadfence();
        cl.sync();
        if (rank == 0 && tid == 0) {
            ctl->status = ST_ACTIVE; ctl->fork_t = t;
            ctl->expanded = pc->expanded; ctl->transitions = pc->transitions;
            atomicAdd(ev.active, 1);
        }
        // n_cur is needed by every CTA below: take it from the copy source instead of ctl (written by rank 0 only)
        vh_src = p;
    }
    const int n_cur_here = status == ST_WAITING ? ev.nlog[(size_t)ctl->primary * ev.Lmax + t - 1] : ctl->n_cur;
    const int ncand = t == 0 ? 1 : 2 * n_cur_here;
    if (rank == 0 && tid == 0) { if (status == ST_WAITING) ctl->n_cur = n_cur_here; ctl->ncand = ncand; ctl->lel_pending = 0; }

    const int per = (((ncand + FCL_CS * NT - 1) / (FCL_CS * NT)) + 3) & ~3;
    const int lo = min(gt * per, ncand), hi = min(lo + per, ncand);

    // ---- B. next_variable: argmin of the vertex histogram (every CTA computes it: n <= 1024 L2 loads, no exchange needed) -----------
    unsigned long long best = ~0ull;
    {
        const uint32_t* vh = ev.vhist + (size_t)vh_src * ev.HN;
        for (int i = tid; i < ev.n; i += NT) {
            const unsigned c = __ldcg(vh + i);
            if (c) best = min(best, ((unsigned long long)c << 32) | (unsigned)i);
        }
    }
    // ---- A. canonical representatives: flags in global memory (a first candidate may belong to another CTA's slice) ------------------
    uint8_t* uniq = ev.gflag + cb;
    for (int c0 = lo; c0 < hi; c0 += 4) *reinterpret_cast<uint32_t*>(uniq + c0) = 0u;
    __threadfence();
    cl.sync();
    for (int c0 = lo; c0 < hi; c0 += 4) {
        const uint4 r4 = *reinterpret_cast<const uint4*>(ev.cand_rep + cb + c0);
        const uint4 f4 = *reinterpret_cast<const uint4*>(ev.cand_first + cb + c0);
        const uint32_t rr[4] = {r4.x, r4.y, r4.z, r4.w}, ff[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t c = (uint32_t)(c0 + j);
            if ((int)c < hi && rr[j] == c) {
                const uint32_t f = ff[j];
                uniq[f] = 1;
                if (f != c) { ev.cand_agg[cb + f] = ev.cand_agg[cb + c]; ev.cand_inex[cb + f] = ev.cand_inex[cb + c]; }
            }
        }
    }
    __threadfence();
    cl.sync();
    // ---- A'. ordered list of the distinct candidates -------------------------------------------------------------------------------
    int cnt = 0;
    for (int c0 = lo; c0 < hi; c0 += 4) {
        uint32_t fl = __ldcg(reinterpret_cast<const uint32_t*>(uniq + c0));
        if (c0 + 4 > hi) fl &= (1u << (8 * (hi - c0))) - 1u;
        cnt += __popc(fl & 0x01010101u);
    }
    int Ublk;
    const int offb = block_excl_scan(cnt, &Ublk, sm.scan);
    int U;
    const int cta_off = X.scan(Ublk, &U);   // first distinct candidate of this CTA in the DD-wide ordered list
    const int off0 = cta_off + offb;
    if (U == 0) {
        if (rank == 0 && tid == 0) { ctl->status = ST_DONE; ctl->t_term = t; ctl->has_best = 0; ctl->has_best_exact = 0; ev.nlog[lb + t] = 0; atomicSub(ev.active, 1); }
        return;
    }
    best = block_reduce(best, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, ~0ull, sm.red64);
    const bool terminal = best == ~0ull;
    const int var = terminal ? -1 : (int)(uint32_t)best;

    // ---- C. width cut --------------------------------------------------------------------------------------------------------------
    const int W = ctl->width, comp = ctl->comp_type;
    bool cut = false; int need = 0;
    if (!terminal) {
        if (comp == DDO_RESTRICTED && U > W) { cut = true; need = W; }
        else if (comp == DDO_RELAXED && U > W && t >= 2) { cut = true; need = W - 1; }
    }
    if (!cut && U > ev.Wcap) {
        if (rank == 0 && tid == 0) { ctl->status = ST_DONE; ctl->overflow = 1; ctl->t_term = t; atomicSub(ev.active, 1); }
        return;
    }
    // local index of a distinct candidate of this CTA: li = ui - cta_off, 0 <= li < Ublk (<= kcap)
    {
        int off = off0;
        for (int c0 = lo; c0 < hi; c0 += 4) {
            uint32_t fl = __ldcg(reinterpret_cast<const uint32_t*>(uniq + c0));
            if (c0 + 4 > hi) fl &= (1u << (8 * (hi - c0))) - 1u;
            if (!fl) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((fl >> (8 * j)) & 0xff) {
                ev.ulist[cb + off] = (uint32_t)(c0 + j);
                if (cut) keys[off - cta_off] = (ev.cand_agg[cb + c0 + j] & 0xFFFFFFFF00000000ull) | ev.cand_rank[cb + c0 + j];
                ++off;
            }
        }
    }
    __syncthreads();
    if (cut) for (int li = tid; li < Ublk; li += NT) stat[li] = 0;
    __syncthreads();
    if (cut) {
        bool done = false;
        if (need == 0) { for (int li = tid; li < Ublk; li += NT) stat[li] = 2; done = true; }
        if (!done) {
            // ---- first pass over a DENSE key.  value_top and the popcount span a few dozen values each in one layer, so the pair
            // (value_top - vmin) * PR + (popcount - pcmin) usually fits FCL_DB bins: ONE histogram pass (summed into every CTA by remote
            // atomics, one cluster barrier) places the cut inside one (value_top, popcount) bucket, where round 1 walked four to five
            // 8-bit digits of the 64-bit key with a barrier and eight serialised DSMEM reads per bin each.
            int mm[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};  // max value, max -value, max popcount, max -popcount
            for (int li = tid; li < Ublk; li += NT) {
                const unsigned long long x = keys[li];
                const int v = key_value(x), pc = (int)(((uint32_t)x) >> 20);
                mm[0] = max(mm[0], v); mm[1] = max(mm[1], -v); mm[2] = max(mm[2], pc); mm[3] = max(mm[3], -pc);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) mm[q] = warp_reduce(mm[q], [](int a, int b) { return a > b ? a : b; });
            __syncthreads();
            if (lane == 0) for (int q = 0; q < 4; ++q) sm.red4[warp][q] = mm[q];
            __syncthreads();
            unsigned long long m4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { int x = INT32_MIN; for (int w2 = 0; w2 < NT / 32; ++w2) x = max(x, sm.red4[w2][q]); m4[q] = (unsigned long long)((uint32_t)x ^ 0x80000000u); }
            X.allreduce(m4, 4, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
            const int vmax = (int)((uint32_t)m4[0] ^ 0x80000000u), vmin = -(int)((uint32_t)m4[1] ^ 0x80000000u);
            const int pcmax = (int)((uint32_t)m4[2] ^ 0x80000000u), pcmin = -(int)((uint32_t)m4[3] ^ 0x80000000u);
            const long long VR = (long long)vmax - vmin + 1, PR = (long long)pcmax - pcmin + 1;
            if (VR * PR <= FCL_DB) {
                const int nbins = (int)(VR * PR);
                for (int i = tid; i < nbins; i += NT) sm.dh[i] = 0;
                __syncthreads();
                for (int li = tid; li < Ublk; li += NT) {
                    const unsigned long long x = keys[li];
                    atomicAdd(&sm.dh[(key_value(x) - vmin) * (int)PR + ((int)(((uint32_t)x) >> 20) - pcmin)], 1u);
                }
                __syncthreads();
                for (int i = tid; i < nbins; i += NT) {
                    const unsigned v = sm.dh[i];
                    if (v) for (unsigned r = 0; r < FCL_CS; ++r) atomicAdd(cl.map_shared_rank(&sm.dgh[i], r), v);
                }
                cl.sync();
                {   // bucket b with  #(bin > b) < need <= #(bin >= b): every thread owns FCL_DB / NT consecutive bins, largest first
                    constexpr int PER = FCL_DB / NT;
                    int c8[PER]; int s8 = 0;
#pragma unroll
                    for (int q = 0; q < PER; ++q) { const int bi = nbins - 1 - (PER * tid + q); c8[q] = bi >= 0 ? (int)sm.dgh[bi] : 0; s8 += c8[q]; }
                    int tot;
                    const int before = block_excl_scan(s8, &tot, sm.scan);
                    if (before < need && need <= before + s8) {
                        int acc = before;
#pragma unroll
                        for (int q = 0; q < PER; ++q) {
                            if (acc < need && need <= acc + c8[q]) { sm.misc[0] = nbins - 1 - (PER * tid + q); sm.misc[1] = acc; sm.misc[2] = c8[q]; }
                            acc += c8[q];
                        }
                    }
                    __syncthreads();
                }
                const int b = sm.misc[0], above = sm.misc[1], inb = sm.misc[2];
                need -= above;
                const bool all_keep = need == inb;
                for (int li = tid; li < Ublk; li += NT) {
                    const unsigned long long x = keys[li];
                    const int bin = (key_value(x) - vmin) * (int)PR + ((int)(((uint32_t)x) >> 20) - pcmin);
                    if (bin > b) stat[li] = 1; else if (bin < b) stat[li] = 2; else if (all_keep) stat[li] = 1;
                }
                __syncthreads();
                if (all_keep) done = true;
            }
        }
        // ---- the rest of the order (the two smallest members in the key's low bits, then BitSet::cmp over member-index chunks): MSD radix
        // select over the still undecided candidates, one cluster barrier per digit
        for (int chunk = 0; chunk <= MemberKey<S>::CHUNKS && !done; ++chunk) {
            if (chunk > 0) {  // the next MK members of every still undecided state (member_key)
                for (int li = tid; li < Ublk; li += NT) if (stat[li] == 0) keys[li] = member_key<S>(ev.cand_state + (cb + ev.ulist[cb + cta_off + li]) * S, MemberKey<S>::MK * (chunk - 1));
                __syncthreads();
            }
            auto key_of = [&](int li) -> unsigned long long { return keys[li]; };
            unsigned long long kk[2] = {0ull, 0ull};  // OR of the undecided keys, OR of their complements (AND = ~kk[1]; one fold serves both)
            for (int li = tid; li < Ublk; li += NT) if (stat[li] == 0) { unsigned long long x = key_of(li); kk[0] |= x; kk[1] |= ~x; }
            kk[0] = block_reduce(kk[0], [](unsigned long long a, unsigned long long b) { return a | b; }, 0ull, sm.red64);
            kk[1] = block_reduce(kk[1], [](unsigned long long a, unsigned long long b) { return a | b; }, 0ull, sm.red64);
            X.allreduce(kk, 2, [](unsigned long long a, unsigned long long b) { return a | b; });
            const unsigned long long diff = kk[0] ^ ~kk[1];
            for (int byte = 7; byte >= 0 && !done; --byte) {
                if (((diff >> (8 * byte)) & 0xff) == 0) continue;
                unsigned int* h = sm.hist;
                unsigned int* gh = sm.ghist[X.hphase];
                for (int i = tid; i < 256; i += NT) h[i] = 0;
                __syncthreads();
                for (int li = tid; li < Ublk; li += NT) if (stat[li] == 0) atomicAdd(&h[(key_of(li) >> (8 * byte)) & 0xff], 1u);
                __syncthreads();
                for (int i = tid; i < 256; i += NT) {
                    const unsigned v = h[i];
                    if (v) for (unsigned r = 0; r < FCL_CS; ++r) atomicAdd(cl.map_shared_rank(&gh[i], r), v);
                }
                cl.sync();
                X.hphase ^= 1;
                if (warp == 0) {
                    int c8[8]; int s8 = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { c8[q] = (int)gh[255 - (lane * 8 + q)]; s8 += c8[q]; }
                    int inc = s8;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { int nn = __shfl_up_sync(FULL_MASK, inc, d); if (lane >= d) inc += nn; }
                    int before = inc - s8;
                    if (before < need && need <= inc) {
                        int acc = before;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (acc < need && need <= acc + c8[q]) { sm.misc[0] = 255 - (lane * 8 + q); sm.misc[1] = acc; sm.misc[2] = c8[q]; }
                            acc += c8[q];
                        }
                    }
                }
                __syncthreads();
                for (int i = tid; i < 256; i += NT) gh[i] = 0;  // (the peers add into this buffer again two barriers from now at the earliest)
                const int b = sm.misc[0], above = sm.misc[1], inb = sm.misc[2];
                need -= above;
                const bool all_keep = (need == inb);
                for (int li = tid; li < Ublk; li += NT) if (stat[li] == 0) {
                    const int d = (int)((key_of(li) >> (8 * byte)) & 0xff);
                    if (d > b) stat[li] = 1; else if (d < b) stat[li] = 2; else if (all_keep) stat[li] = 1;
                }
                __syncthreads();
                if (all_keep) done = true;
            }
        }
        __syncthreads();
    }

    // ---- D. stable positions of the survivors ----------------------------------------------------------------------------------------
    int nkeep, kp;
    if (cut) {
        int kc = 0;
        for (int i = 0; i < cnt; ++i) kc += (stat[off0 - cta_off + i] == 1);
        int kblk;
        const int kpb = block_excl_scan(kc, &kblk, sm.scan);
        kp = X.scan(kblk, &nkeep) + kpb;
        for (int i = 0; i < cnt; ++i) {
            const uint32_t c = ev.ulist[cb + off0 + i];
            if (stat[off0 - cta_off + i] == 1) { ev.pos_of[cb + c] = (uint32_t)kp++; ev.uflag[cb + c] = 2; } else ev.pos_of[cb + c] = NONE32;
        }
    } else {
        nkeep = U; kp = off0;
        for (int c0 = lo; c0 < hi; c0 += 4) {
            uint32_t fl = __ldcg(reinterpret_cast<const uint32_t*>(uniq + c0));
            if (c0 + 4 > hi) fl &= (1u << (8 * (hi - c0))) - 1u;
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((fl >> (8 * j)) & 0xff) { ev.pos_of[cb + c0 + j] = (uint32_t)kp++; ev.uflag[cb + c0 + j] = 2; }
        }
    }
    int n_next = nkeep;
    int s_pos = -1, r_pos = -1;
    __syncthreads();

    // ---- E. relaxation: merge the overflow ------------------------------------------------------------------------------------------
    if (cut && comp == DDO_RELAXED) {
        if (tid < 32) sm.merged[tid] = 0;
        __syncthreads();
        uint64_t acc[S];
#pragma unroll
        for (int j = 0; j < S; ++j) acc[j] = 0;
        unsigned long long mkey = 0;
        for (int li = tid; li < Ublk; li += NT) if (stat[li] == 2) {
            const uint32_t c = ev.ulist[cb + cta_off + li];
            const uint4* sp = reinterpret_cast<const uint4*>(ev.cand_state + (cb + c) * S);
#pragma unroll
            for (int j = 0; j < S / 2; ++j) { const uint4 v4 = sp[j]; acc[2 * j] |= u4lo(v4); acc[2 * j + 1] |= u4hi(v4); }
            mkey = max(mkey, ev.cand_agg[cb + c]);
        }
#pragma unroll
        for (int j = 0; j < S; ++j) {
            uint64_t x = warp_reduce(acc[j], [](uint64_t a, uint64_t b) { return a | b; });
            if (lane == 0 && x) atomicOr(&sm.merged[j], (unsigned long long)x);
        }
        mkey = block_reduce(mkey, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
        X.allreduce(&mkey, 1, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });  // (its barrier also publishes sm.merged and pos_of)
        __threadfence();
        if (tid < S) {  // the cluster-wide union, gathered by every CTA (S <= 16 words x 8 ranks)
            unsigned long long m = 0;
            for (unsigned r = 0; r < FCL_CS; ++r) m |= *cl.map_shared_rank(&sm.merged[tid], r);
            sm.merged[16 + tid] = m;
        }
        __syncthreads();
        // recycled ? a KEPT node whose state equals the merged state (every CTA runs the same lookup: cluster-uniform result)
        if (tid == 0) {
            uint64_t h = 0;
            for (int j = 0; j < S; ++j) h += sm.merged[16 + j] * hash_mul(j);
            h = mix64(h);
            const uint32_t tag = (uint32_t)(h >> 32);
            uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
            const unsigned long long* tab = ev.table + (size_t)k * ev.T;
            int recycled = -1;
            for (;;) {
                const unsigned long long e = tab[slot];
                if (e == EMPTY64) break;
                if ((uint32_t)(e >> 32) == tag) {
                    const uint32_t oc = (uint32_t)e;
                    bool eq = true;
                    for (int j = 0; j < S; ++j) eq = eq && ev.cand_state[(cb + oc) * S + j] == sm.merged[16 + j];
                    if (eq) { const uint32_t f = ev.cand_first[cb + oc]; if (__ldcg(ev.pos_of + cb + f) != NONE32) recycled = (int)f; break; }
                }
                slot = (slot + 1) & (uint32_t)(ev.T - 1);
            }
            sm.misc[4] = recycled;
        }
        __syncthreads();
        const int recycled = sm.misc[4];
        const int mpos = (recycled >= 0) ? (int)__ldcg(ev.pos_of + cb + recycled) : nkeep;
        cl.sync();  // every CTA has finished its lookup (it reads pos_of) before anybody re-points the merged-away candidates below
        if (recycled >= 0) {
            // the best merged-away node ("saved") stays in the layer, un-deleted, next to the recycled node (clean.rs:868-871)
            uint32_t bestc = NONE32;
            for (int li = tid; li < Ublk; li += NT) if (stat[li] == 2) {
                const uint32_t c = ev.ulist[cb + cta_off + li];
                if (bestc == NONE32 || cand_better<S>(ev, cb, c, bestc)) bestc = c;
            }
            __shared__ uint32_t s_best[NT];
            s_best[tid] = bestc;
            __syncthreads();
            for (int d = NT / 2; d > 0; d >>= 1) {
                if (tid < d) {
                    const uint32_t a = s_best[tid], b2 = s_best[tid + d];
                    if (a == NONE32 || (b2 != NONE32 && cand_better<S>(ev, cb, b2, a))) s_best[tid] = b2;
                }
                __syncthreads();
            }
            unsigned long long bc = s_best[0];
            // best over the cluster: every CTA folds the eight local winners in rank order with the same comparator
            if (tid < FCL_CS) *cl.map_shared_rank(&sm.xch[X.phase][0][rank], tid) = bc;
            cl.sync();
            uint32_t saved = NONE32;
            for (unsigned r = 0; r < FCL_CS; ++r) {
                const uint32_t c = (uint32_t)sm.xch[X.phase][0][r];
                if (c != NONE32 && (saved == NONE32 || cand_better<S>(ev, cb, c, saved))) saved = c;
            }
            X.phase ^= 1;
            s_pos = nkeep; r_pos = mpos; n_next = nkeep + 1;
            cl.sync();  // every CTA has read cand_agg[recycled] through cand_better before rank 0 rewrites it
            if (rank == 0 && tid == 0) {
                const unsigned long long rk = ev.cand_agg[cb + recycled];
                if (key_value(mkey) >= key_value(rk)) ev.cand_agg[cb + recycled] = mkey;
                ev.cand_inex[cb + recycled] |= (uint8_t)(NF_INEXACT | NF_RELAXED);
            }
            for (int li = tid; li < Ublk; li += NT) if (stat[li] == 2) {
                const uint32_t c = ev.ulist[cb + cta_off + li];
                if (c == saved) { ev.pos_of[cb + c] = (uint32_t)s_pos; stat[li] = 1; ev.uflag[cb + c] = 2; }
                else ev.pos_of[cb + c] = (uint32_t)r_pos;
            }
        } else {
            n_next = nkeep + 1;
            for (int li = tid; li < Ublk; li += NT) if (stat[li] == 2) ev.pos_of[cb + ev.ulist[cb + cta_off + li]] = (uint32_t)mpos;
            if (rank == 0) {  // new merged node (clean.rs:832-849) written straight into the next layer
                const int nbuf = t & 1;
                const size_t nb = (size_t)k * ev.Wcap + mpos;
                if (tid < S) ev.cur_state[nbuf][nb * S + tid] = sm.merged[16 + tid];
                if (tid == 0) {
                    ev.cur_val[nbuf][nb] = key_value(mkey);
                    ev.cur_flag[nbuf][nb] = (uint8_t)(NF_INEXACT | NF_RELAXED);
                    ev.plog[(lb + t) * ev.Wcap + mpos] = ((uint32_t)mkey & PLOG_CAND_MASK) | PLOG_INEXACT | PLOG_RELAXED;
                }
            }
        }
    }

    // ---- F. terminal layer: best nodes (last maximum in canonical order) -------------------------------------------------------------
    if (terminal) {
        unsigned long long bb[2] = {0ull, 0ull};
        for (int i = 0; i < cnt; ++i) {
            const uint32_t c = ev.ulist[cb + off0 + i];
            const unsigned long long kk = (ev.cand_agg[cb + c] & 0xFFFFFFFF00000000ull) | (unsigned)(ev.pos_of[cb + c] + 1);
            bb[0] = max(bb[0], kk);
            if (!(ev.cand_inex[cb + c] & (NF_INEXACT | NF_RELAXED))) bb[1] = max(bb[1], kk);
        }
        bb[0] = block_reduce(bb[0], [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
        bb[1] = block_reduce(bb[1], [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
        X.allreduce(bb, 2, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
        if (rank == 0 && tid == 0) {
            ctl->has_best = 1; ctl->best_value = key_value(bb[0]); ctl->best_pos = (int)(uint32_t)bb[0] - 1;
            ctl->has_best_exact = bb[1] != 0;
            if (bb[1]) { ctl->best_exact_value = key_value(bb[1]); ctl->best_exact_pos = (int)(uint32_t)bb[1] - 1; }
        }
    }
    if (rank == 0 && tid == 0) {
        ev.nlog[lb + t] = n_next;
        ev.vlog[lb + t] = var;
        ev.rslog[(lb + t) * 2] = s_pos; ev.rslog[(lb + t) * 2 + 1] = r_pos;
        ctl->n_cur = n_next; ctl->var = var;
        if (cut && ctl->lel < 0) { ctl->lel = t - 1; ctl->lel_pending = 1; }
        if (terminal) { ctl->status = ST_TERMINAL; ctl->t_term = t; atomicSub(ev.active, 1); }
    }
}

template <int S>
__global__ void __cluster_dims__(FCL_CS, 1, 1) __launch_bounds__(FCL_NT) k_finish_cl(EV ev, int t, int kcap) {
    pdl_enter();
    __shared__ FinishClSmem sm;
    __shared__ int s_last;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    cg::cluster_group cl = cg::this_cluster();
    finish_body_cl<S>(ev, t, sm, reinterpret_cast<unsigned long long*>(dyn_smem), dyn_smem + (size_t)kcap * 8, kcap);
    cl.sync();  // nobody leaves while a neighbour may still read its shared memory
    if (cl.block_rank() != 0) return;
    // ---- work plan of the two flat kernels that follow (as in k_finish): the last rank-0 CTA scans the per-DD tile counts ----------------
    constexpr int G = S / 2, PER_TILE = 256 / G;
    const int tid = threadIdx.x, count = gridDim.x / FCL_CS;
    __syncthreads();
    if (tid == 0) { __threadfence(); s_last = (atomicAdd(ev.finish_counter, 1u) == (unsigned)count - 1u); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile DDCtl* vc = ev.ctl;
    const int per = (count + blockDim.x - 1) / blockDim.x;
    const int lo = min(tid * per, count), hi = min(lo + per, count);
    int te = 0, tc = 0;
    for (int k = lo; k < hi; ++k) {
        const int st = vc[k].status;
        te += st == ST_ACTIVE ? (vc[k].n_cur + PER_TILE - 1) / PER_TILE : 0;
        tc += (st == ST_ACTIVE || st == ST_TERMINAL) ? (vc[k].ncand + PER_TILE - 1) / PER_TILE : 0;
    }
    int tote, totc;
    int oe = block_excl_scan(te, &tote, sm.scan);
    int oc = block_excl_scan(tc, &totc, sm.scan);
    for (int k = lo; k < hi; ++k) {
        const int st = vc[k].status;
        ev.tile_off_e[k] = oe; ev.tile_off_c[k] = oc;
        oe += st == ST_ACTIVE ? (vc[k].n_cur + PER_TILE - 1) / PER_TILE : 0;
        oc += (st == ST_ACTIVE || st == ST_TERMINAL) ? (vc[k].ncand + PER_TILE - 1) / PER_TILE : 0;
    }
    if (tid == 0) { ev.tile_off_e[count] = tote; ev.tile_off_c[count] = totc; *ev.finish_counter = 0; }
}

// =================================================================================================================
// k_compact: scatter layer t into the ping-pong buffers, write logs, release hash slots, snapshot the LEL.
// =================================================================================================================
template <int S>
__global__ void __launch_bounds__(256) k_compact(EV ev, int t, int count) {
    pdl_enter();
    constexpr int G = S / 2;
    constexpr int CPB = 256 / G;
    const int* off = ev.tile_off_c;  // (L1-resident after the first lookups; staging it in shared memory measured slower)
    const int total = off[count];
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int k = plan_find(off, count, tile);
    const DDCtl* ctl = ev.ctl + k;
    const int ncand = ctl->ncand;
    if (tile == off[k]) {  // first tile of this DD: reset the per-layer accumulators that k_expand fills next
        for (int i = threadIdx.x; i < ev.HN; i += 256) ev.vhist[(size_t)k * ev.HN + i] = 0;
        if (threadIdx.x == 0) ev.ucount[k] = 0;
    }
    const int c = (tile - off[k]) * CPB + threadIdx.x / G;
    if (c >= ncand) continue;
    const int sub = threadIdx.x % G;
    const size_t cb = (size_t)k * ev.C;
    const size_t lb = (size_t)k * ev.Lmax;
    const int nbuf = t & 1;
    const bool relaxed = ctl->comp_type == DDO_RELAXED;
    const uint32_t rep = ev.cand_rep[cb + c];
    uint32_t child = NONE32;
    if (rep != NONE32) {
        const uint32_t f = ev.cand_first[cb + rep];
        child = ev.pos_of[cb + f];
        if (sub == 0 && rep == (uint32_t)c) {  // release the hash slot this candidate claimed
            const uint32_t slot = ev.cand_slot[cb + c];
            if (slot != NONE32) ev.table[(size_t)k * ev.T + slot] = EMPTY64;
        }
        if (ev.uflag[cb + c] == 2) {  // surviving canonical representative: becomes node `pos` of layer t
            const uint32_t pos = ev.pos_of[cb + c];
            const size_t nb = (size_t)k * ev.Wcap + pos;
            const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(ev.cand_state + (cb + c) * S) + sub);
            st_stream_u4(reinterpret_cast<uint4*>(ev.cur_state[nbuf] + nb * S) + sub, v);
            if (sub == 0) {
                const unsigned long long key = ev.cand_agg[cb + c];
                const uint32_t fl = ev.cand_inex[cb + c];
                ev.cur_val[nbuf][nb] = key_value(key);
                ev.cur_flag[nbuf][nb] = (uint8_t)fl;
                ev.plog[(lb + t) * ev.Wcap + pos] = ((uint32_t)key & PLOG_CAND_MASK) | ((fl & NF_INEXACT) ? PLOG_INEXACT : 0u) | ((fl & NF_RELAXED) ? PLOG_RELAXED : 0u);
            }
        }
    }
    if (t > 0) {
        if (relaxed && sub == 0) ev.clog[(lb + t - 1) * ev.C + c] = child;  // edge (parent c/2, decision) -> node `child` of layer t
        if (ctl->lel_pending && relaxed && !(c & 1)) {  // layer t-1 is the last exact layer: keep its nodes for the cutset
            const int i = c >> 1;
            const size_t pb = (size_t)k * ev.Wcap + i;
            const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(ev.cur_state[(t - 1) & 1] + pb * S) + sub);
            st_stream_u4(reinterpret_cast<uint4*>(ev.lel_state + pb * S) + sub, v);
            if (sub == 0) { ev.lel_val[pb] = ev.cur_val[(t - 1) & 1][pb]; ev.lel_rub[pb] = ev.cur_rub[pb]; }
        }
    }
    }  // tile loop
}

// =================================================================================================================
// k_compact1: the same compaction with ONE THREAD per candidate.  k_compact is bound by the latency of its gather chain
// (rep -> first -> pos_of -> row) times the candidates in flight; a thread per candidate keeps G times more of them in flight.
// Grid-stride over units of 32 candidates (one warp each; a tile of the work plan holds 256 / G candidates = 8 / G units).
// =================================================================================================================
template <int S>
__global__ void __launch_bounds__(256) k_compact1(EV ev, int t, int count) {
    pdl_enter();
    constexpr int G = S / 2;
    constexpr int CPB = 256 / G;        // candidates per tile of the work plan
    constexpr int UPT = CPB / 32;       // 32-candidate units per plan tile
    const int* off = ev.tile_off_c;
    const int total_units = off[count] * UPT;
    const int lane = threadIdx.x & 31;
    const int wglobal = (blockIdx.x * 256 + threadIdx.x) >> 5, wstride = (gridDim.x * 256) >> 5;
    for (int u = wglobal; u < total_units; u += wstride) {
        const int tile = u / UPT;
        const int k = plan_find_warp(off, count, tile);
        const DDCtl* ctl = ev.ctl + k;
        const int ncand = ctl->ncand;
        const int c0 = (tile - off[k]) * CPB + (u % UPT) * 32;
        if (c0 == 0) {  // first unit of this DD: reset the per-layer accumulators that the expansion fills next
            for (int i = lane; i < ev.HN; i += 32) ev.vhist[(size_t)k * ev.HN + i] = 0;
            if (lane == 0) ev.ucount[k] = 0;
        }
        const int c = c0 + lane;
        if (c >= ncand) continue;
        const size_t cb = (size_t)k * ev.C;
        const size_t lb = (size_t)k * ev.Lmax;
        const int nbuf = t & 1;
        const bool relaxed = ctl->comp_type == DDO_RELAXED;
        const uint32_t rep = ev.cand_rep[cb + c];
        uint32_t child = NONE32;
        if (rep != NONE32) {
            const uint32_t f = ev.cand_first[cb + rep];
            child = ev.pos_of[cb + f];
            if (rep == (uint32_t)c) {  // release the hash slot this candidate claimed
                const uint32_t slot = ev.cand_slot[cb + c];
                if (slot != NONE32) ev.table[(size_t)k * ev.T + slot] = EMPTY64;
            }
            if (ev.uflag[cb + c] == 2) {  // surviving canonical representative: becomes node `pos` of layer t
                const uint32_t pos = ev.pos_of[cb + c];
                const size_t nb = (size_t)k * ev.Wcap + pos;
                const uint4* src = reinterpret_cast<const uint4*>(ev.cand_state + (cb + c) * S);
                uint4* dst = reinterpret_cast<uint4*>(ev.cur_state[nbuf] + nb * S);
                uint4 v[G];
#pragma unroll
                for (int q = 0; q < G; ++q) v[q] = ld_stream_u4(src + q);
#pragma unroll
                for (int q = 0; q < G; ++q) st_stream_u4(dst + q, v[q]);
                const unsigned long long key = ev.cand_agg[cb + c];
                const uint32_t fl = ev.cand_inex[cb + c];
                ev.cur_val[nbuf][nb] = key_value(key);
                ev.cur_flag[nbuf][nb] = (uint8_t)fl;
                ev.plog[(lb + t) * ev.Wcap + pos] = ((uint32_t)key & PLOG_CAND_MASK) | ((fl & NF_INEXACT) ? PLOG_INEXACT : 0u) | ((fl & NF_RELAXED) ? PLOG_RELAXED : 0u);
            }
        }
        if (t > 0) {
            if (relaxed) ev.clog[(lb + t - 1) * ev.C + c] = child;
            if (ctl->lel_pending && relaxed && !(c & 1)) {  // layer t-1 is the last exact layer: keep its nodes for the cutset
                const int i = c >> 1;
                const size_t pb = (size_t)k * ev.Wcap + i;
                const uint4* src = reinterpret_cast<const uint4*>(ev.cur_state[(t - 1) & 1] + pb * S);
                uint4* dst = reinterpret_cast<uint4*>(ev.lel_state + pb * S);
#pragma unroll
                for (int q = 0; q < G; ++q) st_stream_u4(dst + q, ld_stream_u4(src + q));
                ev.lel_val[pb] = ev.cur_val[(t - 1) & 1][pb]; ev.lel_rub[pb] = ev.cur_rub[pb];
            }
        }
    }
}

// =================================================================================================================
// k_small: one CTA compiles one whole DD in shared memory -- the fast path for the (vast majority of) sub-problems whose layers
// stay narrow.  Valid only while no layer needs a cut (|layer| <= min(max_width, Ws)): then restricted == relaxed == exact DD, node
// order is irrelevant and only (best value, expanded, transitions) are observable (clean.rs:345-381 with _squash_if_needed never
// firing).  A DD that outgrows Ws or its max_width reports `overflow` and is recompiled by the general engine.
// =================================================================================================================

template <int S>
__global__ void __launch_bounds__(128) k_small(EV ev, int count, int Ws, long long best_lb, SmallOut* out) {
    constexpr int G = S / 2, NT = 128, GPB = NT / G, W32 = 2 * S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = blockIdx.x;
    if (k >= count) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, sub = tid % G, grp = tid / G;
    const unsigned gm = group_mask<G>();
    const int TS = 4 * Ws;  // hash slots (power of two: Ws is)
    // carve shared memory
    uint4* cur = reinterpret_cast<uint4*>(smem_raw);                       // [Ws][G]
    uint4* cand = cur + (size_t)Ws * G;                                    // [2 Ws][G]
    int32_t* curv = reinterpret_cast<int32_t*>(cand + (size_t)2 * Ws * G);  // [Ws]
    int32_t* candv = curv + Ws;                                            // [2 Ws]
    uint32_t* table = reinterpret_cast<uint32_t*>(candv + 2 * Ws);         // [TS]  candidate index + 1, 0 = empty
    uint32_t* hist = table + TS;                                           // [64 S]
    uint8_t* candf = reinterpret_cast<uint8_t*>(hist + 64 * S);            // [2 Ws] 0 invalid, 1 valid, 2 claimed (distinct)
    __shared__ unsigned long long s_red[40];
    __shared__ int s_n, s_exp, s_tr;
    const int width = ev.root_width[k];
    const int cap = min(Ws, width);
    if (tid < G) cur[tid] = reinterpret_cast<const uint4*>(ev.root_state + (size_t)k * S)[tid];
    if (tid == 0) { curv[0] = ev.root_val[k]; s_n = 1; }
    unsigned long long expanded = 0, transitions = 0;
    int n = 1, layers = 0, status = 0, has_best = 0, best_value = 0;
    __syncthreads();
    for (;;) {
        // ---- next_variable: occurrences of every vertex among the n states of the layer (misp/main.rs:109-143) ----------
        for (int i = tid; i < 64 * S; i += NT) hist[i] = 0;
        if (tid == 0) { s_exp = 0; s_tr = 0; }
        __syncthreads();
        {
            const uint32_t* rows = reinterpret_cast<const uint32_t*>(cur);
            const int rblocks = (n + 31) / 32;
            for (int job = warp; job < rblocks * W32; job += NT / 32) {
                const int rb = job / W32, col = job % W32, row = rb * 32 + lane;
                const uint32_t x = row < n ? rows[(size_t)row * W32 + col] : 0u;
                const unsigned c = __popc(warp_transpose32(x));
                if (c) atomicAdd(&hist[32 * col + lane], c);
            }
        }
        __syncthreads();
        unsigned long long best = ~0ull;
        for (int i = tid; i < ev.n; i += NT) { const unsigned c = hist[i]; if (c) best = min(best, ((unsigned long long)c << 32) | (unsigned)i); }
        best = block_reduce(best, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, ~0ull, s_red);
        if (best == ~0ull) {  // terminal layer: best node = max value_top (clean.rs:620-632)
            int bv = INT32_MIN;
            for (int i = tid; i < n; i += NT) bv = max(bv, curv[i]);
            unsigned long long r = block_reduce((unsigned long long)((uint32_t)bv ^ 0x80000000u), [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, s_red);
            has_best = 1; best_value = (int32_t)((uint32_t)r ^ 0x80000000u);
            break;
        }
        const int v = (int)(uint32_t)best;
        const int vw = v >> 6;
        const bool owner = (vw >> 1) == sub;
        const uint64_t bit = 1ull << (v & 63);
        const uint4 nc4 = __ldg(reinterpret_cast<const uint4*>(ev.nc + (size_t)v * S) + sub);
        const int wv = ev.weight[v];
        // ---- expansion (clean.rs:360-370, misp/main.rs:77-102,191-193) ---------------------------------------------------
        for (int i = tid; i < TS; i += NT) table[i] = 0;
        for (int i = tid; i < 2 * n; i += NT) candf[i] = 0;
        __syncthreads();
        int my_exp = 0, my_tr = 0;
        for (int base = 0; base < n; base += GPB) {
            const int node = base + grp;
            if (node < n) {
                const uint4 s4 = cur[(size_t)node * G + sub];
                uint64_t w0 = u4lo(s4), w1 = u4hi(s4);
                const int val = curv[node];
                int rub;
                if (ev.unit_weights) rub = __popcll(w0) + __popcll(w1);
                else {
                    rub = 0;
                    uint64_t x = w0; const int32_t* wp = ev.weight + (2 * sub) * 64;
                    while (x) { int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; }
                    x = w1; wp += 64;
                    while (x) { int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; }
                }
                rub = group_sum<G>(rub, gm);
                const bool has_v = group_any<G>(owner && (((vw & 1) ? w1 : w0) & bit), gm);
                if (((long long)rub + (long long)val) > best_lb) {
                    if (owner) { if (vw & 1) w1 &= ~bit; else w0 &= ~bit; }
                    cand[(size_t)(2 * node + 1) * G + sub] = mk_u4(w0, w1);                      // NO
                    if (has_v) cand[(size_t)(2 * node) * G + sub] = mk_u4(w0 & u4lo(nc4), w1 & u4hi(nc4));  // YES
                    if (sub == 0) {
                        candv[2 * node + 1] = val; candf[2 * node + 1] = 1;
                        if (has_v) { candv[2 * node] = val + wv; candf[2 * node] = 1; }
                        ++my_exp; my_tr += has_v ? 2 : 1;
                    }
                }
            }
        }
        if (my_exp) { atomicAdd(&s_exp, my_exp); atomicAdd(&s_tr, my_tr); }
        __syncthreads();
        expanded += (unsigned)s_exp; transitions += (unsigned)s_tr;
        // ---- dedup (next_l.entry(), clean.rs:738-775): value_top = max over the duplicates --------------------------------
        for (int base = 0; base < 2 * n; base += GPB) {
            const int c = base + grp;
            if (c < 2 * n && candf[c]) {
                const uint4 m4 = cand[(size_t)c * G + sub];
                const uint64_t h = mix64(group_xor64<G>(word_hash(u4lo(m4), 2 * sub) ^ word_hash(u4hi(m4), 2 * sub + 1), gm));
                uint32_t slot = (uint32_t)h & (uint32_t)(TS - 1);
                for (;;) {
                    uint32_t old = 0;
                    if (sub == 0) old = atomicCAS(&table[slot], 0u, (uint32_t)c + 1u);
                    old = __shfl_sync(gm, old, lane & ~(G - 1));
                    if (old == 0u) { if (sub == 0) candf[c] = 2; break; }
                    const uint4 o4 = cand[(size_t)(old - 1) * G + sub];
                    if (group_all<G>(o4.x == m4.x && o4.y == m4.y && o4.z == m4.z && o4.w == m4.w, gm)) {
                        if (sub == 0) atomicMax(&candv[old - 1], candv[c]);
                        break;
                    }
                    slot = (slot + 1) & (uint32_t)(TS - 1);
                }
            }
        }
        if (tid == 0) s_n = 0;
        __syncthreads();
        // ---- the distinct candidates become the next layer -----------------------------------------------------------------
        int mine = 0;
        for (int c = tid; c < 2 * n; c += NT) mine += (candf[c] == 2);
        int U;
        int off = block_excl_scan(mine, &U, reinterpret_cast<int*>(s_red));
        ++layers;
        if (U > cap) { status = 1; break; }
        if (U == 0) { has_best = 0; break; }  // every node was pruned: no solution in this DD
        for (int c = tid; c < 2 * n; c += NT) if (candf[c] == 2) {
            for (int j = 0; j < G; ++j) cur[(size_t)off * G + j] = cand[(size_t)c * G + j];
            curv[off] = candv[c];
            ++off;
        }
        n = U;
        __syncthreads();
    }
    if (tid == 0) {
        SmallOut o;
        o.status = status; o.has_best = has_best; o.best_value = best_value; o.layers = layers; o.expanded = expanded; o.transitions = transitions;
        out[k] = o;
    }
}

// parent-log entry of node `pos` of layer `tt` of DD k: layers before a twin's fork belong to its primary
__device__ __forceinline__ uint32_t plog_at(const EV& ev, const DDCtl* ctl, int k, int tt, int pos) {
    const int src = (ctl->primary >= 0 && tt < ctl->fork_t) ? ctl->primary : k;
    return ev.plog[((size_t)src * ev.Lmax + tt) * ev.Wcap + pos];
}

// =================================================================================================================
// k_finalize: exact-best-path walk (clean.rs:634-655) and decision bits of the best / best exact path (clean.rs:329-343)
// =================================================================================================================
static __global__ void k_finalize(EV ev, int count) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    DDCtl* ctl = ev.ctl + k;
    const size_t lb = (size_t)k * ev.Lmax;
    ctl->ebpo = 0;
    if (ctl->overflow) return;
    if (!ctl->has_best) { ctl->ebpo = ctl->comp_type == DDO_RELAXED; return; }  // _has_exact_best_path(None) == true (clean.rs:643-655): an infeasible relaxed DD is exact
    const int T = ctl->t_term;
    if (ctl->comp_type == DDO_RELAXED) {
        int pos = ctl->best_pos, tt = T;
        bool exact = true;
        for (;;) {
            const uint32_t e = plog_at(ev, ctl, k, tt, pos);
            if (!(e & PLOG_INEXACT)) { exact = true; break; }
            if (e & PLOG_RELAXED) { exact = false; break; }
            if (tt == 0) break;
            pos = (int)((e & PLOG_CAND_MASK) >> 1); --tt;
        }
        ctl->ebpo = exact;
        if (exact) { ctl->has_best_exact = 1; ctl->best_exact_pos = ctl->best_pos; ctl->best_exact_value = ctl->best_value; }  // clean.rs:638-640
    }
    for (int which = 0; which < 2; ++which) {
        uint64_t* out = (which == 0 ? ev.best_path : ev.best_exact_path) + (size_t)k * ev.PW;
        for (int w = 0; w < ev.PW; ++w) out[w] = 0;
        if (which == 1 && !ctl->has_best_exact) continue;
        int pos = which == 0 ? ctl->best_pos : ctl->best_exact_pos;
        for (int tt = T; tt >= 1; --tt) {
            const uint32_t cand = plog_at(ev, ctl, k, tt, pos) & PLOG_CAND_MASK;
            if (!(cand & 1u)) out[(tt - 1) >> 6] |= 1ull << ((tt - 1) & 63);  // even candidate = YES
            pos = (int)(cand >> 1);
        }
    }
}

// =================================================================================================================
// k_bottomup: local bounds of a relaxed DD (clean.rs:448-475) as a per-layer GATHER over the child log, then the cutset
// upper bounds ub = min(value_top + rub, value_top + value_bot, best_value) of the last exact layer (clean.rs:426-428).
// =================================================================================================================
constexpr int32_t UNMARKED = INT32_MIN;
static __global__ void __launch_bounds__(1024, 1) k_bottomup(EV ev) {
    const int k = blockIdx.x;
    DDCtl* ctl = ev.ctl + k;
    const int tid = threadIdx.x, NT = blockDim.x;
    if (tid == 0) { ctl->cutset_count = 0; ctl->lel_n = 0; }
    if (ctl->comp_type != DDO_RELAXED || ctl->overflow || !ctl->has_best || ctl->lel < 0) return;
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    const size_t lb = (size_t)k * ev.Lmax;
    const int T = ctl->t_term, L = ctl->lel;
    int32_t* nxt = ev.vb[0] + (size_t)k * ev.Wcap;
    int32_t* cur = ev.vb[1] + (size_t)k * ev.Wcap;
    for (int i = tid; i < ev.nlog[lb + T]; i += NT) nxt[i] = 0;  // terminal layer: value_bot = 0, MARKED
    __syncthreads();
    for (int tt = T - 1; tt >= L; --tt) {
        const int n = ev.nlog[lb + tt];
        const int wv = ev.weight[ev.vlog[lb + tt]];
        const int s = ev.rslog[(lb + tt + 1) * 2], r = ev.rslog[(lb + tt + 1) * 2 + 1];
        const uint32_t* cl = ev.clog + (lb + tt) * ev.C;
        for (int i = tid; i < n; i += NT) {
            int32_t best = UNMARKED;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const uint32_t ch = cl[2 * i + d];
                if (ch == NONE32) continue;
                const int cost = d == 0 ? wv : 0;
                int32_t x = nxt[ch];
                if (x != UNMARKED) best = max(best, x + cost);
                if ((int)ch == s && r >= 0) { x = nxt[r]; if (x != UNMARKED) best = max(best, x + cost); }  // edges of the saved node were also copied to the recycled node
            }
            cur[i] = best;
        }
        __syncthreads();
        int32_t* tmp = nxt; nxt = cur; cur = tmp;
    }
    const int n = ev.nlog[lb + L];
    const size_t nb = (size_t)k * ev.Wcap;
    int local = 0;
    for (int i = tid; i < n; i += NT) {
        const int32_t vbot = nxt[i];
        const bool marked = vbot != UNMARKED;
        ev.cs_marked[nb + i] = marked;
        if (marked) {
            const int val = ev.lel_val[nb + i];
            ev.cs_ub[nb + i] = min(min(val + ev.lel_rub[nb + i], val + vbot), ctl->best_value);
            ++local;
        }
    }
    if (local) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (tid == 0) { ctl->cutset_count = s_cnt; ctl->lel_n = n; }
}

// =================================================================================================================
// drain_cutset (clean.rs:417-445) + the solver-side filter (parallel.rs:460-461) as a batched stream compaction
// =================================================================================================================
static __global__ void __launch_bounds__(1024, 1) k_cutset_count(EV ev, DrainOut o, const long long* ub_cap, const long long* lb_filter, int count) {
    __shared__ int scan[40];
    const int k = blockIdx.x;
    const DDCtl* ctl = ev.ctl + k;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int n = (k < count && ctl->cutset_count > 0) ? ctl->lel_n : 0;
    const size_t nb = (size_t)k * ev.Wcap;
    const int per = (n + NT - 1) / NT, lo = min(tid * per, n), hi = min(lo + per, n);
    const long long cap = ub_cap[k], lbf = lb_filter[k];
    int c = 0;
    for (int i = lo; i < hi; ++i) c += (ev.cs_marked[nb + i] && min((long long)ev.cs_ub[nb + i], cap) > lbf);
    int total;
    int off = block_excl_scan(c, &total, scan);
    for (int i = lo; i < hi; ++i) {
        const bool f = ev.cs_marked[nb + i] && min((long long)ev.cs_ub[nb + i], cap) > lbf;
        o.loc[nb + i] = f ? (uint32_t)off++ : NONE32;
    }
    if (tid == 0) o.count[k] = total;
}
static __global__ void k_cutset_offsets(DrainOut o, int K) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < K; ++k) { o.offset[k] = acc; acc += o.count[k]; }
        o.offset[K] = acc;
    }
}
template <int S>
__global__ void __launch_bounds__(256) k_cutset_write(EV ev, DrainOut o, const long long* ub_cap, int pw) {
    const int k = blockIdx.y;
    const DDCtl* ctl = ev.ctl + k;
    if (o.count[k] == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ctl->lel_n) return;
    const size_t nb = (size_t)k * ev.Wcap;
    const uint32_t loc = o.loc[nb + i];
    if (loc == NONE32) return;
    const size_t rec = (size_t)o.offset[k] + loc;
    for (int j = 0; j < S; ++j) o.state[rec * S + j] = ev.lel_state[(nb + i) * S + j];
    o.val[rec] = ev.lel_val[nb + i];
    o.ub[rec] = (int32_t)min((long long)ev.cs_ub[nb + i], ub_cap[k]);
    o.dd[rec] = k;
    const size_t lb = (size_t)k * ev.Lmax;
    uint64_t bits[8];  // pw <= 8 (lel < 512) is enforced on the host; deeper cutsets use the slow path below
    for (int w = 0; w < 8; ++w) bits[w] = 0;
    int pos = i;
    for (int tt = ctl->lel; tt >= 1; --tt) {
        const uint32_t cand = plog_at(ev, ctl, k, tt, pos) & PLOG_CAND_MASK;
        if (!(cand & 1u)) {
            if (pw <= 8) bits[(tt - 1) >> 6] |= 1ull << ((tt - 1) & 63);
            else o.path[rec * pw + ((tt - 1) >> 6)] |= 1ull << ((tt - 1) & 63);
        }
        pos = (int)(cand >> 1);
    }
    if (pw <= 8) for (int w = 0; w < pw; ++w) o.path[rec * pw + w] = bits[w];
}

}  // namespace ddo

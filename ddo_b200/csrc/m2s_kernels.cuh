// m2s_kernels.cuh -- hand-written sm_100a kernels of the batch DD-compilation engine, MAX2SAT device model
// (ddo/examples/max2sat/model.rs:29-348, relax.rs:43-89, heuristics.rs:30-37).
//
// A state is a row of NW int32 marginal benefits (NW = n rounded up to a multiple of 4; 2 000 B at n = 500): one WARP owns one node and
// moves its row with 128-bit loads / stores, 512 contiguous bytes per instruction.  The branching variable of a layer is a function of
// the depth (model.rs:330-348), so there is no next_variable scan.  One layer step t -> t+1 of every DD of the batch:
//   m2_finish    (one CTA per DD)   canonical representatives, ordered compaction, MSD radix-select of the width cut on
//                                   (value_top, rank = sum |benefit|, then lexicographic benefits), LEL bookkeeping       clean.rs:779-876
//   m2_merge     (flat)             relaxed DDs: column-wise min / max of the merged-away rows                           relax.rs:46-77
//   m2_merge_fin (one CTA per DD)   merged row, its rank, relaxed value_top, recycled-node lookup, merged node           clean.rs:826-876, relax.rs:78-84
//   m2_compact   (flat, warp/cand)  scatter survivors, parent / child / edge-cost logs, hash-slot release, LEL snapshot   clean.rs:657-687
//   m2_expand    (flat, warp/node)  rough upper bound prune, both transitions + costs, rank, hash, dedup insert           clean.rs:360-370,728-776
// After the last layer: m2_finalize, m2_bottomup, m2_cutset_{count,offsets,write}.
#pragma once
#include "kernels.cuh"
#include "m2s_engine.hpp"

namespace ddo {

constexpr uint32_t M2_DROPPED = 0xFFFFFFFEu;  // pos_of marker: merged away (its position is known after m2_merge_fin)

__device__ __forceinline__ int4 ld_stream_i4(const int4* p) { const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(p)); return make_int4((int)v.x, (int)v.y, (int)v.z, (int)v.w); }
__device__ __forceinline__ void st_stream_i4(int4* p, int4 v) { st_stream_u4(reinterpret_cast<uint4*>(p), make_uint4((unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w)); }
__device__ __forceinline__ int4 ld_cg_i4(const int4* p) { const uint4 v = ld_cg_u4(reinterpret_cast<const uint4*>(p)); return make_int4((int)v.x, (int)v.y, (int)v.z, (int)v.w); }
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}
__device__ __forceinline__ int warp_sum32(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
    return v;
}
// State rows live only in the candidate buffers: the candidates of layer t (and its merged node, row C) in buffer t & 1.  A node of the
// current layer is a reference (cur_src) to the candidate row it came from, so a layer step moves every row exactly once (parent read,
// child write) -- there is no separate copy of the current layer.
__device__ __forceinline__ int32_t* m2_row(const M2EV& ev, int b, int k, uint32_t c) { return ev.cand_state[b] + ((size_t)k * (ev.C + 1) + c) * ev.NW; }
__device__ __forceinline__ int iabs(int x) { return x < 0 ? -x : x; }
__device__ __forceinline__ int ipos(int x) { return x > 0 ? x : 0; }

// =================================================================================================================
// m2_init: root of every DD becomes the single "candidate" of layer 0 (clean.rs:383-405)
// =================================================================================================================
__global__ void __launch_bounds__(256) m2_init(M2EV ev, int count, int comp_type, long long best_lb) {
    const int k = blockIdx.x;
    if (k >= count) return;
    __shared__ int s_rank[8];
    const size_t cb = (size_t)k * ev.C;
    int rank = 0;
    for (int i = threadIdx.x; i < ev.NW; i += 256) {
        const int v = ev.root_state[(size_t)k * ev.NW + i];
        m2_row(ev, 0, k, 0)[i] = v;
        rank += iabs(v);
        ev.mrg_min[(size_t)k * ev.NW + i] = INT32_MAX; ev.mrg_max[(size_t)k * ev.NW + i] = INT32_MIN;
    }
    rank = warp_sum32(rank);
    if ((threadIdx.x & 31) == 0) s_rank[threadIdx.x >> 5] = rank;
    __syncthreads();
    if (threadIdx.x == 0) {
        rank = 0;
        for (int w = 0; w < 8; ++w) rank += s_rank[w];
        DDCtl c{};
        c.status = ST_ACTIVE; c.ncand = 1; c.n_cur = 0; c.var = -1;
        c.width = ev.root_width[k]; c.comp_type = comp_type; c.root_depth = ev.root_depth[k]; c.lel = -1;
        c.t_term = -1; c.best_pos = -1; c.best_exact_pos = -1; c.root_value = ev.root_val[k];
        c.best_lb = best_lb; c.primary = -1; c.fork_t = -1;
        ev.ctl[k] = c;
        ev.cand_rep[cb] = 0; ev.cand_first[cb] = 0; ev.cand_agg[cb] = pack_key(ev.root_val[k], PLOG_CAND_MASK); ev.cand_inex[cb] = 0;
        ev.cand_rank[cb] = (uint32_t)rank; ev.cand_slot[cb] = NONE32; ev.cand_cost[cb] = 0; ev.uflag[cb] = 0;
        M2Aux a{}; ev.aux[k] = a;
        if (k == 0) { *ev.active = count; *ev.lel_any = -1; }
    }
}

// =================================================================================================================
// m2_expand: layer t -> candidates of layer t+1.  One warp per node; CH = 128-bit chunks of a row per lane.
// =================================================================================================================
// ---- TMA 1-D bulk copies (cp.async.bulk + mbarrier) ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// Every CTA walks a CONTIGUOUS range of tiles (8 nodes of one DD each), so the branching variable -- a function of the DD's depth -- changes
// rarely: its four clause-weight rows (and, once, the hash multipliers) are staged in shared memory by TMA bulk copies and shared by the
// eight warps, instead of being re-fetched through L1/L2 by every warp for every node.  Warps then run without block barriers.
template <int CH>
__global__ void __launch_bounds__(256, CH <= 4 ? 4 : 2) m2_expand(M2EV ev, int t, int count) {
    const int* off = ev.tile_off_e;
    const int total = off[count];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NW4 = ev.NW4;
    const int buf = t & 1;
    extern __shared__ __align__(128) unsigned char dyn_smem[];  // [5][NW]: PT, QT, PF, QF rows of the current variable, hash multipliers
    int4* sP[2] = {reinterpret_cast<int4*>(dyn_smem), reinterpret_cast<int4*>(dyn_smem) + 2 * NW4};
    int4* sQ[2] = {reinterpret_cast<int4*>(dyn_smem) + NW4, reinterpret_cast<int4*>(dyn_smem) + 3 * NW4};
    const uint4* sM = reinterpret_cast<const uint4*>(dyn_smem) + 4 * NW4;
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t row_bytes = (uint32_t)ev.NW * 4u;
    const int tpb = (total + gridDim.x - 1) / gridDim.x;
    const int tile_lo = min((int)blockIdx.x * tpb, total), tile_hi = min(tile_lo + tpb, total);
    if (tile_lo >= tile_hi) return;
    if (threadIdx.x == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    uint32_t phase = 0;
    int k = plan_find(off, count, tile_lo);
    int cur_var = -1, cur_k = -1;
    unsigned exp_acc = 0;  // nodes of DD cur_k expanded by this warp since the last flush (lane 0)
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
        while (off[k + 1] <= tile) ++k;
        DDCtl* ctl = ev.ctl + k;
        const int depth = ctl->root_depth + t;
        const int var = ev.ord[ev.n - depth - 1];  // model.rs:330-348
        if (k != cur_k) {
            if (lane == 0 && exp_acc) { atomicAdd(&ev.ctl[cur_k].expanded, (unsigned long long)exp_acc); atomicAdd(&ev.ctl[cur_k].transitions, (unsigned long long)(2 * exp_acc)); }
            exp_acc = 0; cur_k = k;
        }
        if (var != cur_var) {  // block-uniform: every warp walks the same tiles
            __syncthreads();   // everybody is done with the previous rows
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const bool first = cur_var < 0;
                mbar_expect_tx(&s_bar, (first ? 5u : 4u) * row_bytes);
                bulk_g2s(sP[0], ev.PT + (size_t)var * ev.NW, row_bytes, &s_bar);
                bulk_g2s(sQ[0], ev.QT + (size_t)var * ev.NW, row_bytes, &s_bar);
                bulk_g2s(sP[1], ev.PF + (size_t)var * ev.NW, row_bytes, &s_bar);
                bulk_g2s(sQ[1], ev.QF + (size_t)var * ev.NW, row_bytes, &s_bar);
                if (first) bulk_g2s(const_cast<uint4*>(sM), ev.hmul, row_bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase); phase ^= 1;
            cur_var = var;
        }
        const int n_cur = ctl->n_cur;
        const int node = (tile - off[k]) * 8 + warp;
        if (node < n_cur) {
            const size_t nb = (size_t)k * ev.Wcap + node;
            const size_t cb = (size_t)k * ev.C;
            const int4* row = reinterpret_cast<const int4*>(m2_row(ev, buf, k, ev.cur_src[buf][nb]));
            int4 s[CH];
#pragma unroll
            for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; s[q] = i < NW4 ? ld_stream_i4(row + i) : make_int4(0, 0, 0, 0); }
            const int val = ev.cur_val[buf][nb];
            const uint32_t fl = ev.cur_flag[buf][nb];
            const int rank = ev.cur_rank[buf][nb];
            // fast_upper_bound (model.rs:240-249): sum |benefit| + estimates[depth] - initial + nk[depth]
            const long long rub = (long long)rank + ev.est[depth] - ev.initial + ev.nk[depth];
            const bool expandable = rub + (long long)val > ctl->best_lb;  // clean.rs:364-365 (no saturation: all terms < 2^31)
            const uint32_t c_t = 2u * node, c_f = 2u * node + 1u;  // for_each_in_domain order: T then F (model.rs:270-273)
            if (lane == 0) {
                ev.cur_rub[nb] = (int32_t)min(rub, (long long)INT32_MAX);
                ev.cand_rep[cb + c_t] = NONE32; ev.cand_rep[cb + c_f] = NONE32;
                ev.uflag[cb + c_t] = 0; ev.uflag[cb + c_f] = 0;
            }
            if (expandable) {
                ++exp_acc;
                // benefit of the branching variable itself (pos(state[k]) / pos(-state[k]) of model.rs:298,313)
                int sv = 0;
                {
                    const int owner_chunk = var >> 2;
#pragma unroll
                    for (int q = 0; q < CH; ++q) if (lane + 32 * q == owner_chunk) { const int e = var & 3; sv = e == 0 ? s[q].x : (e == 1 ? s[q].y : (e == 2 ? s[q].z : s[q].w)); }
                    sv = __shfl_sync(FULL_MASK, sv, owner_chunk & 31);
                }
                // both children are computed and streamed out first, ONE fence makes their rows visible, then both are inserted
                int values[2]; unsigned long long hashes[2];
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const uint32_t c = d == 0 ? c_t : c_f;
                    int4* dst = reinterpret_cast<int4*>(m2_row(ev, buf ^ 1, k, c));
                    int cost = 0, crank = 0;
                    unsigned long long h = 0;
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const int i = lane + 32 * q;
                        if (i < NW4) {
                            const int4 P = sP[d][i], Q = sQ[d][i];
                            const uint4 m = sM[i];
                            int4 x = s[q], r;
                            // transition (model.rs:275-292) and the state-dependent part of transition_cost (model.rs:294-328)
                            r.x = x.x + P.x - Q.x; r.y = x.y + P.y - Q.y; r.z = x.z + P.z - Q.z; r.w = x.w + P.w - Q.w;
                            cost += min(ipos(x.x) + P.x, ipos(-x.x) + Q.x) + min(ipos(x.y) + P.y, ipos(-x.y) + Q.y) + min(ipos(x.z) + P.z, ipos(-x.z) + Q.z) +
                                    min(ipos(x.w) + P.w, ipos(-x.w) + Q.w);
                            if (i == (var >> 2)) { const int e = var & 3; if (e == 0) r.x = 0; else if (e == 1) r.y = 0; else if (e == 2) r.z = 0; else r.w = 0; }  // ret[k] = 0
                            crank += iabs(r.x) + iabs(r.y) + iabs(r.z) + iabs(r.w);
                            h += (unsigned long long)(uint32_t)r.x * m.x + (unsigned long long)(uint32_t)r.y * m.y + (unsigned long long)(uint32_t)r.z * m.z +
                                 (unsigned long long)(uint32_t)r.w * m.w;
                            st_stream_i4(dst + i, r);
                        }
                    }
                    cost = warp_sum32(cost); crank = warp_sum32(crank); h = mix64(warp_sum64(h));
                    cost += (d == 0 ? ev.AT[var] + ipos(sv) : ev.AF[var] + ipos(-sv));
                    const int value = val + cost;
                    values[d] = value; hashes[d] = h;
                    if (lane == 0) {
                        ev.cand_rank[cb + c] = (uint32_t)crank;
                        ev.cand_agg[cb + c] = pack_key(value, c);
                        ev.cand_first[cb + c] = c;
                        ev.cand_inex[cb + c] = (uint8_t)(fl & NF_INEXACT);
                        ev.cand_cost[cb + c] = cost;
                    }
                }
                __threadfence();
                __syncwarp();
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const uint32_t c = d == 0 ? c_t : c_f;
                    const int value = values[d];
                    const unsigned long long h = hashes[d];
                    // open-addressing insert (next_l.entry(), clean.rs:738)
                    const uint32_t tag = (uint32_t)(h >> 32);
                    const unsigned long long entry = ((unsigned long long)tag << 32) | c;
                    uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
                    unsigned long long* tab = ev.table + (size_t)k * ev.T;
                    for (;;) {
                        unsigned long long old = 0;
                        if (lane == 0) old = atomicCAS(tab + slot, EMPTY64, entry);
                        old = __shfl_sync(FULL_MASK, old, 0);
                        if (old == EMPTY64) {  // Entry::Vacant, clean.rs:739-765
                            if (lane == 0) { ev.cand_rep[cb + c] = c; ev.cand_slot[cb + c] = slot; }
                            break;
                        }
                        if ((uint32_t)(old >> 32) == tag) {
                            const uint32_t oc = (uint32_t)old;
                            const int4* orow = reinterpret_cast<const int4*>(m2_row(ev, buf ^ 1, k, oc));
                            const int4* mrow = reinterpret_cast<const int4*>(m2_row(ev, buf ^ 1, k, c));
                            bool eq = true;
                            for (int i = lane; i < NW4; i += 32) { const int4 a = ld_cg_i4(orow + i), b = ld_cg_i4(mrow + i); eq = eq && a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
                            eq = __all_sync(FULL_MASK, eq);
                            if (eq) {  // Entry::Occupied, clean.rs:766-774 + append_edge_to! :199-220
                                if (lane == 0) {
                                    atomicMax(ev.cand_agg + cb + oc, pack_key(value, c));  // value_top = max, `>=`: the last (largest) candidate wins
                                    atomicMin(ev.cand_first + cb + oc, c);                   // canonical identity = first candidate
                                    if (fl & NF_INEXACT) ev.cand_inex[cb + oc] = 1;          // exact &= parent.exact
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
    }
    if (lane == 0 && exp_acc) { atomicAdd(&ev.ctl[cur_k].expanded, (unsigned long long)exp_acc); atomicAdd(&ev.ctl[cur_k].transitions, (unsigned long long)(2 * exp_acc)); }
}

// cut order between two distinct candidates (clean.rs:803-808 + heuristics.rs:33-37 + canonical tie-break): true if a is BETTER than b
__device__ inline bool m2_cand_better(const M2EV& ev, int buf, int k, size_t cb, uint32_t a, uint32_t b) {
    const unsigned long long ka = (ev.cand_agg[cb + a] & 0xFFFFFFFF00000000ull) | ev.cand_rank[cb + a];
    const unsigned long long kb = (ev.cand_agg[cb + b] & 0xFFFFFFFF00000000ull) | ev.cand_rank[cb + b];
    if (ka != kb) return ka > kb;
    const int32_t* ra = m2_row(ev, buf, k, a); const int32_t* rb = m2_row(ev, buf, k, b);
    for (int j = 0; j < ev.n; ++j) if (ra[j] != rb[j]) return ra[j] > rb[j];
    return false;
}

// =================================================================================================================
// m2_finish: one CTA per DD.  Decides everything about layer t (whose candidates were produced by m2_expand(t-1)).
// =================================================================================================================
__global__ void __launch_bounds__(1024, 1) m2_finish(M2EV ev, int t) {
    constexpr int NT = 1024;
    __shared__ FinishSmem sm;
    __shared__ int s_last;
    const int k = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DDCtl* ctl = ev.ctl + k;
    M2Aux* aux = ev.aux + k;
    const size_t cb = (size_t)k * ev.C;
    const size_t lb = (size_t)k * ev.Lmax;
    // cut keys / status bytes of the distinct candidates: shared memory when 2*Wcap of them fit (ev.smem_keys), else global scratch
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    unsigned long long* keys = ev.smem_keys ? reinterpret_cast<unsigned long long*>(dyn_smem) : ev.gkeys + cb;
    uint8_t* stat = ev.smem_keys ? dyn_smem + (size_t)ev.C * 8 : ev.ustat + cb;
    const int status = ctl->status;
    bool live = true;
    if (status == ST_DONE) live = false;
    else if (status == ST_TERMINAL) { __syncthreads(); if (tid == 0) ctl->status = ST_DONE; live = false; }
    if (live) do {
        const int ncand = t == 0 ? 1 : 2 * ctl->n_cur;
        __syncthreads();
        if (tid == 0) { ctl->ncand = ncand; ctl->lel_pending = 0; aux->cut_relaxed = 0; }
        const int per = (ncand + NT - 1) / NT;
        const int lo = min(tid * per, ncand), hi = min(lo + per, ncand);
        // ---- A. canonical representative of every distinct state = its first candidate (rule C1) ---------------------------------
        uint8_t* uniq = stat;
        for (int c = lo; c < hi; ++c) uniq[c] = 0;
        __syncthreads();
        for (int c = lo; c < hi; ++c) {
            if (ev.cand_rep[cb + c] == (uint32_t)c) {
                const uint32_t f = ev.cand_first[cb + c];
                uniq[f] = 1;
                if (f != (uint32_t)c) { ev.cand_agg[cb + f] = ev.cand_agg[cb + c]; ev.cand_inex[cb + f] = ev.cand_inex[cb + c]; }
            }
        }
        __syncthreads();
        int cnt = 0;
        for (int c = lo; c < hi; ++c) cnt += uniq[c];
        int U;
        const int off0 = block_excl_scan(cnt, &U, sm.scan);
        if (U == 0) {  // every node was pruned: empty layer (clean.rs:667-669) -> no best node
            if (tid == 0) { ctl->status = ST_DONE; ctl->t_term = t; ctl->has_best = 0; ctl->has_best_exact = 0; ev.nlog[lb + t] = 0; atomicSub(ev.active, 1); }
            break;
        }
        const int depth = ctl->root_depth + t;
        const bool terminal = depth >= ev.n;  // next_variable == None (model.rs:338-344): the layer is the terminal layer
        const int var = terminal ? -1 : ev.ord[ev.n - depth - 1];
        // ---- C. width cut ---------------------------------------------------------------------------------------------------------
        const int W = ctl->width, comp = ctl->comp_type;
        bool cut = false; int need = 0;
        if (!terminal) {
            if (comp == DDO_RESTRICTED && U > W) { cut = true; need = W; }                  // clean.rs:782-787
            else if (comp == DDO_RELAXED && U > W && t >= 2) { cut = true; need = W - 1; }  // clean.rs:788-793 (layers.len() > 1)
        }
        if (!cut && U > ev.Wcap) {
            if (tid == 0) { ctl->status = ST_DONE; ctl->overflow = 1; ctl->t_term = t; atomicSub(ev.active, 1); }
            break;
        }
        {
            int off = off0;
            for (int c = lo; c < hi; ++c) if (uniq[c]) {
                ev.ulist[cb + off] = (uint32_t)c;
                if (cut) keys[off] = (ev.cand_agg[cb + c] & 0xFFFFFFFF00000000ull) | ev.cand_rank[cb + c];  // (value_top, rank)
                ++off;
            }
        }
        // remember which candidates are canonical (uflag = 1) before the same bytes are reused as the per-distinct-candidate status
        // (0 undecided, 1 keep, 2 drop)
        for (int c = lo; c < hi; ++c) ev.uflag[cb + c] = uniq[c] ? 1 : 0;
        __syncthreads();
        if (cut) for (int ui = tid; ui < U; ui += NT) stat[ui] = 0;
        __syncthreads();
        if (cut) {
            int nactive = U;
            bool done = false;
            if (need == 0) { for (int ui = tid; ui < U; ui += NT) stat[ui] = 2; done = true; }
            const int nchunks = (ev.n + 1) / 2;
            for (int chunk = 0; chunk <= nchunks && !done; ++chunk) {
                // key chunk 0: (value_top, rank); chunk j: benefits 2j-2, 2j-1 as order-preserving unsigned words (canonical tie-break)
                auto key_of = [&](int ui) -> unsigned long long {
                    if (chunk == 0) return keys[ui];
                    const int32_t* r = m2_row(ev, t & 1, k, ev.ulist[cb + ui]) + 2 * (chunk - 1);
                    return ((unsigned long long)((uint32_t)r[0] ^ 0x80000000u) << 32) | ((uint32_t)r[1] ^ 0x80000000u);
                };
                unsigned long long kor = 0, kand = ~0ull;
                for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 0) { unsigned long long x = key_of(ui); kor |= x; kand &= x; }
                kor = block_reduce(kor, [](unsigned long long a, unsigned long long b) { return a | b; }, 0ull, sm.red64);
                kand = block_reduce(kand, [](unsigned long long a, unsigned long long b) { return a & b; }, ~0ull, sm.red64);
                const unsigned long long diff = kor ^ kand;
                for (int byte = 7; byte >= 0 && !done; --byte) {
                    if (((diff >> (8 * byte)) & 0xff) == 0) continue;
                    for (int i = tid; i < 256; i += NT) sm.hist[i] = 0;
                    __syncthreads();
                    for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 0) atomicAdd(&sm.hist[(key_of(ui) >> (8 * byte)) & 0xff], 1u);
                    __syncthreads();
                    if (warp == 0) {
                        int c8[8]; int s8 = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) { c8[q] = (int)sm.hist[255 - (lane * 8 + q)]; s8 += c8[q]; }
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
        // ---- D. stable positions of the survivors (rule C3) ------------------------------------------------------------------------
        int nkeep, kp;
        if (cut) {
            int kc = 0;
            for (int i = 0; i < cnt; ++i) kc += (stat[off0 + i] == 1);
            kp = block_excl_scan(kc, &nkeep, sm.scan);
            for (int i = 0; i < cnt; ++i) {
                const uint32_t c = ev.ulist[cb + off0 + i];
                if (stat[off0 + i] == 1) { ev.pos_of[cb + c] = (uint32_t)kp++; ev.uflag[cb + c] = 2; }
                else { ev.pos_of[cb + c] = comp == DDO_RELAXED ? M2_DROPPED : NONE32; ev.uflag[cb + c] = 0; }
            }
        } else {
            nkeep = U; kp = off0;
            for (int c = lo; c < hi; ++c) if (ev.uflag[cb + c] == 1) { ev.pos_of[cb + c] = (uint32_t)kp++; ev.uflag[cb + c] = 2; }
        }
        int n_next = nkeep;
        __syncthreads();
        // ---- E. relaxation: what the merge kernels need (clean.rs:826-876) ----------------------------------------------------------
        if (cut && comp == DDO_RELAXED) {
            // merged.value_top = max over the merged-away nodes d and their edges of parent.value_top + relax(cost)  (relax.rs:78-84:
            // cost + rank(d) - rank(merged))  =  max_d (value_top(d) + rank(d)) - rank(merged)
            unsigned long long mkey = 0;
            for (int ui = tid; ui < U; ui += NT) if (stat[ui] == 2) {
                const uint32_t c = ev.ulist[cb + ui];
                const unsigned long long a = ev.cand_agg[cb + c];
                mkey = max(mkey, pack_key(key_value(a) + (int32_t)ev.cand_rank[cb + c], (uint32_t)a));
            }
            mkey = block_reduce(mkey, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, sm.red64);
            if (ev.smem_keys) for (int ui = tid; ui < U; ui += NT) ev.ustat[cb + ui] = stat[ui];  // the merge kernels read the status from global memory
            for (int i = tid; i < ev.NW; i += NT) { ev.mrg_min[(size_t)k * ev.NW + i] = INT32_MAX; ev.mrg_max[(size_t)k * ev.NW + i] = INT32_MIN; }
            if (tid == 0) { aux->cut_relaxed = 1; aux->nkeep = nkeep; aux->U = U; aux->mkey = mkey; aux->mpos = nkeep; aux->rank_m = 0; }
            n_next = nkeep + 1;  // the merged node, or (recycled corner case, clean.rs:868-871) the saved node
        }
        // ---- F. terminal layer: best nodes (clean.rs:620-632, rule C4: last maximum) ------------------------------------------------
        if (terminal) {
            unsigned long long b_all = 0, b_ex = 0;
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
            ev.rslog[(lb + t) * 3] = -1; ev.rslog[(lb + t) * 3 + 1] = -1; ev.rslog[(lb + t) * 3 + 2] = 0;
            ctl->n_cur = n_next; ctl->var = var;
            if (cut && ctl->lel < 0) { ctl->lel = t - 1; ctl->lel_pending = 1; if (comp == DDO_RELAXED) *ev.lel_any = t; }  // _maybe_save_lel, clean.rs:796-800
            if (terminal) { ctl->status = ST_TERMINAL; ctl->t_term = t; atomicSub(ev.active, 1); }
        }
    } while (false);
    // ---- work plan of the flat kernels that follow: the last CTA to finish scans the per-DD tile counts (8 warps = 8 items per tile) ----
    const int count = gridDim.x;
    __syncthreads();
    if (tid == 0) { __threadfence(); s_last = (atomicAdd(ev.finish_counter, 1u) == (unsigned)count - 1u); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const volatile DDCtl* vc = ev.ctl;
    const int per = (count + blockDim.x - 1) / blockDim.x;
    const int lo = min(tid * per, count), hi = min(lo + per, count);
    int te = 0, tc = 0;
    for (int kk = lo; kk < hi; ++kk) {
        const int st = vc[kk].status;
        te += st == ST_ACTIVE ? (vc[kk].n_cur + 7) / 8 : 0;
        tc += (st == ST_ACTIVE || st == ST_TERMINAL) ? (vc[kk].ncand + 7) / 8 : 0;
    }
    int tote, totc;
    int oe = block_excl_scan(te, &tote, sm.scan);
    int oc = block_excl_scan(tc, &totc, sm.scan);
    for (int kk = lo; kk < hi; ++kk) {
        const int st = vc[kk].status;
        ev.tile_off_e[kk] = oe; ev.tile_off_c[kk] = oc;
        oe += st == ST_ACTIVE ? (vc[kk].n_cur + 7) / 8 : 0;
        oc += (st == ST_ACTIVE || st == ST_TERMINAL) ? (vc[kk].ncand + 7) / 8 : 0;
    }
    if (tid == 0) { ev.tile_off_e[count] = tote; ev.tile_off_c[count] = totc; *ev.finish_counter = 0; }
}

// =================================================================================================================
// m2_merge: Relaxation::merge (relax.rs:46-77) needs, per variable, min and max of the benefit over the merged-away states:
// merged[v] = min if min > 0 (all positive), max if max < 0 (all negative), else 0 (a zero, or both signs).
// grid = (chunks of 64 distinct candidates, DDs); thread j owns the 128-bit column chunk j of every row of its chunk.
// =================================================================================================================
__global__ void __launch_bounds__(256) m2_merge(M2EV ev, int t) {
    const int k = blockIdx.y;
    const M2Aux* aux = ev.aux + k;
    if (!aux->cut_relaxed) return;
    const int U = aux->U;
    const int u0 = blockIdx.x * 64;
    if (u0 >= U) return;
    const int u1 = min(u0 + 64, U);
    const size_t cb = (size_t)k * ev.C;
    const int j = threadIdx.x;
    if (j >= ev.NW4) return;
    int4 mn = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX), mx = make_int4(INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN);
    bool any = false;
    for (int ui = u0; ui < u1; ++ui) {
        if (ev.ustat[cb + ui] != 2) continue;
        const uint32_t c = ev.ulist[cb + ui];
        const int4 v = ld_stream_i4(reinterpret_cast<const int4*>(m2_row(ev, t & 1, k, c)) + j);
        mn.x = min(mn.x, v.x); mn.y = min(mn.y, v.y); mn.z = min(mn.z, v.z); mn.w = min(mn.w, v.w);
        mx.x = max(mx.x, v.x); mx.y = max(mx.y, v.y); mx.z = max(mx.z, v.z); mx.w = max(mx.w, v.w);
        any = true;
    }
    if (!any) return;
    int32_t* pmn = ev.mrg_min + (size_t)k * ev.NW + 4 * j; int32_t* pmx = ev.mrg_max + (size_t)k * ev.NW + 4 * j;
    atomicMin(pmn, mn.x); atomicMin(pmn + 1, mn.y); atomicMin(pmn + 2, mn.z); atomicMin(pmn + 3, mn.w);
    atomicMax(pmx, mx.x); atomicMax(pmx + 1, mx.y); atomicMax(pmx + 2, mx.z); atomicMax(pmx + 3, mx.w);
}

// =================================================================================================================
// m2_merge_fin: merged state, its rank and value, recycled-node lookup (clean.rs:830), merged node / saved node (clean.rs:832-876)
// =================================================================================================================
__global__ void __launch_bounds__(256) m2_merge_fin(M2EV ev, int t) {
    constexpr int NT = 256;
    const int k = blockIdx.x;
    M2Aux* aux = ev.aux + k;
    if (!aux->cut_relaxed) return;
    DDCtl* ctl = ev.ctl + k;
    __shared__ int s_red[8];
    __shared__ unsigned long long s_h[8];
    __shared__ int s_recycled;
    __shared__ uint32_t s_best[NT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t cb = (size_t)k * ev.C, lb = (size_t)k * ev.Lmax;
    int32_t* mrow = ev.mrg_min + (size_t)k * ev.NW;
    const int32_t* mxr = ev.mrg_max + (size_t)k * ev.NW;
    int rank = 0; unsigned long long h = 0;
    for (int i = tid; i < ev.NW; i += NT) {
        const int mn = mrow[i], mx = mxr[i];
        const int m = (i < ev.n) ? (mn > 0 ? mn : (mx < 0 ? mx : 0)) : 0;
        mrow[i] = m;
        rank += iabs(m);
        h += (unsigned long long)(uint32_t)m * ev.hmul[i];  // 32-bit odd multipliers, 64-bit accumulator
    }
    rank = warp_sum32(rank); h = warp_sum64(h);
    if (lane == 0) { s_red[warp] = rank; s_h[warp] = h; }
    __syncthreads();
    rank = 0; h = 0;
    for (int w = 0; w < NT / 32; ++w) { rank += s_red[w]; h += s_h[w]; }
    h = mix64(h);
    const int rank_m = rank;
    const unsigned long long mkey = aux->mkey;
    const int value_m = key_value(mkey) - rank_m;
    const int nkeep = aux->nkeep, U = aux->U;
    // recycled ? a KEPT node whose state equals the merged state
    if (warp == 0) {
        const uint32_t tag = (uint32_t)(h >> 32);
        uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
        const unsigned long long* tab = ev.table + (size_t)k * ev.T;
        int recycled = -1;
        for (;;) {
            const unsigned long long e = tab[slot];
            if (e == EMPTY64) break;
            if ((uint32_t)(e >> 32) == tag) {
                const uint32_t oc = (uint32_t)e;
                const int32_t* orow = m2_row(ev, t & 1, k, oc);
                bool eq = true;
                for (int i = lane; i < ev.NW; i += 32) eq = eq && orow[i] == mrow[i];
                eq = __all_sync(FULL_MASK, eq);
                if (eq) { const uint32_t f = ev.cand_first[cb + oc]; const uint32_t p = ev.pos_of[cb + f]; if (p != NONE32 && p != M2_DROPPED) recycled = (int)f; break; }
            }
            slot = (slot + 1) & (uint32_t)(ev.T - 1);
        }
        if (lane == 0) s_recycled = recycled;
    }
    __syncthreads();
    const int recycled = s_recycled;
    if (recycled >= 0) {
        // clean.rs:868-871: the best merged-away node ("saved") stays in the layer, un-deleted, next to the recycled node
        uint32_t bestc = NONE32;
        for (int ui = tid; ui < U; ui += NT) if (ev.ustat[cb + ui] == 2) {
            const uint32_t c = ev.ulist[cb + ui];
            if (bestc == NONE32 || m2_cand_better(ev, t & 1, k, cb, c, bestc)) bestc = c;
        }
        s_best[tid] = bestc;
        __syncthreads();
        for (int d = NT / 2; d > 0; d >>= 1) {
            if (tid < d) {
                const uint32_t a = s_best[tid], b2 = s_best[tid + d];
                if (a == NONE32 || (b2 != NONE32 && m2_cand_better(ev, t & 1, k, cb, b2, a))) s_best[tid] = b2;
            }
            __syncthreads();
        }
        const uint32_t saved = s_best[0];
        const int r_pos = (int)ev.pos_of[cb + recycled];
        if (tid == 0) {
            // the recycled node receives every relaxed edge: RELAXED flag, value_top = max (`>=`: the appended edges win ties)
            const unsigned long long rk = ev.cand_agg[cb + recycled];
            if (value_m >= key_value(rk)) ev.cand_agg[cb + recycled] = pack_key(value_m, (uint32_t)mkey);
            ev.cand_inex[cb + recycled] |= (uint8_t)(NF_INEXACT | NF_RELAXED);
            ev.pos_of[cb + saved] = (uint32_t)nkeep; ev.uflag[cb + saved] = 2;
            ev.rslog[(lb + t) * 3] = nkeep; ev.rslog[(lb + t) * 3 + 1] = r_pos; ev.rslog[(lb + t) * 3 + 2] = (int32_t)ev.cand_rank[cb + saved] - rank_m;
            aux->mpos = r_pos; aux->rank_m = rank_m;
        }
    } else {
        // new merged node (clean.rs:832-849) written straight into the next layer
        const int nbuf = t & 1;
        const size_t nb = (size_t)k * ev.Wcap + nkeep;
        int32_t* drow = m2_row(ev, nbuf, k, (uint32_t)ev.C);  // the extra row of the layer's candidate buffer
        for (int i = tid; i < ev.NW; i += NT) drow[i] = mrow[i];
        if (tid == 0) {
            ev.cur_src[nbuf][nb] = (uint32_t)ev.C;
            ev.cur_val[nbuf][nb] = value_m;
            ev.cur_flag[nbuf][nb] = (uint8_t)(NF_INEXACT | NF_RELAXED);
            ev.cur_rank[nbuf][nb] = rank_m;
            ev.plog[(lb + t) * ev.Wcap + nkeep] = ((uint32_t)mkey & PLOG_CAND_MASK) | PLOG_INEXACT | PLOG_RELAXED;
            aux->mpos = nkeep; aux->rank_m = rank_m;
        }
    }
    (void)ctl;
}

// =================================================================================================================
// m2_compact: scatter layer t into the ping-pong buffers, write logs, release hash slots, snapshot the LEL.  One warp per candidate.
// =================================================================================================================
__global__ void __launch_bounds__(256) m2_compact(M2EV ev, int t, int count) {
    // tiles of 8 candidates (the work plan's granularity): 32 tiles = 256 candidates per CTA pass, one THREAD per candidate (rows are not
    // moved any more); the last-exact-layer snapshot, taken once per relaxed DD, copies rows with one warp per row.
    const int total = ev.tile_off_c[count];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int tile0 = blockIdx.x * 32; tile0 < total; tile0 += gridDim.x * 32) {
        {
            const int tile = tile0 + (threadIdx.x >> 3);
            if (tile < total) {
                const int k = plan_find(ev.tile_off_c, count, tile);
                const DDCtl* ctl = ev.ctl + k;
                const M2Aux* aux = ev.aux + k;
                const int c = (tile - ev.tile_off_c[k]) * 8 + (threadIdx.x & 7);
                if (c < ctl->ncand) {
                    const size_t cb = (size_t)k * ev.C;
                    const size_t lb = (size_t)k * ev.Lmax;
                    const int nbuf = t & 1;
                    const bool relaxed = ctl->comp_type == DDO_RELAXED;
                    const uint32_t rep = ev.cand_rep[cb + c];
                    uint32_t child = NONE32;
                    int cost = 0;
                    if (rep != NONE32) {
                        const uint32_t f = ev.cand_first[cb + rep];
                        child = ev.pos_of[cb + f];
                        cost = ev.cand_cost[cb + c];
                        if (child == M2_DROPPED) {  // edge re-pointed to the merged node with its relaxed cost (clean.rs:851-866, relax.rs:78-84)
                            child = (uint32_t)aux->mpos;
                            cost += (int32_t)ev.cand_rank[cb + rep] - aux->rank_m;
                        }
                        if (rep == (uint32_t)c) {  // release the hash slot this candidate claimed
                            const uint32_t slot = ev.cand_slot[cb + c];
                            if (slot != NONE32) ev.table[(size_t)k * ev.T + slot] = EMPTY64;
                        }
                        if (ev.uflag[cb + c] == 2) {  // surviving canonical representative: becomes node `pos` of layer t
                            const uint32_t pos = ev.pos_of[cb + c];
                            const size_t nb = (size_t)k * ev.Wcap + pos;
                            const unsigned long long key = ev.cand_agg[cb + c];
                            const uint32_t fl = ev.cand_inex[cb + c];
                            ev.cur_src[nbuf][nb] = (uint32_t)c;  // the row stays where m2_expand wrote it
                            ev.cur_val[nbuf][nb] = key_value(key);
                            ev.cur_flag[nbuf][nb] = (uint8_t)fl;
                            ev.cur_rank[nbuf][nb] = (int32_t)ev.cand_rank[cb + c];
                            ev.plog[(lb + t) * ev.Wcap + pos] = ((uint32_t)key & PLOG_CAND_MASK) | ((fl & NF_INEXACT) ? PLOG_INEXACT : 0u) | ((fl & NF_RELAXED) ? PLOG_RELAXED : 0u);
                        }
                    }
                    if (t > 0 && relaxed) { ev.clog[(lb + t - 1) * ev.C + c] = child; ev.colog[(lb + t - 1) * ev.C + c] = cost; }
                }
            }
        }
        if (t > 0 && *ev.lel_any == t) {  // layer t-1 is the last exact layer of some relaxed DD: keep its nodes (states, values, rough upper bounds) for the cutset
            for (int q = warp; q < 32 * 4; q += 8) {  // 32 tiles x 4 even candidates
                const int tile = tile0 + (q >> 2);
                if (tile >= total) break;
                const int k = plan_find(ev.tile_off_c, count, tile);
                const DDCtl* ctl = ev.ctl + k;
                if (!ctl->lel_pending || ctl->comp_type != DDO_RELAXED) continue;
                const int c = (tile - ev.tile_off_c[k]) * 8 + 2 * (q & 3);
                if (c >= ctl->ncand) continue;
                const int i = c >> 1;
                const size_t pb = (size_t)k * ev.Wcap + i;
                const int4* src = reinterpret_cast<const int4*>(m2_row(ev, (t - 1) & 1, k, ev.cur_src[(t - 1) & 1][pb]));
                int4* dst = reinterpret_cast<int4*>(ev.lel_state + pb * ev.NW);
                for (int j = lane; j < ev.NW4; j += 32) st_stream_i4(dst + j, ld_stream_i4(src + j));
                if (lane == 0) { ev.lel_val[pb] = ev.cur_val[(t - 1) & 1][pb]; ev.lel_rub[pb] = ev.cur_rub[pb]; }
            }
        }
    }
}

// =================================================================================================================
// m2_finalize: exact-best-path walk (clean.rs:634-655) and decision bits of the best / best exact path (clean.rs:329-343)
// =================================================================================================================
__global__ void m2_finalize(M2EV ev, int count) {
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
            const uint32_t e = ev.plog[(lb + tt) * ev.Wcap + pos];
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
            const uint32_t cand = ev.plog[(lb + tt) * ev.Wcap + pos] & PLOG_CAND_MASK;
            if (!(cand & 1u)) out[(tt - 1) >> 6] |= 1ull << ((tt - 1) & 63);  // even candidate = first decision of the domain = T
            pos = (int)(cand >> 1);
        }
    }
}

// =================================================================================================================
// m2_bottomup: local bounds of a relaxed DD (clean.rs:448-475) as a per-layer GATHER over the child / edge-cost logs, then the
// cutset upper bounds ub = min(value_top + rub, value_top + value_bot, best_value) of the last exact layer (clean.rs:426-428).
// =================================================================================================================
__global__ void __launch_bounds__(1024, 1) m2_bottomup(M2EV ev) {
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
        const int s = ev.rslog[(lb + tt + 1) * 3], r = ev.rslog[(lb + tt + 1) * 3 + 1], delta = ev.rslog[(lb + tt + 1) * 3 + 2];
        const uint32_t* cl = ev.clog + (lb + tt) * ev.C;
        const int32_t* co = ev.colog + (lb + tt) * ev.C;
        for (int i = tid; i < n; i += NT) {
            int32_t best = UNMARKED;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const uint32_t ch = cl[2 * i + d];
                if (ch == NONE32) continue;
                const int cost = co[2 * i + d];
                int32_t x = nxt[ch];
                if (x != UNMARKED) best = max(best, x + cost);
                if ((int)ch == s && r >= 0) { x = nxt[r]; if (x != UNMARKED) best = max(best, x + cost + delta); }  // the saved node's edges were also copied (relaxed) to the recycled node
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
            const long long a = (long long)val + ev.lel_rub[nb + i], b = (long long)val + vbot;
            ev.cs_ub[nb + i] = (int32_t)min(min(a, b), (long long)ctl->best_value);
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
__global__ void __launch_bounds__(1024, 1) m2_cutset_count(M2EV ev, DrainOut o, const long long* ub_cap, const long long* lb_filter, int count) {
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
// one warp per last-exact-layer node: 2 KB state row + value / ub / path bits
__global__ void __launch_bounds__(256) m2_cutset_write(M2EV ev, DrainOut o, const long long* ub_cap, int pw) {
    const int k = blockIdx.y;
    const DDCtl* ctl = ev.ctl + k;
    if (o.count[k] == 0) return;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= ctl->lel_n) return;
    const size_t nb = (size_t)k * ev.Wcap;
    const uint32_t loc = o.loc[nb + i];
    if (loc == NONE32) return;
    const size_t rec = (size_t)o.offset[k] + loc;
    const int4* src = reinterpret_cast<const int4*>(ev.lel_state + (nb + i) * ev.NW);
    int4* dst = reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(o.state) + rec * ev.NW);
    for (int q = lane; q < ev.NW4; q += 32) dst[q] = src[q];
    if (lane != 0) return;
    o.val[rec] = ev.lel_val[nb + i];
    o.ub[rec] = (int32_t)min((long long)ev.cs_ub[nb + i], ub_cap[k]);
    o.dd[rec] = k;
    const size_t lb = (size_t)k * ev.Lmax;
    for (int w = 0; w < pw; ++w) o.path[rec * pw + w] = 0;
    int pos = i;
    for (int tt = ctl->lel; tt >= 1; --tt) {
        const uint32_t cand = ev.plog[(lb + tt) * ev.Wcap + pos] & PLOG_CAND_MASK;
        if (!(cand & 1u)) o.path[rec * pw + ((tt - 1) >> 6)] |= 1ull << ((tt - 1) & 63);
        pos = (int)(cand >> 1);
    }
}

}  // namespace ddo

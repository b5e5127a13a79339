// m2s_frontier.cuh -- FRONTIER cutset of a relaxed DD on the device, MAX2SAT engine (sm_100a).  Same scheme as frontier.cuh:
//   * m2_fc_sweep   Mdd::_compute_frontier_cutset (clean.rs:586-606) fused with Mdd::_compute_local_bounds (clean.rs:448-475) over all
//                   layers: a gather over the child / edge-cost logs, members appended in canonical order (layer descending, position
//                   ascending);
//   * m2_fc_eval    upper bound of every member, ub = min(value_top + rub, value_top + value_bot, best_value) (clean.rs:426-428);
//   * m2_fc_write   Mdd::_drain_cutset (clean.rs:417-445): state, value_top, ub, depth and path bits of the emitted nodes.
// A frontier node is exact, so its state is a function of its best path: one warp replays Max2Sat::transition (examples/max2sat/
// model.rs:275-292) along the path -- two clause-weight rows of 2 KB per layer from the L2-resident instance tables -- and sums the edge
// costs logged by m2_compact (value_top of an exact node = root value + costs of its best path, clean.rs:199-220).
#pragma once
#include "frontier.cuh"
#include "m2s_kernels.cuh"

namespace ddo {

__global__ void __launch_bounds__(1024, 1) m2_fc_sweep(M2EV ev) {
    __shared__ int scan[40];
    const int k = blockIdx.x;
    DDCtl* ctl = ev.ctl + k;
    const int tid = threadIdx.x, NT = blockDim.x;
    const bool run = ctl->comp_type == DDO_RELAXED && !ctl->overflow && ctl->has_best && ctl->lel >= 0;
    const int T = ctl->t_term;
    __syncthreads();
    if (tid == 0) { ctl->cutset_count = 0; ctl->lel_n = 0; }
    if (!run) return;
    const size_t lb = (size_t)k * ev.Lmax;
    const size_t nb = (size_t)k * ev.Wcap;
    int32_t* nxt = ev.vb[0] + nb;
    int32_t* cur = ev.vb[1] + nb;
    uint32_t* out_node = ev.fc_node + (size_t)k * ev.fc_cap;
    int32_t* out_vbot = ev.fc_aux + (size_t)k * ev.fc_cap;
    for (int i = tid; i < ev.nlog[lb + T]; i += NT) nxt[i] = 0;  // terminal layer: value_bot = 0, MARKED (clean.rs:451-455)
    __syncthreads();
    int base = 0;
    for (int tt = T - 1; tt >= 0; --tt) {
        const int n = ev.nlog[lb + tt];
        const int s = ev.rslog[(lb + tt + 1) * 3], r = ev.rslog[(lb + tt + 1) * 3 + 1], delta = ev.rslog[(lb + tt + 1) * 3 + 2];
        const uint32_t* cl = ev.clog + (lb + tt) * ev.C;
        const int32_t* co = ev.colog + (lb + tt) * ev.C;
        const uint32_t* pl = ev.plog + (lb + tt) * ev.Wcap;
        const uint32_t* pl_next = ev.plog + (lb + tt + 1) * ev.Wcap;
        const int per = (n + NT - 1) / NT, lo = min(tid * per, n), hi = min(lo + per, n);
        int c = 0;
        for (int i = lo; i < hi; ++i) {
            int32_t best = UNMARKED;
            bool inexact_child = false;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const uint32_t ch = cl[2 * i + d];
                if (ch == NONE32) continue;
                const int cost = co[2 * i + d];
                int32_t x = nxt[ch];
                if (x != UNMARKED) best = max(best, x + cost);
                if (pl_next[ch] & PLOG_INEXACT) inexact_child = true;
                if ((int)ch == s && r >= 0) {  // the saved node's edges were also copied (relaxed) to the recycled node (clean.rs:851-871)
                    x = nxt[r];
                    if (x != UNMARKED) best = max(best, x + cost + delta);
                    inexact_child = true;
                }
            }
            cur[i] = best;
            const bool member = best != UNMARKED && inexact_child && !(pl[i] & PLOG_INEXACT);
            ev.cs_marked[nb + i] = member;
            c += member;
        }
        int total;
        int off = base + block_excl_scan(c, &total, scan);
        for (int i = lo; i < hi; ++i)
            if (ev.cs_marked[nb + i]) { out_node[off] = ((uint32_t)tt << FC_POS_BITS) | (uint32_t)i; out_vbot[off] = cur[i]; ++off; }
        base += total;
        int32_t* tmp = nxt; nxt = cur; cur = tmp;
        __syncthreads();
    }
    if (tid == 0) { ctl->cutset_count = base; ctl->lel_n = base; }
}

// One warp replays the best path of frontier node `node` of DD k.  Every lane walks the parent log (broadcast loads) for the decision
// bits and the value; the state row is spread over the lanes (CH 128-bit chunks each) and rebuilt root -> node like m2_expand does.
// bits: >= 16 words, zeroed by the caller; bit t = decision T in layer t.
template <int CH>
__device__ __forceinline__ void m2_fc_replay(const M2EV& ev, int k, uint32_t node, int4 (&s)[CH], long long& val, uint64_t* bits, int lane) {
    const size_t lb = (size_t)k * ev.Lmax;
    const DDCtl* ctl = ev.ctl + k;
    const int tt = (int)(node >> FC_POS_BITS);
    int pos = (int)(node & FC_POS_MASK);
    val = ctl->root_value;
    for (int t = tt; t >= 1; --t) {
        const uint32_t cand = ev.plog[(lb + t) * ev.Wcap + pos] & PLOG_CAND_MASK;
        if (!(cand & 1u)) bits[(t - 1) >> 6] |= 1ull << ((t - 1) & 63);  // even candidate = T
        val += ev.colog[(lb + t - 1) * ev.C + cand];
        pos = (int)(cand >> 1);
    }
    const int NW4 = ev.NW4;
    const int4* root = reinterpret_cast<const int4*>(ev.root_state + (size_t)k * ev.NW);
#pragma unroll
    for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; s[q] = i < NW4 ? root[i] : make_int4(0, 0, 0, 0); }
    for (int t = 0; t < tt; ++t) {
        const int var = ev.ord[ev.n - (ctl->root_depth + t) - 1];  // model.rs:330-348
        const bool isT = (bits[t >> 6] >> (t & 63)) & 1ull;
        const int4* P = reinterpret_cast<const int4*>((isT ? ev.PT : ev.PF) + (size_t)var * ev.NW);
        const int4* Q = reinterpret_cast<const int4*>((isT ? ev.QT : ev.QF) + (size_t)var * ev.NW);
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            const int i = lane + 32 * q;
            if (i < NW4) {
                const int4 p = __ldg(P + i), m = __ldg(Q + i);
                int4 x = s[q];
                x.x += p.x - m.x; x.y += p.y - m.y; x.z += p.z - m.z; x.w += p.w - m.w;  // model.rs:282-290
                if (i == (var >> 2)) { const int e = var & 3; if (e == 0) x.x = 0; else if (e == 1) x.y = 0; else if (e == 2) x.z = 0; else x.w = 0; }  // ret[k] = 0
                s[q] = x;
            }
        }
    }
}

template <int CH>
__global__ void __launch_bounds__(256) m2_fc_eval(M2EV ev) {
    const int k = blockIdx.y;
    const DDCtl* ctl = ev.ctl + k;
    const int cnt = ctl->cutset_count;
    const size_t fb = (size_t)k * ev.fc_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < cnt; r += gridDim.x * 8) {
        const uint32_t node = ev.fc_node[fb + r];
        int4 s[CH];
        uint64_t bits[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) bits[q] = 0;
        long long val;
        m2_fc_replay<CH>(ev, k, node, s, val, bits, lane);
        int rank = 0;
#pragma unroll
        for (int q = 0; q < CH; ++q) rank += iabs(s[q].x) + iabs(s[q].y) + iabs(s[q].z) + iabs(s[q].w);
        rank = warp_sum32(rank);
        if (lane == 0) {
            const int depth = ctl->root_depth + (int)(node >> FC_POS_BITS);
            const long long rub = (long long)rank + ev.est[depth] - ev.initial + ev.nk[depth];  // fast_upper_bound, model.rs:240-249
            const long long vbot = ev.fc_aux[fb + r];
            ev.fc_ub[fb + r] = (int32_t)min(min(val + rub, val + vbot), (long long)ctl->best_value);
        }
    }
}

template <int CH>
__global__ void __launch_bounds__(256) m2_fc_write(M2EV ev, DrainOut o, const long long* ub_cap, int pw) {
    const int k = blockIdx.y;
    if (o.count[k] == 0) return;
    const DDCtl* ctl = ev.ctl + k;
    const int cnt = ctl->cutset_count;
    const size_t fb = (size_t)k * ev.fc_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + warp; r < cnt; r += gridDim.x * 8) {
        const int loc = ev.fc_aux[fb + r];
        if (loc < 0) continue;
        const size_t rec = (size_t)o.offset[k] + loc;
        const uint32_t node = ev.fc_node[fb + r];
        int4 s[CH];
        uint64_t bits[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) bits[q] = 0;
        long long val;
        m2_fc_replay<CH>(ev, k, node, s, val, bits, lane);
        int4* dst = reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(o.state) + rec * ev.NW);
#pragma unroll
        for (int q = 0; q < CH; ++q) { const int i = lane + 32 * q; if (i < ev.NW4) dst[i] = s[q]; }
        if (lane == 0) {
            o.val[rec] = (int32_t)val;
            o.ub[rec] = (int32_t)min((long long)ev.fc_ub[fb + r], ub_cap[k]);
            o.dd[rec] = k;
            o.tt[rec] = (int32_t)(node >> FC_POS_BITS);
            for (int q = 0; q < pw; ++q) o.path[rec * pw + q] = bits[q];
        }
    }
}

}  // namespace ddo

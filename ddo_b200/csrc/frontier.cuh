// frontier.cuh -- FRONTIER cutset of a relaxed DD on the device (MISP engine), sm_100a.
//
// Replaces, for engines created with DDO_FRONTIER:
//   * Mdd::_compute_frontier_cutset (ddo/src/implementation/mdd/clean.rs:586-606): the cutset is every EXACT node with at least one edge
//     into an INEXACT node, collected bottom-up over the edges;
//   * Mdd::_compute_local_bounds (clean.rs:448-475) over ALL layers (a frontier node may sit in any layer above the terminal one);
//   * Mdd::_drain_cutset (clean.rs:417-445) for nodes of different depths: state, value_top, rough upper bound and best path of every
//     emitted node, ub = min(value_top + rub, value_top + value_bot, best_value).
//
// The engine keeps no per-layer states (a layer lives in the ping-pong buffers while it is expanded), only the parent log `plog` and the
// child log `clog`.  A frontier node is EXACT, so its state, value_top and rough upper bound are functions of any root path: they are
// re-derived by walking the best-parent chain and replaying misp `transition` / `transition_cost` (examples/misp/main.rs:77-93) and
// `fast_upper_bound` (main.rs:191-193) -- depth x 64..128 B of L2-resident reads per emitted node instead of 64 B x every node of the DD
// in HBM.
//
// Canonical order (C7): the reference pushes frontier nodes in the order the bottom-up edge walk meets them (hash-order dependent); the
// device and the oracle emit them by (layer descending, position in the layer ascending).
#pragma once
#include "kernels.cuh"

namespace ddo {

constexpr uint32_t FC_POS_MASK = (1u << FC_POS_BITS) - 1;

// =================================================================================================================
// k_fc_sweep: one CTA per DD.  Bottom-up over the layers: value_bot of every node as a gather over the child log (as k_bottomup), and in
// the same pass the frontier test of clean.rs:586-606 -- node exact, some child inexact -- restricted to MARKED nodes (the only ones
// drain_cutset emits, clean.rs:424).  Members are appended in canonical order with one block scan per layer.
// =================================================================================================================
static __global__ void __launch_bounds__(1024, 1) k_fc_sweep(EV ev) {
    __shared__ int scan[40];
    const int k = blockIdx.x;
    DDCtl* ctl = ev.ctl + k;
    const int tid = threadIdx.x, NT = blockDim.x;
    const bool run = ctl->comp_type == DDO_RELAXED && !ctl->overflow && ctl->has_best && ctl->lel >= 0;
    const int T = ctl->t_term;
    __syncthreads();
    if (tid == 0) { ctl->cutset_count = 0; ctl->lel_n = 0; }
    if (!run) return;  // restricted / exact DDs and DDs that were never squashed have an empty frontier
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
        const int wv = ev.weight[ev.vlog[lb + tt]];
        const int s = ev.rslog[(lb + tt + 1) * 2], r = ev.rslog[(lb + tt + 1) * 2 + 1];
        const uint32_t* cl = ev.clog + (lb + tt) * ev.C;
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
                const int cost = d == 0 ? wv : 0;
                int32_t x = nxt[ch];
                if (x != UNMARKED) best = max(best, x + cost);
                if (pl_next[ch] & PLOG_INEXACT) inexact_child = true;
                if ((int)ch == s && r >= 0) {  // edges of the saved node were also copied to the recycled node, which is relaxed (clean.rs:851-871)
                    x = nxt[r];
                    if (x != UNMARKED) best = max(best, x + cost);
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

// Replays the best-parent chain of node `node` of DD k from the DD root: state (main.rs:77-85), value_top (main.rs:87-93) and, when
// `bits` is given, the decision bits of the path (bit t = YES in layer t).
template <int S>
__device__ __forceinline__ void fc_walk(const EV& ev, int k, uint32_t node, uint64_t (&w)[S], int& val, uint64_t* bits) {
    const size_t lb = (size_t)k * ev.Lmax;
    int pos = (int)(node & FC_POS_MASK);
#pragma unroll
    for (int j = 0; j < S; ++j) w[j] = ev.root_state[(size_t)k * S + j];
    val = ev.ctl[k].root_value;
    for (int t = (int)(node >> FC_POS_BITS); t >= 1; --t) {
        const uint32_t cand = ev.plog[(lb + t) * ev.Wcap + pos] & PLOG_CAND_MASK;
        const int v = ev.vlog[lb + t - 1];
        const int vw = v >> 6;
        const uint64_t bit = 1ull << (v & 63);
#pragma unroll
        for (int j = 0; j < S; ++j) if (j == vw) w[j] &= ~bit;  // res.remove(var), main.rs:79
        if (!(cand & 1u)) {                                      // even candidate = YES
            const uint64_t* row = ev.nc + (size_t)v * S;
#pragma unroll
            for (int j = 0; j < S; ++j) w[j] &= __ldg(row + j);  // main.rs:82
            val += ev.weight[v];
            if (bits) bits[(t - 1) >> 6] |= 1ull << ((t - 1) & 63);
        }
        pos = (int)(cand >> 1);
    }
}

// =================================================================================================================
// k_fc_eval: one thread per frontier record: ub = min(value_top + rub, value_top + value_bot, best_value) (clean.rs:426-428)
// =================================================================================================================
template <int S>
__global__ void __launch_bounds__(256) k_fc_eval(EV ev) {
    const int k = blockIdx.y;
    const DDCtl* ctl = ev.ctl + k;
    const int cnt = ctl->cutset_count;
    const size_t fb = (size_t)k * ev.fc_cap;
    for (int r = blockIdx.x * 256 + threadIdx.x; r < cnt; r += gridDim.x * 256) {
        uint64_t w[S];
        int val;
        fc_walk<S>(ev, k, ev.fc_node[fb + r], w, val, nullptr);
        int rub = 0;
        if (ev.unit_weights) {
#pragma unroll
            for (int j = 0; j < S; ++j) rub += __popcll(w[j]);
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) { uint64_t x = w[j]; const int32_t* wp = ev.weight + j * 64; while (x) { const int b = __ffsll((long long)x) - 1; rub += wp[b]; x &= x - 1; } }
        }
        const int vbot = ev.fc_aux[fb + r];
        ev.fc_ub[fb + r] = min(min(val + rub, val + vbot), ctl->best_value);
    }
}

// =================================================================================================================
// drain: count (the solver's filter min(ub, ub_cap) > lb_filter, parallel.rs:460-461) -> offsets (k_cutset_offsets) -> write
// =================================================================================================================
template <class EVT>
__global__ void __launch_bounds__(1024, 1) k_fc_count(EVT ev, DrainOut o, const long long* ub_cap, const long long* lb_filter, int count) {
    __shared__ int scan[40];
    const int k = blockIdx.x;
    const DDCtl* ctl = ev.ctl + k;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int n = k < count ? ctl->cutset_count : 0;
    const size_t fb = (size_t)k * ev.fc_cap;
    const int per = (n + NT - 1) / NT, lo = min(tid * per, n), hi = min(lo + per, n);
    const long long cap = ub_cap[k], lbf = lb_filter[k];
    int c = 0;
    for (int i = lo; i < hi; ++i) c += min((long long)ev.fc_ub[fb + i], cap) > lbf;
    int total;
    int off = block_excl_scan(c, &total, scan);
    for (int i = lo; i < hi; ++i) {
        const bool f = min((long long)ev.fc_ub[fb + i], cap) > lbf;
        ev.fc_aux[fb + i] = f ? off++ : -1;
    }
    if (tid == 0) o.count[k] = total;
}

template <int S>
__global__ void __launch_bounds__(256) k_fc_write(EV ev, DrainOut o, const long long* ub_cap, int pw) {
    const int k = blockIdx.y;
    if (o.count[k] == 0) return;
    const DDCtl* ctl = ev.ctl + k;
    const int cnt = ctl->cutset_count;
    const size_t fb = (size_t)k * ev.fc_cap;
    for (int r = blockIdx.x * 256 + threadIdx.x; r < cnt; r += gridDim.x * 256) {
        const int loc = ev.fc_aux[fb + r];
        if (loc < 0) continue;
        const size_t rec = (size_t)o.offset[k] + loc;
        const uint32_t node = ev.fc_node[fb + r];
        uint64_t w[S], bits[16];  // a frontier node lies above the terminal layer: at most n - 1 <= 1023 decisions
        int val;
#pragma unroll
        for (int q = 0; q < 16; ++q) bits[q] = 0;
        fc_walk<S>(ev, k, node, w, val, bits);
#pragma unroll
        for (int j = 0; j < S; ++j) o.state[rec * S + j] = w[j];
        o.val[rec] = val;
        o.ub[rec] = (int32_t)min((long long)ev.fc_ub[fb + r], ub_cap[k]);
        o.dd[rec] = k;
        o.tt[rec] = (int32_t)(node >> FC_POS_BITS);
        for (int q = 0; q < pw; ++q) o.path[rec * pw + q] = bits[q];
    }
}

}  // namespace ddo

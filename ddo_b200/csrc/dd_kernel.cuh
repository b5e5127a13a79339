// dd_kernel.cuh -- k_dd: the whole of Mdd::compile (ddo/src/implementation/mdd/clean.rs:345-381 and everything it calls per layer) for
// one decision diagram inside ONE persistent kernel (MISP device model, sm_100a).
//
// Round 1 ran a layer step of the whole batch as three dependent launches (k_finish / k_compact / k_expand); a solve of BASELINE config 2
// is ~2 800 such steps and the chain, not bandwidth, was its critical path.  Here a thread-block CLUSTER owns one DD from its root to its
// terminal layer: the layer loop runs inside the kernel, the phases of a layer are separated by cluster barriers instead of launches, and
// the DDs of a batch advance independently -- clusters pull DD slots from a device-side queue, a restricted DD that takes its first width
// cut publishes its relaxed twin to the same queue (parallel.rs:419-430: the relaxed DD is compiled only when the restricted one is
// inexact).  The cluster size is chosen per launch: wide clusters when the batch holds a few wide DDs (latency), single CTAs when it holds
// hundreds (throughput).
//
// The layer step is restated around one property of MISP under `next_variable` = vertex occurring in the fewest states
// (misp/main.rs:109-143): most nodes of a layer do NOT contain the branching vertex (mean out-degree 1.18 on config 2), so their only child
// has their own state, value and exactness (main.rs:77-102).  Such an IDENTITY candidate is never materialised: the dedup table entry, the
// width-cut keys and the commit of the next layer refer to the parent's row; its hash accumulator, popcount and rough upper bound are
// carried over; only nodes that contain the vertex (and pruned nodes) touch their 64..128-byte rows before the commit.  The per-vertex
// occurrence counts behind `next_variable` are maintained INCREMENTALLY (rows that leave the layer are subtracted, rows that enter are
// added) instead of being recounted over every distinct state of every layer.
//
// Everything a layer step decides from lives in SHARED memory, distributed over the CTAs of the cluster: the node records of the two
// layer buffers (hash accumulator, value_top, popcount, flags) and the candidate records (value_top / best parent, first candidate,
// exactness, claimer-or-slot word, position) belong to the CTA that owns the node; a duplicate found in another CTA's slice is merged by
// remote atomics on that CTA's shared memory (atom.shared::cluster).  Only the state rows, the dedup table and the logs stay in global
// memory (L2-resident), so the dependent chain of a phase is one L2 round trip per row instead of one per record.
//
// Synchronisation budget of a layer (cluster barriers): 3 without a width cut (candidates inserted / counts exchanged / layer committed),
// 6-7 with one.  Everything exchanged between the CTAs of a cluster is PUSHED into the peers' shared memory before a barrier (remote
// stores / remote atomics, fire and forget) and read locally after it -- a pulled value costs a ~215-cycle DSMEM round trip per peer.  The
// width cut is one histogram pass over a dense (value_top, popcount) key; the candidates of the boundary bucket are gathered into rank 0,
// which resolves BitSet::cmp among them with block barriers only (a generic MSD radix select over the whole cluster remains as the
// fall-back for key ranges or buckets that do not fit).
//
// Semantics are those of kernels.cuh (same canonical rules C1-C4, same logs plog / clog / nlog / vlog / rslog / lel_* and DDCtl fields), so
// k_finalize, k_bottomup, the cutset drains and the FRONTIER kernels consume a DD compiled here unchanged.
#pragma once
#include "kernels.cuh"

namespace ddo {

constexpr int DD_NT = 512, DD_NW = DD_NT / 32;
constexpr uint32_t DD_IDENT = 1u << 30;          // candidate flag: the state row of this candidate is its parent's row in the current layer
constexpr uint32_t DD_CMASK = DD_IDENT - 1u;
constexpr int DD_NB = 2048;                      // bins of the dense (value_top, popcount) histogram of the width cut
constexpr int DD_GCAP = 2048;                    // boundary-bucket candidates rank 0 can resolve alone
constexpr int DD_MAXCS = 16;
constexpr uint32_t RS_CLAIM = 1u << 31;          // claimer-or-slot word of a candidate: NONE32 | RS_CLAIM | ident | hash slot | ident | claimer

// byte offsets into the dynamic shared memory of a CTA (computed on the host, dd_layout())
struct DDLayout {
    int slice, maxch, capc, weighted;  // nodes of a layer per CTA, 32-node chunks per CTA, candidates per CTA, weighted instance (rough upper bounds kept)
    unsigned o_D, o_master, o_stage, o_cnt, o_off, o_koff, o_fb, o_kb, o_keys, o_ulist, o_stat, o_lh, o_gh;
    unsigned o_agg, o_first, o_rs, o_f, o_pos, o_inex, o_pcy, o_nm, o_rub, total;
};

struct DDFixed {
    int scan[40];
    unsigned long long red64[40];
    unsigned int hist[2][256];
    unsigned int ghist[256];
    unsigned long long merged[17];               // OR of the merged-away states of this CTA (+ their best key in [16])
    unsigned long long mx[DD_MAXCS][17];         // ... of every CTA of the cluster (pushed)
    unsigned long long xch[2][8][DD_MAXCS];      // exchange slots [phase][slot][source rank] (pushed, double-buffered across barriers)
    int kbx[2][DD_MAXCS];                        // boundary-bucket candidates kept per rank (pushed by rank 0)
    int kbc[DD_MAXCS];
    int misc[16];
    int red4[DD_NW][4];
    int job;
    int gcnt;                                    // rank 0: candidates gathered so far
    int ucnt;                                    // undecided candidates of this CTA
    unsigned long long cnt[2];                   // expanded nodes / transitions of the DD being compiled (this CTA's share)
    unsigned int tour[DD_NT];
    long long prof[16]; long long prof_t;       // DDO_DD_PROF: cycles of rank 0 / thread 0 per phase
};

// ---- hashing: multilinear over the 32-bit halves, h = fin(sum_j w32[j] * m32[j]); the accumulator is stored with every node so that the
// NO child of a node (one bit cleared) and an identity child cost no pass over the row
__device__ __forceinline__ uint32_t dd_mul32(int j) { return (uint32_t)(mix64(0x9E3779B97F4A7C15ULL * (uint64_t)(j + 1)) >> 32) | 1u; }
template <int S> __device__ __forceinline__ uint64_t dd_hacc(const uint64_t (&w)[S]) {
    uint64_t a = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) a += (uint64_t)(uint32_t)w[j] * dd_mul32(2 * j) + (uint64_t)(uint32_t)(w[j] >> 32) * dd_mul32(2 * j + 1);
    return a;
}
__device__ __forceinline__ uint64_t dd_hfin(uint64_t a) { a ^= a >> 32; a *= 0x9E3779B97F4A7C15ULL; a ^= a >> 29; return a; }

__device__ __forceinline__ int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }

// ---- device work queue: q[0] next primary, q[1] / q[2] head / tail of the twin queue, q[3] DDs finished or never needed, q[4] total
__device__ int dd_fetch_job(const EV& ev, int count) {
    int* q = ev.dq;
    const long long t0 = clock64();
    for (;;) {
        if (clock64() - t0 > 20000000000ll) { q[5] = 1; return -1; }  // watchdog (~10 s): a lost job must not hang the device
        const int t2 = ld_volatile_i32(q + 2), h2 = ld_volatile_i32(q + 1);
        if (h2 < t2) {  // twins first: they are the wide DDs of the batch
            if (atomicCAS(q + 1, h2, h2 + 1) == h2) {
                int s;
                while ((s = ld_volatile_i32(ev.dq_jobs + h2)) < 0) __nanosleep(100);
                return s;
            }
            continue;
        }
        const int h1 = ld_volatile_i32(q);
        if (h1 < count) {
            if (atomicCAS(q, h1, h1 + 1) == h1) return h1;
            continue;
        }
        if (ld_volatile_i32(q + 3) >= ld_volatile_i32(q + 4)) return -1;
        __nanosleep(500);
    }
}

// =================================================================================================================
// k_dd_init: DD control blocks (clean.rs:383-405) and the queue
// =================================================================================================================
static __global__ void k_dd_init(EV ev, int count, int comp_type, long long best_lb, int dual) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int slots = dual ? 2 * count : count;
    if (k == 0) { ev.dq[0] = 0; ev.dq[1] = 0; ev.dq[2] = 0; ev.dq[3] = 0; ev.dq[4] = slots; ev.dq[5] = 0; *ev.active = 0; }
    if (k >= slots) return;
    ev.dq_jobs[k] = -1;
    const bool twin = k >= count;
    const int p = twin ? k - count : k;
    DDCtl c{};
    c.status = twin ? ST_WAITING : ST_ACTIVE; c.ncand = 0; c.n_cur = 0; c.var = -1;
    c.width = ev.root_width[p]; c.comp_type = twin ? DDO_RELAXED : comp_type; c.root_depth = ev.root_depth[p]; c.lel = -1;
    c.t_term = -1; c.best_pos = -1; c.best_exact_pos = -1; c.root_value = ev.root_val[p];
    c.best_lb = best_lb; c.primary = -1; c.fork_t = -1;  // a twin compiled here owns all its logs (it restarts from the root)
    ev.ctl[k] = c;
}

// =================================================================================================================
// the per-DD compiler
// =================================================================================================================
template <int S>
struct DDC {
    static constexpr int W32 = 2 * S;          // 32-bit words of a state row
    static constexpr int SROW = W32 + 1;       // padded row stride of the histogram staging buffers (conflict-free column reads)
    const EV& ev; const DDLayout& L; DDFixed& fx; unsigned char* dsm;
    cg::cluster_group cl;
    unsigned CS, rank;
    int tid, lane, warp;
    int xphase = 0, hphase = 0, kphase = 0;
    // shared-memory views
    int* D[2]; unsigned int* master; uint32_t* stage; int* cnt_s; int* off_s; int* koff_s; uint2* fb_s; uint2* kb_s;
    unsigned long long* keys; uint32_t* ulist; uint8_t* stat;
    unsigned int* lh; unsigned int* gh; unsigned long long* garr; unsigned long long* gkeys; uint8_t* gstat; uint32_t* und;
    // records of this CTA's slice: candidates [cbase, cbase + capc), nodes [node0, node0 + slice)
    uint32_t* s_val32; uint32_t* s_bp; uint32_t* s_first; uint32_t* s_rs; uint32_t* s_f; uint32_t* s_pos; uint8_t* s_inex; uint16_t* s_pcy;
    uint4* s_nm[2]; int32_t* s_rub[2];
    unsigned cbase; int node0;
    // the DD
    int k, rk; size_t cb, lb, nb; DDCtl* ctl;
    int comp, W; long long best_lb;
    unsigned long long* tabs[2];
    // histogram staging of this warp
    int st_cnt = 0; unsigned st_minus = 0; int hcnt[W32];

    __device__ DDC(const EV& e, const DDLayout& l, DDFixed& f, unsigned char* d) : ev(e), L(l), fx(f), dsm(d), cl(cg::this_cluster()) {
        CS = cl.num_blocks(); rank = cl.block_rank();
        tid = threadIdx.x; lane = tid & 31; warp = tid >> 5;
        D[0] = reinterpret_cast<int*>(dsm + L.o_D); D[1] = D[0] + ev.HN;
        master = reinterpret_cast<unsigned int*>(dsm + L.o_master);
        stage = reinterpret_cast<uint32_t*>(dsm + L.o_stage) + (size_t)warp * 32 * SROW;
        cnt_s = reinterpret_cast<int*>(dsm + L.o_cnt); off_s = reinterpret_cast<int*>(dsm + L.o_off); koff_s = reinterpret_cast<int*>(dsm + L.o_koff);
        fb_s = reinterpret_cast<uint2*>(dsm + L.o_fb); kb_s = reinterpret_cast<uint2*>(dsm + L.o_kb);
        keys = reinterpret_cast<unsigned long long*>(dsm + L.o_keys); ulist = reinterpret_cast<uint32_t*>(dsm + L.o_ulist); stat = dsm + L.o_stat;
        lh = reinterpret_cast<unsigned int*>(dsm + L.o_lh); gh = reinterpret_cast<unsigned int*>(dsm + L.o_gh);
        // scratch of the boundary-bucket resolution, aliased with regions that are idle during the width cut: the gathered list (rank 0)
        // over the staging buffers (empty between the expansion and the commit), its keys over the cut keys, the undecided list over lh
        garr = reinterpret_cast<unsigned long long*>(dsm + L.o_stage); gstat = dsm + L.o_stage + (size_t)DD_GCAP * 8;
        gkeys = keys; und = lh;
        s_val32 = reinterpret_cast<uint32_t*>(dsm + L.o_agg); s_bp = s_val32 + L.capc; s_first = reinterpret_cast<uint32_t*>(dsm + L.o_first);
        s_rs = reinterpret_cast<uint32_t*>(dsm + L.o_rs); s_f = reinterpret_cast<uint32_t*>(dsm + L.o_f); s_pos = reinterpret_cast<uint32_t*>(dsm + L.o_pos);
        s_inex = dsm + L.o_inex; s_pcy = reinterpret_cast<uint16_t*>(dsm + L.o_pcy);
        s_nm[0] = reinterpret_cast<uint4*>(dsm + L.o_nm); s_nm[1] = s_nm[0] + L.slice;
        s_rub[0] = reinterpret_cast<int32_t*>(dsm + L.o_rub); s_rub[1] = s_rub[0] + L.slice;
        cbase = rank * (unsigned)L.capc; node0 = (int)rank * L.slice;
    }

    // ---- records of ANY candidate / node of the DD: the owner's shared memory (local or remote through DSMEM) --------------------------
    template <class T> __device__ __forceinline__ T* cptr(T* arr, uint32_t c) const {
        const unsigned o = (c >> 1) / (unsigned)L.slice;
        T* p = arr + (c - o * (unsigned)L.capc);
        return o == rank ? p : cl.map_shared_rank(p, o);
    }
    template <class T> __device__ __forceinline__ T* nptr(T* arr, int i) const {
        const unsigned o = (unsigned)i / (unsigned)L.slice;
        T* p = arr + (i - (int)o * L.slice);
        return o == rank ? p : cl.map_shared_rank(p, o);
    }
    // (biased value_top << 32 | best parent candidate) of claimer `rep`.  The two halves are separate 32-bit words updated by NATIVE 32-bit
    // atomics: a 64-bit atomic on shared memory is a lock-based CAS loop (ATOMS.CAST.SPIN) that atomics arriving from other CTAs of the
    // cluster do not respect -- measured as lost value_top updates.  value_top = max is taken while the candidates are inserted, the best
    // parent (largest candidate among those reaching the final value: `>=`, last tie wins, clean.rs:215) one phase later.
    __device__ __forceinline__ unsigned long long agg_of(uint32_t rep) const { return ((unsigned long long)*cptr(s_val32, rep) << 32) | *cptr(s_bp, rep); }
    __device__ __forceinline__ static uint32_t vbias(int v) { return (uint32_t)v ^ 0x80000000u; }
    // claimer (= the candidate that holds value_top / exactness of the state) of candidate c, from its claimer-or-slot word
    __device__ __forceinline__ static uint32_t rep_of(uint32_t c, uint32_t rs) { return (rs & RS_CLAIM) ? c : (rs & DD_CMASK); }

    __device__ __forceinline__ void csync() {
        if (ev.dd_dbg & 4) { __threadfence(); fence_cluster(); __syncthreads(); }
        cl.sync();
        if (ev.dd_dbg & 4) { fence_cluster(); __syncthreads(); }
    }
    // ---- cluster exchange of up to 6 block-uniform values: pushed into every peer, read locally after the barrier ---------------------
    __device__ __forceinline__ void push(int nvals, unsigned long long v0, unsigned long long v1 = 0, unsigned long long v2 = 0, unsigned long long v3 = 0,
                                         unsigned long long v4 = 0, unsigned long long v5 = 0) {
        if (tid < nvals * (int)CS) {
            const int slot = tid / (int)CS; const unsigned r = tid % CS;
            const unsigned long long v = slot == 0 ? v0 : slot == 1 ? v1 : slot == 2 ? v2 : slot == 3 ? v3 : slot == 4 ? v4 : v5;
            *cl.map_shared_rank(&fx.xch[xphase][slot][rank], r) = v;
        }
    }
    __device__ __forceinline__ unsigned long long got(int slot, unsigned r) const { return fx.xch[xphase][slot][r]; }
    __device__ __forceinline__ void xnext() { xphase ^= 1; }
    template <class Op> __device__ unsigned long long allreduce1(unsigned long long v, Op op) {
        push(1, v); csync();
        unsigned long long acc = got(0, 0);
        for (unsigned r = 1; r < CS; ++r) acc = op(acc, got(0, r));
        xnext();
        return acc;
    }
    // sum / max over the block of 4 ints per thread (results broadcast)
    __device__ void block_sum4(int (&v)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = warp_reduce(v[q], [](int a, int b) { return a + b; });
        __syncthreads();
        if (lane == 0) for (int q = 0; q < 4; ++q) fx.red4[warp][q] = v[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) { int s = 0; for (int w2 = 0; w2 < DD_NW; ++w2) s += fx.red4[w2][q]; v[q] = s; }
    }
    __device__ void block_max4(int (&v)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = warp_reduce(v[q], [](int a, int b) { return a > b ? a : b; });
        __syncthreads();
        if (lane == 0) for (int q = 0; q < 4; ++q) fx.red4[warp][q] = v[q];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) { int s = INT32_MIN; for (int w2 = 0; w2 < DD_NW; ++w2) s = max(s, fx.red4[w2][q]); v[q] = s; }
    }

    // ---- rows ---------------------------------------------------------------------------------------------------------------
    // state row of candidate c of the layer being built (parents in cur_state[buf])
    __device__ __forceinline__ const uint64_t* cand_row(uint32_t c_with_flag, int buf) const {
        const uint32_t c = c_with_flag & DD_CMASK;
        return (c_with_flag & DD_IDENT) ? ev.cur_state[buf] + (nb + (c >> 1)) * S : ev.cand_state + (cb + c) * S;
    }
    __device__ __forceinline__ void load_row_cg(const uint64_t* p, uint64_t (&w)[S]) const {
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) { const uint4 x = ld_cg_u4(q + j); w[2 * j] = u4lo(x); w[2 * j + 1] = u4hi(x); }
    }
    __device__ __forceinline__ void store_row(uint64_t* p, const uint64_t (&w)[S]) const {
        uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
        for (int j = 0; j < S / 2; ++j) q[j] = mk_u4(w[2 * j], w[2 * j + 1]);
    }
    __device__ __forceinline__ int row_rub(const uint64_t (&w)[S], int pc) const {  // fast_upper_bound, misp/main.rs:191-193
        if (ev.unit_weights) return pc;
        int r = 0;
#pragma unroll
        for (int j = 0; j < S; ++j) { uint64_t x = w[j]; const int32_t* wp = ev.weight + j * 64; while (x) { const int b = __ffsll((long long)x) - 1; r += wp[b]; x &= x - 1; } }
        return r;
    }

    // ---- histogram staging: rows that enter (+) or leave (-) the set of distinct states are buffered per warp and bit-transposed 32 at a
    // time into per-lane counters (lane b of column j counts vertex 32 j + b): ONE transpose serves both signs (after it, bit r of a lane's
    // word belongs to staged row r, and st_minus marks the rows that leave).  Flushed into D[e] at the end of an expansion.
    __device__ void stage_flush() {
        if (st_cnt == 0) return;
        __syncwarp();
        const bool mine = lane < st_cnt;
        const unsigned minus = st_minus;
#pragma unroll
        for (int j = 0; j < W32; ++j) {
            const uint32_t tr = warp_transpose32(mine ? stage[lane * SROW + j] : 0u);
            hcnt[j] += __popc(tr & ~minus) - __popc(tr & minus);
        }
        __syncwarp();
        st_cnt = 0; st_minus = 0;
    }
    // warp-collective: every lane with `has` appends its row (minus: it leaves the layer)
    __device__ void stage_rows(bool has, const uint64_t (&w)[S], bool minus) {
        const unsigned m = __ballot_sync(FULL_MASK, has);
        if (!m) return;
        const int add = __popc(m);
        if (st_cnt + add > 32) stage_flush();
        if (has) {
            const int r = st_cnt + __popc(m & ((1u << lane) - 1u));
#pragma unroll
            for (int j = 0; j < S; ++j) { stage[r * SROW + 2 * j] = (uint32_t)w[j]; stage[r * SROW + 2 * j + 1] = (uint32_t)(w[j] >> 32); }
        }
        if (minus) st_minus |= (add == 32 ? 0xFFFFFFFFu : ((1u << add) - 1u)) << st_cnt;
        st_cnt += add;
    }
    __device__ void stage_commit(int e) {  // end of an expansion: counters of this lane -> D[e]
        stage_flush();
#pragma unroll
        for (int j = 0; j < W32; ++j) if (hcnt[j]) { atomicAdd(&D[e][32 * j + lane], hcnt[j]); hcnt[j] = 0; }
    }

    // ---- open-addressing insert (next_l.entry(), clean.rs:738-775 + append_edge_to! :199-220) ---------------------------------------
    // returns the claimer-or-slot word of candidate c: RS_CLAIM | slot when it claimed a slot (a new distinct state), else the claimer
    template <class RowFn>
    __device__ __forceinline__ uint32_t insert(unsigned long long* tab, uint64_t hacc, uint32_t c, uint32_t ident, int value, uint32_t fl, int buf, RowFn my_row) {
        const uint64_t h = dd_hfin(hacc);
        const uint32_t tag = (uint32_t)(h >> 32);
        const unsigned long long entry = ((unsigned long long)tag << 32) | c | ident;
        uint32_t slot = (uint32_t)h & (uint32_t)(ev.T - 1);
        uint64_t w[S]; bool loaded = false;
        for (;;) {
            const unsigned long long old = atomicCAS(tab + slot, EMPTY64, entry);
            if (old == EMPTY64) return RS_CLAIM | ident | slot;
            if ((uint32_t)(old >> 32) == tag) {
                if (!loaded) { my_row(w); loaded = true; }
                const uint32_t oe = (uint32_t)old;
                const uint4* orow = reinterpret_cast<const uint4*>(cand_row(oe, buf));
                bool eq = true;
#pragma unroll
                for (int q = 0; q < S / 2; ++q) { const uint4 o4 = ld_cg_u4(orow + q); eq = eq && u4lo(o4) == w[2 * q] && u4hi(o4) == w[2 * q + 1]; }
                if (eq) {
                    const uint32_t oc = oe & DD_CMASK;
                    atomicMax(cptr(s_val32, oc), vbias(value));                 // value_top = max
                    atomicMin(cptr(s_first, oc), c);                            // canonical identity = first candidate (rule C1)
                    if (fl & NF_INEXACT) *cptr(s_inex, oc) = 1;                 // exact &= parent.exact
                    return ident | oc;
                }
            }
            slot = (slot + 1) & (uint32_t)(ev.T - 1);
        }
    }

    // ---- child log of the PREVIOUS layer step (relaxed DDs: edge (parent c / 2, decision) -> node of the layer committed last) and release
    // of the hash slots its candidates claimed; runs at the start of the next step, off the critical path (the tables alternate by layer parity)
    __device__ void trail(int tp, int n_prev, bool relaxed) {
        const int slice_lo = min(node0, n_prev), slice_hi = min(slice_lo + L.slice, n_prev);
        const int nch = (slice_hi - slice_lo + 31) >> 5;
        unsigned long long* tab = tabs[tp & 1];
        for (int ch = warp; ch < nch; ch += DD_NW) {
            const int i = slice_lo + 32 * ch + lane;
            if (i < slice_hi) {
                const unsigned lc = 2u * (unsigned)(i - node0);
                const uint2 rs = *reinterpret_cast<const uint2*>(s_rs + lc);
                if (rs.x != NONE32 && (rs.x & RS_CLAIM)) tab[rs.x & DD_CMASK] = EMPTY64;
                if (rs.y != NONE32 && (rs.y & RS_CLAIM)) tab[rs.y & DD_CMASK] = EMPTY64;
                if (relaxed) {
                    const uint2 f = *reinterpret_cast<const uint2*>(s_f + lc);
                    uint32_t cy = NONE32, cn = NONE32;
                    if (f.x != NONE32) cy = *cptr(s_pos, f.x);
                    if (f.y != NONE32) cn = *cptr(s_pos, f.y);
                    *reinterpret_cast<uint2*>(ev.clog + (lb + tp) * ev.C + 2u * i) = make_uint2(cy, cn);
                }
            }
        }
    }

    // ---- E: expansion of layer t (clean.rs:360-370, :728-776; misp/main.rs:77-102,191-193) --------------------------------------------
    __device__ void expand(int t, int n, int var) {
        const int buf = t & 1, e = t & 1;
        unsigned long long* tab = tabs[t & 1];
        const int slice_lo = min(node0, n), slice_hi = min(slice_lo + L.slice, n);
        const int nch = (slice_hi - slice_lo + 31) >> 5;
        const int vw = var >> 6;
        const uint64_t bit = 1ull << (var & 63);
        const int wv = ev.weight[var];
        const uint64_t hdelta = (uint64_t)(1u << (var & 31)) * (uint64_t)dd_mul32(var >> 5);  // hash accumulator of the branching vertex's bit
        uint64_t ncr[S];  // complement-adjacency row of the branching vertex (misp/main.rs:82), once per thread and layer
        {
            const uint4* q = reinterpret_cast<const uint4*>(ev.nc + (size_t)var * S);
#pragma unroll
            for (int j = 0; j < S / 2; ++j) { const uint4 x = __ldg(q + j); ncr[2 * j] = u4lo(x); ncr[2 * j + 1] = u4hi(x); }
        }
        int my_exp = 0, my_tr = 0, no_claims = 0;
        for (int ch = warp; ch < nch; ch += DD_NW) {
            const int i = slice_lo + 32 * ch + lane;
            const bool active = i < slice_hi;
            uint64_t w[S];
            uint32_t rs_y = NONE32, rs_n = NONE32;
            bool minus_parent = false, plus_yes = false;
            uint64_t wy[S];
            if (active) {
                const int il = i - node0;
                const unsigned lc = 2u * (unsigned)il;
                const uint4 m = s_nm[buf][il];
                const uint64_t hacc = (uint64_t)m.x | ((uint64_t)m.y << 32);
                const int val = (int)m.z, pc = (int)(m.w & 0xFFFFu);
                const uint32_t fl = m.w >> 16;
                if (ev.dd_dbgbuf) { atomicAdd(ev.dd_dbgbuf + 4 * ev.Lmax + 4 * t, val); atomicAdd(ev.dd_dbgbuf + 4 * ev.Lmax + 4 * t + 1, pc); atomicAdd(ev.dd_dbgbuf + 4 * ev.Lmax + 4 * t + 2, (int)fl); atomicAdd(ev.dd_dbgbuf + 4 * ev.Lmax + 4 * t + 3, (int)(hacc & 0xFFFF)); }
                const int rub = ev.unit_weights ? pc : s_rub[buf][il];
                const bool expandable = ((long long)rub + (long long)val) > best_lb;  // clean.rs:364-365
                const uint64_t* prow = ev.cur_state[buf] + (nb + i) * S;
                if (!expandable) {
                    load_row_cg(prow, w); minus_parent = true;  // the node leaves without a child
                } else {
                    const bool has_v = (__ldcg(prow + vw) & bit) != 0;  // misp/main.rs:96
                    ++my_exp; my_tr += has_v ? 2 : 1;
                    const uint32_t c_yes = 2u * i, c_no = 2u * i + 1u;  // for_each_in_domain order: YES then NO (main.rs:95-102)
                    if (!has_v) {
                        // identity candidate: the child IS the parent's state (value, exactness, hash, popcount carried over)
                        s_val32[lc + 1] = vbias(val); s_bp[lc + 1] = 0;
                        s_first[lc + 1] = c_no;
                        s_inex[lc + 1] = (uint8_t)(fl & NF_INEXACT);
                        fence_cluster();
                        rs_n = insert(tab, hacc, c_no, DD_IDENT, val, fl, buf, [&](uint64_t (&r)[S]) { load_row_cg(prow, r); });
                        if (!(rs_n & RS_CLAIM)) { load_row_cg(prow, w); minus_parent = true; }  // its state is already counted through the claimer
                    } else {
                        load_row_cg(prow, w);
#pragma unroll
                        for (int j = 0; j < S; ++j) if (j == vw) w[j] &= ~bit;  // res.remove(var), main.rs:79: w is now the NO child
                        int pcy = 0;
#pragma unroll
                        for (int j = 0; j < S; ++j) { wy[j] = w[j] & ncr[j]; pcy += __popcll(wy[j]); }  // main.rs:82
                        const uint64_t hy = dd_hacc<S>(wy), hn = hacc - hdelta;
                        store_row(ev.cand_state + (cb + c_yes) * S, wy);
                        store_row(ev.cand_state + (cb + c_no) * S, w);
                        const int valy = val + wv;  // main.rs:87-93
                        *reinterpret_cast<uint2*>(s_val32 + lc) = make_uint2(vbias(valy), vbias(val)); *reinterpret_cast<uint2*>(s_bp + lc) = make_uint2(0u, 0u);
                        *reinterpret_cast<uint2*>(s_first + lc) = make_uint2(c_yes, c_no);
                        *reinterpret_cast<uchar2*>(s_inex + lc) = make_uchar2((uint8_t)(fl & NF_INEXACT), (uint8_t)(fl & NF_INEXACT));
                        s_pcy[il] = (uint16_t)pcy;
                        fence_cluster();
                        rs_y = insert(tab, hy, c_yes, 0u, valy, fl, buf, [&](uint64_t (&r)[S]) {
#pragma unroll
                            for (int j = 0; j < S; ++j) r[j] = wy[j]; });
                        rs_n = insert(tab, hn, c_no, 0u, val, fl, buf, [&](uint64_t (&r)[S]) {
#pragma unroll
                            for (int j = 0; j < S; ++j) r[j] = w[j]; });
                        plus_yes = (rs_y & RS_CLAIM) != 0;
                        if (rs_n & RS_CLAIM) ++no_claims;  // parent row out, NO row in: only the branching vertex loses an occurrence
                        else minus_parent = true;
#pragma unroll
                        for (int j = 0; j < S; ++j) if (j == vw) w[j] |= bit;  // back to the parent's row for the histogram
                    }
                }
                *reinterpret_cast<uint2*>(s_rs + lc) = make_uint2(rs_y, rs_n);
            }
            stage_rows(minus_parent, w, true);
            stage_rows(plus_yes, wy, false);
        }
        stage_commit(e);
        // counters
        my_exp = warp_reduce(my_exp, [](int a, int b) { return a + b; });
        my_tr = warp_reduce(my_tr, [](int a, int b) { return a + b; });
        no_claims = warp_reduce(no_claims, [](int a, int b) { return a + b; });
        if (lane == 0) {
            if (no_claims) atomicAdd(&D[e][var], -no_claims);
            if (my_exp) { atomicAdd(&fx.cnt[0], (unsigned long long)my_exp); atomicAdd(&fx.cnt[1], (unsigned long long)my_tr); }
        }
    }

    // popcount / claimer of ANY candidate of the layer being built (with its DD_IDENT flag)
    __device__ int cand_pc(uint32_t cf, int buf) const {
        const uint32_t c = cf & DD_CMASK;
        const int i = (int)(c >> 1);
        if (!(cf & DD_IDENT) && !(c & 1u)) return (int)*nptr(s_pcy, i);
        const int pc = (int)(nptr(s_nm[buf], i)->w & 0xFFFFu);
        return (cf & DD_IDENT) ? pc : pc - 1;
    }
    __device__ uint32_t cand_rep(uint32_t cf) const { const uint32_t c = cf & DD_CMASK; return rep_of(c, *cptr(s_rs, c)); }
    // compare two distinct candidates (with their DD_IDENT flags) by the cut order (clean.rs:803-808 + misp/main.rs:205-208)
    __device__ bool better(uint32_t a, uint32_t b, int buf) const {
        const int va = key_value(agg_of(cand_rep(a))), vb2 = key_value(agg_of(cand_rep(b)));
        if (va != vb2) return va > vb2;
        const int pca = cand_pc(a, buf), pcb = cand_pc(b, buf);
        if (pca != pcb) return pca > pcb;
        const uint64_t* ra = cand_row(a, buf); const uint64_t* rb = cand_row(b, buf);
        for (int j = 0; j < S; ++j) { const uint64_t xa = lex_word(__ldcg(ra + j)), xb = lex_word(__ldcg(rb + j)); if (xa != xb) return xa > xb; }
        return false;
    }

    // ---- MSD radix select of the `need` best among the undecided (st == 0) entries of kk / st / list[0 .. n): 1 = keep, 2 = drop.
    // Chunk 0 of a key is what kk holds on entry (used when first_chunk == 0), chunk j >= 1 packs the next MK members of the state
    // (member_key).  CLUSTER: the entries are spread over the CTAs of the cluster (digit histograms summed through DSMEM, one cluster
    // barrier per digit); otherwise this CTA alone, block barriers only.
    template <bool CLUSTER>
    __device__ void radix_select(unsigned long long* kk, uint8_t* st, const uint32_t* list, const unsigned long long* list64, int n, int need, int first_chunk, int buf) {
        bool done = false;
        if (need == 0) { for (int li = tid; li < n; li += DD_NT) if (st[li] == 0) st[li] = 2; done = true; }
        for (int chunk = first_chunk; chunk <= MemberKey<S>::CHUNKS && !done; ++chunk) {
            if (chunk > 0) {  // the next MK members of every still undecided state
                for (int li = tid; li < n; li += DD_NT) if (st[li] == 0) {
                    uint64_t w[S];
                    load_row_cg(cand_row(list ? list[li] : (uint32_t)list64[li], buf), w);
                    kk[li] = member_key<S>(w, MemberKey<S>::MK * (chunk - 1));
                }
                __syncthreads();
            }
            unsigned long long k0 = 0ull, k1 = 0ull;
            for (int li = tid; li < n; li += DD_NT) if (st[li] == 0) { const unsigned long long x = kk[li]; k0 |= x; k1 |= ~x; }
            k0 = block_reduce(k0, [](unsigned long long a, unsigned long long b) { return a | b; }, 0ull, fx.red64);
            k1 = block_reduce(k1, [](unsigned long long a, unsigned long long b) { return a | b; }, 0ull, fx.red64);
            if (CLUSTER) {
                push(2, k0, k1); csync();
                k0 = 0; k1 = 0;
                for (unsigned r = 0; r < CS; ++r) { k0 |= got(0, r); k1 |= got(1, r); }
                xnext();
            }
            const unsigned long long diff = k0 ^ ~k1;
            for (int byte = 7; byte >= 0 && !done; --byte) {
                if (((diff >> (8 * byte)) & 0xff) == 0) continue;
                unsigned int* h = fx.hist[hphase];
                for (int i = tid; i < 256; i += DD_NT) h[i] = 0;
                __syncthreads();
                for (int li = tid; li < n; li += DD_NT) if (st[li] == 0) atomicAdd(&h[(kk[li] >> (8 * byte)) & 0xff], 1u);
                if (CLUSTER) {
                    csync();
                    for (int i = tid; i < 256; i += DD_NT) {
                        unsigned int a = 0;
                        for (unsigned r = 0; r < CS; ++r) a += *cl.map_shared_rank(&h[i], r);
                        fx.ghist[i] = a;
                    }
                    hphase ^= 1;
                } else {
                    __syncthreads();
                    for (int i = tid; i < 256; i += DD_NT) fx.ghist[i] = h[i];
                }
                __syncthreads();
                if (warp == 0) {
                    int c8[8]; int s8 = 0;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { c8[q] = (int)fx.ghist[255 - (lane * 8 + q)]; s8 += c8[q]; }
                    int inc = s8;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int nn = __shfl_up_sync(FULL_MASK, inc, d); if (lane >= d) inc += nn; }
                    const int before = inc - s8;
                    if (before < need && need <= inc) {
                        int acc = before;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (acc < need && need <= acc + c8[q]) { fx.misc[0] = 255 - (lane * 8 + q); fx.misc[1] = acc; fx.misc[2] = c8[q]; }
                            acc += c8[q];
                        }
                    }
                }
                __syncthreads();
                const int b = fx.misc[0], above = fx.misc[1], inb = fx.misc[2];
                need -= above;
                const bool all_keep = (need == inb);
                for (int li = tid; li < n; li += DD_NT) if (st[li] == 0) {
                    const int d = (int)((kk[li] >> (8 * byte)) & 0xff);
                    if (d > b) st[li] = 1; else if (d < b) st[li] = 2; else if (all_keep) st[li] = 1;
                }
                __syncthreads();
                if (all_keep) done = true;
            }
        }
        __syncthreads();
    }

    // phase timer (thread 0 of rank 0; ev.dd_prof != nullptr only in profiling runs)
    __device__ __forceinline__ void prof(int phase) {
        if (ev.dd_prof && rank == 0 && tid == 0) { const long long c = clock64(); fx.prof[phase] += c - fx.prof_t; fx.prof_t = c; }
    }
    __device__ void run(int slot, int count, int dual);
};

template <int S>
__device__ void DDC<S>::run(int slot, int count, int dual) {
    k = slot; rk = k >= count ? k - count : k;
    cb = (size_t)k * ev.C; lb = (size_t)k * ev.Lmax; nb = (size_t)k * ev.Wcap;
    ctl = ev.ctl + k;
    comp = ctl->comp_type; W = ctl->width; best_lb = ctl->best_lb;
    tabs[0] = ev.table + (size_t)k * ev.T; tabs[1] = ev.table + ((size_t)ev.K + k) * ev.T;
    const bool relaxed = comp == DDO_RELAXED;
#pragma unroll
    for (int j = 0; j < W32; ++j) hcnt[j] = 0;
    st_cnt = 0; st_minus = 0;
    const int HS = ev.HN / (int)CS;          // vertices whose occurrence counts this CTA owns
    const int u_lo = (int)rank * HS, u_hi = u_lo + HS;

    // ---- root (clean.rs:383-405): node 0 of layer 0 ----------------------------------------------------------------------------
    for (int i = tid; i < 2 * ev.HN; i += DD_NT) D[0][i] = 0;
    for (int i = tid; i < DD_NB; i += DD_NT) gh[i] = 0;
    if (tid == 0) { fx.cnt[0] = 0; fx.cnt[1] = 0; fx.gcnt = 0; }
    unsigned long long best = ~0ull;
    for (int u = u_lo + tid; u < u_hi; u += DD_NT) {
        const unsigned c = (unsigned)((ev.root_state[(size_t)rk * S + (u >> 6)] >> (u & 63)) & 1ull);
        master[u] = c;
        if (c) best = min(best, ((unsigned long long)c << 32) | (unsigned)u);
    }
    if (rank == 0 && tid == 0) {
        uint64_t w[S]; int pc = 0;
#pragma unroll
        for (int j = 0; j < S; ++j) { w[j] = ev.root_state[(size_t)rk * S + j]; pc += __popcll(w[j]); }
        store_row(ev.cur_state[0] + nb * S, w);
        const uint64_t ha = dd_hacc<S>(w);
        s_nm[0][0] = make_uint4((uint32_t)ha, (uint32_t)(ha >> 32), (uint32_t)ctl->root_value, (uint32_t)pc);
        if (L.weighted) s_rub[0][0] = row_rub(w, pc);
        ev.plog[lb * ev.Wcap] = PLOG_CAND_MASK;
        ev.nlog[lb] = 1; ev.rslog[lb * 2] = -1; ev.rslog[lb * 2 + 1] = -1;
    }
    best = block_reduce(best, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, ~0ull, fx.red64);
    best = allreduce1(best, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; });
    int var = best == ~0ull ? -1 : (int)(uint32_t)best;
    int n = 1, t = 0, lel = -1, n_prev = 0;
    bool twin_pushed = false, have_trail = false;
    if (rank == 0 && tid == 0) ev.vlog[lb] = var;
    if (var < 0) {  // the root is a terminal node (clean.rs:350,608-632)
        if (rank == 0 && tid == 0) {
            ctl->has_best = 1; ctl->best_value = ctl->root_value; ctl->best_pos = 0;
            ctl->has_best_exact = 1; ctl->best_exact_value = ctl->root_value; ctl->best_exact_pos = 0;
            ctl->status = ST_DONE; ctl->t_term = 0; ctl->n_cur = 1; ctl->var = -1; ctl->ncand = 1;
        }
    }
    csync();

    if (ev.dd_prof && rank == 0 && tid == 0) { for (int i = 0; i < 16; ++i) fx.prof[i] = 0; fx.prof_t = clock64(); }
    while (var >= 0) {
        const int buf = t & 1, nbuf = buf ^ 1, e = t & 1;
        // D[e ^ 1] was last read by the peers two barriers ago: clear it for the next epoch
        for (int i = tid; i < ev.HN; i += DD_NT) D[e ^ 1][i] = 0;
        if (have_trail) trail(t - 1, n_prev, relaxed);
        prof(0);
        expand(t, n, var);
        prof(1);
        csync();  // ---- S1: every candidate of layer t+1 is inserted ------------------------------------------------------------
        prof(2);

        // ---- next_variable (misp/main.rs:109-143): occurrence counts of the distinct states of layer t+1, argmin, lowest index on ties.
        // Thread (u, r) fetches rank r's delta of vertex u; the CS lanes of a vertex fold them with shuffles.
        best = ~0ull;
        for (int base = 0; base < HS * (int)CS; base += DD_NT) {
            const int idx = base + tid;
            const int u = u_lo + idx / (int)CS; const unsigned r = idx % CS;
            int d = idx < HS * (int)CS ? *cl.map_shared_rank(&D[e][u], r) : 0;
            for (unsigned s = 1; s < CS; s <<= 1) d += __shfl_xor_sync(FULL_MASK, d, s);
            if (r == 0 && idx < HS * (int)CS) {
                const unsigned m = master[u] + (unsigned)d;
                master[u] = m;
                if (m) best = min(best, ((unsigned long long)m << 32) | (unsigned)u);
            }
        }
        // ---- first candidates (rule C1): candidate c represents its state iff it is the smallest candidate that produced it -----------
        const int ncand = 2 * n;
        const int slice_lo = min(node0, n), slice_hi = min(slice_lo + L.slice, n);
        const int nch = (slice_hi - slice_lo + 31) >> 5;
        const int wv_t = ev.weight[var];
        for (int ch = warp; ch < nch; ch += DD_NW) {
            const int i = slice_lo + 32 * ch + lane;
            bool fy = false, fn = false;
            if (i < slice_hi) {
                const unsigned lc = 2u * (unsigned)(i - node0);
                const uint2 rs = *reinterpret_cast<const uint2*>(s_rs + lc);
                uint32_t f0 = NONE32, f1 = NONE32;
                const int pval = (int)s_nm[buf][i - node0].z;
                if (rs.x != NONE32) {
                    const uint32_t rep = rep_of(2u * i, rs.x);
                    f0 = *cptr(s_first, rep); fy = f0 == 2u * i;
                    if (*cptr(s_val32, rep) == vbias(pval + wv_t)) atomicMax(cptr(s_bp, rep), 2u * i);       // a best parent of its state
                }
                if (rs.y != NONE32) {
                    const uint32_t rep = rep_of(2u * i + 1u, rs.y);
                    f1 = *cptr(s_first, rep); fn = f1 == 2u * i + 1u;
                    if (*cptr(s_val32, rep) == vbias(pval)) atomicMax(cptr(s_bp, rep), 2u * i + 1u);
                }
                *reinterpret_cast<uint2*>(s_f + lc) = make_uint2(f0, f1);
                if (ev.dd_dbgbuf) {
                    const int ncl = ((rs.x != NONE32 && (rs.x & RS_CLAIM)) ? 1 : 0) + ((rs.y != NONE32 && (rs.y & RS_CLAIM)) ? 1 : 0);
                    const int ncd = (rs.x != NONE32 ? 1 : 0) + (rs.y != NONE32 ? 1 : 0);
                    if (ncl) atomicAdd(ev.dd_dbgbuf + 4 * t, ncl);
                    if (fy || fn) atomicAdd(ev.dd_dbgbuf + 4 * t + 1, (fy ? 1 : 0) + (fn ? 1 : 0));
                    if (ncd) atomicAdd(ev.dd_dbgbuf + 4 * t + 2, ncd);
                    atomicAdd(ev.dd_dbgbuf + 4 * t + 3, 1);
                }
            }
            const unsigned by = __ballot_sync(FULL_MASK, fy), bn = __ballot_sync(FULL_MASK, fn);
            if (lane == 0) { fb_s[ch] = make_uint2(by, bn); cnt_s[ch] = __popc(by) + __popc(bn); }
        }
        __syncthreads();
        int Ublk;
        {
            const int per = (nch + DD_NT - 1) / DD_NT;  // (1 unless the slice holds more than 16 384 nodes)
            int c = 0;
            for (int q = 0; q < per; ++q) { const int ch = tid * per + q; if (ch < nch) c += cnt_s[ch]; }
            int o = block_excl_scan(c, &Ublk, fx.scan);
            for (int q = 0; q < per; ++q) { const int ch = tid * per + q; if (ch < nch) { off_s[ch] = o; o += cnt_s[ch]; } }
        }
        best = block_reduce(best, [](unsigned long long a, unsigned long long b) { return a < b ? a : b; }, ~0ull, fx.red64);
        // ---- ordered list of the distinct candidates of this CTA and the keys of the width cut (only when a cut is possible: U <= 2 n) ---
        const bool may_cut = comp != DDO_EXACT && ncand > W;
        int mm[4] = {INT32_MIN, INT32_MIN, INT32_MIN, INT32_MIN};  // max value, max -value, max popcount, max -popcount
        if (may_cut) {
            for (int ch = warp; ch < nch; ch += DD_NW) {
                const uint2 fb = fb_s[ch];
                const int i = slice_lo + 32 * ch + lane;
                const unsigned lt = (1u << lane) - 1u;
                const int base = off_s[ch] + __popc(fb.x & lt) + __popc(fb.y & lt);
                const bool fy = (fb.x >> lane) & 1u, fn = (fb.y >> lane) & 1u;
                if (fy | fn) {
                    const int il = i - node0;
                    const unsigned lc = 2u * (unsigned)il;
                    const uint2 rs = *reinterpret_cast<const uint2*>(s_rs + lc);
                    const int ppc = (int)(s_nm[buf][il].w & 0xFFFFu);
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        if (!(d == 0 ? fy : fn)) continue;
                        const uint32_t rsx = d == 0 ? rs.x : rs.y;
                        const int li = base + (d == 1 && fy ? 1 : 0);
                        const uint32_t c = 2u * i + d;
                        const int pc = (rsx & DD_IDENT) ? ppc : (d == 0 ? (int)s_pcy[il] : ppc - 1);
                        const uint32_t vb = *cptr(s_val32, rep_of(c, rsx));
                        const int value = (int)(vb ^ 0x80000000u);
                        ulist[li] = c | (rsx & DD_IDENT); stat[li] = 0;
                        keys[li] = ((unsigned long long)vb << 32) | (unsigned)pc;
                        mm[0] = max(mm[0], value); mm[1] = max(mm[1], -value); mm[2] = max(mm[2], pc); mm[3] = max(mm[3], -pc);
                    }
                }
            }
            block_max4(mm);
        }
        push(6, (unsigned long long)(unsigned)Ublk, best, (unsigned long long)(uint32_t)mm[0], (unsigned long long)(uint32_t)mm[1], (unsigned long long)(uint32_t)mm[2],
             (unsigned long long)(uint32_t)mm[3]);
        prof(3);
        csync();  // ---- S2 ---------------------------------------------------------------------------------------------------------
        prof(4);
        int U = 0, cta_off = 0;
        best = ~0ull;
        for (unsigned r = 0; r < CS; ++r) {
            const int y = (int)got(0, r);
            if (r < rank) cta_off += y;
            U += y; best = min(best, got(1, r));
            mm[0] = max(mm[0], (int)(uint32_t)got(2, r)); mm[1] = max(mm[1], (int)(uint32_t)got(3, r));
            mm[2] = max(mm[2], (int)(uint32_t)got(4, r)); mm[3] = max(mm[3], (int)(uint32_t)got(5, r));
        }
        xnext();
        const int tn = t + 1;  // the layer being decided
        if (U == 0) {  // every node was pruned: empty layer (clean.rs:667-669) -> no best node
            if (rank == 0 && tid == 0) { ctl->status = ST_DONE; ctl->t_term = tn; ctl->has_best = 0; ctl->has_best_exact = 0; ev.nlog[lb + tn] = 0; }
            break;
        }
        const bool terminal = best == ~0ull;  // next_variable == None: layer t+1 is the terminal layer
        const int var_next = terminal ? -1 : (int)(uint32_t)best;
        bool cut = false; int need = 0;
        if (!terminal) {
            if (comp == DDO_RESTRICTED && U > W) { cut = true; need = W; }                   // clean.rs:782-787
            else if (relaxed && U > W && tn >= 2) { cut = true; need = W - 1; }             // clean.rs:788-793 (layers.len() > 1)
        }
        if (!cut && U > ev.Wcap) {
            if (rank == 0 && tid == 0) { ctl->status = ST_DONE; ctl->overflow = 1; ctl->t_term = tn; }
            break;
        }
        if (cut && comp == DDO_RESTRICTED && dual && !twin_pushed) {  // the restricted DD is inexact: its relaxed twin is needed (parallel.rs:425-430)
            twin_pushed = true;
            if (rank == 0 && tid == 0) { const int idx = atomicAdd(ev.dq + 2, 1); ev.dq_jobs[idx] = count + k; __threadfence(); }
        }

        // ---- width cut: the `need` best by (value_top, popcount, BitSet::cmp) -- clean.rs:803-808 / :819-824 ----------------------------
        int nkeep = U, kcta = cta_off;
        if (cut) {
            const int vmin = -mm[1], pcmin = -mm[3];
            const long long VR = (long long)mm[0] - vmin + 1, PR = (long long)mm[2] - pcmin + 1;
            const bool dense = !ev.dd_generic && VR * PR <= DD_NB;
            bool fallback = !dense;
            int first_chunk = 0;
            if (need == 0) {  // relaxed DD of width 1: every candidate is merged away
                for (int li = tid; li < Ublk; li += DD_NT) stat[li] = 2;
                __syncthreads();
                nkeep = 0; kcta = 0; fallback = false;
            } else if (dense) {
                // one pass: histogram of the dense key (value_top - vmin) * PR + (popcount - pcmin), summed into every CTA by remote atomics
                const int nbins = (int)(VR * PR);
                for (int i = tid; i < nbins; i += DD_NT) lh[i] = 0;
                __syncthreads();
                for (int li = tid; li < Ublk; li += DD_NT) {
                    const unsigned long long x = keys[li];
                    atomicAdd(&lh[(key_value(x) - vmin) * (int)PR + ((int)(uint32_t)x - pcmin)], 1u);
                }
                __syncthreads();
                for (int i = tid; i < nbins; i += DD_NT) {
                    const unsigned v = lh[i];
                    if (v) for (unsigned r = 0; r < CS; ++r) atomicAdd(cl.map_shared_rank(&gh[i], r), v);
                }
                csync();  // ---- S3 ---------------------------------------------------------------------------------------------------
                {   // bucket b with  #(bin > b) < need <= #(bin >= b): every thread owns 4 consecutive bins, largest first
                    int c4[4]; int s4 = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const int bi = nbins - 1 - (4 * tid + q); c4[q] = bi >= 0 ? (int)gh[bi] : 0; s4 += c4[q]; }
                    int tot;
                    const int before = block_excl_scan(s4, &tot, fx.scan);
                    if (before < need && need <= before + s4) {
                        int acc = before;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (acc < need && need <= acc + c4[q]) { fx.misc[0] = nbins - 1 - (4 * tid + q); fx.misc[1] = acc; fx.misc[2] = c4[q]; }
                            acc += c4[q];
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < nbins; i += DD_NT) gh[i] = 0;  // (the next remote adds come two barriers later at the earliest)
                }
                const int b = fx.misc[0], above = fx.misc[1], inb = fx.misc[2];
                need -= above;
                const bool all_keep = need == inb;
                if (tid == 0) fx.ucnt = 0;
                __syncthreads();  // (lh is read no more: its memory now holds the undecided list)
                int cnts[4] = {0, 0, 0, 0};  // kept for sure, in the boundary bucket
                for (int li = tid; li < Ublk; li += DD_NT) {
                    const unsigned long long x = keys[li];
                    const int bin = (key_value(x) - vmin) * (int)PR + ((int)(uint32_t)x - pcmin);
                    if (bin > b) { stat[li] = 1; ++cnts[0]; }
                    else if (bin < b) stat[li] = 2;
                    else {
                        ++cnts[1];
                        if (all_keep) stat[li] = 1;
                        else if (inb <= DD_GCAP) und[atomicAdd(&fx.ucnt, 1)] = (uint32_t)li;
                    }
                }
                block_sum4(cnts);
                const int above_l = cnts[0], inb_l = cnts[1];
                if (all_keep) {
                    push(1, (unsigned long long)(unsigned)(above_l + inb_l));
                    csync();  // ---- S4' -------------------------------------------------------------------------------------------
                    nkeep = 0; kcta = 0;
                    for (unsigned r = 0; r < CS; ++r) { const int y = (int)got(0, r); if (r < rank) kcta += y; nkeep += y; }
                    xnext();
                } else if (inb <= DD_GCAP) {
                    // the candidates of the boundary bucket go to rank 0, which orders them by BitSet::cmp alone
                    if (tid == 0) fx.misc[3] = inb_l ? atomicAdd(cl.map_shared_rank(&fx.gcnt, 0), inb_l) : 0;
                    __syncthreads();
                    const int gbase = fx.misc[3];
                    for (int j = tid; j < inb_l; j += DD_NT) {
                        const uint32_t li = und[j];
                        *cl.map_shared_rank(&garr[gbase + j], 0) = (unsigned long long)ulist[li] | ((unsigned long long)li << 32) | ((unsigned long long)rank << 56);
                    }
                    push(1, (unsigned long long)(unsigned)above_l);
                    csync();  // ---- S4 --------------------------------------------------------------------------------------------
                    int above_r[DD_MAXCS];
                    for (unsigned r = 0; r < CS; ++r) above_r[r] = (int)got(0, r);
                    xnext();
                    if (rank == 0) {
                        for (int j = tid; j < inb; j += DD_NT) gstat[j] = 0;
                        if (tid < DD_MAXCS) fx.kbc[tid] = 0;
                        __syncthreads();
                        radix_select<false>(gkeys, gstat, nullptr, garr, inb, need, 1, buf);
                        for (int j = tid; j < inb; j += DD_NT) {
                            const unsigned long long g = garr[j];
                            const unsigned owner = (unsigned)(g >> 56); const uint32_t li = (uint32_t)(g >> 32) & 0xFFFFFFu;
                            const uint8_t s = gstat[j];
                            *cl.map_shared_rank(&stat[li], owner) = s;
                            if (s == 1) atomicAdd(&fx.kbc[owner], 1);
                        }
                        __syncthreads();
                        if (tid < (int)(CS * CS)) *cl.map_shared_rank(&fx.kbx[kphase][tid % CS], tid / CS) = fx.kbc[tid % CS];
                        if (tid == 0) fx.gcnt = 0;
                    }
                    csync();  // ---- S5 --------------------------------------------------------------------------------------------
                    nkeep = 0; kcta = 0;
                    for (unsigned r = 0; r < CS; ++r) { const int y = above_r[r] + fx.kbx[kphase][r]; if (r < rank) kcta += y; nkeep += y; }
                    kphase ^= 1;
                } else {
                    fallback = true; first_chunk = 1;  // a boundary bucket too large for one CTA: the generic select continues from chunk 1
                }
            }
            if (fallback) {
                radix_select<true>(keys, stat, ulist, nullptr, Ublk, need, first_chunk, buf);
                int cnts[4] = {0, 0, 0, 0};
                for (int li = tid; li < Ublk; li += DD_NT) cnts[0] += stat[li] == 1;
                block_sum4(cnts);
                push(1, (unsigned long long)(unsigned)cnts[0]);
                csync();
                nkeep = 0; kcta = 0;
                for (unsigned r = 0; r < CS; ++r) { const int y = (int)got(0, r); if (r < rank) kcta += y; nkeep += y; }
                xnext();
            }
            // keep flags and local offsets of the survivors
            for (int ch = warp; ch < nch; ch += DD_NW) {
                const uint2 fb = fb_s[ch];
                const unsigned lt = (1u << lane) - 1u;
                const int base = off_s[ch] + __popc(fb.x & lt) + __popc(fb.y & lt);
                const bool fy = (fb.x >> lane) & 1u, fn = (fb.y >> lane) & 1u;
                const bool ky = fy && stat[base] == 1, kn = fn && stat[base + (fy ? 1 : 0)] == 1;
                const unsigned by = __ballot_sync(FULL_MASK, ky), bn = __ballot_sync(FULL_MASK, kn);
                if (lane == 0) { kb_s[ch] = make_uint2(by, bn); cnt_s[ch] = __popc(by) + __popc(bn); }
            }
            __syncthreads();
            {
                const int per = (nch + DD_NT - 1) / DD_NT;
                int c = 0, kblk;
                for (int q = 0; q < per; ++q) { const int ch = tid * per + q; if (ch < nch) c += cnt_s[ch]; }
                int o = block_excl_scan(c, &kblk, fx.scan);
                for (int q = 0; q < per; ++q) { const int ch = tid * per + q; if (ch < nch) { koff_s[ch] = o; o += cnt_s[ch]; } }
            }
            __syncthreads();  // koff_s is read by other threads in the commit (a restricted cut has no barrier in between)
        }
        prof(5);
        int n_next = nkeep, sv_pos = -1, r_pos = -1;
        int mpos = -1;          // position of the node that receives the merged-away states (relaxed cut)

        // ---- relaxation: merge the overflow (clean.rs:826-876; misp/main.rs:172-178 union) -------------------------------------------
        if (cut && relaxed) {
            if (tid < 17) fx.merged[tid] = 0;
            __syncthreads();
            uint64_t acc[S];
#pragma unroll
            for (int j = 0; j < S; ++j) acc[j] = 0;
            unsigned long long mkey = 0;
            for (int li = tid; li < Ublk; li += DD_NT) if (stat[li] == 2) {
                const uint32_t cf = ulist[li];
                uint64_t w[S];
                load_row_cg(cand_row(cf, buf), w);
#pragma unroll
                for (int j = 0; j < S; ++j) acc[j] |= w[j];
                const uint32_t c = cf & DD_CMASK;
                const uint32_t rs = s_rs[c - cbase];
                mkey = max(mkey, agg_of(rep_of(c, rs)));
            }
#pragma unroll
            for (int j = 0; j < S; ++j) {
                const uint64_t x = warp_reduce(acc[j], [](uint64_t a, uint64_t b) { return a | b; });
                if (lane == 0 && x) atomicOr(&fx.merged[j], (unsigned long long)x);
            }
            mkey = warp_reduce(mkey, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
            if (lane == 0 && mkey) atomicMax(&fx.merged[16], mkey);
            __syncthreads();
            if (tid < (S + 1) * (int)CS) {  // every CTA's union and best key, pushed to every peer
                const int j = tid / (int)CS; const unsigned r = tid % CS;
                const int src = j < S ? j : 16;
                *cl.map_shared_rank(&fx.mx[rank][src], r) = fx.merged[src];
            }
            csync();  // ---- S6 -----------------------------------------------------------------------------------------------------
            if (tid < 17) {
                unsigned long long m = 0;
                if (tid < S) for (unsigned r = 0; r < CS; ++r) m |= fx.mx[r][tid];
                else if (tid == 16) for (unsigned r = 0; r < CS; ++r) m = max(m, fx.mx[r][16]);
                fx.merged[tid] = m;
            }
            __syncthreads();
            mkey = fx.merged[16];
            // recycled ? (clean.rs:830): a KEPT node whose state equals the merged state.  Every CTA runs the same lookup (cluster-uniform);
            // the table of this parity is only released by the next step's trail.
            if (tid == 0) {
                uint64_t mw[S];
#pragma unroll
                for (int j = 0; j < S; ++j) mw[j] = fx.merged[j];
                const uint64_t h = dd_hfin(dd_hacc<S>(mw));
                const uint32_t tag = (uint32_t)(h >> 32);
                uint32_t sl = (uint32_t)h & (uint32_t)(ev.T - 1);
                const unsigned long long* tab = tabs[t & 1];
                int recycled = -1, rec_rep = -1;
                for (;;) {
                    const unsigned long long en = __ldcg(tab + sl);
                    if (en == EMPTY64) break;
                    if ((uint32_t)(en >> 32) == tag) {
                        const uint4* orow = reinterpret_cast<const uint4*>(cand_row((uint32_t)en, buf));
                        bool eq = true;
#pragma unroll
                        for (int q = 0; q < S / 2; ++q) { const uint4 o4 = ld_cg_u4(orow + q); eq = eq && u4lo(o4) == mw[2 * q] && u4hi(o4) == mw[2 * q + 1]; }
                        if (eq) { rec_rep = (int)((uint32_t)en & DD_CMASK); recycled = (int)*cptr(s_first, (uint32_t)rec_rep); break; }
                    }
                    sl = (sl + 1) & (uint32_t)(ev.T - 1);
                }
                fx.misc[4] = recycled; fx.misc[5] = rec_rep;
            }
            __syncthreads();
            int recycled = fx.misc[4];
            const int rec_rep = fx.misc[5];
            if (recycled >= 0) {
                // (rare) is the state's first candidate kept ?  Its owner CTA knows; everybody learns it through one exchange.
                unsigned long long rinfo = 0;  // (kept ? position + 1 : 0) published by the owner
                const int ri = recycled >> 1;
                if (ri >= slice_lo && ri < slice_hi) {
                    const int ch = (ri - slice_lo) >> 5, ln = (ri - slice_lo) & 31;
                    const uint2 kb = kb_s[ch];
                    const bool kept = (((recycled & 1) ? kb.y : kb.x) >> ln) & 1u;
                    if (kept) {
                        const unsigned lt = (1u << ln) - 1u;
                        const int p = kcta + koff_s[ch] + __popc(kb.x & lt) + __popc(kb.y & lt) + (((recycled & 1) && ((kb.x >> ln) & 1u)) ? 1 : 0);
                        rinfo = (unsigned long long)(p + 1);
                    }
                }
                rinfo = allreduce1(rinfo, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; });
                if (rinfo == 0) recycled = -1;
                else r_pos = (int)rinfo - 1;
            }
            if (recycled >= 0) {
                // the best merged-away node ("saved") stays in the layer, un-deleted, next to the recycled node (clean.rs:868-871)
                uint32_t bestc = NONE32;
                for (int li = tid; li < Ublk; li += DD_NT) if (stat[li] == 2) {
                    const uint32_t cf = ulist[li];
                    if (bestc == NONE32 || better(cf, bestc, buf)) bestc = cf;
                }
                fx.tour[tid] = bestc;
                __syncthreads();
                for (int d = DD_NT / 2; d > 0; d >>= 1) {
                    if (tid < d) {
                        const uint32_t a = fx.tour[tid], b2 = fx.tour[tid + d];
                        if (b2 != NONE32 && (a == NONE32 || better(b2, a, buf))) fx.tour[tid] = b2;
                    }
                    __syncthreads();
                }
                push(1, (unsigned long long)fx.tour[0]); csync();
                uint32_t saved = NONE32;
                for (unsigned r = 0; r < CS; ++r) {
                    const uint32_t c2 = (uint32_t)got(0, r);
                    if (c2 != NONE32 && (saved == NONE32 || better(c2, saved, buf))) saved = c2;
                }
                xnext();
                sv_pos = nkeep; n_next = nkeep + 1; mpos = r_pos;
                csync();  // every CTA has read the recycled node's record through better() before rank 0 rewrites it
                if (rank == 0 && tid == 0) {
                    // the recycled node receives every relaxed edge: RELAXED flag, value_top = max (`>=`: the appended edges win ties)
                    if (key_value(mkey) >= key_value(agg_of((uint32_t)rec_rep))) { *cptr(s_val32, (uint32_t)rec_rep) = (uint32_t)(mkey >> 32); *cptr(s_bp, (uint32_t)rec_rep) = (uint32_t)mkey; }
                    *cptr(s_inex, (uint32_t)rec_rep) |= (uint8_t)(NF_INEXACT | NF_RELAXED);
                }
                csync();  // ... and the commit below reads the rewritten record
                for (int li = tid; li < Ublk; li += DD_NT) if (stat[li] == 2 && ulist[li] == saved) stat[li] = 3;  // kept at s_pos
            } else {
                r_pos = -1;
                mpos = nkeep; n_next = nkeep + 1;
                if (rank == 0 && tid < 32) {  // new merged node (clean.rs:832-849) written straight into the next layer
                    uint64_t mw[S]; int pc = 0;
#pragma unroll
                    for (int j = 0; j < S; ++j) { mw[j] = fx.merged[j]; pc += __popcll(mw[j]); }
                    if (tid == 0) {
                        store_row(ev.cur_state[nbuf] + (nb + mpos) * S, mw);
                        const uint64_t ha = dd_hacc<S>(mw);
                        *nptr(s_nm[nbuf], mpos) = make_uint4((uint32_t)ha, (uint32_t)(ha >> 32), (uint32_t)key_value(mkey), (uint32_t)pc | ((NF_INEXACT | NF_RELAXED) << 16));
                        if (L.weighted) *nptr(s_rub[nbuf], mpos) = row_rub(mw, pc);
                        ev.plog[(lb + tn) * ev.Wcap + mpos] = ((uint32_t)mkey & PLOG_CAND_MASK) | PLOG_INEXACT | PLOG_RELAXED;
                    }
                    // its occurrences enter the histogram of the next epoch (one lane per 32-bit column)
                    for (int j = lane; j < W32; j += 32) {
                        uint32_t x = (uint32_t)(fx.merged[j >> 1] >> ((j & 1) * 32));
                        while (x) { const int b = __ffs((int)x) - 1; atomicAdd(&D[e ^ 1][32 * j + b], 1); x &= x - 1; }
                    }
                }
            }
            __syncthreads();
        }

        prof(6);
        // ---- commit of layer t+1 (_move_to_next_layer, clean.rs:657-687): rows, node records, parent log, positions ----------------------
        const uint64_t hdelta = (uint64_t)(1u << (var & 31)) * (uint64_t)dd_mul32(var >> 5);
        const int wv = ev.weight[var];
        unsigned long long b_all = 0, b_ex = 0;  // terminal layer: (biased value, pos + 1), last maximum (rule C4)
        for (int ch = warp; ch < nch; ch += DD_NW) {
            const uint2 fb = fb_s[ch];
            const int i = slice_lo + 32 * ch + lane;
            const int il = i - node0;
            const unsigned lc = 2u * (unsigned)il;
            const unsigned lt = (1u << lane) - 1u;
            const int fbase = off_s[ch] + __popc(fb.x & lt) + __popc(fb.y & lt);
            const bool fy = (fb.x >> lane) & 1u, fn = (fb.y >> lane) & 1u;
            uint2 kb = fb; int kbase = cta_off + fbase;
            if (cut) { kb = kb_s[ch]; kbase = kcta + koff_s[ch] + __popc(kb.x & lt) + __popc(kb.y & lt); }
            uint64_t w0[S], w1[S];
            bool drop0 = false, drop1 = false;
            uint2 rs = make_uint2(NONE32, NONE32);
            uint4 pm = make_uint4(0, 0, 0, 0);
            if (fy | fn) { rs = *reinterpret_cast<const uint2*>(s_rs + lc); pm = s_nm[buf][il]; }
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const bool isf = d == 0 ? fy : fn;
                uint64_t (&w)[S] = d == 0 ? w0 : w1;
                bool dropped = false;
                if (isf) {
                    const uint32_t c = 2u * i + d;
                    const uint32_t rsx = d == 0 ? rs.x : rs.y;
                    const uint32_t cf = c | (rsx & DD_IDENT);
                    const int li = fbase + (d == 1 && fy ? 1 : 0);
                    bool keep = (((d == 0 ? kb.x : kb.y) >> lane) & 1u) != 0;
                    int pos = kbase + (d == 1 && ((kb.x >> lane) & 1u) ? 1 : 0);
                    if (cut && !keep) {
                        if (stat[li] == 3) { keep = true; pos = sv_pos; }   // the saved node of the recycled corner case
                        else { dropped = true; pos = relaxed ? mpos : -1; }
                    }
                    s_pos[lc + d] = pos < 0 ? NONE32 : (uint32_t)pos;
                    load_row_cg(cand_row(cf, buf), w);  // (a dropped state: its occurrences leave the histogram)
                    if (keep) {
                        unsigned long long key; uint32_t fl;
                        { const uint32_t rep = rep_of(c, rsx); key = agg_of(rep); fl = *cptr(s_inex, rep); }
                        const int value = key_value(key);
                        store_row(ev.cur_state[nbuf] + (nb + pos) * S, w);
                        const uint64_t pha = (uint64_t)pm.x | ((uint64_t)pm.y << 32);
                        const int ppc = (int)(pm.w & 0xFFFFu);
                        uint64_t ha; int pc;
                        if (rsx & DD_IDENT) { ha = pha; pc = ppc; }
                        else if (d == 1) { ha = pha - hdelta; pc = ppc - 1; }
                        else { ha = dd_hacc<S>(w); pc = (int)s_pcy[il]; }
                        *nptr(s_nm[nbuf], pos) = make_uint4((uint32_t)ha, (uint32_t)(ha >> 32), (uint32_t)value, (uint32_t)pc | (fl << 16));
                        if (L.weighted) {
                            const int prub = s_rub[buf][il];
                            *nptr(s_rub[nbuf], pos) = (rsx & DD_IDENT) ? prub : (d == 1 ? prub - wv : row_rub(w, pc));
                        }
                        ev.plog[(lb + tn) * ev.Wcap + pos] = ((uint32_t)key & PLOG_CAND_MASK) | ((fl & NF_INEXACT) ? PLOG_INEXACT : 0u) | ((fl & NF_RELAXED) ? PLOG_RELAXED : 0u);
                        if (terminal) {
                            const unsigned long long kk2 = (key & 0xFFFFFFFF00000000ull) | (unsigned)(pos + 1);
                            b_all = max(b_all, kk2);
                            if (!(fl & (NF_INEXACT | NF_RELAXED))) b_ex = max(b_ex, kk2);
                        }
                    }
                }
                if (d == 0) drop0 = dropped; else drop1 = dropped;
            }
            stage_rows(drop0, w0, true);   // (flushed into the next epoch's counters by the next expansion)
            stage_rows(drop1, w1, true);
        }
        if (terminal) {
            b_all = block_reduce(b_all, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, fx.red64);
            b_ex = block_reduce(b_ex, [](unsigned long long a, unsigned long long b) { return a > b ? a : b; }, 0ull, fx.red64);
            push(2, b_all, b_ex); csync();
            b_all = 0; b_ex = 0;
            for (unsigned r = 0; r < CS; ++r) { b_all = max(b_all, got(0, r)); b_ex = max(b_ex, got(1, r)); }
            xnext();
            if (rank == 0 && tid == 0) {
                ctl->has_best = 1; ctl->best_value = key_value(b_all); ctl->best_pos = (int)(uint32_t)b_all - 1;
                ctl->has_best_exact = b_ex != 0;
                if (b_ex) { ctl->best_exact_value = key_value(b_ex); ctl->best_exact_pos = (int)(uint32_t)b_ex - 1; }
            }
        }
        const bool first_cut = cut && lel < 0;
        if (first_cut) lel = t;  // _maybe_save_lel, clean.rs:796-800: the parent layer of the first squashed layer
        if (rank == 0 && tid == 0) {
            ev.nlog[lb + tn] = n_next; ev.vlog[lb + tn] = var_next;
            ev.rslog[(lb + tn) * 2] = sv_pos; ev.rslog[(lb + tn) * 2 + 1] = r_pos;
            if (first_cut) ctl->lel = t;
        }
        if (first_cut && relaxed) {  // layer t is the last exact layer: keep its nodes for the cutset (clean.rs:566-583)
            for (int i = slice_lo + tid; i < slice_hi; i += DD_NT) {
                uint64_t w[S];
                load_row_cg(ev.cur_state[buf] + (nb + i) * S, w);
                store_row(ev.lel_state + (nb + i) * S, w);
                const uint4 m = s_nm[buf][i - node0];
                ev.lel_val[nb + i] = (int)m.z;
                ev.lel_rub[nb + i] = ev.unit_weights ? (int)(m.w & 0xFFFFu) : s_rub[buf][i - node0];
            }
        }
        prof(7);
        csync();  // ---- end of the layer step: layer t+1 and the positions of its first candidates are published ---------------------------
        prof(8);
        n_prev = n; have_trail = true;
        if (terminal) {
            trail(t, n, relaxed);
            if (rank == 0 && tid == 0) { ctl->status = ST_DONE; ctl->t_term = tn; ctl->n_cur = n_next; ctl->var = -1; ctl->ncand = ncand; }
            have_trail = false;
            break;
        }
        t = tn; n = n_next; var = var_next;
    }
    // ---- the DD is complete -----------------------------------------------------------------------------------------------------------------
    // (a DD that ended on an empty or overflowing layer leaves its last candidates in the tables: the host clears them before the next batch)
    __syncthreads();
    if (tid == 0 && fx.cnt[0]) { atomicAdd(&ctl->expanded, fx.cnt[0]); atomicAdd(&ctl->transitions, fx.cnt[1]); }
    if (ev.dd_prof && rank == 0 && tid == 0) for (int i = 0; i < 16; ++i) atomicAdd((unsigned long long*)ev.dd_prof + i, (unsigned long long)fx.prof[i]);
    __threadfence();
    csync();  // (nobody starts the next DD while a peer may still read this one's records)
    if (rank == 0 && tid == 0) {
        // a restricted DD that never needed a cut is exact: its relaxed twin will never be compiled (parallel.rs:421-423)
        const int fin = (dual && k < count && !twin_pushed) ? 2 : 1;
        __threadfence();
        atomicAdd(ev.dq + 3, fin);
    }
}

template <int S>
__global__ void __launch_bounds__(DD_NT, 1) k_dd(const __grid_constant__ EV ev, const __grid_constant__ DDLayout L, int count, int dual) {
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ DDFixed fx;
    DDC<S> c(ev, L, fx, dsm);
    for (;;) {
        if (c.rank == 0 && c.tid == 0) fx.job = dd_fetch_job(ev, count);
        c.cl.sync();
        const int slot = *c.cl.map_shared_rank(&fx.job, 0);
        c.cl.sync();  // everybody has read the job before rank 0 fetches the next one (or leaves)
        if (slot < 0) break;
        c.run(slot, count, dual);
    }
}

}  // namespace ddo

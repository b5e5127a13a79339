// m2s_engine.cu -- host side of the MAX2SAT device model: instance tables in HBM, arenas, the per-layer launch loop.
#include "m2s_kernels.cuh"
#include "m2s_frontier.cuh"
#include "m2s_engine.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace ddo {

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                               \
            return DDO_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static uint64_t host_mix64(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// ---------------------------------------------------------------------------------------------------------------
// model: ddo/examples/max2sat/model.rs:98-249 as device tables
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
static cudaError_t upload(T** dst, const std::vector<T>& src) {
    cudaError_t e = cudaMalloc((void**)dst, std::max<size_t>(src.size() * sizeof(T), 16));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

int model_create_max2sat(int32_t n, int64_t m, const int64_t* clauses, int device, M2Model** out) {
    if (n <= 0 || m < 0 || (m > 0 && !clauses) || !out) { set_error("ddo_model_create_max2sat: invalid argument"); return DDO_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (there is no CPU fallback)"); return DDO_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_error("invalid device ordinal"); return DDO_ERR_INVALID; }
    if (n > 1024) { set_error("MAX2SAT device model supports n <= 1024 variables"); return DDO_ERR_UNSUPPORTED; }
    auto* M = new M2Model();
    M->n = n; M->words = (n + 1) / 2; M->NW = (n + 3) & ~3; M->device = device;
    const size_t L = 2 * (size_t)n;
    auto mk_lit = [](int64_t x) -> size_t { const size_t a = (size_t)((x < 0 ? -x : x) - 1); return a + a + (x > 0 ? 1 : 0); };  // model.rs:116-121
    // data.rs:99,106: a repeated clause keeps the LAST weight
    std::vector<int64_t> w(L * L, 0);
    std::vector<char> present(L * L, 0);
    std::vector<size_t> keys;
    for (int64_t i = 0; i < m; ++i) {
        const int64_t wt = clauses[3 * i], x = clauses[3 * i + 1], y = clauses[3 * i + 2];
        if (x == 0 || y == 0 || (x < 0 ? -x : x) > n || (y < 0 ? -y : y) > n) { delete M; set_error("clause literal out of range"); return DDO_ERR_INVALID; }
        const int64_t a = std::min(x, y), b = std::max(x, y);
        const size_t key = mk_lit(a) * L + mk_lit(b);
        if (!present[key]) { present[key] = 1; keys.push_back(key); }
        w[key] = wt;
    }
    int64_t abs_sum = 0;
    std::vector<int64_t> socw(n, 0);
    M->initial = 0;
    for (size_t key : keys) {  // model.rs:136-147
        const size_t la = key / L, lb = key % L;
        const int va = (int)(la / 2), vb = (int)(lb / 2);
        const int64_t wt = w[key];
        abs_sum += wt < 0 ? -wt : wt;
        socw[va] += wt;
        if (la != lb) socw[vb] += wt;
        if (va == vb && la != lb) M->initial += wt;  // tautology: x == -y
    }
    if (abs_sum >= (1ll << 28)) { delete M; set_error("sum of |clause weights| must be < 2^28 (values are 32-bit on the device)"); return DDO_ERR_UNSUPPORTED; }
    auto weight = [&](int64_t x, int64_t y) -> int64_t { const int64_t a = std::min(x, y), b = std::max(x, y); return w[mk_lit(a) * L + mk_lit(b)]; };
    auto tl = [](int v) -> int64_t { return (int64_t)v + 1; };
    auto fl = [](int v) -> int64_t { return -((int64_t)v + 1); };
    // variable order (model.rs:149-151): ascending sum of clause weights; canonical: stable (ties by variable id)
    M->h_ord.resize(n);
    for (int i = 0; i < n; ++i) M->h_ord[i] = i;
    std::stable_sort(M->h_ord.begin(), M->h_ord.end(), [&](int a, int b) { return socw[a] < socw[b]; });
    std::vector<int> posn(n);
    for (int i = 0; i < n; ++i) posn[M->h_ord[i]] = i;
    // fast_upper_bound tables (model.rs:183-238)
    std::vector<long long> est(n, 0), nk(n, 0);
    {
        long long acc = 0;
        for (int i = n; i-- > 0;) {
            const int vi = M->h_ord[i];
            for (int j = i + 1; j < n; ++j) {
                const int vj = M->h_ord[j];
                const int64_t tt = weight(tl(vi), tl(vj)), tf = weight(tl(vi), fl(vj)), ft = weight(fl(vi), tl(vj)), ff = weight(fl(vi), fl(vj));
                acc += std::max(std::max(tt + tf + ft, tt + tf + ff), std::max(tt + ft + ff, tf + ft + ff));
            }
            acc += weight(tl(vi), fl(vi)) + std::max(weight(tl(vi), tl(vi)), weight(fl(vi), fl(vi)));
            est[i] = acc;
        }
        long long sum = 0;
        for (int k = 0; k < n; ++k) { nk[k] = sum; sum += weight(tl(M->h_ord[k]), fl(M->h_ord[k])); }
    }
    // Rows of the branching variable k (the "remaining" variables l are those before k in the order, model.rs:173-181):
    //   decision T: child = s + PT[k] - QT[k],  cost = pos(s[k])  + AT[k] + sum_l min(pos(s[l]) + PT[k][l], pos(-s[l]) + QT[k][l])
    //   decision F: child = s + PF[k] - QF[k],  cost = pos(-s[k]) + AF[k] + sum_l min(pos(s[l]) + PF[k][l], pos(-s[l]) + QF[k][l])
    //   PT = w(fk,tl), QT = w(fk,fl), AT = w(tk,tk) + sum_l w(tk,fl) + w(tk,tl);  PF = w(tk,tl), QF = w(tk,fl), AF = w(fk,fk) + sum_l w(fk,fl) + w(fk,tl)
    // (model.rs:275-328); rows are zero outside the remaining variables, where min(pos(s), pos(-s)) = 0 contributes nothing.
    const size_t NW = M->NW;
    std::vector<int32_t> PT((size_t)n * NW, 0), QT((size_t)n * NW, 0), PF((size_t)n * NW, 0), QF((size_t)n * NW, 0), AT(n, 0), AF(n, 0);
    for (int k = 0; k < n; ++k) {
        int64_t at = weight(tl(k), tl(k)), af = weight(fl(k), fl(k));
        for (int l = 0; l < n; ++l) {
            if (posn[l] >= posn[k]) continue;
            PT[(size_t)k * NW + l] = (int32_t)weight(fl(k), tl(l)); QT[(size_t)k * NW + l] = (int32_t)weight(fl(k), fl(l));
            PF[(size_t)k * NW + l] = (int32_t)weight(tl(k), tl(l)); QF[(size_t)k * NW + l] = (int32_t)weight(tl(k), fl(l));
            at += weight(tl(k), fl(l)) + weight(tl(k), tl(l));
            af += weight(fl(k), fl(l)) + weight(fl(k), tl(l));
        }
        AT[k] = (int32_t)at; AF[k] = (int32_t)af;
    }
    std::vector<uint32_t> hmul(NW);
    for (size_t i = 0; i < NW; ++i) hmul[i] = (uint32_t)(host_mix64(0x9E3779B97F4A7C15ULL * (i + 1)) >> 32) | 1u;
    bool ok = cudaSetDevice(device) == cudaSuccess && upload(&M->d_ord, M->h_ord) == cudaSuccess && upload(&M->d_PT, PT) == cudaSuccess &&
              upload(&M->d_QT, QT) == cudaSuccess && upload(&M->d_PF, PF) == cudaSuccess && upload(&M->d_QF, QF) == cudaSuccess &&
              upload(&M->d_AT, AT) == cudaSuccess && upload(&M->d_AF, AF) == cudaSuccess && upload(&M->d_est, est) == cudaSuccess &&
              upload(&M->d_nk, nk) == cudaSuccess && upload(&M->d_hmul, hmul) == cudaSuccess;
    if (!ok) { set_error(std::string("model upload: ") + cudaGetErrorString(cudaGetLastError())); model_destroy(M); return DDO_ERR_CUDA; }
    *out = M;
    return DDO_OK;
}
void model_destroy(M2Model* M) {
    if (!M) return;
    for (void* p : {(void*)M->d_ord, (void*)M->d_PT, (void*)M->d_QT, (void*)M->d_PF, (void*)M->d_QF, (void*)M->d_AT, (void*)M->d_AF, (void*)M->d_est, (void*)M->d_nk, (void*)M->d_hmul})
        if (p) cudaFree(p);
    delete M;
}

// ---------------------------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
static int dev_alloc(Engine* E, T** p, size_t count) {
    void* q = nullptr;
    size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) { set_error(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e)); return DDO_ERR_CUDA; }
    E->allocations.push_back(q);
    E->bytes_allocated += bytes;
    *p = (T*)q;
    return DDO_OK;
}
#define ALLOC(ptr, count)                                   \
    do { int _r = dev_alloc(this, &(ptr), (size_t)(count)); if (_r != DDO_OK) return _r; } while (0)

int M2Engine::create_m2s(const M2Model* m, int dev, uint64_t max_width_cap, int batch_cap, int cutset) {
    if (!m || batch_cap < 1 || max_width_cap < 1) { set_error("ddo_mdd_create: invalid argument"); return DDO_ERR_INVALID; }
    if (cutset != DDO_LAST_EXACT_LAYER && cutset != DDO_FRONTIER) { set_error("cutset_type must be DDO_LAST_EXACT_LAYER or DDO_FRONTIER (mdd.rs:24-28)"); return DDO_ERR_INVALID; }
    if (max_width_cap > (1u << 24)) { set_error("max_width_cap too large"); return DDO_ERR_INVALID; }
    if (cutset == DDO_FRONTIER && (max_width_cap + 2 >= (1u << FC_POS_BITS) || (uint64_t)(m->n + 1) * (max_width_cap + 2) >= (1ull << 30))) {
        set_error("FRONTIER cutset: max_width_cap * (n + 1) must stay below 2^30 frontier records per DD"); return DDO_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (there is no CPU fallback)"); return DDO_ERR_NO_DEVICE; }
    if (dev != m->device) { set_error("model and mdd must live on the same device"); return DDO_ERR_INVALID; }
    m2 = m; model = nullptr; device = dev; cutset_type = cutset; n_vars = m->n; abi_words = m->words;
    bit_value[0] = -1; bit_value[1] = 1;  // model.rs:30-32: F = -1, T = 1; the first decision of the domain (T) is the even candidate
    small_ws = 0; dual_enabled = false;
    K = batch_cap; Wcap = (int)((std::max<uint64_t>(max_width_cap, 2) + 1) & ~1ull); C = 2 * Wcap; T = next_pow2(std::max(3 * Wcap, 64));
    const int NW = m->NW;
    S = NW / 2;  // uint64 words of a device state row (the drain buffers and root staging of the base class are sized with it)
    Lmax = m->n + 1; PW = (Lmax + 63) / 64;
    Klog = K; pool_layers = (size_t)K * Lmax; Lcur = Lmax;  // (this engine keeps one full-depth log per slot)
    CUDA_TRY(cudaSetDevice(dev));
    CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    const size_t KW = (size_t)K * Wcap, KC = (size_t)K * C, KL = (size_t)K * Lmax;
    M2EV& v = mv;
    v.K = K; v.Wcap = Wcap; v.C = C; v.T = T; v.Lmax = Lmax; v.n = m->n; v.NW = NW; v.NW4 = NW / 4; v.PW = PW;
    v.ord = m->d_ord; v.PT = m->d_PT; v.QT = m->d_QT; v.PF = m->d_PF; v.QF = m->d_QF; v.AT = m->d_AT; v.AF = m->d_AF;
    v.est = m->d_est; v.nk = m->d_nk; v.initial = m->initial; v.hmul = m->d_hmul;
    ALLOC(v.ctl, K); ALLOC(v.aux, K); ALLOC(v.active, 4); ALLOC(v.lel_any, 4); ALLOC(v.tile_off_e, K + 1); ALLOC(v.tile_off_c, K + 1); ALLOC(v.finish_counter, 4);
    for (int b = 0; b < 2; ++b) { ALLOC(v.cur_src[b], KW); ALLOC(v.cand_state[b], (size_t)K * (C + 1) * NW); ALLOC(v.cur_val[b], KW); ALLOC(v.cur_flag[b], KW); ALLOC(v.cur_rank[b], KW); ALLOC(v.vb[b], KW); }
    ALLOC(v.cur_rub, KW);
    ALLOC(v.cand_rep, KC); ALLOC(v.cand_first, KC); ALLOC(v.cand_agg, KC); ALLOC(v.cand_inex, KC);
    ALLOC(v.cand_rank, KC); ALLOC(v.cand_slot, KC); ALLOC(v.cand_cost, KC);
    ALLOC(v.uflag, KC); ALLOC(v.ulist, KC); ALLOC(v.ustat, KC); ALLOC(v.pos_of, KC); ALLOC(v.gkeys, KC);
    finish_smem = (size_t)C * 9 + 16;  // keys (8 B) + status (1 B) of up to C distinct candidates next to ~9 KB of static shared memory
    v.smem_keys = finish_smem <= 200 * 1024;
    if (!v.smem_keys) finish_smem = 0;
    ALLOC(v.table, (size_t)K * T);
    ALLOC(v.mrg_min, (size_t)K * NW); ALLOC(v.mrg_max, (size_t)K * NW);
    ALLOC(v.plog, KL * Wcap); ALLOC(v.clog, KL * C); ALLOC(v.colog, KL * C); ALLOC(v.nlog, KL); ALLOC(v.vlog, KL); ALLOC(v.rslog, KL * 3);
    ALLOC(v.lel_state, KW * NW); ALLOC(v.lel_val, KW); ALLOC(v.lel_rub, KW);
    ALLOC(v.cs_ub, KW); ALLOC(v.cs_marked, KW);
    ALLOC(v.best_path, (size_t)K * PW); ALLOC(v.best_exact_path, (size_t)K * PW);
    out_cap = KW;
    v.fc_node = nullptr; v.fc_ub = nullptr; v.fc_aux = nullptr; v.fc_cap = 0; d_out.tt = nullptr;
    if (cutset == DDO_FRONTIER) { int fr = alloc_frontier(&v.fc_node, &v.fc_ub, &v.fc_aux, &v.fc_cap); if (fr != DDO_OK) return fr; }
    ALLOC(d_out.state, out_cap * S); ALLOC(d_out.val, out_cap); ALLOC(d_out.ub, out_cap); ALLOC(d_out.dd, out_cap); ALLOC(d_out.path, out_cap * PW);
    ALLOC(d_out.count, K + 1); ALLOC(d_out.offset, K + 1); ALLOC(d_out.loc, KW);
    ALLOC(d_ub_cap, K); ALLOC(d_lb_filter, K);
    // what the base class reads (fetch_ctl, best_solution, layer_trace, fetch_vars)
    ev.K = K; ev.Wcap = Wcap; ev.C = C; ev.T = T; ev.Lmax = Lmax; ev.n = m->n; ev.S = S; ev.PW = PW;
    ev.ctl = v.ctl; ev.active = v.active; ev.nlog = v.nlog; ev.vlog = v.vlog; ev.best_path = v.best_path; ev.best_exact_path = v.best_exact_path;
    CUDA_TRY(cudaMemsetAsync(v.table, 0xFF, (size_t)K * T * 8, stream));
    CUDA_TRY(cudaMemsetAsync(v.finish_counter, 0, 16, stream));
    { int rr = reserve_roots(K); if (rr != DDO_OK) return rr; }
    CUDA_TRY(cudaMallocHost(&h_ctl, (size_t)K * sizeof(DDCtl)));
    CUDA_TRY(cudaMallocHost(&h_active, 16));
    CUDA_TRY(cudaMallocHost(&h_caps, (size_t)K * 16));
    CUDA_TRY(cudaMallocHost(&h_counts, (size_t)(K + 1) * 8));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return DDO_OK;
}

int M2Engine::reserve_roots(int count) {
    const int rc = Engine::reserve_roots(count);
    mv.root_state = reinterpret_cast<int32_t*>(ev.root_state); mv.root_val = ev.root_val; mv.root_depth = ev.root_depth; mv.root_width = ev.root_width;
    return rc;
}

// dynamic shared memory of m2_expand: the four clause-weight rows of the branching variable + the hash multipliers (TMA staging buffers)
template <int CH>
static cudaError_t launch_expand(M2Engine* E, int grid, int t, int count) {
    const size_t smem = (size_t)5 * E->mv.NW * 4;
    if (!E->expand_attr_set) {
        cudaError_t e = cudaFuncSetAttribute(m2_expand<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        E->expand_attr_set = true;
    }
    m2_expand<CH><<<grid, 256, smem, E->stream>>>(E->mv, t, count);
    return cudaSuccess;
}

int M2Engine::compile_staged(int count, int comp_type, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms) {
    if (count < 1 || count > K || count > staged) { set_error("compile: batch not staged"); return DDO_ERR_INVALID; }
    if (comp_type != DDO_EXACT && comp_type != DDO_RELAXED && comp_type != DDO_RESTRICTED) { set_error("bad compilation type"); return DDO_ERR_INVALID; }
    for (int i = 0; i < count; ++i)
        if (comp_type == DDO_RELAXED && h_root_width[i] < 1) { set_error("max_width must be >= 1 for a relaxed DD (the reference panics at clean.rs:827)"); return DDO_ERR_INVALID; }
    if (cutoff_flag && *cutoff_flag) return DDO_CUTOFF;
    CUDA_TRY(cudaSetDevice(device));
    cudaStream_t st = stream;
    const M2EV& v = mv;
    CUDA_TRY(cudaMemsetAsync(v.table, 0xFF, (size_t)count * T * 8, st));
    CUDA_TRY(cudaEventRecord(ev0, st));
    if (finish_smem && !finish_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(m2_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)finish_smem));
        finish_attr_set = true;
    }
    m2_init<<<count, 256, 0, st>>>(v, count, comp_type, (long long)best_lb);
    ++g_kernel_launches;
    prof_mark(-1);
    const long long max_tiles = (long long)count * ((C + 7) / 8);
    const int flat_grid = (int)std::min<long long>(max_tiles, (long long)num_sms * 8);
    const int compact_grid = (int)std::min<long long>((max_tiles + 31) / 32, (long long)num_sms * 8);
    const int ch = (v.NW4 + 31) / 32;
    const bool relaxed = comp_type == DDO_RELAXED;
    const dim3 merge_grid((C + 63) / 64, count);
    const int CHUNK = 16;
    int rc = DDO_OK;
    for (int t = 0; t < Lmax; ++t) {
        m2_finish<<<count, 1024, finish_smem, st>>>(v, t);
        ++g_kernel_launches;
        prof_mark(1);
        if (relaxed && t >= 2) {
            m2_merge<<<merge_grid, 256, 0, st>>>(v, t);
            m2_merge_fin<<<count, 256, 0, st>>>(v, t);
            g_kernel_launches += 2;
            prof_mark(5);
        }
        m2_compact<<<compact_grid, 256, 0, st>>>(v, t, count);
        prof_mark(2);
        switch (ch) {
            case 1: CUDA_TRY(launch_expand<1>(this, flat_grid, t, count)); break;
            case 2: CUDA_TRY(launch_expand<2>(this, flat_grid, t, count)); break;
            case 3: case 4: CUDA_TRY(launch_expand<4>(this, flat_grid, t, count)); break;
            default: CUDA_TRY(launch_expand<8>(this, flat_grid, t, count)); break;
        }
        prof_mark(0);
        g_kernel_launches += 2; ++layer_steps;
        if ((t % CHUNK) == CHUNK - 1 || t == Lmax - 1) {
            bytes_d2h += sizeof(int); CUDA_TRY(cudaMemcpyAsync(h_active, v.active, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (*h_active <= 0) break;
            if (cutoff_flag && *cutoff_flag) { rc = DDO_CUTOFF; break; }  // Cutoff::must_stop polled between layers (clean.rs:352)
        }
    }
    if (rc == DDO_OK) {
        m2_finalize<<<(count + 63) / 64, 64, 0, st>>>(v, count);
        ++g_kernel_launches;
        if (relaxed && cutset_type == DDO_FRONTIER) {  // clean.rs:586-606 + 448-475 in one sweep, then the upper bound of every member
            m2_fc_sweep<<<count, 1024, 0, st>>>(v);
            const dim3 eg(64, count);
            switch (ch) {
                case 1: m2_fc_eval<1><<<eg, 256, 0, st>>>(v); break;
                case 2: m2_fc_eval<2><<<eg, 256, 0, st>>>(v); break;
                case 3: case 4: m2_fc_eval<4><<<eg, 256, 0, st>>>(v); break;
                default: m2_fc_eval<8><<<eg, 256, 0, st>>>(v); break;
            }
            g_kernel_launches += 2;
        } else if (relaxed) { m2_bottomup<<<count, 1024, 0, st>>>(v); ++g_kernel_launches; }
        prof_mark(3);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (device_ms) CUDA_TRY(cudaEventElapsedTime(device_ms, ev0, ev1));
    { int prc = prof_collect(); if (prc != DDO_OK) return prc; }
    last_count = count; last_comp_type = comp_type; ctl_fetched = false;
    return rc;
}

void M2Engine::fc_launch_count(int count) { k_fc_count<M2EV><<<last_count, 1024, 0, stream>>>(mv, d_out, d_ub_cap, d_lb_filter, count); }
void M2Engine::fc_launch_write(int pw) {
    const dim3 grid(64, last_count);
    switch ((mv.NW4 + 31) / 32) {
        case 1: m2_fc_write<1><<<grid, 256, 0, stream>>>(mv, d_out, d_ub_cap, pw); break;
        case 2: m2_fc_write<2><<<grid, 256, 0, stream>>>(mv, d_out, d_ub_cap, pw); break;
        case 3: case 4: m2_fc_write<4><<<grid, 256, 0, stream>>>(mv, d_out, d_ub_cap, pw); break;
        default: m2_fc_write<8><<<grid, 256, 0, stream>>>(mv, d_out, d_ub_cap, pw); break;
    }
}

int M2Engine::drain_all(int count, const int64_t* ub_cap, const int64_t* lb_filter, int* pw_out) {
    if (last_comp_type != DDO_RELAXED) { set_error("drain_cutset: the last batch was not a relaxed compilation (mdd.rs:103-110)"); return DDO_ERR_INVALID; }
    if (count > last_count) { set_error("drain_cutset: bad count"); return DDO_ERR_INVALID; }
    if (cutset_type == DDO_FRONTIER) return drain_all_frontier(count, ub_cap, lb_filter, pw_out);
    int rc = fetch_ctl(last_count);
    if (rc != DDO_OK) return rc;
    int max_lel = 0;
    for (int i = 0; i < count; ++i) max_lel = std::max(max_lel, h_ctl[i].lel);
    const int pw = std::max(1, (max_lel + 63) / 64);
    *pw_out = pw;
    long long* caps = (long long*)h_caps;
    for (int i = 0; i < K; ++i) { caps[i] = i < count ? ub_cap[i] : 0; caps[K + i] = i < count ? lb_filter[i] : INT64_MAX; }
    CUDA_TRY(cudaSetDevice(device));
    bytes_h2d += (unsigned long long)((size_t)K * 16);
    CUDA_TRY(cudaMemcpyAsync(d_ub_cap, caps, (size_t)K * 8, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_lb_filter, caps + K, (size_t)K * 8, cudaMemcpyHostToDevice, stream));
    prof_mark(-1);
    m2_cutset_count<<<last_count, 1024, 0, stream>>>(mv, d_out, d_ub_cap, d_lb_filter, count);
    k_cutset_offsets<<<1, 32, 0, stream>>>(d_out, last_count);
    g_kernel_launches += 2;
    bytes_d2h += (unsigned long long)((size_t)(last_count + 1) * 4); CUDA_TRY(cudaMemcpyAsync(h_counts, d_out.offset, (size_t)(last_count + 1) * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const int total = ((int32_t*)h_counts)[last_count];
    if (total == 0) { prof_used = 0; return 0; }
    const dim3 grid((Wcap + 7) / 8, last_count);
    m2_cutset_write<<<grid, 256, 0, stream>>>(mv, d_out, d_ub_cap, pw);
    ++g_kernel_launches;
    prof_mark(4);
    if (!h_out_state) {
        const size_t KW = (size_t)K * Wcap;
        CUDA_TRY(cudaMallocHost(&h_out_state, KW * S * 8));
        CUDA_TRY(cudaMallocHost(&h_out_val, KW * 4));
        CUDA_TRY(cudaMallocHost(&h_out_ub, KW * 4));
        CUDA_TRY(cudaMallocHost(&h_out_dd, KW * 4));
        CUDA_TRY(cudaMallocHost(&h_out_path, KW * PW * 8));
    }
    bytes_d2h += (unsigned long long)((size_t)total * (S * 8 + 12 + pw * 8));
    CUDA_TRY(cudaMemcpyAsync(h_out_state, d_out.state, (size_t)total * S * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_val, d_out.val, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_ub, d_out.ub, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_dd, d_out.dd, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_path, d_out.path, (size_t)total * pw * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    { int prc = prof_collect(); if (prc != DDO_OK) return prc; }
    return total;
}

}  // namespace ddo

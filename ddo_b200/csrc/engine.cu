// engine.cu -- host side of the batch DD-compilation engine: arenas in HBM, the per-layer launch loop, result fetch.
#include "kernels.cuh"
#include "frontier.cuh"
#include "dd_kernel.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cstdio>

namespace ddo {

thread_local std::string g_last_error;
unsigned long long g_kernel_launches = 0;
void set_error(const std::string& msg) { g_last_error = msg; }

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                               \
            return DDO_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// ---------------------------------------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------------------------------------
int model_create_misp(int32_t n, const int64_t* weights, int64_t m, const int32_t* src, const int32_t* dst, int device, MispModel** out) {
    if (n <= 0 || m < 0 || (m > 0 && (!src || !dst)) || !out) { set_error("ddo_model_create_misp: invalid argument"); return DDO_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (there is no CPU fallback)"); return DDO_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { set_error("invalid device ordinal"); return DDO_ERR_INVALID; }
    if (n > 1024) { set_error("MISP device model supports n <= 1024 vertices"); return DDO_ERR_UNSUPPORTED; }
    auto* M = new MispModel();
    M->n = n; M->words = (n + 63) / 64; M->S = std::max(2, next_pow2(M->words)); M->device = device;
    M->h_weight.assign(n, 1);
    if (weights) for (int i = 0; i < n; ++i) M->h_weight[i] = weights[i];
    M->unit_weights = true; M->weight_abs_sum = 0;
    for (int i = 0; i < n; ++i) { if (M->h_weight[i] != 1) M->unit_weights = false; M->weight_abs_sum += M->h_weight[i] < 0 ? -M->h_weight[i] : M->h_weight[i]; }
    if (M->weight_abs_sum >= (1ll << 30)) { delete M; set_error("sum of |weights| must be < 2^30 (values are 32-bit on the device)"); return DDO_ERR_UNSUPPORTED; }
    // complement adjacency: full rows, then every edge removed in both directions (misp/main.rs:280-307)
    M->h_nc.assign((size_t)n * M->S, 0);
    for (int v = 0; v < n; ++v)
        for (int u = 0; u < n; ++u) M->h_nc[(size_t)v * M->S + (u >> 6)] |= 1ull << (u & 63);
    for (int64_t e = 0; e < m; ++e) {
        int a = src[e], b = dst[e];
        if (a < 0 || b < 0 || a >= n || b >= n) { delete M; set_error("edge endpoint out of range"); return DDO_ERR_INVALID; }
        M->h_nc[(size_t)a * M->S + (b >> 6)] &= ~(1ull << (b & 63));
        M->h_nc[(size_t)b * M->S + (a >> 6)] &= ~(1ull << (a & 63));
    }
    std::vector<int32_t> w32((size_t)M->S * 64, 0);
    for (int i = 0; i < n; ++i) w32[i] = (int32_t)M->h_weight[i];
    if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&M->d_weight, w32.size() * 4) != cudaSuccess ||
        cudaMalloc(&M->d_nc, M->h_nc.size() * 8) != cudaSuccess ||
        cudaMemcpy(M->d_weight, w32.data(), w32.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(M->d_nc, M->h_nc.data(), M->h_nc.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error(std::string("model upload: ") + cudaGetErrorString(cudaGetLastError()));
        model_destroy(M);
        return DDO_ERR_CUDA;
    }
    *out = M;
    return DDO_OK;
}
void model_destroy(MispModel* M) {
    if (!M) return;
    if (M->d_weight) cudaFree(M->d_weight);
    if (M->d_nc) cudaFree(M->d_nc);
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

// every node above the terminal layer may be a frontier node; the drain buffers hold DDO_FC_OUT_FACTOR (default 8) layers' worth
int Engine::alloc_frontier(uint32_t** node, int32_t** ub, int32_t** aux, unsigned long long* cap) {
    *cap = (unsigned long long)(Lmax - 1) * Wcap;
    const size_t KF = (size_t)K * *cap, KW = (size_t)K * Wcap;
    ALLOC(*node, KF); ALLOC(*ub, KF); ALLOC(*aux, KF);
    int factor = 8;
    if (const char* e = getenv("DDO_FC_OUT_FACTOR")) factor = std::max(1, atoi(e));
    out_cap = std::min(KF, std::max<size_t>(KW * (size_t)factor, 1u << 16));
    ALLOC(d_out.tt, out_cap);
    return DDO_OK;
}

int Engine::create(const MispModel* m, int dev, uint64_t max_width_cap, int batch_cap, int cutset) {
    if (!m || batch_cap < 1 || max_width_cap < 1) { set_error("ddo_mdd_create: invalid argument"); return DDO_ERR_INVALID; }
    if (cutset != DDO_LAST_EXACT_LAYER && cutset != DDO_FRONTIER) { set_error("cutset_type must be DDO_LAST_EXACT_LAYER or DDO_FRONTIER (mdd.rs:24-28)"); return DDO_ERR_INVALID; }
    if (max_width_cap > (1u << 27)) { set_error("max_width_cap too large"); return DDO_ERR_INVALID; }
    if (cutset == DDO_FRONTIER && (max_width_cap + 2 >= (1u << FC_POS_BITS) || (uint64_t)(m->n + 1) * (max_width_cap + 2) >= (1ull << 30))) {
        set_error("FRONTIER cutset: max_width_cap * (n + 1) must stay below 2^30 frontier records per DD"); return DDO_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device (there is no CPU fallback)"); return DDO_ERR_NO_DEVICE; }
    if (dev != m->device) { set_error("model and mdd must live on the same device"); return DDO_ERR_INVALID; }
    model = m; device = dev; cutset_type = cutset; n_vars = m->n; abi_words = m->words;
    K = batch_cap; Wcap = (int)((std::max<uint64_t>(max_width_cap, 2) + 1) & ~1ull); C = 2 * Wcap; T = next_pow2(std::max(3 * Wcap, 64)); S = m->S;
    Lmax = m->n + 1; PW = (Lmax + 63) / 64;
    Klog = std::min(K, 512);
    if (const char* e = getenv("DDO_LOG_SLOTS")) Klog = std::max(1, std::min(K, atoi(e)));
    if (cutset == DDO_FRONTIER) Klog = K;  // the frontier records are per slot and per layer: no pooling
    pool_layers = (size_t)Klog * Lmax; Lcur = Lmax; staged_layers = Lmax;
    CUDA_TRY(cudaSetDevice(dev));
    CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    const size_t KW = (size_t)K * Wcap, KC = (size_t)K * C, KL = (size_t)K * Lmax;
    ev.K = K; ev.Wcap = Wcap; ev.C = C; ev.T = T; ev.Lmax = Lmax; ev.n = m->n; ev.S = S; ev.PW = PW;
    ev.HN = 64 * S; ev.unit_weights = m->unit_weights; ev.weight = m->d_weight; ev.nc = m->d_nc;
    ALLOC(ev.ctl, K); ALLOC(ev.active, 4); ALLOC(ev.tile_off_e, K + 1); ALLOC(ev.tile_off_c, K + 1); ALLOC(ev.finish_counter, 4); ALLOC(ev.fin_list, 2 * (size_t)K); ALLOC(ev.fin_cnt, 4);
    for (int b = 0; b < 2; ++b) { ALLOC(ev.cur_state[b], KW * S); ALLOC(ev.cur_val[b], KW); ALLOC(ev.cur_flag[b], KW); ALLOC(ev.vb[b], KW); }
    ALLOC(ev.cur_rub, KW);
    ALLOC(ev.cand_state, KC * S); ALLOC(ev.cand_rep, KC); ALLOC(ev.cand_first, KC); ALLOC(ev.cand_agg, KC); ALLOC(ev.cand_inex, KC);
    ALLOC(ev.cand_rank, KC); ALLOC(ev.cand_slot, KC);
    ALLOC(ev.uflag, KC); ALLOC(ev.ulist, KC); ALLOC(ev.pos_of, KC); ALLOC(ev.gflag, KC + 16);
    {   // cluster finish: per-CTA capacity of distinct candidates = its slice of the candidates
        const int per = (((C + FCL_CS * FCL_NT - 1) / (FCL_CS * FCL_NT)) + 3) & ~3;
        finish_cl_kcap = per * FCL_NT;
        finish_cl_smem = (size_t)finish_cl_kcap * 9 + 16;
        if (finish_cl_smem > 200 * 1024) finish_cl_max = 0;  // slices of very wide layers do not fit shared memory: one-CTA finish with global keys
        if (const char* e = getenv("DDO_FINISH_CL_MAX")) finish_cl_max = std::min(finish_cl_max > 0 ? 1 << 20 : 0, atoi(e));
        if (const char* e = getenv("DDO_EXPAND1_MIN")) expand1_min = atoi(e);
        if (const char* e = getenv("DDO_COMPACT1_MIN")) compact1_min = atoi(e);
    }
    // keys (8 B) + status (1 B) of up to C distinct candidates: shared memory when they fit next to the 19 KB of static smem
    finish_smem = (size_t)C * 9 + 16;
    ev.smem_keys = finish_smem <= 200 * 1024;
    if (ev.smem_keys) { ev.gkeys = nullptr; ev.ustat = nullptr; }
    else { finish_smem = 0; ALLOC(ev.gkeys, KC); ALLOC(ev.ustat, KC); }
    ALLOC(ev.dq, 8); ALLOC(ev.dq_jobs, K);
    ev.dd_generic = 0; ev.dd_prof = nullptr; ev.dd_dbg = 0;
    if (const char* e = getenv("DDO_DD_DBG")) ev.dd_dbg = atoi(e);
    ev.dd_dbgbuf = nullptr;
    if (ev.dd_dbg & 8) { ALLOC(ev.dd_dbgbuf, 8 * (size_t)Lmax); CUDA_TRY(cudaMemsetAsync(ev.dd_dbgbuf, 0, 32 * (size_t)Lmax, stream)); }
    if (const char* e = getenv("DDO_DD_PROF")) if (atoi(e)) { ALLOC(ev.dd_prof, 16); CUDA_TRY(cudaMemsetAsync(ev.dd_prof, 0, 128, stream)); }
    if (const char* e = getenv("DDO_DD_GENERIC")) ev.dd_generic = atoi(e) != 0;
    ALLOC(ev.table, (size_t)2 * K * T); /* k_dd alternates between two tables by layer parity */ ALLOC(ev.vhist, (size_t)K * 64 * S); ALLOC(ev.ucount, K);
    ALLOC(ev.plog, pool_layers * Wcap); ALLOC(ev.clog, pool_layers * C); ALLOC(ev.nlog, KL); ALLOC(ev.vlog, KL); ALLOC(ev.rslog, KL * 2);
    ALLOC(ev.lel_state, KW * S); ALLOC(ev.lel_val, KW); ALLOC(ev.lel_rub, KW);
    ALLOC(ev.cs_ub, KW); ALLOC(ev.cs_marked, KW);
    ALLOC(ev.best_path, (size_t)K * PW); ALLOC(ev.best_exact_path, (size_t)K * PW);
    // drain buffers: the cutsets of one batch (at most Klog slots' worth of records: a batch of more, shallower DDs drains fewer nodes each)
    out_cap = (size_t)Klog * Wcap;
    ev.fc_node = nullptr; ev.fc_ub = nullptr; ev.fc_aux = nullptr; ev.fc_cap = 0; d_out.tt = nullptr;
    if (cutset == DDO_FRONTIER) { int fr = alloc_frontier(&ev.fc_node, &ev.fc_ub, &ev.fc_aux, &ev.fc_cap); if (fr != DDO_OK) return fr; }
    ALLOC(d_out.state, out_cap * S); ALLOC(d_out.val, out_cap); ALLOC(d_out.ub, out_cap); ALLOC(d_out.dd, out_cap); ALLOC(d_out.path, out_cap * PW);
    ALLOC(d_out.count, K + 1); ALLOC(d_out.offset, K + 1); ALLOC(d_out.loc, KW);
    ALLOC(d_ub_cap, K); ALLOC(d_lb_filter, K);
    if (const char* e = getenv("DDO_FINISH_SPLIT_MIN")) finish_split_min = atoi(e);
    if (const char* e = getenv("DDO_EXPAND2")) expand2 = atoi(e) != 0;
    if (const char* e = getenv("DDO_DD")) dd_enabled = atoi(e) != 0;
    if (const char* e = getenv("DDO_DD_CS")) dd_cs = atoi(e);
    if (const char* e = getenv("DDO_DUAL")) dual_enabled = atoi(e) != 0;
    if (const char* e = getenv("DDO_PDL")) pdl_enabled = atoi(e) != 0;
    if (const char* e = getenv("DDO_LAYER_CHUNK")) layer_chunk = std::max(1, atoi(e));
    if (cutset == DDO_FRONTIER) dual_enabled = false;  // the twin's logs start at its fork layer; the frontier sweep reads whole DDs
    if (const char* e = getenv("DDO_SMALL_WS")) { int v = atoi(e); if (v == 0 || v == 64 || v == 128 || v == 256 || v == 512 || v == 1024) small_ws = v; }
    if (const char* e = getenv("DDO_SMALL_WS_FIRST")) { int v = atoi(e); if (v == 0 || v == 32 || v == 64 || v == 128) small_ws_first = v; }
    CUDA_TRY(cudaMemsetAsync(ev.table, 0xFF, (size_t)2 * K * T * 8, stream));
    CUDA_TRY(cudaMemsetAsync(ev.finish_counter, 0, 16, stream));
    { int rr = reserve_roots(K); if (rr != DDO_OK) return rr; }
    CUDA_TRY(cudaMallocHost(&h_ctl, (size_t)K * sizeof(DDCtl)));
    CUDA_TRY(cudaMallocHost(&h_active, 16));
    CUDA_TRY(cudaMallocHost(&h_caps, (size_t)K * 16));
    CUDA_TRY(cudaMallocHost(&h_counts, (size_t)(K + 1) * 8));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return DDO_OK;
}

void Engine::destroy() {
    if (ev.dd_dbgbuf) {
        std::vector<int> h(8 * (size_t)Lmax);
        if (cudaMemcpy(h.data(), ev.dd_dbgbuf, h.size() * 4, cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int t = 0; t < Lmax; ++t) if (h[4 * t + 3]) fprintf(stderr, "[dbg] t=%d nodes=%d cands=%d claimers=%d firsts=%d%s\n", t, h[4 * t + 3], h[4 * t + 2], h[4 * t], h[4 * t + 1], h[4 * t] != h[4 * t + 1] ? "  <-- MISMATCH" : "");
        for (int t = 0; t < Lmax; ++t) if (h[4 * t + 3]) fprintf(stderr, "[sum] t=%d val=%d pc=%d fl=%d h=%d\n", t, h[4 * Lmax + 4 * t], h[4 * Lmax + 4 * t + 1], h[4 * Lmax + 4 * t + 2], h[4 * Lmax + 4 * t + 3]);
    }
    if (ev.dd_prof) {  // DDO_DD_PROF=1: cycles per phase of k_dd over the life of the engine
        long long h[16];
        if (cudaMemcpy(h, ev.dd_prof, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const char* names[9] = {"trail", "expand", "S1 wait", "first/list", "S2 wait", "cut", "merge", "commit", "S_end wait"};
            long long tot = 0; for (int i = 0; i < 9; ++i) tot += h[i];
            fprintf(stderr, "[k_dd phases, rank 0 thread 0 cycles]");
            for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.1f%%", names[i], tot ? 100.0 * h[i] / tot : 0.0);
            fprintf(stderr, " | total %.2f ms @1.9GHz\n", tot / 1.9e6);
        }
    }
    for (cudaEvent_t e : prof_events) cudaEventDestroy(e);
    prof_events.clear();
    for (void* p : allocations) cudaFree(p);
    allocations.clear();
    for (void* p : {(void*)ev.root_state, (void*)ev.root_val, (void*)ev.root_depth, (void*)ev.root_width, (void*)d_small}) if (p) cudaFree(p);
    for (void* p : {(void*)h_root_state, (void*)h_root_val, (void*)h_root_depth, (void*)h_root_width, (void*)h_ctl, (void*)h_active, (void*)h_caps,
                    (void*)h_counts, (void*)h_small, (void*)h_out_state, (void*)h_out_val, (void*)h_out_ub, (void*)h_out_dd, (void*)h_out_path,
                    (void*)h_out_tt})
        if (p) cudaFreeHost(p);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
}

// (re)allocates the root staging area for `count` roots: the general engine uses the first K, the fast path all of them
int Engine::reserve_roots(int count) {
    if (count <= root_cap) return DDO_OK;
    CUDA_TRY(cudaSetDevice(device));
    if (stream) CUDA_TRY(cudaStreamSynchronize(stream));
    for (void* p : {(void*)ev.root_state, (void*)ev.root_val, (void*)ev.root_depth, (void*)ev.root_width, (void*)d_small}) if (p) cudaFree(p);
    for (void* p : {(void*)h_root_state, (void*)h_root_val, (void*)h_root_depth, (void*)h_root_width, (void*)h_small}) if (p) cudaFreeHost(p);
    CUDA_TRY(cudaMalloc((void**)&ev.root_state, (size_t)count * S * 8));
    CUDA_TRY(cudaMalloc((void**)&ev.root_val, (size_t)count * 4));
    CUDA_TRY(cudaMalloc((void**)&ev.root_depth, (size_t)count * 4));
    CUDA_TRY(cudaMalloc((void**)&ev.root_width, (size_t)count * 4));
    CUDA_TRY(cudaMalloc((void**)&d_small, (size_t)count * sizeof(SmallOut)));
    CUDA_TRY(cudaMallocHost(&h_root_state, (size_t)count * S * 8));
    CUDA_TRY(cudaMallocHost(&h_root_val, (size_t)count * 4));
    CUDA_TRY(cudaMallocHost(&h_root_depth, (size_t)count * 4));
    CUDA_TRY(cudaMallocHost(&h_root_width, (size_t)count * 4));
    CUDA_TRY(cudaMallocHost(&h_small, (size_t)count * sizeof(SmallOut)));
    root_cap = count;
    return DDO_OK;
}

int Engine::slots_for(int layers_needed) const {
    if (cutset_type == DDO_FRONTIER) return std::min(K, Klog);
    return (int)std::min<size_t>((size_t)K, pool_layers / (size_t)std::max(1, layers_needed));
}

int Engine::layers_bound(const uint64_t* state, int depth) const {
    int b = n_vars - depth + 1;
    if (model && state) {  // the MISP model
        int pc = 0;
        for (int j = 0; j < abi_words; ++j) pc += __builtin_popcountll(state[j]);
        b = std::min(b, pc + 2);
    }
    return std::max(1, std::min(b, Lmax));
}

int Engine::stage_roots(int count, const uint64_t* widths, const uint64_t* states, const int64_t* values, const int32_t* depths) {
    if (count < 1 || count > root_cap) { set_error("batch larger than batch_cap"); return DDO_ERR_CAPACITY; }
    const int words = abi_words;
    for (int i = 0; i < count; ++i) {
        if (widths[i] > (uint64_t)Wcap) { set_error("max_width larger than max_width_cap"); return DDO_ERR_CAPACITY; }
        if (values[i] < -(1ll << 30) || values[i] > (1ll << 30)) { set_error("root value outside the 31-bit device range"); return DDO_ERR_UNSUPPORTED; }
        if (depths[i] < 0 || depths[i] > n_vars) { set_error("root depth out of range"); return DDO_ERR_INVALID; }
        for (int j = 0; j < S; ++j) h_root_state[(size_t)i * S + j] = j < words ? states[(size_t)i * words + j] : 0;
        if (model && (model->n & 63)) {
            uint64_t mask = (1ull << (model->n & 63)) - 1;
            if (h_root_state[(size_t)i * S + words - 1] & ~mask) { set_error("root state has bits beyond nb_variables"); return DDO_ERR_INVALID; }
        }
        h_root_val[i] = (int32_t)values[i]; h_root_depth[i] = depths[i]; h_root_width[i] = (int32_t)widths[i];
    }
    CUDA_TRY(cudaSetDevice(device));
    bytes_h2d += (unsigned long long)((size_t)count * S * 8); CUDA_TRY(cudaMemcpyAsync(ev.root_state, h_root_state, (size_t)count * S * 8, cudaMemcpyHostToDevice, stream));
    bytes_h2d += (unsigned long long)((size_t)count * 4); CUDA_TRY(cudaMemcpyAsync(ev.root_val, h_root_val, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
    bytes_h2d += (unsigned long long)((size_t)count * 4); CUDA_TRY(cudaMemcpyAsync(ev.root_depth, h_root_depth, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
    bytes_h2d += (unsigned long long)((size_t)count * 4); CUDA_TRY(cudaMemcpyAsync(ev.root_width, h_root_width, (size_t)count * 4, cudaMemcpyHostToDevice, stream));
    staged = count;
    return DDO_OK;
}

void Engine::prof_mark(int kind) {
    if (!profiling) return;
    if (prof_used == prof_events.size()) { cudaEvent_t e; cudaEventCreate(&e); prof_events.push_back(e); prof_kinds.push_back(0); }
    prof_kinds[prof_used] = kind;
    cudaEventRecord(prof_events[prof_used++], stream);
}
int Engine::prof_collect() {
    if (!profiling || prof_used == 0) { prof_used = 0; return DDO_OK; }
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (size_t i = 1; i < prof_used; ++i) {
        const int kind = prof_kinds[i];
        if (kind < 0) continue;  // interval start marker
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, prof_events[i - 1], prof_events[i]));
        prof_ms[kind] += ms; prof_launches[kind] += 1;
    }
    prof_used = 0;
    return DDO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// persistent whole-DD kernel (dd_kernel.cuh): one cluster per DD, the layer loop inside the kernel
// ---------------------------------------------------------------------------------------------------------------
static DDLayout dd_layout(const Engine* E, int cs) {
    DDLayout L{};
    const int S = E->S, HN = 64 * S;
    L.slice = (((E->Wcap + cs - 1) / cs) + 31) & ~31;
    L.maxch = L.slice / 32;
    L.capc = 2 * L.slice;
    L.weighted = E->model && !E->model->unit_weights;
    unsigned o = 0;
    auto take = [&](size_t bytes) { const unsigned at = o; o += (unsigned)((bytes + 15) & ~(size_t)15); return at; };
    L.o_D = take((size_t)2 * HN * 4); L.o_master = take((size_t)HN * 4);
    L.o_stage = take(std::max((size_t)DD_NW * 32 * (2 * S + 1) * 4, (size_t)DD_GCAP * 9));  // (the gathered boundary bucket aliases the staging rows)
    L.o_cnt = take((size_t)L.maxch * 4); L.o_off = take((size_t)L.maxch * 4); L.o_koff = take((size_t)L.maxch * 4);
    L.o_fb = take((size_t)L.maxch * 8); L.o_kb = take((size_t)L.maxch * 8);
    L.o_keys = take(std::max((size_t)L.capc, (size_t)DD_GCAP) * 8); L.o_ulist = take((size_t)L.capc * 4); L.o_stat = take((size_t)L.capc);
    L.o_lh = take((size_t)DD_NB * 4); L.o_gh = take((size_t)DD_NB * 4);
    L.o_agg = take((size_t)L.capc * 8); L.o_first = take((size_t)L.capc * 4); L.o_rs = take((size_t)L.capc * 4); L.o_f = take((size_t)L.capc * 4);
    L.o_pos = take((size_t)L.capc * 4); L.o_inex = take((size_t)L.capc); L.o_pcy = take((size_t)L.slice * 2);
    L.o_nm = take((size_t)2 * L.slice * 16); L.o_rub = take(L.weighted ? (size_t)2 * L.slice * 4 : 16);
    L.total = o;
    return L;
}

// smallest cluster size whose per-CTA slice of the records fits the shared memory of an SM (0: none does, the layer-by-layer kernels run)
static int dd_min_cs(const Engine* E, size_t max_dyn) {
    for (int cs = 1; cs <= DD_MAXCS; cs <<= 1) if (dd_layout(E, cs).total <= max_dyn) return cs;
    return 0;
}

template <int S>
static int run_dd(Engine* E, int count, int slots, int comp_type, int64_t best_lb) {
    const EV& ev = E->ev;
    cudaStream_t st = E->stream;
    const int dual = slots > count;
    // cluster size: the smallest whose slice of the records fits an SM; doubled while every DD of the batch (twins included) still gets
    // a cluster of its own (latency of a lone wide DD)
    int cs = E->dd_min_cs;
    while (2 * cs <= DD_MAXCS && slots * 2 * cs <= E->num_sms) cs *= 2;
    if (E->dd_cs >= E->dd_min_cs && E->dd_cs <= DD_MAXCS) cs = E->dd_cs;
    const DDLayout L = dd_layout(E, cs);
    int nclusters = std::min(slots, std::max(1, E->num_sms / cs));
    k_dd_init<<<(slots + 127) / 128, 128, 0, st>>>(ev, count, comp_type, (long long)best_lb, dual);
    E->prof_mark(-1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nclusters * cs); cfg.blockDim = dim3(DD_NT); cfg.dynamicSmemBytes = L.total; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (ev.dd_prof) {
        int maxc = -1;
        cudaOccupancyMaxActiveClusters(&maxc, k_dd<S>, &cfg);
        fprintf(stderr, "[k_dd launch] slots %d cluster size %d clusters %d (max active clusters %d) dynamic smem %u B\n", slots, cs, nclusters, maxc, L.total);
    }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k_dd<S>, ev, L, count, dual));
    E->prof_mark(0);
    g_kernel_launches += 2; ++E->dd_launches;
    return DDO_OK;
}

// launch with (pdl) or without programmatic stream serialization: see pdl_enter() in kernels.cuh
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// `count` DDs are initialised from the staged roots; `slots` (= count, or 2*count in dual mode) DD slots take part in every launch
template <int S>
static int run_layers(Engine* E, int count, int slots, int comp_type, int64_t best_lb, const volatile int32_t* cutoff_flag) {
    constexpr int G = S / 2;
    const EV& ev = E->ev;
    cudaStream_t st = E->stream;
    if (E->finish_smem && !E->finish_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(k_finish<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E->finish_smem));
        E->finish_attr_set = true;
    }
    // one thread per node (k_expand1) measured faster than G lanes per node (k_expand) at every batch size; DDO_EXPAND1_MIN=<slots>
    // brings the lane-group kernel back for batches below that size (A/B runs)
    const bool use_e1 = slots >= E->expand1_min;
    const bool use_c1 = slots >= E->compact1_min;
    const size_t e1_smem = (size_t)512 * G * 16;
    const int e1_grid = E->num_sms * DDO_EXPAND1_MINB;  // one resident wave of CTAs (contiguous tile ranges)
    if (use_e1 && !E->expand1_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(k_expand1<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e1_smem));
        E->expand1_attr_set = true;
    }
    const bool use_cl = slots <= E->finish_cl_max;
    if (use_cl && !E->finish_cl_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(k_finish_cl<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E->finish_cl_smem));
        E->finish_cl_attr_set = true;
    }
    if (E->dd_enabled && E->model && !E->dd_attr_set) {  // once per engine: can k_dd hold a slice of this engine's widest layer in shared memory ?
        cudaFuncAttributes fa{};
        CUDA_TRY(cudaFuncGetAttributes(&fa, k_dd<S>));
        int optin = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, E->device));
        const size_t max_dyn = (size_t)optin - fa.sharedSizeBytes;
        E->dd_min_cs = dd_min_cs(E, max_dyn);
        if (E->dd_min_cs > 0) {
            CUDA_TRY(cudaFuncSetAttribute(k_dd<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn));
            CUDA_TRY(cudaFuncSetAttribute(k_dd<S>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        E->dd_attr_set = true;
    }
    bool use_dd = E->dd_enabled && E->model != nullptr && E->dd_min_cs > 0;
    for (int i = 0; i < count && use_dd; ++i) if (E->h_root_width[i] < 1) use_dd = false;  // max_width 0 (restricted): the layer-by-layer kernels keep that corner
    if (use_dd) {
        if (cutoff_flag && *cutoff_flag) return DDO_CUTOFF;
        const int rc = run_dd<S>(E, count, slots, comp_type, best_lb);
        if (rc != DDO_OK) return rc;
    } else {
    k_init<S><<<slots, 64, 0, st>>>(ev, count, comp_type, (long long)best_lb, slots > count);
    ++g_kernel_launches;
    E->prof_mark(-1);
    const int npb = 256 / G;
    // flat kernels run grid-stride over the per-layer work plan; the grid only has to be large enough to fill the machine
    const long long max_tiles = (long long)slots * ((E->C + npb - 1) / npb);
    const int flat_grid = (int)std::min<long long>(max_tiles, (long long)E->num_sms * 8);
    const int CHUNK = E->layer_chunk;  // layers launched between two polls of the device's `active` counter (and of the cutoff flag)
    const bool pdl = E->pdl_enabled && !E->profiling;  // the profiling events between launches serialise the stream anyway
    for (int t = 0; t < E->Lcur; ++t) {
        if (use_cl) CUDA_TRY(launch_k(pdl, k_finish_cl<S>, dim3(slots * FCL_CS), dim3(FCL_NT), E->finish_cl_smem, st, ev, t, E->finish_cl_kcap));
        else if (slots < E->finish_split_min) CUDA_TRY(launch_k(pdl, k_finish<S>, dim3(slots), dim3(1024), E->finish_smem, st, ev, t, 0));
        else {  // a large batch: the wide DDs one per SM, the narrow ones five to an SM (kernels.cuh, k_finish_s)
            CUDA_TRY(launch_k(pdl, k_finish<S>, dim3(std::min(slots, E->num_sms)), dim3(1024), E->finish_smem, st, ev, t, 1));
            CUDA_TRY(launch_k(pdl, k_finish_s<S>, dim3(std::min(slots, E->num_sms * (FIN_S_NT == 128 ? 7 : 6))), dim3(FIN_S_NT), 0, st, ev, t, slots));
            ++g_kernel_launches;
        }
        E->prof_mark(1);
        if (use_c1) CUDA_TRY(launch_k(pdl, k_compact1<S>, dim3(flat_grid), dim3(256), 0, st, ev, t, slots));
        else CUDA_TRY(launch_k(pdl, k_compact<S>, dim3(flat_grid), dim3(256), 0, st, ev, t, slots));
        E->prof_mark(2);
        if (E->expand2) CUDA_TRY(launch_k(pdl, k_expand2<S>, dim3(E->num_sms * 3), dim3(256), 0, st, ev, t, slots));
        else if (use_e1) CUDA_TRY(launch_k(pdl, k_expand1<S>, dim3(e1_grid), dim3(256), e1_smem, st, ev, t, slots));
        else CUDA_TRY(launch_k(pdl, k_expand<S>, dim3(flat_grid), dim3(256), 0, st, ev, t, slots));
        E->prof_mark(0);
        g_kernel_launches += 3; ++E->layer_steps;
        if ((t % CHUNK) == CHUNK - 1 || t == E->Lcur - 1) {
            E->bytes_d2h += sizeof(int); CUDA_TRY(cudaMemcpyAsync(E->h_active, ev.active, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (*E->h_active <= 0) break;
            if (cutoff_flag && *cutoff_flag) return DDO_CUTOFF;  // Cutoff::must_stop polled between layers (clean.rs:352)
        }
    }
    }
    // one more k_finish turns TERMINAL into DONE; harmless otherwise
    k_finalize<<<(slots + 63) / 64, 64, 0, st>>>(ev, slots);
    ++g_kernel_launches;
    if (E->cutset_type == DDO_FRONTIER) {
        // clean.rs:586-606 + 448-475: frontier membership and local bounds in one bottom-up sweep, then the upper bound of every member
        if (comp_type == DDO_RELAXED) {
            k_fc_sweep<<<slots, 1024, 0, st>>>(ev);
            k_fc_eval<S><<<dim3(64, slots), 256, 0, st>>>(ev);
            g_kernel_launches += 2;
        }
    } else if (comp_type == DDO_RELAXED || slots > count) { k_bottomup<<<slots, 1024, 0, st>>>(ev); ++g_kernel_launches; }
    E->prof_mark(3);
    CUDA_TRY(cudaGetLastError());
    return DDO_OK;
}

static int compile_impl(Engine* E, int count, int slots, int comp_type, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms) {
    // the log stride of this batch: the layers its deepest DD can have (every kernel and every host-side read of the logs uses ev.Lmax)
    E->staged_layers = 1;  // (from the host copies of the staged roots: the fast path stages every wave and never needs it)
    for (int i = 0; i < count; ++i) E->staged_layers = std::max(E->staged_layers, E->layers_bound(E->h_root_state + (size_t)i * E->S, E->h_root_depth[i]));
    E->Lcur = std::min(E->Lmax, std::max(1, E->staged_layers));
    if ((size_t)slots * (size_t)E->Lcur > E->pool_layers) { set_error("batch too large for the log pool: slots x layers of its deepest DD exceed it (see Engine::slots_for, Engine::layers_bound)"); return DDO_ERR_CAPACITY; }
    E->ev.Lmax = E->Lcur;
    const EV& ev = E->ev;
    CUDA_TRY(cudaSetDevice(E->device));
    CUDA_TRY(cudaMemsetAsync(ev.table, 0xFF, (size_t)slots * E->T * 8, E->stream));
    if (E->dd_enabled && E->model) CUDA_TRY(cudaMemsetAsync(ev.table + (size_t)E->K * E->T, 0xFF, (size_t)slots * E->T * 8, E->stream));
    CUDA_TRY(cudaMemsetAsync(ev.vhist, 0, (size_t)slots * 64 * E->S * 4, E->stream));
    CUDA_TRY(cudaMemsetAsync(ev.ucount, 0, (size_t)slots * 4, E->stream));
    CUDA_TRY(cudaEventRecord(E->ev0, E->stream));
    int rc;
    switch (E->S) {
        case 2: rc = run_layers<2>(E, count, slots, comp_type, best_lb, cutoff_flag); break;
        case 4: rc = run_layers<4>(E, count, slots, comp_type, best_lb, cutoff_flag); break;
        case 8: rc = run_layers<8>(E, count, slots, comp_type, best_lb, cutoff_flag); break;
        case 16: rc = run_layers<16>(E, count, slots, comp_type, best_lb, cutoff_flag); break;
        default: set_error("unsupported state width"); return DDO_ERR_UNSUPPORTED;
    }
    CUDA_TRY(cudaEventRecord(E->ev1, E->stream));
    CUDA_TRY(cudaStreamSynchronize(E->stream));
    if (device_ms) CUDA_TRY(cudaEventElapsedTime(device_ms, E->ev0, E->ev1));
    { int prc = E->prof_collect(); if (prc != DDO_OK) return prc; }
    E->last_count = slots; E->last_comp_type = slots > count ? DDO_RELAXED : comp_type; E->ctl_fetched = false;
    return rc;
}

int Engine::compile_staged(int count, int comp_type, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms) {
    if (count < 1 || count > K || count > staged) { set_error("compile: batch not staged"); return DDO_ERR_INVALID; }
    if (comp_type != DDO_EXACT && comp_type != DDO_RELAXED && comp_type != DDO_RESTRICTED) { set_error("bad compilation type"); return DDO_ERR_INVALID; }
    for (int i = 0; i < count; ++i)
        if (comp_type == DDO_RELAXED && h_root_width[i] < 1) { set_error("max_width must be >= 1 for a relaxed DD (the reference panics at clean.rs:827)"); return DDO_ERR_INVALID; }
    if (cutoff_flag && *cutoff_flag) return DDO_CUTOFF;
    return compile_impl(this, count, count, comp_type, best_lb, cutoff_flag, device_ms);
}

int Engine::compile_dual(int half, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms) {
    if (half < 1 || 2 * half > K || half > staged) { set_error("compile_dual: needs two DD slots per sub-problem"); return DDO_ERR_INVALID; }
    for (int i = 0; i < half; ++i)
        if (h_root_width[i] < 1) { set_error("max_width must be >= 1"); return DDO_ERR_INVALID; }
    if (cutoff_flag && *cutoff_flag) return DDO_CUTOFF;
    return compile_impl(this, half, 2 * half, DDO_RESTRICTED, best_lb, cutoff_flag, device_ms);
}

template <int S>
static int launch_small(Engine* E, int count, int64_t best_lb, int Ws) {
    const int G = S / 2;
    auto smem_of = [&](int ws) { return (size_t)ws * G * 16 * 3 + (size_t)ws * 4 * 3 + (size_t)4 * ws * 4 + (size_t)64 * S * 4 + (size_t)2 * ws; };
    const size_t smem = smem_of(Ws);
    if (!E->small_attr_set) {
        CUDA_TRY(cudaFuncSetAttribute(k_small<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(E->small_ws)));
        E->small_attr_set = true;
    }
    E->prof_mark(-1);
    k_small<S><<<count, 128, smem, E->stream>>>(E->ev, count, Ws, (long long)best_lb, E->d_small);
    ++g_kernel_launches;
    E->prof_mark(5);
    CUDA_TRY(cudaGetLastError());
    return DDO_OK;
}

int Engine::compile_small(int count, int64_t best_lb, float* device_ms) {
    const int rc = compile_small_launch(count, best_lb);
    if (rc != DDO_OK) return rc;
    return compile_small_wait(device_ms);
}
int Engine::compile_small_wait(float* device_ms) {
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (device_ms) CUDA_TRY(cudaEventElapsedTime(device_ms, ev0, ev1));
    { int prc = prof_collect(); if (prc != DDO_OK) return prc; }
    return DDO_OK;
}
// launch + result copy are only enqueued: the host may prepare the next wave while the device works (Solver::wave)
int Engine::compile_small_launch(int count, int64_t best_lb, int ws) {
    if (ws <= 0 || ws > small_ws) ws = small_ws;
    if (count < 1 || count > root_cap || count > staged) { set_error("compile_small: batch not staged"); return DDO_ERR_INVALID; }
    if (small_ws <= 0) { set_error("small path disabled"); return DDO_ERR_INVALID; }
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaEventRecord(ev0, stream));
    int rc;
    switch (S) {
        case 2: rc = launch_small<2>(this, count, best_lb, ws); break;
        case 4: rc = launch_small<4>(this, count, best_lb, ws); break;
        case 8: rc = launch_small<8>(this, count, best_lb, ws); break;
        case 16: rc = launch_small<16>(this, count, best_lb, ws); break;
        default: set_error("unsupported state width"); return DDO_ERR_UNSUPPORTED;
    }
    if (rc != DDO_OK) return rc;
    CUDA_TRY(cudaEventRecord(ev1, stream));
    CUDA_TRY(cudaMemcpyAsync(h_small, d_small, (size_t)count * sizeof(SmallOut), cudaMemcpyDeviceToHost, stream));
    return DDO_OK;
}

int Engine::fetch_ctl(int count) {
    if (ctl_fetched && count <= last_count) {
        if (ctl_overflow) { set_error("a layer outgrew max_width_cap (Exact compilation needs a larger cap)"); return DDO_ERR_CAPACITY; }
        return DDO_OK;
    }
    CUDA_TRY(cudaSetDevice(device));
    bytes_d2h += (unsigned long long)((size_t)last_count * sizeof(DDCtl)); CUDA_TRY(cudaMemcpyAsync(h_ctl, ev.ctl, (size_t)last_count * sizeof(DDCtl), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    ctl_fetched = true; ctl_overflow = false;
    for (int i = 0; i < last_count; ++i)
        if (h_ctl[i].overflow) { ctl_overflow = true; set_error("a layer outgrew max_width_cap (Exact compilation needs a larger cap)"); return DDO_ERR_CAPACITY; }
    return DDO_OK;
}

void Engine::fill_completion(int i, ddo_completion* out) const {
    const DDCtl& c = h_ctl[i];
    std::memset(out, 0, sizeof(*out));
    out->is_exact = (c.lel < 0) || c.ebpo;  // clean.rs:241-243: is_exact || has_exact_best_path
    out->has_best_value = c.has_best; out->best_value = c.has_best ? c.best_value : 0;
    out->has_best_exact = c.has_best_exact; out->best_exact_value = c.has_best_exact ? c.best_exact_value : 0;
    out->cutset_size = c.cutset_count;
    out->lel_depth = c.lel < 0 ? -1 : c.root_depth + c.lel;
    out->n_layers = c.t_term + (c.has_best ? 1 : 0);
    out->expanded = c.expanded; out->transitions = c.transitions;
}

int Engine::best_solution(int index, int exact, ddo_decision* out, int32_t* len) {
    if (index < 0 || index >= last_count || !len) { set_error("best_solution: bad index"); return DDO_ERR_INVALID; }
    int rc = fetch_ctl(last_count);
    if (rc != DDO_OK) return rc;
    const DDCtl& c = h_ctl[index];
    if (!(exact ? c.has_best_exact : c.has_best)) { set_error("no such solution"); return DDO_ERR_INVALID; }
    const int L = c.t_term;  // decisions of layers 0..L-1
    if (*len < L) { *len = L; set_error("best_solution: buffer too small"); return DDO_ERR_CAPACITY; }
    std::vector<uint64_t> bits(PW);
    std::vector<int32_t> vars(Lmax);
    bytes_d2h += (unsigned long long)(PW * 8); CUDA_TRY(cudaMemcpyAsync(bits.data(), (exact ? ev.best_exact_path : ev.best_path) + (size_t)index * PW, PW * 8, cudaMemcpyDeviceToHost, stream));
    bytes_d2h += (unsigned long long)((size_t)Lmax * 4); CUDA_TRY(cudaMemcpyAsync(vars.data(), ev.vlog + (size_t)index * Lcur, (size_t)Lcur * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int i = 0; i < L; ++i) {  // reference order: terminal -> root (clean.rs:337-341)
        const int tt = L - 1 - i;
        out[i].variable = vars[tt];
        out[i].value = bit_value[(bits[tt >> 6] >> (tt & 63)) & 1];
    }
    *len = L;
    return DDO_OK;
}

int Engine::layer_trace(int index, int32_t* vars, int32_t* widths, int cap) {
    if (index < 0 || index >= last_count) { set_error("layer_trace: bad index"); return DDO_ERR_INVALID; }
    int rc = fetch_ctl(last_count);
    if (rc != DDO_OK) return rc;
    const DDCtl& c = h_ctl[index];
    const int L = std::max(0, c.t_term);  // expanded layers: 0..t_term-1
    std::vector<int32_t> v(Lmax), w(Lmax);
    bytes_d2h += (unsigned long long)((size_t)Lmax * 4); CUDA_TRY(cudaMemcpyAsync(v.data(), ev.vlog + (size_t)index * Lcur, (size_t)Lcur * 4, cudaMemcpyDeviceToHost, stream));
    bytes_d2h += (unsigned long long)((size_t)Lmax * 4); CUDA_TRY(cudaMemcpyAsync(w.data(), ev.nlog + (size_t)index * Lcur, (size_t)Lcur * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int i = 0; i < L && i < cap; ++i) { vars[i] = v[i]; widths[i] = w[i]; }
    return L;
}

// Batched drain: every DD i < count emits its MARKED cutset nodes with min(ub, ub_cap[i]) > lb_filter[i].
// Results land in the pinned h_out_* arrays; returns the total number of records (<0: error).
int Engine::drain_all(int count, const int64_t* ub_cap, const int64_t* lb_filter, int* pw_out) {
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
    for (int i = 0; i < K; ++i) { caps[2 * i] = i < count ? ub_cap[i] : 0; caps[2 * i + 1] = i < count ? lb_filter[i] : INT64_MAX; }
    // de-interleave on the device side via two strided copies
    std::vector<long long> a(K), b(K);
    for (int i = 0; i < K; ++i) { a[i] = caps[2 * i]; b[i] = caps[2 * i + 1]; }
    std::memcpy(caps, a.data(), (size_t)K * 8); std::memcpy(caps + K, b.data(), (size_t)K * 8);
    CUDA_TRY(cudaSetDevice(device));
    bytes_h2d += (unsigned long long)((size_t)K * 8); CUDA_TRY(cudaMemcpyAsync(d_ub_cap, caps, (size_t)K * 8, cudaMemcpyHostToDevice, stream));
    bytes_h2d += (unsigned long long)((size_t)K * 8); CUDA_TRY(cudaMemcpyAsync(d_lb_filter, caps + K, (size_t)K * 8, cudaMemcpyHostToDevice, stream));
    prof_mark(-1);
    k_cutset_count<<<last_count, 1024, 0, stream>>>(ev, d_out, d_ub_cap, d_lb_filter, count);
    k_cutset_offsets<<<1, 32, 0, stream>>>(d_out, last_count);
    g_kernel_launches += 2;
    bytes_d2h += (unsigned long long)((size_t)(last_count + 1) * 4); CUDA_TRY(cudaMemcpyAsync(h_counts, d_out.offset, (size_t)(last_count + 1) * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const int total = ((int32_t*)h_counts)[last_count];
    if (total == 0) { prof_used = 0; return 0; }
    if ((size_t)total > out_cap) { prof_used = 0; set_error("cutsets of the batch exceed the drain buffers (lower batch_cap or raise DDO_LOG_SLOTS)"); return DDO_ERR_CAPACITY; }
    if (pw > 8) CUDA_TRY(cudaMemsetAsync(d_out.path, 0, (size_t)total * pw * 8, stream));
    const dim3 grid((Wcap + 255) / 256, last_count);
    switch (S) {
        case 2: k_cutset_write<2><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
        case 4: k_cutset_write<4><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
        case 8: k_cutset_write<8><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
        default: k_cutset_write<16><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
    }
    ++g_kernel_launches;
    prof_mark(4);
    if (!h_out_state) {
        const size_t KW = out_cap;
        CUDA_TRY(cudaMallocHost(&h_out_state, KW * S * 8));
        CUDA_TRY(cudaMallocHost(&h_out_val, KW * 4));
        CUDA_TRY(cudaMallocHost(&h_out_ub, KW * 4));
        CUDA_TRY(cudaMallocHost(&h_out_dd, KW * 4));
        CUDA_TRY(cudaMallocHost(&h_out_path, KW * PW * 8));
        CUDA_TRY(cudaMallocHost(&h_out_tt, KW * 4));
    }
    bytes_d2h += (unsigned long long)((size_t)total * S * 8); CUDA_TRY(cudaMemcpyAsync(h_out_state, d_out.state, (size_t)total * S * 8, cudaMemcpyDeviceToHost, stream));
    bytes_d2h += (unsigned long long)((size_t)total * 4); CUDA_TRY(cudaMemcpyAsync(h_out_val, d_out.val, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    bytes_d2h += (unsigned long long)((size_t)total * 4); CUDA_TRY(cudaMemcpyAsync(h_out_ub, d_out.ub, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    bytes_d2h += (unsigned long long)((size_t)total * 4); CUDA_TRY(cudaMemcpyAsync(h_out_dd, d_out.dd, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    bytes_d2h += (unsigned long long)((size_t)total * pw * 8); CUDA_TRY(cudaMemcpyAsync(h_out_path, d_out.path, (size_t)total * pw * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int r = 0; r < total; ++r) h_out_tt[r] = h_ctl[h_out_dd[r]].lel;  // a LEL cutset is one layer of its DD
    { int prc = prof_collect(); if (prc != DDO_OK) return prc; }
    return total;
}

void Engine::fc_launch_count(int count) { k_fc_count<EV><<<last_count, 1024, 0, stream>>>(ev, d_out, d_ub_cap, d_lb_filter, count); }
void Engine::fc_launch_write(int pw) {
    const dim3 grid(64, last_count);
    switch (S) {
        case 2: k_fc_write<2><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
        case 4: k_fc_write<4><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
        case 8: k_fc_write<8><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
        default: k_fc_write<16><<<grid, 256, 0, stream>>>(ev, d_out, d_ub_cap, pw); break;
    }
}

// The same for a FRONTIER engine (frontier.cuh): the records of a DD come from different layers, h_out_tt holds the layer of each.
int Engine::drain_all_frontier(int count, const int64_t* ub_cap, const int64_t* lb_filter, int* pw_out) {
    int rc = fetch_ctl(last_count);
    if (rc != DDO_OK) return rc;
    int max_t = 0;
    for (int i = 0; i < count; ++i) max_t = std::max(max_t, h_ctl[i].t_term);
    const int pw = std::min(16, std::max(1, (max_t + 63) / 64));
    *pw_out = pw;
    long long* caps = (long long*)h_caps;
    for (int i = 0; i < K; ++i) { caps[i] = i < count ? ub_cap[i] : 0; caps[K + i] = i < count ? lb_filter[i] : INT64_MAX; }
    CUDA_TRY(cudaSetDevice(device));
    bytes_h2d += (unsigned long long)((size_t)K * 16);
    CUDA_TRY(cudaMemcpyAsync(d_ub_cap, caps, (size_t)K * 8, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(d_lb_filter, caps + K, (size_t)K * 8, cudaMemcpyHostToDevice, stream));
    prof_mark(-1);
    fc_launch_count(count);
    k_cutset_offsets<<<1, 32, 0, stream>>>(d_out, last_count);
    g_kernel_launches += 2;
    bytes_d2h += (unsigned long long)((size_t)(last_count + 1) * 4); CUDA_TRY(cudaMemcpyAsync(h_counts, d_out.offset, (size_t)(last_count + 1) * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    const int total = ((int32_t*)h_counts)[last_count];
    if (total == 0) { prof_used = 0; return 0; }
    if ((size_t)total > out_cap) { prof_used = 0; set_error("frontier cutset larger than the drain buffers (raise DDO_FC_OUT_FACTOR)"); return DDO_ERR_CAPACITY; }
    fc_launch_write(pw);
    ++g_kernel_launches;
    prof_mark(4);
    if (!h_out_state) {
        CUDA_TRY(cudaMallocHost(&h_out_state, out_cap * S * 8));
        CUDA_TRY(cudaMallocHost(&h_out_val, out_cap * 4));
        CUDA_TRY(cudaMallocHost(&h_out_ub, out_cap * 4));
        CUDA_TRY(cudaMallocHost(&h_out_dd, out_cap * 4));
        CUDA_TRY(cudaMallocHost(&h_out_path, out_cap * PW * 8));
        CUDA_TRY(cudaMallocHost(&h_out_tt, out_cap * 4));
    }
    bytes_d2h += (unsigned long long)((size_t)total * (S * 8 + 16 + pw * 8));
    CUDA_TRY(cudaMemcpyAsync(h_out_state, d_out.state, (size_t)total * S * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_val, d_out.val, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_ub, d_out.ub, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_dd, d_out.dd, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_tt, d_out.tt, (size_t)total * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(h_out_path, d_out.path, (size_t)total * pw * 8, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    { int prc = prof_collect(); if (prc != DDO_OK) return prc; }
    return total;
}

// branching variables of every DD of the last batch in ONE copy: [slots][Lmax] (a per-DD copy + synchronize costs ~10 us each, and a wide
// wave drains hundreds of DDs)
int Engine::fetch_vars_all(int slots, std::vector<int32_t>& vars) {
    vars.resize((size_t)slots * Lcur);
    bytes_d2h += (unsigned long long)((size_t)slots * Lcur * 4);
    CUDA_TRY(cudaMemcpyAsync(vars.data(), ev.vlog, (size_t)slots * Lcur * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return DDO_OK;
}

int Engine::fetch_vars(int index, std::vector<int32_t>& vars) {
    vars.resize(Lmax);
    bytes_d2h += (unsigned long long)((size_t)Lmax * 4); CUDA_TRY(cudaMemcpyAsync(vars.data(), ev.vlog + (size_t)index * Lcur, (size_t)Lcur * 4, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return DDO_OK;
}

int Engine::drain_cutset(int index, int64_t ub_cap, int64_t lb_filter, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depth_out,
                         int32_t* path_len_out, ddo_decision* paths, int32_t* count) {
    if (index < 0 || index >= last_count || !count) { set_error("drain_cutset: bad index"); return DDO_ERR_INVALID; }
    if (cutset_type == DDO_FRONTIER) {
        set_error("a FRONTIER cutset has one depth per node: use ddo_mdd_drain_cutset_batch + ddo_mdd_drain_layer_index");
        return DDO_ERR_UNSUPPORTED;
    }
    std::vector<int64_t> caps(last_count, 0), lbs(last_count, INT64_MAX);
    caps[index] = ub_cap; lbs[index] = lb_filter;
    int pw = 1;
    int total = drain_all(last_count, caps.data(), lbs.data(), &pw);
    if (total < 0) return total;
    const DDCtl& c = h_ctl[index];
    const int lel = std::max(0, c.lel);
    if (depth_out) *depth_out = c.root_depth + lel;
    if (path_len_out) *path_len_out = lel;
    if (total > *count) { *count = total; set_error("drain_cutset: buffer too small"); return DDO_ERR_CAPACITY; }
    std::vector<int32_t> vars;
    if (paths && total > 0) { int rc = fetch_vars(index, vars); if (rc != DDO_OK) return rc; }
    const int words = abi_words;
    for (int r = 0; r < total; ++r) {
        if (states) for (int j = 0; j < words; ++j) states[(size_t)r * words + j] = h_out_state[(size_t)r * S + j];
        if (values) values[r] = h_out_val[r];
        if (ubs) ubs[r] = h_out_ub[r];
        if (paths)
            for (int i = 0; i < lel; ++i) {  // terminal -> root order (clean.rs:329-343)
                const int tt = lel - 1 - i;
                paths[(size_t)r * lel + i].variable = vars[tt];
                paths[(size_t)r * lel + i].value = bit_value[(h_out_path[(size_t)r * pw + (tt >> 6)] >> (tt & 63)) & 1];
            }
    }
    *count = total;
    return DDO_OK;
}

}  // namespace ddo

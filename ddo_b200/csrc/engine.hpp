// engine.hpp -- host-visible declarations of the B200 batch DD-compilation engine (MISP device model).
// The engine compiles up to `K` decision diagrams in lock-step, one layer per step, entirely on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ddo_b200.h"

namespace ddo {

void set_error(const std::string& msg);
extern thread_local std::string g_last_error;
extern unsigned long long g_kernel_launches;

// ---------------------------------------------------------------------------------------------
// Model: ddo/examples/misp/main.rs:37-209 as device-resident arrays.
// ---------------------------------------------------------------------------------------------
struct MispModel {
    int n = 0;            // vertices
    int words = 0;        // ceil(n/64): words of a packed state at the ABI
    int S = 0;            // device words per state: power of two >= 2 (pad words are always zero)
    int device = 0;
    bool unit_weights = true;
    int64_t weight_abs_sum = 0;
    std::vector<int64_t> h_weight;
    std::vector<uint64_t> h_nc;  // n x S complement-adjacency rows (main.rs:40-45)
    int32_t* d_weight = nullptr;
    uint64_t* d_nc = nullptr;
};

enum : int32_t { ST_ACTIVE = 0, ST_TERMINAL = 1, ST_DONE = 2, ST_WAITING = 3 };  // WAITING: relaxed twin slot that has not forked (yet)
constexpr uint32_t NONE32 = 0xFFFFFFFFu;
constexpr uint64_t EMPTY64 = 0xFFFFFFFFFFFFFFFFull;
// node flag bits kept in cur_flag / uinex / parent log
constexpr uint32_t NF_INEXACT = 1u;  // !F_EXACT (node_flags.rs:51)
constexpr uint32_t NF_RELAXED = 2u;  // F_RELAXED (node_flags.rs:53)
constexpr uint32_t PLOG_INEXACT = 1u << 31, PLOG_RELAXED = 1u << 30, PLOG_CAND_MASK = (1u << 30) - 1;

// Per-DD control block (device resident; copied back once per compile)
struct DDCtl {
    int32_t status, ncand, n_cur, var;
    int32_t width, comp_type, root_depth, lel;  // lel: layer index of the last exact layer, -1 = never squashed
    int32_t lel_pending, t_term, has_best, has_best_exact;
    int32_t best_pos, best_exact_pos, best_value, best_exact_value;
    int32_t ebpo, overflow, cutset_count, lel_n;
    int32_t root_value, pad0;
    int32_t primary, fork_t;  // relaxed twin of a restricted DD: slot of the primary and layer at which it forked (-1: ordinary DD)
    int64_t best_lb;
    unsigned long long expanded, transitions;
};

// Everything a kernel needs, passed by value.
struct EV {
    int K, Wcap, C, T, Lmax, n, S, PW;  // PW = uint64 words of a packed decision-bit path
    int HN;                              // counters per DD in vhist (= 64 * S)
    int unit_weights;
    const int32_t* weight;
    const uint64_t* nc;
    DDCtl* ctl;
    int* active;  // number of DDs still compiling
    int* fin_list; int* fin_cnt;  // [2][K] DD slots whose next finish is wide / narrow, [2] their counts (written by the plan step of the previous layer)
    int* tile_off_e; int* tile_off_c; unsigned int* finish_counter;  // per-layer work plan of k_expand / k_compact (exclusive tile offsets, [K+1])
    // staged roots
    uint64_t* root_state; int32_t* root_val; int32_t* root_depth; int32_t* root_width;
    // current layer (ping-pong)
    uint64_t* cur_state[2]; int32_t* cur_val[2]; uint8_t* cur_flag[2]; int32_t* cur_rub;
    // candidates of the next layer
    uint64_t* cand_state; uint32_t* cand_rep; uint32_t* cand_first; unsigned long long* cand_agg; uint8_t* cand_inex; uint32_t* cand_rank; uint32_t* cand_slot;
    // unique nodes of the next layer
    uint8_t* uflag; uint32_t* ulist; uint8_t* ustat; uint32_t* pos_of;
    uint8_t* gflag;   // [K][C] canonical-candidate flags of the cluster finish (k_finish_cl)
    unsigned long long* gkeys; int smem_keys;  // cut keys of the distinct candidates: in k_finish's shared memory when 2*Wcap of them fit, else here
    unsigned long long* table;
    uint32_t* ucount; // [K] number of distinct states among the candidates being built (hash-slot claims of k_expand)
    uint32_t* vhist;  // [K][HN] occurrences of every vertex among the distinct states of the layer being built
    // logs
    uint32_t* plog;   // [K][Lmax][Wcap] best parent candidate + flags
    uint32_t* clog;   // [K][Lmax][C]    child position of every candidate (relaxed only)
    int32_t* nlog;    // [K][Lmax]       layer sizes
    int32_t* vlog;    // [K][Lmax]       branching variables
    int32_t* rslog;   // [K][Lmax][2]    (saved pos, recycled pos) of the recycled-merge corner case, else -1
    // last exact layer snapshot
    uint64_t* lel_state; int32_t* lel_val; int32_t* lel_rub;
    // bottom-up scratch + cutset outputs
    int32_t* vb[2];
    int32_t* cs_ub; uint8_t* cs_marked;
    // best paths (decision bits, one per layer) for best / best exact terminal node
    uint64_t* best_path; uint64_t* best_exact_path;
    // FRONTIER cutset (frontier.cuh; allocated by engines created with DDO_FRONTIER only): [K][fc_cap] records of the MARKED cutset nodes
    // in canonical order (layer descending, position ascending): node = layer << FC_POS_BITS | position, upper bound, scratch (value_bot, then
    // the local output index of a drain)
    uint32_t* fc_node; int32_t* fc_ub; int32_t* fc_aux; unsigned long long fc_cap;
    // persistent whole-DD kernel (dd_kernel.cuh): the device work queue and the twin jobs published by restricted DDs (the node and
    // candidate records of a DD live in the shared memory of its cluster)
    int* dq; int* dq_jobs;
    long long* dd_prof; // [16] cycles per phase of k_dd (rank 0 / thread 0), DDO_DD_PROF=1 only
    int* dd_dbgbuf;     // (debug, DDO_DD_DBG & 8) per layer: claimers, first candidates, candidates, nodes
    int dd_dbg;         // debug switches of k_dd (DDO_DD_DBG)
    int dd_generic;     // force the generic cluster-wide radix select of the width cut (DDO_DD_GENERIC=1, tests)
};
constexpr int FC_POS_BITS = 21;

// result of the shared-memory fast path (k_small), one per DD
struct SmallOut {
    int32_t status;      // 0 = compiled, 1 = overflow (recompile with the general engine)
    int32_t has_best, best_value, layers;
    unsigned long long expanded, transitions;
};

// batched drain_cutset output (device side)
struct DrainOut {
    uint64_t* state; int32_t* val; int32_t* ub; int32_t* dd; uint64_t* path;  // [total] records
    int32_t* tt;                      // [total] layer of the record inside its DD (FRONTIER engines only; a LEL cutset has one layer per DD)
    int32_t* count; int32_t* offset;  // [K+1]; offset[K] = total
    uint32_t* loc;                    // [K][Wcap] local index of each emitted node
};

struct Engine {
    const MispModel* model = nullptr;  // null for engines of other device models (M2Engine)
    int n_vars = 0, abi_words = 0;     // nb_variables and uint64 words of a packed state at the ABI
    int32_t bit_value[2] = {0, 1};     // decision value of a path bit: MISP NO / YES; MAX2SAT F = -1 / T = +1
    int device = 0;
    int K = 0, Wcap = 0, C = 0, T = 0, Lmax = 0, S = 0, PW = 0;
    // The parent / child logs (12 B per node and layer: 60 of the 65 MB of a DD slot at W = 10 000, n = 500) are ONE pool of `pool_layers`
    // layer records shared by the DDs of a batch: a batch of sub-problems deep in the search needs n - depth + 1 layers per DD, not n + 1,
    // so the same memory holds several times more DDs in lock-step (Lcur = log stride of the current batch, set by stage_roots).
    int Klog = 0, Lcur = 0; size_t pool_layers = 0; int staged_layers = 0;
    virtual int slots_for(int layers_needed) const;  // DD slots a batch whose deepest DD has `layers_needed` layers may use
    // Layers (log entries) a DD rooted at `state` / `depth` can have.  Every model: one per undecided variable, plus the terminal layer.
    // MISP: next_variable only returns a vertex that some state of the layer still holds (misp/main.rs:109-143), every state is a subset
    // of the root state and a branched vertex leaves every descendant, so a DD has at most popcount(root state) layers -- for the
    // sub-problems of G(500, 0.5) a quarter of n - depth, i.e. four times the DDs per batch in the same log pool.
    int layers_bound(const uint64_t* state, int depth) const;
    int cutset_type = DDO_LAST_EXACT_LAYER;
    int num_sms = 148;
    int layer_chunk = 16;      // layer steps launched between two host polls of the `active` counter / cutoff flag (DDO_LAYER_CHUNK)
    bool pdl_enabled = true;   // programmatic dependent launch of the layer-step kernels (DDO_PDL=0 disables; pdl_enter() in kernels.cuh)
    bool dual_enabled = true;  // fork the relaxed twin at the first cut of a restricted DD (DDO_DUAL=0 disables)
    size_t finish_smem = 0; bool finish_attr_set = false;
    int compact1_min = 1;  // thread-per-candidate compaction (k_compact1) for batches of >= compact1_min DD slots (DDO_COMPACT1_MIN)
    int expand1_min = 1; bool expand1_attr_set = false;  // thread-per-node expansion (k_expand1) for batches of >= expand1_min DD slots
    bool dd_enabled = false; int dd_cs = 0; bool dd_attr_set = false; int dd_min_cs = 0;  // persistent whole-DD kernel k_dd (opt-in: DDO_DD=1; see DESIGN.md section 4b for where it stands); dd_cs: forced cluster size (DDO_DD_CS), 0 = by batch size
    unsigned long long dd_launches = 0;
    bool expand2 = false;        // warp-autonomous expansion kernel k_expand2 (DDO_EXPAND2=1; measured slower than k_expand1: 279 vs 167 ms per config-2 solve -- its per-warp histogram flushes cost more global atomics than the block barriers they remove)
    int finish_split_min = 160;  // batches of at least this many DD slots run the finish in two size classes (k_finish + k_finish_s; DDO_FINISH_SPLIT_MIN)
    int finish_cl_max = 128; int finish_cl_kcap = 0; size_t finish_cl_smem = 0; bool finish_cl_attr_set = false;  // cluster finish: used for batches of <= finish_cl_max DD slots
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    EV ev{};
    std::vector<void*> allocations;
    // pinned host staging
    uint64_t* h_root_state = nullptr; int32_t* h_root_val = nullptr; int32_t* h_root_depth = nullptr; int32_t* h_root_width = nullptr;
    DDCtl* h_ctl = nullptr; int* h_active = nullptr;
    void* h_caps = nullptr; void* h_counts = nullptr;
    DrainOut d_out{}; long long* d_ub_cap = nullptr; long long* d_lb_filter = nullptr;
    // pinned results of the last drain_all (records [0,total))
    uint64_t* h_out_state = nullptr; int32_t* h_out_val = nullptr; int32_t* h_out_ub = nullptr; int32_t* h_out_dd = nullptr; uint64_t* h_out_path = nullptr;
    int32_t* h_out_tt = nullptr;   // layer of every record inside its DD (LEL: the DD's last exact layer)
    size_t out_cap = 0;            // records the drain buffers hold (LEL: K * Wcap)
    int last_drain_total = 0;      // records of the last ddo_mdd_drain_cutset_batch
    int last_count = 0; int last_comp_type = -1; int staged = 0; bool ctl_fetched = false; bool ctl_overflow = false;  // verdict of the last fetch_ctl: an overflowed batch keeps failing
    size_t bytes_allocated = 0;
    unsigned long long layer_steps = 0;  // layer steps (k_finish + k_compact + k_expand) launched so far
    unsigned long long bytes_h2d = 0, bytes_d2h = 0;  // traffic over PCIe / NVLink-C2C issued by this engine
    // optional per-kernel timing
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events; std::vector<int> prof_kinds; size_t prof_used = 0;
    double prof_ms[6] = {0, 0, 0, 0, 0, 0}; uint64_t prof_launches[6] = {0, 0, 0, 0, 0, 0};
    void prof_mark(int kind);   // record an event; the interval since the previous mark is attributed to `kind`
    int prof_collect();

    virtual ~Engine() = default;
    int create(const MispModel* m, int device, uint64_t max_width_cap, int batch_cap, int cutset_type);
    virtual void destroy();
    int root_cap = 0;  // roots that can be staged at once (>= K): the wave size of the solver's fast path
    virtual int reserve_roots(int count);
    int stage_roots(int count, const uint64_t* widths, const uint64_t* states, const int64_t* values, const int32_t* depths);
    virtual int compile_staged(int count, int comp_type, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms);
    // dual mode: `half` restricted DDs in slots [0,half); the relaxed twin of DD j forks into slot half+j at the first width cut and then
    // advances in the same launches (both against best_lb).  Results: ctl[j] restricted, ctl[half+j] relaxed (status WAITING = never forked).
    int compile_dual(int half, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms);
    int fetch_ctl(int count);
    void fill_completion(int i, ddo_completion* out) const;
    int best_solution(int index, int exact, ddo_decision* out, int32_t* len);
    int drain_cutset(int index, int64_t ub_cap, int64_t lb_filter, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depth_out,
                     int32_t* path_len_out, ddo_decision* paths, int32_t* count);
    int layer_trace(int index, int32_t* vars, int32_t* widths, int cap);
    // batched drain for the solver: records of every DD in h_out_*; returns total (<0 error); *pw = uint64 words of path bits per record
    virtual int drain_all(int count, const int64_t* ub_cap, const int64_t* lb_filter, int* pw);
    int drain_all_frontier(int count, const int64_t* ub_cap, const int64_t* lb_filter, int* pw);  // DDO_FRONTIER engines
    virtual void fc_launch_count(int count);   // the model's frontier drain kernels (frontier.cuh / m2s_frontier.cuh)
    virtual void fc_launch_write(int pw);
    int alloc_frontier(uint32_t** node, int32_t** ub, int32_t** aux, unsigned long long* cap);  // record arrays + drain buffers sized for a frontier
    // shared-memory fast path: every staged root compiled by one CTA (exact DDs only); results in h_small[0..count)
    int small_ws_first = 64;  // first-tier capacity of the fast path (more CTAs per SM); 0 = single tier
    int small_ws = 256; SmallOut* d_small = nullptr; SmallOut* h_small = nullptr; bool small_attr_set = false;
    int compile_small(int count, int64_t best_lb, float* device_ms);
    int compile_small_launch(int count, int64_t best_lb, int ws = 0);  // ws: fast-path capacity of this launch (0 = small_ws)
    int compile_small_wait(float* device_ms);
    int fetch_vars(int index, std::vector<int32_t>& vars);
    int fetch_vars_all(int slots, std::vector<int32_t>& vars);
};

int model_create_misp(int32_t n, const int64_t* weights, int64_t m, const int32_t* src, const int32_t* dst, int device, MispModel** out);
void model_destroy(MispModel*);

}  // namespace ddo

// m2s_engine.hpp -- MAX2SAT device model (ddo/examples/max2sat) behind the same engine interface as the MISP model.
#pragma once
#include "engine.hpp"

namespace ddo {

// ddo/examples/max2sat/model.rs:98-249 as device-resident tables (built by model_create_max2sat, m2s_engine.cu)
struct M2Model {
    int n = 0;        // variables
    int words = 0;    // uint64 words of a packed state at the ABI: n int32 benefits, two per word
    int NW = 0;       // int32 entries of a device state row (n rounded up to a multiple of 4; pad entries are always zero)
    int device = 0;
    long long initial = 0;  // sum of the tautological clause weights = initial_value(), model.rs:145-147,266-269
    std::vector<int32_t> h_ord;
    int32_t* d_ord = nullptr; int32_t *d_PT = nullptr, *d_QT = nullptr, *d_PF = nullptr, *d_QF = nullptr, *d_AT = nullptr, *d_AF = nullptr;
    long long *d_est = nullptr, *d_nk = nullptr;
    uint32_t* d_hmul = nullptr;
};

// ---- kernel argument blocks (passed by value) ---------------------------------------------------------------------------------
struct M2Aux {  // per DD, per layer step: what m2_finish hands to the merge kernels and to m2_compact
    unsigned long long mkey;  // max over the merged-away nodes of (value_top + rank, best candidate)
    int32_t cut_relaxed, nkeep, U, mpos, rank_m, pad0, pad1, pad2;
};

struct M2EV {
    int K, Wcap, C, T, Lmax, n, NW, NW4, PW, smem_keys;
    // model (device resident, immutable)
    const int32_t* ord;                                   // vars_by_sum_of_clause_weights, model.rs:149-151
    const int32_t *PT, *QT, *PF, *QF;                     // [n][NW] clause-weight rows of the branching variable (see m2s_engine.cu)
    const int32_t *AT, *AF;                               // [n] state-independent part of the transition cost
    const long long *est, *nk; long long initial;         // fast_upper_bound tables, model.rs:183-249
    const uint32_t* hmul;                                 // [NW] odd 32-bit multipliers of the multilinear state hash (64-bit accumulator)
    DDCtl* ctl; M2Aux* aux; int* active; int* lel_any; int* tile_off_e; int* tile_off_c; unsigned int* finish_counter;
    int32_t* root_state; int32_t* root_val; int32_t* root_depth; int32_t* root_width;
    uint32_t* cur_src[2]; int32_t* cur_val[2]; uint8_t* cur_flag[2]; int32_t* cur_rank[2]; int32_t* cur_rub;  // cur_src: candidate row holding the node's state
    // cand_state[b]: [K][C + 1][NW] state rows of the candidates of the layers with parity b (+ the merged node in row C)
    int32_t* cand_state[2]; uint32_t* cand_rep; uint32_t* cand_first; unsigned long long* cand_agg; uint8_t* cand_inex; uint32_t* cand_rank;
    uint32_t* cand_slot; int32_t* cand_cost;
    uint8_t* uflag; uint32_t* ulist; uint8_t* ustat; uint32_t* pos_of; unsigned long long* gkeys;
    unsigned long long* table;
    int32_t *mrg_min, *mrg_max;                           // [K][NW]; mrg_min holds the merged row after m2_merge_fin
    uint32_t* plog; uint32_t* clog; int32_t* colog; int32_t* nlog; int32_t* vlog; int32_t* rslog;  // rslog: [K][Lmax][3] (saved pos, recycled pos, cost delta)
    int32_t* lel_state; int32_t* lel_val; int32_t* lel_rub;
    int32_t* vb[2]; int32_t* cs_ub; uint8_t* cs_marked;
    uint64_t* best_path; uint64_t* best_exact_path;
    uint32_t* fc_node; int32_t* fc_ub; int32_t* fc_aux; unsigned long long fc_cap;  // FRONTIER cutset records, as in EV (m2s_frontier.cuh)
};


struct M2Engine : Engine {
    const M2Model* m2 = nullptr;
    M2EV mv{};
    bool expand_attr_set = false;
    int create_m2s(const M2Model* m, int device, uint64_t max_width_cap, int batch_cap, int cutset_type);
    int reserve_roots(int count) override;
    int compile_staged(int count, int comp_type, int64_t best_lb, const volatile int32_t* cutoff_flag, float* device_ms) override;
    int drain_all(int count, const int64_t* ub_cap, const int64_t* lb_filter, int* pw) override;
    void fc_launch_count(int count) override;
    void fc_launch_write(int pw) override;
};

int model_create_max2sat(int32_t n, int64_t m, const int64_t* clauses, int device, M2Model** out);
void model_destroy(M2Model*);

}  // namespace ddo

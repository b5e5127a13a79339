// capi.cu -- the extern "C" boundary declared in include/ddo_b200.h.
#include <algorithm>
#include <cstring>
#include <new>

#include "m2s_engine.hpp"
#include "solver.hpp"

using namespace ddo;

struct ddo_model { int kind; MispModel* m; M2Model* m2; };
struct ddo_mdd { int kind; Engine* ep; };
struct ddo_solver { Solver* s; };

#define GUARD_BEGIN try {
#define GUARD_END                                                                      \
    } catch (const std::bad_alloc&) { set_error("out of host memory"); return DDO_ERR_INVALID; } \
    catch (const std::exception& ex) { set_error(ex.what()); return DDO_ERR_INVALID; }

extern "C" {

const char* ddo_last_error(void) { return g_last_error.c_str(); }
int ddo_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
uint64_t ddo_kernel_launches(void) { return g_kernel_launches; }

int ddo_model_create_misp(int32_t n, const int64_t* weights, int64_t m, const int32_t* src, const int32_t* dst, int device, ddo_model** out) {
    GUARD_BEGIN
    MispModel* M = nullptr;
    int rc = model_create_misp(n, weights, m, src, dst, device, &M);
    if (rc != DDO_OK) return rc;
    *out = new ddo_model{DDO_MODEL_MISP, M, nullptr};
    return DDO_OK;
    GUARD_END
}
int ddo_model_create_max2sat(int32_t n, int64_t m, const int64_t* clauses, int device, ddo_model** out) {
    GUARD_BEGIN
    M2Model* M = nullptr;
    int rc = model_create_max2sat(n, m, clauses, device, &M);
    if (rc != DDO_OK) return rc;
    *out = new ddo_model{DDO_MODEL_MAX2SAT, nullptr, M};
    return DDO_OK;
    GUARD_END
}
int32_t ddo_model_kind(const ddo_model* m) { return m ? m->kind : -1; }
void ddo_model_destroy(ddo_model* m) { if (m) { if (m->m) model_destroy(m->m); if (m->m2) model_destroy(m->m2); delete m; } }
int32_t ddo_model_nb_variables(const ddo_model* m) { return !m ? 0 : (m->kind == DDO_MODEL_MISP ? m->m->n : m->m2->n); }
int32_t ddo_model_state_words(const ddo_model* m) { return !m ? 0 : (m->kind == DDO_MODEL_MISP ? m->m->words : m->m2->words); }
int ddo_model_initial_state(const ddo_model* m, uint64_t* state_out, int64_t* value_out) {
    if (!m || !state_out) { set_error("null argument"); return DDO_ERR_INVALID; }
    if (m->kind == DDO_MODEL_MAX2SAT) {  // model.rs:259-269: all benefits zero, value = weight of the tautologies
        for (int j = 0; j < m->m2->words; ++j) state_out[j] = 0;
        if (value_out) *value_out = m->m2->initial;
        return DDO_OK;
    }
    for (int j = 0; j < m->m->words; ++j) state_out[j] = 0;
    for (int i = 0; i < m->m->n; ++i) state_out[i >> 6] |= 1ull << (i & 63);
    if (value_out) *value_out = 0;
    return DDO_OK;
}

int ddo_mdd_create(const ddo_model* m, int device, uint64_t max_width_cap, int32_t batch_cap, int32_t cutset_type, ddo_mdd** out) {
    GUARD_BEGIN
    if (!m || !out) { set_error("null argument"); return DDO_ERR_INVALID; }
    ddo_mdd* d = new ddo_mdd{m->kind, nullptr};
    int rc;
    if (m->kind == DDO_MODEL_MAX2SAT) { M2Engine* me = new M2Engine(); d->ep = me; rc = me->create_m2s(m->m2, device, max_width_cap, batch_cap, cutset_type); }
    else { d->ep = new Engine(); rc = d->ep->create(m->m, device, max_width_cap, batch_cap, cutset_type); }
    if (rc != DDO_OK) { d->ep->destroy(); delete d->ep; delete d; return rc; }
    *out = d;
    return DDO_OK;
    GUARD_END
}
void ddo_mdd_destroy(ddo_mdd* d) { if (d) { d->ep->destroy(); delete d->ep; delete d; } }

int ddo_mdd_compile_batch(ddo_mdd* d, int32_t count, int32_t comp_type, const uint64_t* max_widths, const uint64_t* root_states,
                          const int64_t* root_values, const int32_t* root_depths, int64_t best_lb, const volatile int32_t* cutoff_flag,
                          ddo_completion* out) {
    GUARD_BEGIN
    if (!d || !max_widths || !root_states || !root_values || !root_depths) { set_error("null argument"); return DDO_ERR_INVALID; }
    int rc = d->ep->stage_roots(count, max_widths, root_states, root_values, root_depths);
    if (rc != DDO_OK) return rc;
    rc = d->ep->compile_staged(count, comp_type, best_lb, cutoff_flag, nullptr);
    if (rc != DDO_OK) return rc;
    rc = d->ep->fetch_ctl(count);
    if (rc != DDO_OK) return rc;
    if (out) for (int i = 0; i < count; ++i) d->ep->fill_completion(i, out + i);
    return DDO_OK;
    GUARD_END
}
int ddo_mdd_compile(ddo_mdd* d, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value, int32_t root_depth,
                    int64_t best_lb, const volatile int32_t* cutoff_flag, ddo_completion* out) {
    return ddo_mdd_compile_batch(d, 1, comp_type, &max_width, root_state, &root_value, &root_depth, best_lb, cutoff_flag, out);
}
int ddo_mdd_best_solution(ddo_mdd* d, int32_t index, int32_t exact, ddo_decision* out, int32_t* len) {
    GUARD_BEGIN
    if (!d) { set_error("null argument"); return DDO_ERR_INVALID; }
    return d->ep->best_solution(index, exact, out, len);
    GUARD_END
}
int ddo_mdd_drain_cutset(ddo_mdd* d, int32_t index, int64_t ub_cap, int64_t lb_filter, uint64_t* states, int64_t* values, int64_t* ubs,
                         int32_t* depth_out, int32_t* path_len_out, ddo_decision* paths, int32_t* count) {
    GUARD_BEGIN
    if (!d) { set_error("null argument"); return DDO_ERR_INVALID; }
    return d->ep->drain_cutset(index, ub_cap, lb_filter, states, values, ubs, depth_out, path_len_out, paths, count);
    GUARD_END
}
int ddo_mdd_drain_cutset_batch(ddo_mdd* d, int32_t count, const int64_t* ub_caps, const int64_t* lb_filters, uint64_t* states, int64_t* values,
                               int64_t* ubs, int32_t* dd_index, uint64_t* path_bits, int32_t* path_words, int64_t* total) {
    GUARD_BEGIN
    if (!d || !ub_caps || !lb_filters || !total) { set_error("null argument"); return DDO_ERR_INVALID; }
    Engine& e = *d->ep;
    int pw = 1;
    const int n = e.drain_all(count, ub_caps, lb_filters, &pw);
    if (n < 0) return n;
    e.last_drain_total = n;
    if (path_words) *path_words = pw;
    if ((int64_t)n > *total) { *total = n; set_error("drain_cutset_batch: buffer too small"); return DDO_ERR_CAPACITY; }
    const int words = e.abi_words;
    for (int r = 0; r < n; ++r) {
        if (states) for (int j = 0; j < words; ++j) states[(size_t)r * words + j] = e.h_out_state[(size_t)r * e.S + j];
        if (values) values[r] = e.h_out_val[r];
        if (ubs) ubs[r] = e.h_out_ub[r];
        if (dd_index) dd_index[r] = e.h_out_dd[r];
    }
    if (path_bits && n > 0) std::memcpy(path_bits, e.h_out_path, (size_t)n * pw * 8);
    *total = n;
    return DDO_OK;
    GUARD_END
}
int ddo_mdd_drain_layer_index(ddo_mdd* d, int32_t* layer_index, int64_t cap) {
    GUARD_BEGIN
    if (!d || !layer_index) { set_error("null argument"); return DDO_ERR_INVALID; }
    Engine& e = *d->ep;
    const int n = e.last_drain_total;
    if ((int64_t)n > cap) { set_error("drain_layer_index: buffer too small"); return DDO_ERR_CAPACITY; }
    for (int r = 0; r < n; ++r) layer_index[r] = e.cutset_type == DDO_FRONTIER ? e.h_out_tt[r] : e.h_ctl[e.h_out_dd[r]].lel;
    return n;
    GUARD_END
}
int ddo_mdd_set_profiling(ddo_mdd* d, int32_t on) {
    if (!d) return DDO_ERR_INVALID;
    d->ep->profiling = on != 0; d->ep->prof_used = 0;
    for (int i = 0; i < 6; ++i) { d->ep->prof_ms[i] = 0; d->ep->prof_launches[i] = 0; }
    return DDO_OK;
}
int ddo_mdd_kernel_times(ddo_mdd* d, double ms[6], uint64_t launches[6]) {
    if (!d || !ms || !launches) return DDO_ERR_INVALID;
    for (int i = 0; i < 6; ++i) { ms[i] = d->ep->prof_ms[i]; launches[i] = d->ep->prof_launches[i]; }
    return DDO_OK;
}
int ddo_mdd_layer_trace(ddo_mdd* d, int32_t index, int32_t* vars, int32_t* widths, int32_t cap) {
    GUARD_BEGIN
    if (!d) { set_error("null argument"); return DDO_ERR_INVALID; }
    return d->ep->layer_trace(index, vars, widths, cap);
    GUARD_END
}
int ddo_mdd_stage_roots(ddo_mdd* d, int32_t count, const uint64_t* max_widths, const uint64_t* root_states, const int64_t* root_values,
                        const int32_t* root_depths) {
    GUARD_BEGIN
    if (!d) { set_error("null argument"); return DDO_ERR_INVALID; }
    int rc = d->ep->stage_roots(count, max_widths, root_states, root_values, root_depths);
    if (rc != DDO_OK) return rc;
    if (cudaStreamSynchronize(d->ep->stream) != cudaSuccess) { set_error("stage_roots: sync failed"); return DDO_ERR_CUDA; }
    return DDO_OK;
    GUARD_END
}
int ddo_mdd_compile_staged(ddo_mdd* d, int32_t count, int32_t comp_type, int64_t best_lb, float* device_ms) {
    GUARD_BEGIN
    if (!d) { set_error("null argument"); return DDO_ERR_INVALID; }
    return d->ep->compile_staged(count, comp_type, best_lb, nullptr, device_ms);
    GUARD_END
}
int ddo_mdd_fetch_completions(ddo_mdd* d, int32_t count, ddo_completion* out) {
    GUARD_BEGIN
    if (!d || !out) { set_error("null argument"); return DDO_ERR_INVALID; }
    int rc = d->ep->fetch_ctl(count);
    if (rc != DDO_OK) return rc;
    for (int i = 0; i < count; ++i) d->ep->fill_completion(i, out + i);
    return DDO_OK;
    GUARD_END
}

int ddo_solver_create(const ddo_model* m, ddo_mdd* d, int32_t width_kind, uint64_t width, int32_t wave_size, ddo_solver** out) {
    GUARD_BEGIN
    if (!m || !d || !out) { set_error("null argument"); return DDO_ERR_INVALID; }
    if (wave_size < 1) { set_error("wave_size must be >= 1"); return DDO_ERR_INVALID; }
    { int rr = d->ep->reserve_roots(wave_size); if (rr != DDO_OK) return rr; }  // the wave may exceed batch_cap: only DDs that need a cut go through the general engine, batch_cap at a time
    if (width_kind == DDO_WIDTH_FIXED && (width < 1 || width > (uint64_t)d->ep->Wcap)) { set_error("width must be in [1, max_width_cap]"); return DDO_ERR_INVALID; }
    if (m->kind != d->kind) { set_error("model and mdd are of different kinds"); return DDO_ERR_INVALID; }
    if (width_kind < DDO_WIDTH_FIXED || width_kind > DDO_WIDTH_DIVBY_NB_UNASSIGNED) { set_error("unknown width heuristic"); return DDO_ERR_INVALID; }
    if ((width_kind == DDO_WIDTH_NB_UNASSIGNED || width_kind == DDO_WIDTH_DIVBY_NB_UNASSIGNED) && ddo_model_nb_variables(m) > d->ep->Wcap) { set_error("NbUnassignedWidth needs max_width_cap >= nb_variables"); return DDO_ERR_INVALID; }
    if (width_kind == DDO_WIDTH_TIMES_NB_UNASSIGNED && (width < 1 || width * (uint64_t)ddo_model_nb_variables(m) > (uint64_t)d->ep->Wcap)) { set_error("Times(k, NbUnassignedWidth) needs max_width_cap >= k * nb_variables"); return DDO_ERR_INVALID; }
    if (width_kind == DDO_WIDTH_DIVBY_NB_UNASSIGNED && width < 1) { set_error("DivBy(k, ...) needs k >= 1"); return DDO_ERR_INVALID; }
    std::vector<uint64_t> rs((size_t)ddo_model_state_words(m));
    int64_t rv = 0;
    ddo_model_initial_state(m, rs.data(), &rv);
    *out = new ddo_solver{new Solver(d->ep, m->kind, rs.data(), rv, width_kind, width, wave_size)};
    return DDO_OK;
    GUARD_END
}
void ddo_solver_destroy(ddo_solver* s) { if (s) { delete s->s; delete s; } }
int ddo_solver_maximize(ddo_solver* s, double time_budget_s, uint64_t max_waves, int32_t* is_exact, int32_t* has_value, int64_t* best_value) {
    GUARD_BEGIN
    if (!s) { set_error("null argument"); return DDO_ERR_INVALID; }
    return s->s->maximize(time_budget_s, max_waves, is_exact, has_value, best_value);
    GUARD_END
}
int ddo_solver_init(ddo_solver* s, int32_t push_root) { GUARD_BEGIN if (!s) return DDO_ERR_INVALID; return s->s->init(push_root != 0); GUARD_END }
int ddo_solver_wave(ddo_solver* s, const volatile int32_t* cutoff_flag, int64_t out3[3]) {
    GUARD_BEGIN
    if (!s || !out3) { set_error("null argument"); return DDO_ERR_INVALID; }
    return s->s->wave(cutoff_flag, out3);
    GUARD_END
}
int ddo_solver_set_lower_bound(ddo_solver* s, int64_t lb) { if (!s) return DDO_ERR_INVALID; if (lb > s->s->best_lb) s->s->best_lb = lb; return DDO_OK; }
int ddo_solver_retain_share(ddo_solver* s, int32_t rank, int32_t nranks) { GUARD_BEGIN if (!s) return DDO_ERR_INVALID; return s->s->retain_share(rank, nranks); GUARD_END }
int32_t ddo_solver_node_words(const ddo_solver* s) { return s ? s->s->node_words() : 0; }
int ddo_solver_export_open(ddo_solver* s, int32_t max_nodes, int64_t* rows, int32_t* count) {
    GUARD_BEGIN
    if (!s) { set_error("null argument"); return DDO_ERR_INVALID; }
    return s->s->export_open(max_nodes, rows, count);
    GUARD_END
}
int ddo_solver_import_open(ddo_solver* s, int32_t count, const int64_t* rows) {
    GUARD_BEGIN
    if (!s) { set_error("null argument"); return DDO_ERR_INVALID; }
    return s->s->import_open(count, rows);
    GUARD_END
}
int ddo_solver_maximize_sharded(ddo_solver* s, ddo_comm* comm, double time_budget_s, uint64_t max_waves, int32_t rebalance, int64_t out[8]) {
    GUARD_BEGIN
    if (!s || !comm || !out) { set_error("null argument"); return DDO_ERR_INVALID; }
    ShardComm cm;
    cm.ctx = comm; cm.nranks = ddo_comm_size(comm); cm.rank = ddo_comm_rank(comm);
    cm.allgather = [](void* c, const int64_t* v, int32_t n, int64_t* r) { return ddo_comm_allgather((ddo_comm*)c, v, n, r); };
    cm.send = [](void* c, const void* b, int64_t n, int32_t p) { return ddo_comm_send((ddo_comm*)c, b, n, p); };
    cm.recv = [](void* c, void* b, int64_t n, int32_t p) { return ddo_comm_recv((ddo_comm*)c, b, n, p); };
    return s->s->maximize_sharded(cm, time_budget_s, max_waves, rebalance != 0, out);
    GUARD_END
}
int ddo_solver_finish(ddo_solver* s) {
    if (!s) return DDO_ERR_INVALID;
    s->s->finish();
    std::stable_sort(s->s->best_sol.begin(), s->s->best_sol.end(), [](const ddo_decision& a, const ddo_decision& b) { return a.variable < b.variable; });
    return DDO_OK;
}
int64_t ddo_solver_best_lower_bound(const ddo_solver* s) { return s->s->best_lb; }
int64_t ddo_solver_best_upper_bound(const ddo_solver* s) { return s->s->best_ub; }
int ddo_solver_best_value(const ddo_solver* s, int32_t* has, int64_t* value) {
    if (!s) return DDO_ERR_INVALID;
    if (has) *has = s->s->has_sol;
    if (value) *value = s->s->has_sol ? s->s->sol_value : 0;  // the objective of the solution THIS solver holds (best_lb may come from another rank)
    return DDO_OK;
}
int ddo_solver_best_solution(const ddo_solver* s, ddo_decision* out, int32_t* len) {
    if (!s || !len) return DDO_ERR_INVALID;
    if (!s->s->has_sol) { set_error("no solution"); return DDO_ERR_INVALID; }
    const int n = (int)s->s->best_sol.size();
    if (*len < n) { *len = n; return DDO_ERR_CAPACITY; }
    for (int i = 0; i < n; ++i) out[i] = s->s->best_sol[i];
    *len = n;
    return DDO_OK;
}
uint64_t ddo_solver_explored(const ddo_solver* s) { return s->s->explored; }
uint64_t ddo_solver_fringe_len(const ddo_solver* s) { return s->s->open_len(); }
int ddo_solver_stats(const ddo_solver* s, double stats[8]) {
    if (!s || !stats) return DDO_ERR_INVALID;
    stats[0] = (double)s->s->expanded; stats[1] = (double)s->s->transitions; stats[2] = (double)s->s->compilations; stats[3] = (double)s->s->waves;
    stats[4] = s->s->device_ms; stats[5] = s->s->fringe_ms;
    stats[6] = (double)s->s->eng->bytes_h2d; stats[7] = (double)s->s->eng->bytes_d2h;
    return DDO_OK;
}

}  // extern "C"

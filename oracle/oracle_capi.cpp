// ============================================================================
// oracle_capi.cpp -- TEST INFRASTRUCTURE.  C entry points (ctypes) over the CPU
// oracle.  Loaded only by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs.
// ============================================================================
#include "models.hpp"
#include "tsptw.hpp"
#include <cstdio>
#include <sstream>

using namespace ddo_oracle;

namespace {
// ---- model adapters: everything the generic entry points need to know about a model -------------------------------------------
struct MispHandle {
    using State = BitState; using Hash = BitStateHash; using Eq = BitStateEq;
    Misp pb;
    MispRelax rlx;
    MispRanking rk;
    MispHandle(size_t n, const isize* w, size_t m, const int32_t* s, const int32_t* d) : pb(n, w, m, s, d), rlx(&pb) {}
    size_t abi_words() const { return pb.words; }  // uint64 words of a packed state at the ABI
    static isize other_decision() { return 0; }    // NO (the decision a 0 path bit stands for)
    State state_from_abi(const uint64_t* p, size_t /*depth*/) const { BitState s; s.w.assign(p, p + pb.words); return s; }
    void state_to_abi(const State& s, uint64_t* p) const { std::memcpy(p, s.w.data(), pb.words * 8); }
};
// MAX2SAT states cross the ABI as n int32 benefits packed two per uint64 word (little endian), the depth travels separately
struct M2Handle {
    using State = M2State; using Hash = M2Hash; using Eq = M2Eq;
    Max2Sat pb;
    Max2SatRelax rlx;
    Max2SatRanking rk;
    M2Handle(size_t n, const std::vector<M2Clause>& c) : pb(n, c), rlx(&pb) {}
    size_t abi_words() const { return (pb.nb_vars + 1) / 2; }
    static isize other_decision() { return -1; }   // F
    State state_from_abi(const uint64_t* p, size_t depth) const {
        const int32_t* q = (const int32_t*)p;
        M2State s{depth, std::vector<isize>(pb.nb_vars)};
        for (size_t i = 0; i < pb.nb_vars; ++i) s.sub[i] = q[i];
        return s;
    }
    void state_to_abi(const State& s, uint64_t* p) const {
        std::memset(p, 0, abi_words() * 8);
        int32_t* q = (int32_t*)p;
        for (size_t i = 0; i < pb.nb_vars; ++i) q[i] = (int32_t)s.sub[i];
    }
};
// TSPTW states cross the ABI as 16 uint64 words (the depth travels separately, like MAX2SAT):
//   [0] position: bit 63 set = Position::Virtual, else the node id in the low 16 bits      [1..4]  the Virtual pool (Set256)
//   [5] earliest (= duration of a FixedAmount)   [6] latest (= earliest when FixedAmount)   [7] flags: bit 0 FuzzyAmount, bit 1 maybe_visit is Some
//   [8..11] must_visit (Set256)                  [12..15] maybe_visit (Set256, zero when None)
struct TsptwHandle {
    using State = TsptwState; using Hash = TsptwHash; using Eq = TsptwEq;
    Tsptw pb;
    TsptwRelax rlx;
    TsptwRanking rk;
    explicit TsptwHandle(TsptwInstance inst) : pb(std::move(inst)), rlx(&pb) {}
    size_t abi_words() const { return 16; }
    State state_from_abi(const uint64_t* p, size_t depth) const {
        TsptwState s;
        s.virtual_pos = (p[0] >> 63) & 1; s.node = (uint16_t)(p[0] & 0xFFFF);
        for (int j = 0; j < 4; ++j) { s.pool.w[j] = s.virtual_pos ? p[1 + j] : 0; s.must_visit.w[j] = p[8 + j]; }
        s.earliest = (size_t)p[5]; s.fuzzy = p[7] & 1; s.latest = s.fuzzy ? (size_t)p[6] : s.earliest;
        s.has_maybe = (p[7] >> 1) & 1;
        for (int j = 0; j < 4; ++j) s.maybe_visit.w[j] = s.has_maybe ? p[12 + j] : 0;
        s.depth = (uint16_t)depth;
        return s;
    }
    void state_to_abi(const State& s, uint64_t* p) const {
        std::memset(p, 0, 16 * 8);
        p[0] = s.virtual_pos ? (1ull << 63) : (uint64_t)s.node;
        for (int j = 0; j < 4; ++j) { if (s.virtual_pos) p[1 + j] = s.pool.w[j]; p[8 + j] = s.must_visit.w[j]; if (s.has_maybe) p[12 + j] = s.maybe_visit.w[j]; }
        p[5] = s.earliest; p[6] = s.fuzzy ? s.latest : s.earliest; p[7] = (s.fuzzy ? 1u : 0u) | (s.has_maybe ? 2u : 0u);
    }
};
template <class H>
struct DD {
    using S = typename H::State;
    H* h;
    Mdd<S, typename H::Hash, typename H::Eq> mdd;
    std::vector<SubProblem<S>> cutset;  // drained copy
    SubProblem<S> root;
    DD(H* h_, int cutset_type) : h(h_), mdd(cutset_type) {}
};
template <class H>
struct Stepper {
    using S = typename H::State;
    H* h;
    FixedWidth<S> fw; NbUnassignedWidth<S> nw; NoCutoff nocut; EmptyDominanceChecker<S> dom; EmptyCache<S> cache;
    MaxUB<S> mx; NoDupFringe<S, typename H::Hash, typename H::Eq> fringe;
    std::unique_ptr<WaveSolver<S, typename H::Hash, typename H::Eq>> solver;
    Stepper(H* h_, int k, int width_kind, uint64_t width)
        : h(h_), fw((size_t)width), nw(h_->pb.nb_variables()), mx{&h_->rk}, fringe(mx) {
        const WidthHeuristic<S>* wh = width_kind == 0 ? (const WidthHeuristic<S>*)&fw : (const WidthHeuristic<S>*)&nw;
        SolverConfig<S> cfg{&h->pb, &h->rlx, &h->rk, wh, &dom, &nocut, &fringe, &cache, LAST_EXACT_LAYER};
        solver.reset(new WaveSolver<S, typename H::Hash, typename H::Eq>(cfg, (size_t)k));
    }
};
using MispDD = DD<MispHandle>;
using MispStepper = Stepper<MispHandle>;
using MispMdd = Mdd<BitState, BitStateHash, BitStateEq>;
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

extern "C" {
struct oracle_dd_result {
    int32_t has_best; int32_t is_exact; int32_t has_best_exact; int32_t lel;  // lel: -1 = none (never squashed)
    int64_t best_value; int64_t best_exact_value;
    uint64_t expanded; uint64_t transitions;
    int32_t n_layers; int32_t cutset_size; int32_t cutoff; int32_t pad;
};
struct oracle_solve_result {
    int32_t has_value; int32_t is_exact;
    int64_t best_value; int64_t best_lb; int64_t best_ub;
    uint64_t explored; uint64_t expanded; uint64_t transitions; uint64_t compilations; uint64_t waves;
    double seconds;
};
}

namespace {
// comp_type: 0 Exact, 1 Relaxed, 2 Restricted (abstraction/mdd.rs:40-47 order).  Returns 0 ok, 1 cutoff.
template <class H>
int32_t dd_compile(DD<H>* dd, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value, uint64_t root_depth,
                   int64_t best_lb, int32_t cutoff_now, oracle_dd_result* out) {
    using S = typename H::State;
    H* h = dd->h;
    dd->root = SubProblem<S>{std::make_shared<const S>(h->state_from_abi(root_state, (size_t)root_depth)), root_value, {}, ISIZE_MAX, (size_t)root_depth};
    EmptyCache<S> cache; EmptyDominanceChecker<S> dom; FlagCutoff cut(cutoff_now != 0);
    CompilationType t = comp_type == 0 ? CompilationType::Exact : (comp_type == 1 ? CompilationType::Relaxed : CompilationType::Restricted);
    CompilationInput<S> in{t, &h->pb, &h->rlx, &h->rk, &cut, (size_t)max_width, &dd->root, best_lb, &cache, &dom};
    Completion c;
    std::memset(out, 0, sizeof(*out));
    if (!dd->mdd.compile(in, &c)) { out->cutoff = 1; return 1; }
    out->has_best = dd->mdd.best_value().has_value();
    out->best_value = dd->mdd.best_value().value_or(0);
    out->is_exact = dd->mdd.is_exact();
    out->has_best_exact = dd->mdd.best_exact_value().has_value();
    out->best_exact_value = dd->mdd.best_exact_value().value_or(0);
    out->expanded = dd->mdd.expanded; out->transitions = dd->mdd.transitions;
    out->n_layers = (int32_t)dd->mdd.layers.size();
    auto lel = dd->mdd.lel();
    out->lel = (lel && *lel < dd->mdd.layers.size()) ? (int32_t)*lel : -1;
    dd->cutset.clear();
    dd->mdd.drain_cutset([&](SubProblem<S> n) { dd->cutset.push_back(std::move(n)); });
    out->cutset_size = (int32_t)dd->cutset.size();
    return 0;
}
// per expanded layer: branching variable and |curr_l| after the cut
template <class H>
int32_t dd_layers(DD<H>* dd, int32_t* vars, int32_t* widths, int32_t cap) {
    int32_t n = (int32_t)dd->mdd.layer_vars.size();
    for (int32_t i = 0; i < n && i < cap; ++i) { vars[i] = (int32_t)dd->mdd.layer_vars[i]; widths[i] = (int32_t)dd->mdd.layer_widths[i]; }
    return n;
}
// drained cutset (MARKED nodes only, clean.rs:417-445) in drain order.  paths: path_stride (var,value) int32 pairs per node.
template <class H>
int32_t dd_cutset(DD<H>* dd, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depths, int32_t* path_lens, int32_t* paths, int32_t cap,
                  int32_t path_stride) {
    size_t W = dd->h->abi_words();
    int32_t n = (int32_t)dd->cutset.size();
    for (int32_t i = 0; i < n && i < cap; ++i) {
        const auto& sp = dd->cutset[i];
        dd->h->state_to_abi(*sp.state, states + (size_t)i * W);
        values[i] = sp.value; ubs[i] = sp.ub; depths[i] = (int32_t)sp.depth;
        path_lens[i] = (int32_t)sp.path.size();
        if (paths)
            for (int32_t j = 0; j < (int32_t)sp.path.size() && j < path_stride; ++j) {
                paths[((size_t)i * path_stride + j) * 2] = (int32_t)sp.path[j].variable;
                paths[((size_t)i * path_stride + j) * 2 + 1] = (int32_t)sp.path[j].value;
            }
    }
    return n;
}
// best (exact != 0: best exact) solution as (var,value) pairs in path order; returns length or -1 if None
template <class H>
int32_t dd_solution(DD<H>* dd, int32_t exact, int32_t* vars, int32_t* vals, int32_t cap) {
    auto sol = exact ? dd->mdd.best_exact_solution() : dd->mdd.best_solution();
    if (!sol) return -1;
    for (int32_t i = 0; i < (int32_t)sol->size() && i < cap; ++i) { vars[i] = (int32_t)(*sol)[i].variable; vals[i] = (int32_t)(*sol)[i].value; }
    return (int32_t)sol->size();
}
// mode: 0 SequentialSolver, 1 WaveSolver(k), 2 ParallelSolver(k threads).  width_kind: 0 FixedWidth(width), 1 NbUnassignedWidth.
// sol_vars / sol_vals: the best solution sorted by variable (cap nb_variables).  trace (mode 1): 4 int64 per wave (popped, best_lb, fringe_len, top_ub).
template <class H>
int32_t solve(H* h, int32_t mode, int32_t k, int32_t width_kind, uint64_t width, int32_t cutset_type, double time_budget_s, uint64_t max_waves,
              oracle_solve_result* out, int32_t* sol_vars, int32_t* sol_vals, int32_t* sol_len, int64_t* trace, int32_t trace_cap, int32_t* trace_len) {
    using S = typename H::State; using HS = typename H::Hash; using EQ = typename H::Eq;
    // width_kind: 0 FixedWidth(width), 1 NbUnassignedWidth, 2 Times(width, NbUnassignedWidth), 3 DivBy(width, NbUnassignedWidth)  (width.rs)
    FixedWidth<S> fw((size_t)width); NbUnassignedWidth<S> nw(h->pb.nb_variables());
    Times<S> tw((size_t)width, &nw); DivBy<S> dw((size_t)std::max<uint64_t>(width, 1), &nw);
    const WidthHeuristic<S>* wh = width_kind == 0 ? (const WidthHeuristic<S>*)&fw : width_kind == 2 ? (const WidthHeuristic<S>*)&tw
                                  : width_kind == 3 ? (const WidthHeuristic<S>*)&dw : (const WidthHeuristic<S>*)&nw;
    NoCutoff nocut; std::unique_ptr<TimeBudget> tb;
    const Cutoff* cut = &nocut;
    if (time_budget_s > 0) { tb.reset(new TimeBudget(time_budget_s)); cut = tb.get(); }
    EmptyDominanceChecker<S> dom; EmptyCache<S> cache;
    MaxUB<S> mx{&h->rk};
    NoDupFringe<S, HS, EQ> fringe(mx);
    SolverConfig<S> cfg{&h->pb, &h->rlx, &h->rk, wh, &dom, cut, &fringe, &cache, cutset_type};
    std::memset(out, 0, sizeof(*out));
    if (trace_len) *trace_len = 0;
    double t0 = now_s();
    Completion c; SolverStats st; isize lb, ub; std::optional<Solution> sol;
    if (mode == 0) { SequentialSolver<S, HS, EQ> s(cfg); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    else if (mode == 1) {
        WaveSolver<S, HS, EQ> s(cfg, (size_t)k); s.max_waves = max_waves ? max_waves : UINT64_MAX;
        c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol;
        if (trace) {
            int32_t n = 0;
            for (auto& t : s.trace) { if (n >= trace_cap) break; trace[4 * n] = (int64_t)t.popped; trace[4 * n + 1] = t.best_lb; trace[4 * n + 2] = (int64_t)t.fringe_len; trace[4 * n + 3] = t.top_ub; ++n; }
            *trace_len = n;
        }
    } else { ParallelSolver<S, HS, EQ> s(cfg, (size_t)k); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    out->seconds = now_s() - t0;
    out->has_value = c.best_value.has_value(); out->best_value = c.best_value.value_or(0); out->is_exact = c.is_exact;
    out->best_lb = lb; out->best_ub = ub;
    out->explored = st.explored; out->expanded = st.expanded; out->transitions = st.transitions; out->compilations = st.compilations; out->waves = st.waves;
    int32_t n = 0;
    if (sol && sol_vars && sol_vals) for (auto& d : *sol) { sol_vars[n] = (int32_t)d.variable; sol_vals[n] = (int32_t)d.value; ++n; }
    if (sol_len) *sol_len = n;
    return 0;
}
template <class H> void stepper_state(Stepper<H>* s, int64_t out[6]) {
    auto& w = *s->solver;
    out[0] = w.best_lb; out[1] = w.best_ub; out[2] = (int64_t)w.fringe_len(); out[3] = (int64_t)w.stats.explored; out[4] = (int64_t)w.stats.expanded; out[5] = w.best_sol.has_value();
}
// packed open nodes, the layout of ddo_solver_export_open: int64 words [value, ub, depth, state words ..., decisions four per word, 16 bits
// each = variable | path bit << 15]; path bit 1 = the model's "positive" decision (MISP YES = 1, MAX2SAT T = +1), 0 = the other one
template <class H> size_t stepper_node_words(Stepper<H>* s) { return 3 + s->h->abi_words() + (s->h->pb.nb_variables() + 3) / 4; }
template <class H> int32_t stepper_export(Stepper<H>* s, int32_t max_nodes, int64_t* rows) {
    const size_t W = s->h->abi_words(), RW = stepper_node_words(s);
    auto nodes = s->solver->export_open((size_t)max_nodes);
    for (size_t i = 0; i < nodes.size(); ++i) {
        int64_t* r = rows + i * RW;
        std::memset(r, 0, RW * 8);
        r[0] = nodes[i].value; r[1] = nodes[i].ub; r[2] = (int64_t)nodes[i].depth;
        s->h->state_to_abi(*nodes[i].state, (uint64_t*)(r + 3));
        uint16_t* d16 = (uint16_t*)(r + 3 + W);
        for (size_t j = 0; j < nodes[i].path.size(); ++j) d16[j] = (uint16_t)(nodes[i].path[j].variable | (nodes[i].path[j].value == 1 ? 0x8000 : 0));
    }
    return (int32_t)nodes.size();
}
template <class H> void stepper_import(Stepper<H>* s, int32_t count, const int64_t* rows) {
    using S = typename H::State;
    const size_t W = s->h->abi_words(), RW = stepper_node_words(s);
    const isize other = s->h->other_decision();
    std::vector<SubProblem<S>> nodes;
    for (int32_t i = 0; i < count; ++i) {
        const int64_t* r = rows + (size_t)i * RW;
        SubProblem<S> sp;
        sp.value = r[0]; sp.ub = r[1]; sp.depth = (size_t)r[2];
        sp.state = std::make_shared<const S>(s->h->state_from_abi((const uint64_t*)(r + 3), sp.depth));
        const uint16_t* d16 = (const uint16_t*)(r + 3 + W);
        for (size_t j = 0; j < sp.depth; ++j) sp.path.push_back(Decision{(size_t)(d16[j] & 0x7FFF), (d16[j] & 0x8000) ? (isize)1 : other});
        nodes.push_back(std::move(sp));
    }
    s->solver->import_open(std::move(nodes));
}
template <class H> int32_t stepper_wave(Stepper<H>* s, int64_t out3[3]) { isize o[3]; bool ok = s->solver->wave(o); out3[0] = o[0]; out3[1] = o[1]; out3[2] = o[2]; return ok ? 0 : 1; }
}  // namespace

extern "C" {

// ---- MISP (BASELINE configs 2, 5) -------------------------------------------------------------------------------------------------
void* oracle_misp_new(int32_t n, const int64_t* weights, int32_t m, const int32_t* src, const int32_t* dst) {
    return new MispHandle((size_t)n, weights, (size_t)m, src, dst);
}
void oracle_misp_free(void* h) { delete (MispHandle*)h; }
int32_t oracle_misp_words(void* h) { return (int32_t)((MispHandle*)h)->pb.words; }
void* oracle_misp_dd_new(void* h, int32_t cutset_type) { return new MispDD((MispHandle*)h, cutset_type); }
void oracle_misp_dd_free(void* dd) { delete (MispDD*)dd; }
int32_t oracle_misp_dd_compile(void* ddp, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value,
                               uint64_t root_depth, int64_t best_lb, int32_t cutoff_now, oracle_dd_result* out) {
    return dd_compile((MispDD*)ddp, comp_type, max_width, root_state, root_value, root_depth, best_lb, cutoff_now, out);
}
int32_t oracle_misp_dd_layers(void* ddp, int32_t* vars, int32_t* widths, int32_t cap) { return dd_layers((MispDD*)ddp, vars, widths, cap); }
int32_t oracle_misp_dd_cutset(void* ddp, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depths, int32_t* path_lens,
                              int32_t* paths, int32_t cap, int32_t path_stride) {
    return dd_cutset((MispDD*)ddp, states, values, ubs, depths, path_lens, paths, cap, path_stride);
}
int32_t oracle_misp_dd_solution(void* ddp, int32_t exact, int32_t* vars, int32_t* vals, int32_t cap) { return dd_solution((MispDD*)ddp, exact, vars, vals, cap); }
// sol_yes: vertices with decision YES (cap n)
int32_t oracle_misp_solve(void* hp, int32_t mode, int32_t k, int32_t width_kind, uint64_t width, int32_t cutset_type, double time_budget_s,
                          uint64_t max_waves, oracle_solve_result* out, int32_t* sol_yes, int32_t* sol_len, int64_t* trace, int32_t trace_cap,
                          int32_t* trace_len) {
    MispHandle* h = (MispHandle*)hp;
    std::vector<int32_t> vars(h->pb.nb_vars + 1), vals(h->pb.nb_vars + 1);
    int32_t len = 0;
    int32_t rc = solve(h, mode, k, width_kind, width, cutset_type, time_budget_s, max_waves, out, vars.data(), vals.data(), &len, trace, trace_cap, trace_len);
    int32_t n = 0;
    if (sol_yes) for (int32_t i = 0; i < len; ++i) if (vals[i] == MISP_YES) sol_yes[n++] = vars[i];
    if (sol_len) *sol_len = n;
    return rc;
}

// ---- MAX2SAT (BASELINE config 3).  clauses: m triples (weight, literal x, literal y), literals +-(variable + 1), x == y for a unit clause ----
void* oracle_m2s_new(int32_t n, int32_t m, const int64_t* clauses) {
    std::vector<M2Clause> c((size_t)m);
    for (int32_t i = 0; i < m; ++i) c[i] = M2Clause{clauses[3 * i], clauses[3 * i + 1], clauses[3 * i + 2]};
    return new M2Handle((size_t)n, c);
}
void oracle_m2s_free(void* h) { delete (M2Handle*)h; }
int32_t oracle_m2s_words(void* h) { return (int32_t)((M2Handle*)h)->abi_words(); }
int64_t oracle_m2s_initial_value(void* h) { return ((M2Handle*)h)->pb.initial; }
void oracle_m2s_order(void* h, int32_t* out) { auto& o = ((M2Handle*)h)->pb.order; for (size_t i = 0; i < o.size(); ++i) out[i] = (int32_t)o[i]; }
// model-level known answers (model.rs:396-448): transition / cost / rank / rub of one state
void oracle_m2s_transition(void* hp, const int32_t* sub, int32_t depth, int32_t var, int32_t value, int32_t* out_sub, int64_t* cost, int64_t* rank, int64_t* rub) {
    M2Handle* h = (M2Handle*)hp;
    M2State s{(size_t)depth, std::vector<isize>(sub, sub + h->pb.nb_vars)};
    Decision d{(size_t)var, value};
    M2State r = h->pb.transition(s, d);
    for (size_t i = 0; i < h->pb.nb_vars; ++i) out_sub[i] = (int32_t)r.sub[i];
    if (cost) *cost = h->pb.transition_cost(s, r, d);
    if (rank) *rank = Max2SatRanking::rank(r);
    if (rub) *rub = (size_t)depth < h->pb.nb_vars ? h->pb.fast_upper_bound(s) : 0;
}
void* oracle_m2s_dd_new(void* h, int32_t cutset_type) { return new DD<M2Handle>((M2Handle*)h, cutset_type); }
void oracle_m2s_dd_free(void* dd) { delete (DD<M2Handle>*)dd; }
int32_t oracle_m2s_dd_compile(void* ddp, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value,
                              uint64_t root_depth, int64_t best_lb, int32_t cutoff_now, oracle_dd_result* out) {
    return dd_compile((DD<M2Handle>*)ddp, comp_type, max_width, root_state, root_value, root_depth, best_lb, cutoff_now, out);
}
int32_t oracle_m2s_dd_layers(void* ddp, int32_t* vars, int32_t* widths, int32_t cap) { return dd_layers((DD<M2Handle>*)ddp, vars, widths, cap); }
int32_t oracle_m2s_dd_cutset(void* ddp, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depths, int32_t* path_lens, int32_t* paths,
                             int32_t cap, int32_t path_stride) {
    return dd_cutset((DD<M2Handle>*)ddp, states, values, ubs, depths, path_lens, paths, cap, path_stride);
}
int32_t oracle_m2s_dd_solution(void* ddp, int32_t exact, int32_t* vars, int32_t* vals, int32_t cap) { return dd_solution((DD<M2Handle>*)ddp, exact, vars, vals, cap); }
int32_t oracle_m2s_solve(void* hp, int32_t mode, int32_t k, int32_t width_kind, uint64_t width, int32_t cutset_type, double time_budget_s,
                         uint64_t max_waves, oracle_solve_result* out, int32_t* sol_vars, int32_t* sol_vals, int32_t* sol_len, int64_t* trace,
                         int32_t trace_cap, int32_t* trace_len) {
    return solve((M2Handle*)hp, mode, k, width_kind, width, cutset_type, time_budget_s, max_waves, out, sol_vars, sol_vals, sol_len, trace, trace_cap, trace_len);
}
void* oracle_m2s_stepper_new(void* hp, int32_t k, int32_t width_kind, uint64_t width) { return new Stepper<M2Handle>((M2Handle*)hp, k, width_kind, width); }
void oracle_m2s_stepper_free(void* s) { delete (Stepper<M2Handle>*)s; }
void oracle_m2s_stepper_init(void* s, int32_t push_root) { ((Stepper<M2Handle>*)s)->solver->init(push_root != 0); }
int32_t oracle_m2s_stepper_wave(void* s, int64_t out3[3]) { return stepper_wave((Stepper<M2Handle>*)s, out3); }
void oracle_m2s_stepper_state(void* s, int64_t out[6]) { stepper_state((Stepper<M2Handle>*)s, out); }
void oracle_m2s_stepper_set_lb(void* s, int64_t lb) { ((Stepper<M2Handle>*)s)->solver->set_lower_bound(lb); }
void oracle_m2s_stepper_retain_share(void* s, int32_t rank, int32_t nranks) { ((Stepper<M2Handle>*)s)->solver->retain_share((size_t)rank, (size_t)nranks); }
void oracle_m2s_stepper_finish(void* s) { ((Stepper<M2Handle>*)s)->solver->finish(); }
int32_t oracle_m2s_stepper_node_words(void* s) { return (int32_t)stepper_node_words((Stepper<M2Handle>*)s); }
int32_t oracle_m2s_stepper_export(void* s, int32_t max_nodes, int64_t* rows) { return stepper_export((Stepper<M2Handle>*)s, max_nodes, rows); }
void oracle_m2s_stepper_import(void* s, int32_t count, const int64_t* rows) { stepper_import((Stepper<M2Handle>*)s, count, rows); }

// CPU baseline of a batch of independent sub-problems: for each root, restricted DD then (if inexact) relaxed DD, all against the same
// best_lb, on `threads` worker threads each owning one Mdd (the ParallelSolver worker body, parallel.rs:391-437, without the fringe).
// time_budget_s > 0: a TimeBudget cutoff (cutoff.rs:302-323) stops every worker; the layers expanded before the cutoff still count.
// per-root outputs (all optional): best exact value of the restricted DD (or INT64_MIN), best value of the relaxed DD (or INT64_MIN),
// cutset size.  Returns total expanded nodes; *seconds = wall-clock.
}  // extern "C"
namespace {
template <class H>
uint64_t compile_many(H* h, int32_t threads, int32_t n_roots, const uint64_t* root_states, const int64_t* root_values, const int32_t* root_depths,
                      const uint64_t* widths, int64_t best_lb, int32_t cutset_type, double time_budget_s, int64_t* restricted_best,
                      int64_t* relaxed_best, int32_t* cutset_sizes, uint64_t* transitions_out, double* seconds) {
    using S = typename H::State;
    size_t W = h->abi_words();
    std::atomic<int32_t> next{0};
    std::atomic<uint64_t> expanded{0}, transitions{0};
    NoCutoff nocut; std::unique_ptr<TimeBudget> tb;
    const Cutoff* cut = &nocut;
    if (time_budget_s > 0) { tb.reset(new TimeBudget(time_budget_s)); cut = tb.get(); }
    double t0 = now_s();
    auto work = [&]() {
        Mdd<S, typename H::Hash, typename H::Eq> mdd(cutset_type);
        EmptyCache<S> cache; EmptyDominanceChecker<S> dom;
        uint64_t exp = 0, tr = 0;
        for (;;) {
            int32_t i = next.fetch_add(1);
            if (i >= n_roots) break;
            SubProblem<S> root{std::make_shared<const S>(h->state_from_abi(root_states + (size_t)i * W, (size_t)root_depths[i])), root_values[i], {}, ISIZE_MAX,
                               (size_t)root_depths[i]};
            CompilationInput<S> in{CompilationType::Restricted, &h->pb, &h->rlx, &h->rk, cut, (size_t)widths[i], &root, best_lb, &cache, &dom};
            Completion c;
            bool ok = mdd.compile(in, &c);
            exp += mdd.expanded; tr += mdd.transitions;
            if (!ok) break;
            if (restricted_best) restricted_best[i] = mdd.best_exact_value().value_or(ISIZE_MIN);
            if (relaxed_best) relaxed_best[i] = ISIZE_MIN;
            if (cutset_sizes) cutset_sizes[i] = 0;
            if (c.is_exact) continue;
            in.comp_type = CompilationType::Relaxed;
            ok = mdd.compile(in, &c);
            exp += mdd.expanded; tr += mdd.transitions;
            if (!ok) break;
            if (relaxed_best) relaxed_best[i] = mdd.best_value().value_or(ISIZE_MIN);
            int32_t n = 0;
            if (!c.is_exact) mdd.drain_cutset([&](SubProblem<S>) { ++n; });
            if (cutset_sizes) cutset_sizes[i] = n;
        }
        expanded += exp; transitions += tr;
    };
    std::vector<std::thread> ts;
    for (int32_t t = 0; t < threads; ++t) ts.emplace_back(work);
    for (auto& t : ts) t.join();
    if (seconds) *seconds = now_s() - t0;
    if (transitions_out) *transitions_out = transitions.load();
    return expanded.load();
}
}  // namespace
extern "C" {
uint64_t oracle_misp_compile_many(void* hp, int32_t threads, int32_t n_roots, const uint64_t* root_states, const int64_t* root_values,
                                  const int32_t* root_depths, const uint64_t* widths, int64_t best_lb, int32_t cutset_type,
                                  int64_t* restricted_best, int64_t* relaxed_best, int32_t* cutset_sizes, uint64_t* transitions_out, double* seconds) {
    return compile_many((MispHandle*)hp, threads, n_roots, root_states, root_values, root_depths, widths, best_lb, cutset_type, 0.0, restricted_best, relaxed_best,
                        cutset_sizes, transitions_out, seconds);
}
uint64_t oracle_m2s_compile_many(void* hp, int32_t threads, int32_t n_roots, const uint64_t* root_states, const int64_t* root_values,
                                 const int32_t* root_depths, const uint64_t* widths, int64_t best_lb, int32_t cutset_type, double time_budget_s,
                                 int64_t* restricted_best, int64_t* relaxed_best, int32_t* cutset_sizes, uint64_t* transitions_out, double* seconds) {
    return compile_many((M2Handle*)hp, threads, n_roots, root_states, root_values, root_depths, widths, best_lb, cutset_type, time_budget_s, restricted_best,
                        relaxed_best, cutset_sizes, transitions_out, seconds);
}

// stepwise wave solver (CPU stand-in for the device solver in the gloo tests of the fringe-sharded driver)
void* oracle_misp_stepper_new(void* hp, int32_t k, int32_t width_kind, uint64_t width) { return new MispStepper((MispHandle*)hp, k, width_kind, width); }
void oracle_misp_stepper_free(void* s) { delete (MispStepper*)s; }
void oracle_misp_stepper_init(void* s, int32_t push_root) { ((MispStepper*)s)->solver->init(push_root != 0); }
int32_t oracle_misp_stepper_wave(void* s, int64_t out3[3]) { isize o[3]; bool ok = ((MispStepper*)s)->solver->wave(o); out3[0] = o[0]; out3[1] = o[1]; out3[2] = o[2]; return ok ? 0 : 1; }
void oracle_misp_stepper_set_lb(void* s, int64_t lb) { ((MispStepper*)s)->solver->set_lower_bound(lb); }
void oracle_misp_stepper_retain_share(void* s, int32_t rank, int32_t nranks) { ((MispStepper*)s)->solver->retain_share((size_t)rank, (size_t)nranks); }
void oracle_misp_stepper_finish(void* s) { ((MispStepper*)s)->solver->finish(); }
int64_t oracle_misp_stepper_sol_value(void* s) { return ((MispStepper*)s)->solver->sol_value; }
int32_t oracle_misp_stepper_node_words(void* s) { return (int32_t)stepper_node_words((MispStepper*)s); }
int32_t oracle_misp_stepper_export(void* s, int32_t max_nodes, int64_t* rows) { return stepper_export((MispStepper*)s, max_nodes, rows); }
void oracle_misp_stepper_import(void* s, int32_t count, const int64_t* rows) { stepper_import((MispStepper*)s, count, rows); }
// solution of the stepper's solver: returns its length (-1: none), decisions sorted by variable; *value = objective of that solution
int32_t oracle_misp_stepper_solution(void* s, int32_t* vars, int32_t* vals, int32_t cap) {
    auto& w = *((MispStepper*)s)->solver;
    if (!w.best_sol) return -1;
    int32_t n = 0;
    for (const Decision& d : *w.best_sol) { if (n < cap) { vars[n] = (int32_t)d.variable; vals[n] = (int32_t)d.value; } ++n; }
    return n;
}
// out[0..5] = best_lb, best_ub, fringe_len, explored, expanded, has_solution
void oracle_misp_stepper_state(void* s, int64_t out[6]) {
    auto& w = *((MispStepper*)s)->solver;
    out[0] = w.best_lb; out[1] = w.best_ub; out[2] = (int64_t)w.fringe_len(); out[3] = (int64_t)w.stats.explored; out[4] = (int64_t)w.stats.expanded; out[5] = w.best_sol.has_value();
}

// Knapsack (BASELINE config 1).  solver: 0 sequential, 2 parallel(k).  caching != 0: SimpleCache + KPDominance (SeqCachingSolverFc, knapsack/main.rs:329)
int32_t oracle_knapsack_solve(int32_t n, int64_t capacity, const int64_t* profit, const int64_t* weight, int32_t solver, int32_t k, int32_t width_kind,
                              uint64_t width, int32_t cutset_type, int32_t caching, oracle_solve_result* out, int32_t* taken) {
    std::vector<isize> p(profit, profit + n); std::vector<size_t> w(weight, weight + n);
    Knapsack pb((size_t)capacity, p, w);
    KPRelax rlx(&pb); KPRanking rk; KPDominance kd;
    FixedWidth<KnapsackState> fw((size_t)width); NbUnassignedWidth<KnapsackState> nw(pb.nb_variables());
    const WidthHeuristic<KnapsackState>* wh = width_kind == 0 ? (const WidthHeuristic<KnapsackState>*)&fw : (const WidthHeuristic<KnapsackState>*)&nw;
    NoCutoff nocut;
    EmptyDominanceChecker<KnapsackState> edom; SimpleDominanceChecker<KnapsackState> sdom(&kd, pb.nb_variables());
    EmptyCache<KnapsackState> ec; SimpleCache<KnapsackState, KnapsackHash, KnapsackEq> sc;
    MaxUB<KnapsackState> mx{&rk};
    NoDupFringe<KnapsackState, KnapsackHash, KnapsackEq> fringe(mx);
    SolverConfig<KnapsackState> cfg{&pb, &rlx, &rk, wh, caching ? (DominanceChecker<KnapsackState>*)&sdom : (DominanceChecker<KnapsackState>*)&edom,
                                    &nocut, &fringe, caching ? (Cache<KnapsackState>*)&sc : (Cache<KnapsackState>*)&ec, cutset_type};
    std::memset(out, 0, sizeof(*out));
    double t0 = now_s();
    Completion c; SolverStats st; isize lb, ub; std::optional<Solution> sol;
    if (solver == 0) { SequentialSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    else { ParallelSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg, (size_t)k); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    out->seconds = now_s() - t0;
    out->has_value = c.best_value.has_value(); out->best_value = c.best_value.value_or(0); out->is_exact = c.is_exact;
    out->best_lb = lb; out->best_ub = ub; out->explored = st.explored; out->expanded = st.expanded; out->transitions = st.transitions; out->compilations = st.compilations;
    if (taken && sol) { for (int32_t i = 0; i < n; ++i) taken[i] = 0; for (auto& d : *sol) taken[d.variable] = (int32_t)d.value; }
    return 0;
}

// TSPTW at the DD level (the checker of the device model to come): same entry points as the MISP / MAX2SAT models
static TsptwInstance make_tsptw_instance(int32_t n, const int64_t* dist, const int64_t* tw) {
    TsptwInstance inst;
    inst.nb_nodes = (size_t)n;
    inst.distances.assign((size_t)n, std::vector<size_t>((size_t)n, 0));
    for (int32_t i = 0; i < n; ++i) for (int32_t j = 0; j < n; ++j) inst.distances[i][j] = (size_t)dist[(size_t)i * n + j];
    for (int32_t i = 0; i < n; ++i) inst.timewindows.push_back(TimeWindow{(size_t)tw[2 * i], (size_t)tw[2 * i + 1]});
    return inst;
}
void* oracle_tsptw_new(int32_t n, const int64_t* dist, const int64_t* tw) { return new TsptwHandle(make_tsptw_instance(n, dist, tw)); }
void oracle_tsptw_free(void* h) { delete (TsptwHandle*)h; }
int32_t oracle_tsptw_words(void* h) { return (int32_t)((TsptwHandle*)h)->abi_words(); }
void oracle_tsptw_initial_state(void* hp, uint64_t* out) { TsptwHandle* h = (TsptwHandle*)hp; h->state_to_abi(h->pb.initial_state(), out); }
void* oracle_tsptw_dd_new(void* h, int32_t cutset_type) { return new DD<TsptwHandle>((TsptwHandle*)h, cutset_type); }
void oracle_tsptw_dd_free(void* dd) { delete (DD<TsptwHandle>*)dd; }
int32_t oracle_tsptw_dd_compile(void* ddp, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value,
                                uint64_t root_depth, int64_t best_lb, int32_t cutoff_now, oracle_dd_result* out) {
    return dd_compile((DD<TsptwHandle>*)ddp, comp_type, max_width, root_state, root_value, root_depth, best_lb, cutoff_now, out);
}
int32_t oracle_tsptw_dd_layers(void* ddp, int32_t* vars, int32_t* widths, int32_t cap) { return dd_layers((DD<TsptwHandle>*)ddp, vars, widths, cap); }
int32_t oracle_tsptw_dd_cutset(void* ddp, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depths, int32_t* path_lens, int32_t* paths,
                               int32_t cap, int32_t path_stride) {
    return dd_cutset((DD<TsptwHandle>*)ddp, states, values, ubs, depths, path_lens, paths, cap, path_stride);
}
int32_t oracle_tsptw_dd_solution(void* ddp, int32_t exact, int32_t* vars, int32_t* vals, int32_t cap) { return dd_solution((DD<TsptwHandle>*)ddp, exact, vars, vals, cap); }

// TSPTW (BASELINE config 4), examples/tsptw/main.rs:66-84 and tests.rs:33-57: TsptwWidth(nb_vars, factor), SimpleDominanceChecker(TsptwDominance),
// NoDupFringe(MaxUB(TsptwRanking)), DefaultCachingSolver = ParCachingSolverFc (FRONTIER cutset + SimpleCache, solver/mod.rs:30,37).
// dist: n x n row-major, tw: n x (earliest, latest), both already scaled to integers (instance.rs:86-98).  solver: 0 sequential, 2 parallel(k).
// perm[variable] = city visited at that step (main.rs:124-139).
int32_t oracle_tsptw_solve(int32_t n, const int64_t* dist, const int64_t* tw, int32_t factor, int32_t solver, int32_t k, int32_t cutset_type,
                           int32_t caching, double time_budget_s, oracle_solve_result* out, int32_t* perm) {
    TsptwInstance inst;
    inst.nb_nodes = (size_t)n;
    inst.distances.assign((size_t)n, std::vector<size_t>((size_t)n, 0));
    for (int32_t i = 0; i < n; ++i) for (int32_t j = 0; j < n; ++j) inst.distances[i][j] = (size_t)dist[(size_t)i * n + j];
    for (int32_t i = 0; i < n; ++i) inst.timewindows.push_back(TimeWindow{(size_t)tw[2 * i], (size_t)tw[2 * i + 1]});
    Tsptw pb(inst);
    TsptwRelax rlx(&pb); TsptwRanking rk; TsptwDominance td;
    TsptwWidth wh(pb.nb_variables(), (size_t)std::max(1, factor));
    NoCutoff nocut; std::unique_ptr<TimeBudget> tb;
    const Cutoff* cut = &nocut;
    if (time_budget_s > 0) { tb.reset(new TimeBudget(time_budget_s)); cut = tb.get(); }
    EmptyDominanceChecker<TsptwState> edom; SimpleDominanceChecker<TsptwState> sdom(&td, pb.nb_variables());
    EmptyCache<TsptwState> ec; SimpleCache<TsptwState, TsptwHash, TsptwEq> sc;
    MaxUB<TsptwState> mx{&rk};
    NoDupFringe<TsptwState, TsptwHash, TsptwEq> fringe(mx);
    SolverConfig<TsptwState> cfg{&pb, &rlx, &rk, &wh, caching ? (DominanceChecker<TsptwState>*)&sdom : (DominanceChecker<TsptwState>*)&edom,
                                 cut, &fringe, caching ? (Cache<TsptwState>*)&sc : (Cache<TsptwState>*)&ec, cutset_type};
    std::memset(out, 0, sizeof(*out));
    double t0 = now_s();
    Completion c; SolverStats st; isize lb, ub; std::optional<Solution> sol;
    if (solver == 0) { SequentialSolver<TsptwState, TsptwHash, TsptwEq> s(cfg); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    else { ParallelSolver<TsptwState, TsptwHash, TsptwEq> s(cfg, (size_t)k); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    out->seconds = now_s() - t0;
    out->has_value = c.best_value.has_value(); out->best_value = c.best_value.value_or(0); out->is_exact = c.is_exact;
    out->best_lb = lb; out->best_ub = ub; out->explored = st.explored; out->expanded = st.expanded; out->transitions = st.transitions; out->compilations = st.compilations;
    if (perm) { for (int32_t i = 0; i < n; ++i) perm[i] = -1; if (sol) for (auto& d : *sol) perm[d.variable] = (int32_t)d.value; }
    return 0;
}

// The LocBoundsAndThresholds example (clean.rs:2056-2181), relaxed, W = 3: textual dump of every node / edge so that the python
// test can compare it with resources/visualisation_tests/*.dot (clean.rs:2401-2546).  One line per node:
//   N <label> val locb rub theta exact relaxed cutset deleted      (locb/theta: "none" when unset; +inf when isize::MAX)
// one line per edge:  E <from-label> <to-label> <var> <value> <cost> <is_best>
int32_t oracle_locbounds_dump(int32_t cutset_type, int64_t best_lb, char* buf, int32_t cap) {
    LocBoundsPb pb; LocBoundsRelax rlx; CmpChar rk; NoCutoff nocut; EmptyDominanceChecker<char> dom;
    SimpleCache<char, CharHash, CharEq> cache; cache.initialize(pb);
    SubProblem<char> root{std::make_shared<const char>('r'), 0, {}, ISIZE_MAX, 0};
    CompilationInput<char> in{CompilationType::Relaxed, &pb, &rlx, &rk, &nocut, 3, &root, best_lb, &cache, &dom};
    Mdd<char, CharHash, CharEq> m(cutset_type);
    Completion c; m.compile(in, &c);
    std::ostringstream os;
    auto fmt = [](isize v) { return v == ISIZE_MAX ? std::string("+inf") : (v == ISIZE_MIN ? std::string("-inf") : std::to_string(v)); };
    for (size_t id = 0; id < m.nodes.size(); ++id) {
        const auto& n = m.nodes[id];
        os << "N " << *n.state << " " << fmt(n.value_top) << " " << fmt(n.value_bot) << " " << fmt(n.rub) << " "
           << (n.theta ? fmt(*n.theta) : std::string("none")) << " " << n.flags.is_exact() << " " << n.flags.is_relaxed() << " " << n.flags.is_cutset() << " "
           << n.flags.is_deleted() << "\n";
    }
    for (size_t id = 0; id < m.nodes.size(); ++id)
        m.foreach_edge_of(id, [&](const Mdd<char, CharHash, CharEq>::Edge& e) {
            bool best = m.nodes[id].best >= 0 && m.edges[m.nodes[id].best].from == e.from && m.edges[m.nodes[id].best].decision == e.decision &&
                        m.edges[m.nodes[id].best].cost == e.cost;
            os << "E " << *m.nodes[e.from].state << " " << *m.nodes[e.to].state << " " << e.decision.variable << " " << e.decision.value << " " << e.cost << " " << best << "\n";
        });
    std::string s = os.str();
    if ((int32_t)s.size() + 1 > cap) return -(int32_t)s.size() - 1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int32_t)s.size();
}

}  // extern "C"

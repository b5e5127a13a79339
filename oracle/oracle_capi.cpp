// ============================================================================
// oracle_capi.cpp -- TEST INFRASTRUCTURE.  C entry points (ctypes) over the CPU
// oracle.  Loaded only by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs.
// ============================================================================
#include "models.hpp"
#include <cstdio>
#include <sstream>

using namespace ddo_oracle;

namespace {
using MispMdd = Mdd<BitState, BitStateHash, BitStateEq>;
using MispFringe = NoDupFringe<BitState, BitStateHash, BitStateEq>;

struct MispHandle {
    Misp pb;
    MispRelax rlx;
    MispRanking rk;
    MispHandle(size_t n, const isize* w, size_t m, const int32_t* s, const int32_t* d) : pb(n, w, m, s, d), rlx(&pb) {}
};
struct MispDD {
    MispHandle* h;
    MispMdd mdd;
    std::vector<SubProblem<BitState>> cutset;  // drained copy
    SubProblem<BitState> root;
    MispDD(MispHandle* h_, int cutset_type) : h(h_), mdd(cutset_type) {}
};
struct MispStepper {
    MispHandle* h;
    FixedWidth<BitState> fw; NbUnassignedWidth<BitState> nw; NoCutoff nocut; EmptyDominanceChecker<BitState> dom; EmptyCache<BitState> cache;
    MaxUB<BitState> mx; MispFringe fringe;
    std::unique_ptr<WaveSolver<BitState, BitStateHash, BitStateEq>> solver;
    MispStepper(MispHandle* h_, int k, int width_kind, uint64_t width)
        : h(h_), fw((size_t)width), nw(h_->pb.nb_vars), mx{&h_->rk}, fringe(mx) {
        const WidthHeuristic<BitState>* wh = width_kind == 0 ? (const WidthHeuristic<BitState>*)&fw : (const WidthHeuristic<BitState>*)&nw;
        SolverConfig<BitState> cfg{&h->pb, &h->rlx, &h->rk, wh, &dom, &nocut, &fringe, &cache, LAST_EXACT_LAYER};
        solver.reset(new WaveSolver<BitState, BitStateHash, BitStateEq>(cfg, (size_t)k));
    }
};
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

void* oracle_misp_new(int32_t n, const int64_t* weights, int32_t m, const int32_t* src, const int32_t* dst) {
    return new MispHandle((size_t)n, weights, (size_t)m, src, dst);
}
void oracle_misp_free(void* h) { delete (MispHandle*)h; }
int32_t oracle_misp_words(void* h) { return (int32_t)((MispHandle*)h)->pb.words; }

void* oracle_misp_dd_new(void* h, int32_t cutset_type) { return new MispDD((MispHandle*)h, cutset_type); }
void oracle_misp_dd_free(void* dd) { delete (MispDD*)dd; }

// comp_type: 0 Exact, 1 Relaxed, 2 Restricted (abstraction/mdd.rs:40-47 order).  Returns 0 ok, 1 cutoff.
int32_t oracle_misp_dd_compile(void* ddp, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value,
                               uint64_t root_depth, int64_t best_lb, int32_t cutoff_now, oracle_dd_result* out) {
    MispDD* dd = (MispDD*)ddp;
    MispHandle* h = dd->h;
    BitState s; s.w.assign(root_state, root_state + h->pb.words);
    dd->root = SubProblem<BitState>{std::make_shared<const BitState>(std::move(s)), root_value, {}, ISIZE_MAX, (size_t)root_depth};
    EmptyCache<BitState> cache; EmptyDominanceChecker<BitState> dom; FlagCutoff cut(cutoff_now != 0);
    CompilationType t = comp_type == 0 ? CompilationType::Exact : (comp_type == 1 ? CompilationType::Relaxed : CompilationType::Restricted);
    CompilationInput<BitState> in{t, &h->pb, &h->rlx, &h->rk, &cut, (size_t)max_width, &dd->root, best_lb, &cache, &dom};
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
    dd->mdd.drain_cutset([&](SubProblem<BitState> n) { dd->cutset.push_back(std::move(n)); });
    out->cutset_size = (int32_t)dd->cutset.size();
    return 0;
}
// per expanded layer: branching variable and |curr_l| after the cut
int32_t oracle_misp_dd_layers(void* ddp, int32_t* vars, int32_t* widths, int32_t cap) {
    MispDD* dd = (MispDD*)ddp;
    int32_t n = (int32_t)dd->mdd.layer_vars.size();
    for (int32_t i = 0; i < n && i < cap; ++i) { vars[i] = (int32_t)dd->mdd.layer_vars[i]; widths[i] = (int32_t)dd->mdd.layer_widths[i]; }
    return n;
}
// drained cutset (MARKED nodes only, clean.rs:417-445) in drain order.  paths: path_stride (var,value) int32 pairs per node.
int32_t oracle_misp_dd_cutset(void* ddp, uint64_t* states, int64_t* values, int64_t* ubs, int32_t* depths, int32_t* path_lens,
                              int32_t* paths, int32_t cap, int32_t path_stride) {
    MispDD* dd = (MispDD*)ddp;
    size_t W = dd->h->pb.words;
    int32_t n = (int32_t)dd->cutset.size();
    for (int32_t i = 0; i < n && i < cap; ++i) {
        const auto& sp = dd->cutset[i];
        std::memcpy(states + (size_t)i * W, sp.state->w.data(), W * 8);
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
int32_t oracle_misp_dd_solution(void* ddp, int32_t exact, int32_t* vars, int32_t* vals, int32_t cap) {
    MispDD* dd = (MispDD*)ddp;
    auto sol = exact ? dd->mdd.best_exact_solution() : dd->mdd.best_solution();
    if (!sol) return -1;
    for (int32_t i = 0; i < (int32_t)sol->size() && i < cap; ++i) { vars[i] = (int32_t)(*sol)[i].variable; vals[i] = (int32_t)(*sol)[i].value; }
    return (int32_t)sol->size();
}

// mode: 0 SequentialSolver, 1 WaveSolver(k), 2 ParallelSolver(k threads).  width_kind: 0 FixedWidth(width), 1 NbUnassignedWidth.
// sol_yes: vertices with decision YES (cap n).  trace (mode 1): 4 int64 per wave (popped, best_lb, fringe_len, top_ub).
int32_t oracle_misp_solve(void* hp, int32_t mode, int32_t k, int32_t width_kind, uint64_t width, int32_t cutset_type, double time_budget_s,
                          uint64_t max_waves, oracle_solve_result* out, int32_t* sol_yes, int32_t* sol_len, int64_t* trace, int32_t trace_cap,
                          int32_t* trace_len) {
    MispHandle* h = (MispHandle*)hp;
    FixedWidth<BitState> fw((size_t)width); NbUnassignedWidth<BitState> nw(h->pb.nb_vars);
    const WidthHeuristic<BitState>* wh = width_kind == 0 ? (const WidthHeuristic<BitState>*)&fw : (const WidthHeuristic<BitState>*)&nw;
    NoCutoff nocut; std::unique_ptr<TimeBudget> tb;
    const Cutoff* cut = &nocut;
    if (time_budget_s > 0) { tb.reset(new TimeBudget(time_budget_s)); cut = tb.get(); }
    EmptyDominanceChecker<BitState> dom; EmptyCache<BitState> cache;
    MaxUB<BitState> mx{&h->rk};
    MispFringe fringe(mx);
    SolverConfig<BitState> cfg{&h->pb, &h->rlx, &h->rk, wh, &dom, cut, &fringe, &cache, cutset_type};
    std::memset(out, 0, sizeof(*out));
    if (trace_len) *trace_len = 0;
    double t0 = now_s();
    Completion c; SolverStats st; isize lb, ub; std::optional<Solution> sol;
    if (mode == 0) { SequentialSolver<BitState, BitStateHash, BitStateEq> s(cfg); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    else if (mode == 1) {
        WaveSolver<BitState, BitStateHash, BitStateEq> s(cfg, (size_t)k); s.max_waves = max_waves ? max_waves : UINT64_MAX;
        c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol;
        if (trace) {
            int32_t n = 0;
            for (auto& t : s.trace) { if (n >= trace_cap) break; trace[4 * n] = (int64_t)t.popped; trace[4 * n + 1] = t.best_lb; trace[4 * n + 2] = (int64_t)t.fringe_len; trace[4 * n + 3] = t.top_ub; ++n; }
            *trace_len = n;
        }
    } else { ParallelSolver<BitState, BitStateHash, BitStateEq> s(cfg, (size_t)k); c = s.maximize(); st = s.stats; lb = s.best_lb; ub = s.best_ub; sol = s.best_sol; }
    out->seconds = now_s() - t0;
    out->has_value = c.best_value.has_value(); out->best_value = c.best_value.value_or(0); out->is_exact = c.is_exact;
    out->best_lb = lb; out->best_ub = ub;
    out->explored = st.explored; out->expanded = st.expanded; out->transitions = st.transitions; out->compilations = st.compilations; out->waves = st.waves;
    int32_t n = 0;
    if (sol && sol_yes) for (auto& d : *sol) if (d.value == MISP_YES) sol_yes[n++] = (int32_t)d.variable;
    if (sol_len) *sol_len = n;
    return 0;
}

// CPU baseline of one bench "step": for each root, restricted DD then (if inexact) relaxed DD, all against the same best_lb,
// on `threads` worker threads each owning one Mdd (the ParallelSolver worker body, parallel.rs:391-437, without the fringe).
// per-root outputs (all optional): best exact value of the restricted DD (or INT64_MIN), best value of the relaxed DD (or INT64_MIN),
// cutset size.  Returns total expanded nodes; *seconds = wall-clock.
uint64_t oracle_misp_compile_many(void* hp, int32_t threads, int32_t n_roots, const uint64_t* root_states, const int64_t* root_values,
                                  const int32_t* root_depths, const uint64_t* widths, int64_t best_lb, int32_t cutset_type,
                                  int64_t* restricted_best, int64_t* relaxed_best, int32_t* cutset_sizes, uint64_t* transitions_out, double* seconds) {
    MispHandle* h = (MispHandle*)hp;
    size_t W = h->pb.words;
    std::atomic<int32_t> next{0};
    std::atomic<uint64_t> expanded{0}, transitions{0};
    double t0 = now_s();
    auto work = [&]() {
        MispMdd mdd(cutset_type);
        EmptyCache<BitState> cache; EmptyDominanceChecker<BitState> dom; NoCutoff nocut;
        uint64_t exp = 0, tr = 0;
        for (;;) {
            int32_t i = next.fetch_add(1);
            if (i >= n_roots) break;
            BitState s; s.w.assign(root_states + (size_t)i * W, root_states + (size_t)(i + 1) * W);
            SubProblem<BitState> root{std::make_shared<const BitState>(std::move(s)), root_values[i], {}, ISIZE_MAX, (size_t)root_depths[i]};
            CompilationInput<BitState> in{CompilationType::Restricted, &h->pb, &h->rlx, &h->rk, &nocut, (size_t)widths[i], &root, best_lb, &cache, &dom};
            Completion c;
            mdd.compile(in, &c);
            exp += mdd.expanded; tr += mdd.transitions;
            if (restricted_best) restricted_best[i] = mdd.best_exact_value().value_or(ISIZE_MIN);
            if (relaxed_best) relaxed_best[i] = ISIZE_MIN;
            if (cutset_sizes) cutset_sizes[i] = 0;
            if (c.is_exact) continue;
            in.comp_type = CompilationType::Relaxed;
            mdd.compile(in, &c);
            exp += mdd.expanded; tr += mdd.transitions;
            if (relaxed_best) relaxed_best[i] = mdd.best_value().value_or(ISIZE_MIN);
            int32_t n = 0;
            if (!c.is_exact) mdd.drain_cutset([&](SubProblem<BitState>) { ++n; });
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

// stepwise wave solver (CPU stand-in for the device solver in the gloo tests of the fringe-sharded driver)
void* oracle_misp_stepper_new(void* hp, int32_t k, int32_t width_kind, uint64_t width) { return new MispStepper((MispHandle*)hp, k, width_kind, width); }
void oracle_misp_stepper_free(void* s) { delete (MispStepper*)s; }
void oracle_misp_stepper_init(void* s, int32_t push_root) { ((MispStepper*)s)->solver->init(push_root != 0); }
int32_t oracle_misp_stepper_wave(void* s, int64_t out3[3]) { isize o[3]; bool ok = ((MispStepper*)s)->solver->wave(o); out3[0] = o[0]; out3[1] = o[1]; out3[2] = o[2]; return ok ? 0 : 1; }
void oracle_misp_stepper_set_lb(void* s, int64_t lb) { ((MispStepper*)s)->solver->set_lower_bound(lb); }
void oracle_misp_stepper_retain_share(void* s, int32_t rank, int32_t nranks) { ((MispStepper*)s)->solver->retain_share((size_t)rank, (size_t)nranks); }
void oracle_misp_stepper_finish(void* s) { ((MispStepper*)s)->solver->finish(); }
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

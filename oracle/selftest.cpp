// ============================================================================
// selftest.cpp -- TEST INFRASTRUCTURE.  Pins the oracle against the
// expectations hard-coded in the reference's own unit tests.  Each CHECK cites
// the reference test it restates (paths relative to /root/reference/ddo/src).
// Exit code 0 iff every check passes; one line per failed check.
// ============================================================================
#include "models.hpp"
#include <cstdio>
#include <map>

using namespace ddo_oracle;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, what)                                                     \
    do {                                                                      \
        ++g_checks;                                                           \
        if (!(cond)) { ++g_fail; std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, what); } \
    } while (0)

using DummyMdd = Mdd<DummyState, DummyHash, DummyEq>;
using CharMdd = Mdd<char, CharHash, CharEq>;
using DummyCache = SimpleCache<DummyState, DummyHash, DummyEq>;
using CharCache = SimpleCache<char, CharHash, CharEq>;

struct DummyEnv {
    DummyProblem pb; DummyRelax rlx; DummyRanking rk; NoCutoff nocut; FlagCutoff always{true};
    EmptyCache<DummyState> ecache; EmptyDominanceChecker<DummyState> dom;
    SubProblem<DummyState> root{std::make_shared<const DummyState>(DummyState{0, 0}), 0, {}, ISIZE_MAX, 0};
    CompilationInput<DummyState> input(CompilationType t, size_t w, isize lb, Cache<DummyState>* cache = nullptr, const Cutoff* c = nullptr) {
        return CompilationInput<DummyState>{t, &pb, &rlx, &rk, c ? c : &nocut, w, &root, lb, cache ? cache : &ecache, &dom};
    }
};

static bool same_path(const Solution& s, std::vector<std::pair<size_t, isize>> exp) {
    if (s.size() != exp.size()) return false;
    for (size_t i = 0; i < s.size(); ++i) if (s[i].variable != exp[i].first || s[i].value != exp[i].second) return false;
    return true;
}

static void test_dummy_dd() {
    DummyEnv e; Completion c;
    {   // clean.rs:1120-1150 root_remembers_the_pa_from_the_fringe_node
        SubProblem<DummyState> r{std::make_shared<const DummyState>(DummyState{42, 1}), 42, {Decision{0, 42}}, ISIZE_MAX, 1};
        for (auto t : {CompilationType::Exact, CompilationType::Relaxed, CompilationType::Restricted}) {
            auto in = e.input(t, 3, ISIZE_MIN); in.residual = &r;
            DummyMdd m; CHECK(m.compile(in, &c), "compile ok");
            auto sol = m.best_solution();
            CHECK(sol && sol->size() >= 1 && (*sol)[0] == (Decision{0, 42}), "root path kept in front of the best path");
        }
    }
    {   // clean.rs:1152-1188 exact_completely_unrolls_the_mdd_no_matter_its_width
        DummyMdd m; auto in = e.input(CompilationType::Exact, 1, ISIZE_MIN);
        CHECK(m.compile(in, &c), "exact compile");
        CHECK(m.best_value() == std::optional<isize>(6), "exact best 6");
        CHECK(same_path(*m.best_solution(), {{2, 2}, {1, 2}, {0, 2}}), "exact best path");
        CHECK(c.is_exact == m.is_exact() && c.best_value == m.best_value(), "completion coherent (clean.rs:1226-1254)");
        CHECK(m.is_exact(), "an_exact_mdd_must_be_exact clean.rs:1473");
    }
    {   // clean.rs:1190-1224 restricted_drops_the_less_interesting_nodes
        DummyMdd m; auto in = e.input(CompilationType::Restricted, 1, ISIZE_MIN);
        CHECK(m.compile(in, &c), "restricted compile");
        CHECK(m.best_value() == std::optional<isize>(6), "restricted best 6");
        CHECK(same_path(*m.best_solution(), {{2, 2}, {1, 2}, {0, 2}}), "restricted best path");
        CHECK(c.is_exact == m.is_exact() && c.best_value == m.best_value(), "completion coherent (clean.rs:1256-1284)");
        CHECK(!m.is_exact(), "restricted W=1 is inexact");
    }
    {   // clean.rs:1405-1440 relaxed_merges_the_less_interesting_nodes
        DummyMdd m; auto in = e.input(CompilationType::Relaxed, 1, ISIZE_MIN);
        CHECK(m.compile(in, &c), "relaxed compile");
        CHECK(m.best_value() == std::optional<isize>(24), "relaxed best 24");
        CHECK(same_path(*m.best_solution(), {{2, 2}, {1, 0}, {0, 2}}), "relaxed best path [x2=2,x1=0,x0=2]");
        CHECK(c.is_exact == m.is_exact() && c.best_value == m.best_value(), "completion coherent (clean.rs:1286-1314)");
        // clean.rs:1442-1471 relaxed_populates_the_cutset_and_will_not_squash_first_layer
        size_t n = 0; m.drain_cutset([&](SubProblem<DummyState>) { ++n; });
        CHECK(n == 3, "cutset size 3: L1 not squashed");
        CHECK(!m.is_exact(), "a_relaxed_mdd_is_not_exact_when_a_merge_occurred clean.rs:1530");
    }
    {   // clean.rs:1322-1403 cutoff
        for (auto t : {CompilationType::Exact, CompilationType::Relaxed, CompilationType::Restricted}) {
            DummyMdd m; auto in = e.input(t, 1, ISIZE_MIN, nullptr, &e.always);
            CHECK(!m.compile(in, &c), "Err(CutoffOccurred)");
        }
    }
    {   // clean.rs:1501-1528, 1559-1586 exact as long as no squash
        DummyMdd m; auto in = e.input(CompilationType::Relaxed, 10, ISIZE_MIN);
        CHECK(m.compile(in, &c) && m.is_exact(), "relaxed W=10 exact");
        auto in2 = e.input(CompilationType::Restricted, 10, ISIZE_MIN);
        CHECK(m.compile(in2, &c) && m.is_exact(), "restricted W=10 exact");
    }
    {   // clean.rs:1615-1668 infeasible
        DummyEnv f; f.pb.infeasible = true;
        DummyMdd m; auto in = f.input(CompilationType::Exact, SIZE_MAX, ISIZE_MIN);
        CHECK(m.compile(in, &c), "infeasible compile");
        CHECK(!m.best_solution() && !m.best_value(), "infeasible: no solution / value");
    }
    {   // clean.rs:1669-1749 rub pruning vs best_lb = 1000
        for (auto t : {CompilationType::Exact, CompilationType::Relaxed, CompilationType::Restricted}) {
            DummyMdd m; auto in = e.input(t, SIZE_MAX, 1000);
            CHECK(m.compile(in, &c) && !m.best_solution(), "ub <= best_lb: nothing expanded");
        }
    }
    {   // clean.rs:1750-1842 cache thresholds prune
        for (auto t : {CompilationType::Exact, CompilationType::Relaxed, CompilationType::Restricted}) {
            DummyCache cache; cache.initialize(e.pb);
            for (isize v = 0; v <= 2; ++v) cache.update_threshold(std::make_shared<const DummyState>(DummyState{v, 1}), 1, v, true);
            DummyMdd m; auto in = e.input(t, SIZE_MAX, ISIZE_MIN, &cache);
            CHECK(m.compile(in, &c) && !m.best_solution(), "cache-threshold pruning");
        }
    }
    {   // clean.rs:1844-1948 thresholds when exact
        for (auto t : {CompilationType::Restricted, CompilationType::Relaxed}) {
            DummyCache cache; cache.initialize(e.pb);
            DummyMdd m; auto in = e.input(t, 10, ISIZE_MIN, &cache);
            CHECK(m.compile(in, &c) && m.is_exact(), "exact W=10");
            isize exp_by_depth[4] = {0, 2, 4, 6};
            size_t maxv[4] = {0, 2, 4, 6};
            for (size_t d = 0; d <= 3; ++d)
                for (size_t v = 0; v <= maxv[d]; ++v) {
                    auto th = cache.get_threshold(DummyState{(isize)v, d}, d);
                    CHECK(th && th->value == exp_by_depth[d] && th->explored, "threshold table (exact)");
                }
        }
    }
    {   // clean.rs:1950-2054 thresholds when all pruned (best_lb = 15)
        for (auto t : {CompilationType::Restricted, CompilationType::Relaxed}) {
            DummyCache cache; cache.initialize(e.pb);
            DummyMdd m; auto in = e.input(t, 10, 15, &cache);
            CHECK(m.compile(in, &c) && m.is_exact(), "exact W=10 lb=15");
            isize exp_by_depth[3] = {1, 3, 5};
            size_t maxv[4] = {0, 2, 4, 6};
            for (size_t d = 0; d <= 2; ++d)
                for (size_t v = 0; v <= maxv[d]; ++v) {
                    auto th = cache.get_threshold(DummyState{(isize)v, d}, d);
                    CHECK(th && th->value == exp_by_depth[d] && th->explored, "threshold table (pruned)");
                }
            for (size_t v = 0; v <= 6; ++v) CHECK(!cache.get_threshold(DummyState{(isize)v, 3}, 3), "depth 3: none");
        }
    }
}

static void test_locbounds() {
    LocBoundsPb pb; LocBoundsRelax rlx; CmpChar rk; NoCutoff nocut; EmptyDominanceChecker<char> dom;
    SubProblem<char> root{std::make_shared<const char>('r'), 0, {}, ISIZE_MAX, 0};
    auto run = [&](int cutset, isize lb, std::map<char, isize>& ubs, CharCache& cache, Completion& c, CharMdd& m) {
        cache.initialize(pb);
        CompilationInput<char> in{CompilationType::Relaxed, &pb, &rlx, &rk, &nocut, 3, &root, lb, &cache, &dom};
        bool ok = m.compile(in, &c);
        m.drain_cutset([&](SubProblem<char> n) { ubs[*n.state] = n.ub; });
        return ok;
    };
    auto thr = [](CharCache& c, char s, size_t d) { return c.get_threshold(s, d); };
    {   // clean.rs:2183-2242 (LEL)
        std::map<char, isize> v; CharCache cache; Completion c; CharMdd m(LAST_EXACT_LAYER);
        CHECK(run(LAST_EXACT_LAYER, 0, v, cache, c, m), "compile");
        CHECK(!m.is_exact() && m.best_value() == std::optional<isize>(16), "LEL: inexact, best 16");
        CHECK(v.size() == 2 && v['a'] == 16 && v['b'] == 14, "LEL cutset ubs a=16 b=14");
        CHECK(thr(cache, 'r', 0) && thr(cache, 'a', 1) && thr(cache, 'b', 1), "thresholds r,a,b present");
        for (auto p : std::vector<std::pair<char, size_t>>{{'M', 2}, {'e', 2}, {'f', 2}, {'g', 3}, {'h', 3}, {'i', 3}, {'t', 4}})
            CHECK(!thr(cache, p.first, p.second), "no threshold below the LEL cutset");
        CHECK(thr(cache, 'r', 0)->value == 0 && thr(cache, 'r', 0)->explored, "theta r = 0 explored");
        CHECK(thr(cache, 'a', 1)->value == 10 && !thr(cache, 'a', 1)->explored, "theta a = 10");
        CHECK(thr(cache, 'b', 1)->value == 7 && !thr(cache, 'b', 1)->explored, "theta b = 7");
    }
    {   // clean.rs:2244-2321 (frontier)
        std::map<char, isize> v; CharCache cache; Completion c; CharMdd m(FRONTIER);
        CHECK(run(FRONTIER, 0, v, cache, c, m), "compile");
        CHECK(!m.is_exact() && m.best_value() == std::optional<isize>(16), "FC: inexact, best 16");
        CHECK(v.size() == 4 && v['a'] == 16 && v['b'] == 14 && v['h'] == 13 && v['i'] == 14, "FC cutset ubs");
        CHECK(!thr(cache, 'M', 2) && !thr(cache, 'g', 3) && !thr(cache, 't', 4), "no threshold for relaxed nodes");
        struct E { char s; size_t d; isize v; bool ex; };
        for (E x : std::vector<E>{{'r', 0, 0, true}, {'a', 1, 10, false}, {'b', 1, 7, false}, {'e', 2, 13, true}, {'f', 2, 12, true}, {'h', 3, 13, false}, {'i', 3, 14, false}}) {
            auto t = thr(cache, x.s, x.d);
            CHECK(t && t->value == x.v && t->explored == x.ex, "FC threshold table");
        }
    }
    {   // clean.rs:2323-2398 (frontier, best_lb = 15)
        std::map<char, isize> v; CharCache cache; Completion c; CharMdd m(FRONTIER);
        CHECK(run(FRONTIER, 15, v, cache, c, m), "compile");
        CHECK(!m.is_exact() && m.best_value() == std::optional<isize>(16), "FC lb=15: inexact, best 16");
        CHECK(v.size() == 2 && v['a'] == 16 && v['b'] == 14, "FC lb=15 cutset ubs");
        struct E { char s; size_t d; isize v; bool ex; };
        for (E x : std::vector<E>{{'r', 0, 0, true}, {'a', 1, 10, false}, {'b', 1, 8, false}, {'e', 2, 15, true}, {'f', 2, 13, true}, {'h', 3, 15, true}, {'i', 3, 15, true}}) {
            auto t = thr(cache, x.s, x.d);
            CHECK(t && t->value == x.v && t->explored == x.ex, "FC lb=15 threshold table");
        }
    }
}

static void test_flags() {  // node_flags.rs:191-719 (semantics)
    NodeFlags f = NodeFlags::new_exact();
    CHECK(f.is_exact() && !f.is_relaxed() && !f.is_marked() && !f.is_cutset() && !f.is_deleted(), "new_exact");
    f.set_relaxed(true); CHECK(!f.is_exact() && f.is_relaxed(), "relaxed masks exact (node_flags.rs:88)");
    f.set_relaxed(false); CHECK(f.is_exact(), "un-relax");
    NodeFlags r = NodeFlags::new_relaxed();
    CHECK(!r.is_exact() && r.is_relaxed(), "new_relaxed");
    r.set_exact(true); CHECK(!r.is_exact(), "relaxed stays inexact even with F_EXACT");
    f.set_marked(true); f.set_cutset(true); f.set_deleted(true); f.set_pruned_by_cache(true); f.set_above_cutset(true);
    CHECK(f.is_marked() && f.is_cutset() && f.is_deleted() && f.is_pruned_by_cache() && f.is_above_cutset() && f.is_exact(), "independent bits");
    CHECK(NodeFlags::F_EXACT == 1 && NodeFlags::F_RELAXED == 2 && NodeFlags::F_MARKED == 4 && NodeFlags::F_CUTSET == 8 && NodeFlags::F_DELETED == 16 && NodeFlags::F_CACHE == 32 && NodeFlags::F_ABOVE_CUTSET == 64, "bit values node_flags.rs:51-63");
}

static void test_fringe() {  // no_duplicate.rs:326-663 + subproblem_ranking.rs doc-test :50-74
    CmpChar rk; MaxUB<char> mx{&rk};
    auto sp = [](char s, isize value, isize ub) { return SubProblem<char>{std::make_shared<const char>(s), value, {}, ub, 0}; };
    {
        NoDupFringe<char, CharHash, CharEq> q(mx);
        q.push(sp('a', 10, 300)); q.push(sp('b', 2, 100)); q.push(sp('c', 24, 150));
        q.push(sp('d', 13, 13)); q.push(sp('e', 65, 700)); q.push(sp('f', 19, 100));
        std::string order;
        while (auto n = q.pop()) order.push_back(*n->state);
        CHECK(order == "eacfbd", "MaxUB pop order e,a,c,f,b,d");
    }
    {
        NoDupFringe<char, CharHash, CharEq> q(mx);
        CHECK(q.is_empty() && q.len() == 0, "empty");
        q.push(sp('x', 5, 50)); q.push(sp('x', 7, 40));  // same state: longer path wins, ub = max
        CHECK(q.len() == 1, "one entry per state");
        auto n = q.pop();
        CHECK(n && n->value == 7 && n->ub == 50, "longest path kept, ub = max(old,new)");
        q.push(sp('x', 7, 40)); q.push(sp('x', 5, 60));
        n = q.pop();
        CHECK(n && n->value == 7 && n->ub == 60, "ub raised, value kept");
        CHECK(!q.pop(), "pop on empty -> None");
        q.push(sp('a', 1, 1)); q.push(sp('b', 1, 2)); q.clear();
        CHECK(q.is_empty(), "clear");
    }
    {
        SimpleFringe<char> q(mx);
        q.push(sp('a', 10, 300)); q.push(sp('b', 2, 100)); q.push(sp('e', 65, 700));
        CHECK(*q.pop()->state == 'e' && *q.pop()->state == 'a' && *q.pop()->state == 'b' && !q.pop(), "SimpleFringe order");
    }
}

// parallel.rs:1256-1337 / sequential.rs:750-909: the 3-item / 7-item knapsacks solved by every solver flavour -> 220
static void test_solvers_knapsack() {
    using KMdd = Mdd<KnapsackState, KnapsackHash, KnapsackEq>;
    (void)sizeof(KMdd);
    for (int cutset : {LAST_EXACT_LAYER, FRONTIER})
        for (int flavour = 0; flavour < 3; ++flavour)
            for (int caching = 0; caching < 2; ++caching) {
                Knapsack pb(50, {60, 100, 120}, {10, 20, 30});
                KPRelax rlx(&pb); KPRanking rk; NbUnassignedWidth<KnapsackState> w(pb.nb_variables()); NoCutoff nocut;
                EmptyDominanceChecker<KnapsackState> edom;
                MaxUB<KnapsackState> mx{&rk};
                NoDupFringe<KnapsackState, KnapsackHash, KnapsackEq> fr(mx);
                EmptyCache<KnapsackState> ec; SimpleCache<KnapsackState, KnapsackHash, KnapsackEq> sc;
                Cache<KnapsackState>* cache = caching ? (Cache<KnapsackState>*)&sc : (Cache<KnapsackState>*)&ec;
                SolverConfig<KnapsackState> cfg{&pb, &rlx, &rk, &w, &edom, &nocut, &fr, cache, cutset};
                Completion c; isize lb = 0, ub = 0; Solution sol;
                if (flavour == 0) { SequentialSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg); c = s.maximize(); lb = s.best_lb; ub = s.best_ub; sol = *s.best_sol; }
                else if (flavour == 1) { ParallelSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg, 4); c = s.maximize(); lb = s.best_lb; ub = s.best_ub; sol = *s.best_sol; }
                else { if (caching) continue; WaveSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg, 3); c = s.maximize(); lb = s.best_lb; ub = s.best_ub; sol = *s.best_sol; }
                CHECK(c.is_exact && c.best_value == std::optional<isize>(220) && lb == 220 && ub == 220, "3-item knapsack -> 220 (parallel.rs:902-950)");
                CHECK(same_path(sol, {{0, 0}, {1, 1}, {2, 1}}), "decisions x0=0,x1=1,x2=1 (parallel.rs:941-948)");
            }
}

static void test_solvers_knapsack7_and_primal() {  // parallel.rs:981-1254, sequential.rs (same vectors): 7 items, set_primal, gap
    for (int cutset : {LAST_EXACT_LAYER, FRONTIER})
        for (int flavour = 0; flavour < 2; ++flavour) {
            Knapsack pb(50, {60, 210, 12, 5, 100, 120, 110}, {10, 45, 20, 4, 20, 30, 50});
            KPRelax rlx(&pb); KPRanking rk; NbUnassignedWidth<KnapsackState> w(pb.nb_variables()); NoCutoff nocut;
            EmptyDominanceChecker<KnapsackState> edom;
            MaxUB<KnapsackState> mx{&rk};
            NoDupFringe<KnapsackState, KnapsackHash, KnapsackEq> fr(mx);
            EmptyCache<KnapsackState> ec; SimpleCache<KnapsackState, KnapsackHash, KnapsackEq> sc;
            Cache<KnapsackState>* cache = cutset == FRONTIER ? (Cache<KnapsackState>*)&sc : (Cache<KnapsackState>*)&ec;  // DdLel / DdFc, parallel.rs:658-659
            SolverConfig<KnapsackState> cfg{&pb, &rlx, &rk, &w, &edom, &nocut, &fr, cache, cutset};
            Completion c; Solution sol;
            if (flavour == 0) { SequentialSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg); c = s.maximize(); sol = *s.best_sol; }
            else { ParallelSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg, 1); c = s.maximize(); sol = *s.best_sol; }
            CHECK(c.is_exact && c.best_value == std::optional<isize>(220), "7-item knapsack -> 220 (parallel.rs:981-1150)");
            CHECK(same_path(sol, {{0, 0}, {1, 0}, {2, 0}, {3, 0}, {4, 1}, {5, 1}, {6, 0}}), "decisions: items 4 and 5 (parallel.rs:1012-1021)");
        }
    {   // set_primal_overwrites_best_value_and_sol_if_it_improves (parallel.rs:1153-1199), gap (:1201-1254)
        Knapsack pb(50, {60, 100, 120}, {10, 20, 30});
        KPRelax rlx(&pb); KPRanking rk; NbUnassignedWidth<KnapsackState> w(pb.nb_variables()); NoCutoff nocut;
        EmptyDominanceChecker<KnapsackState> edom; MaxUB<KnapsackState> mx{&rk};
        NoDupFringe<KnapsackState, KnapsackHash, KnapsackEq> fr(mx); EmptyCache<KnapsackState> ec;
        SolverConfig<KnapsackState> cfg{&pb, &rlx, &rk, &w, &edom, &nocut, &fr, &ec, LAST_EXACT_LAYER};
        SequentialSolver<KnapsackState, KnapsackHash, KnapsackEq> s(cfg);
        CHECK(s.best_lb == ISIZE_MIN && s.best_ub == ISIZE_MAX && !s.best_sol && s.gap() == 1.0, "defaults: lb -inf, ub +inf, no solution, gap 1 (parallel.rs:662-876,1201-1225)");
        Solution one{Decision{0, 10}};
        s.set_primal(10, one); CHECK(s.best_sol && s.best_lb == 10, "set_primal(10)");
        s.set_primal(5, one); CHECK(s.best_sol && s.best_lb == 10, "set_primal(5) does not improve");
        s.set_primal(10000, one); CHECK(s.best_sol && s.best_lb == 10000, "set_primal(10000)");
        Completion c = s.maximize();
        CHECK(c.is_exact && c.best_value == std::optional<isize>(10000) && s.best_sol, "a better primal survives maximize()");
        NoDupFringe<KnapsackState, KnapsackHash, KnapsackEq> fr2(mx);
        SolverConfig<KnapsackState> cfg2{&pb, &rlx, &rk, &w, &edom, &nocut, &fr2, &ec, LAST_EXACT_LAYER};
        SequentialSolver<KnapsackState, KnapsackHash, KnapsackEq> s2(cfg2);
        Completion c2 = s2.maximize();
        CHECK(c2.is_exact && c2.best_value == std::optional<isize>(220) && s2.gap() == 0.0, "gap 0 once the optimum is proven (parallel.rs:1227-1254)");
    }
}

static void test_maxub_and_defaults() {  // subproblem_ranking.rs:134-175, abstraction/dp.rs:128-137, cutoff.rs (NoCutoff / TimeBudget)
    CmpChar rk; MaxUB<char> cmp{&rk};
    auto sp = [](char c, isize value, isize ub) { return SubProblem<char>{std::make_shared<const char>(c), value, {}, ub, 0}; };
    auto sg = [](int c) { return c < 0 ? -1 : (c > 0 ? 1 : 0); };
    CHECK(sg(cmp.compare(sp('a', 42, 300), sp('b', 42, 100))) == 1 && sg(cmp.compare(sp('b', 42, 100), sp('a', 42, 300))) == -1, "MaxUB: upper bound first");
    CHECK(sg(cmp.compare(sp('a', 42, 300), sp('b', 2, 300))) == 1 && sg(cmp.compare(sp('b', 2, 300), sp('a', 42, 300))) == -1, "MaxUB: then the longest path");
    CHECK(sg(cmp.compare(sp('a', 42, 300), sp('b', 42, 300))) == -1 && sg(cmp.compare(sp('a', 42, 300), sp('a', 42, 300))) == 0, "MaxUB: then the state ranking; equal to itself");
    LocBoundsRelax lr; struct NoRub : Relaxation<char> {
        char merge(const std::vector<const char*>&) const override { return 'M'; }
        isize relax(const char&, const char&, const char&, Decision, isize c) const override { return c; }
    } nr;
    LocBoundsPb pb;
    CHECK(nr.fast_upper_bound('x') == ISIZE_MAX && pb.is_impacted_by(Variable{10}, 'x'), "defaults: fast_upper_bound = +inf, every state impacted by every variable");
    (void)lr;
    NoCutoff nc; TimeBudget far(3600.0), gone(0.0);
    CHECK(!nc.must_stop() && !far.must_stop() && gone.must_stop(), "NoCutoff never stops; TimeBudget stops only once elapsed");
}

static void test_width() {  // heuristics/width.rs:884-1075 (test_nbunassigned, test_fixedwidth, test_adapters)
    auto sub = [](size_t decided) {
        SubProblem<char> s{std::make_shared<const char>('a'), 10, {}, 100, decided};
        for (size_t i = 0; i < decided; ++i) s.path.push_back(Decision{i, (isize)i});
        return s;
    };
    NbUnassignedWidth<char> nb(5);
    CHECK(nb.max_width(sub(1)) == 4 && nb.max_width(sub(0)) == 5 && nb.max_width(sub(5)) == 0, "NbUnassignedWidth 4 / 5 / 0");
    FixedWidth<char> f5(5);
    CHECK(f5.max_width(sub(1)) == 5 && f5.max_width(sub(0)) == 5 && f5.max_width(sub(5)) == 5, "FixedWidth is constant");
    CHECK(Times<char>(2, &f5).max_width(sub(5)) == 10 && Times<char>(3, &f5).max_width(sub(5)) == 15 && Times<char>(1, &f5).max_width(sub(5)) == 5 &&
          Times<char>(10, &f5).max_width(sub(5)) == 50, "Times 10 / 15 / 5 / 50");
    FixedWidth<char> f4(4), f9(9), f10(10), f0(0);
    CHECK(DivBy<char>(2, &f4).max_width(sub(5)) == 2 && DivBy<char>(3, &f9).max_width(sub(5)) == 3 && DivBy<char>(1, &f10).max_width(sub(5)) == 10, "DivBy 2 / 3 / 10");
    CHECK(Times<char>(0, &f10).max_width(sub(5)) == 1 && Times<char>(10, &f0).max_width(sub(5)) == 1, "wrappers never return a zero max_width");
}

using VecState = std::vector<isize>;
struct DummyDominance : Dominance<VecState> {  // dominance/simple.rs:226-242
    std::optional<isize> get_key(const VecState& s) const override { return s[0]; }
    size_t nb_dimensions(const VecState& s) const override { return s.size(); }
    isize get_coordinate(const VecState& s, size_t i) const override { return s[i]; }
};
struct DummyDominanceWithValue : DummyDominance { bool use_value() const override { return true; } };  // :244-263
static void test_dominance() {  // dominance/simple.rs:119-224
    auto st = [](std::initializer_list<isize> v) { return std::make_shared<const VecState>(v); };
    auto none = [](const DominanceCheckResult& r) { return !r.dominated && !r.threshold; };
    DummyDominance dd; DummyDominanceWithValue dv;
    {   // not_dominated_when_keys_are_different
        SimpleDominanceChecker<VecState> d(&dv, 0);
        CHECK(none(d.is_dominated_or_insert(st({3, 0}), 0, 3)) && none(d.is_dominated_or_insert(st({2, 0}), 0, 2)) && none(d.is_dominated_or_insert(st({1, 0}), 0, 1)) &&
              none(d.is_dominated_or_insert(st({0, 0}), 0, 0)), "different keys never dominate");
    }
    {   // dominated_when_keys_are_equal
        SimpleDominanceChecker<VecState> d(&dd, 0);
        CHECK(none(d.is_dominated_or_insert(st({0, 3}), 0, 0)), "first entry");
        CHECK(d.is_dominated_or_insert(st({0, 2}), 0, 2).dominated && d.is_dominated_or_insert(st({0, 1}), 0, 1).dominated && d.is_dominated_or_insert(st({0, 0}), 0, 0).dominated,
              "dominated by coordinates");
        SimpleDominanceChecker<VecState> e(&dv, 0);
        CHECK(none(e.is_dominated_or_insert(st({0, 0}), 0, 3)), "first entry (value)");
        CHECK(e.is_dominated_or_insert(st({0, 0}), 0, 2).dominated && e.is_dominated_or_insert(st({0, 0}), 0, 1).dominated && e.is_dominated_or_insert(st({0, 0}), 0, 0).dominated,
              "dominated by value");
    }
    {   // not_dominated_when_keys_are_equal
        SimpleDominanceChecker<VecState> d(&dd, 0);
        CHECK(none(d.is_dominated_or_insert(st({0, 0, 3}), 0, 3)) && none(d.is_dominated_or_insert(st({0, 0, 3}), 0, 1)) && none(d.is_dominated_or_insert(st({0, 1, 1}), 0, 5)) &&
              none(d.is_dominated_or_insert(st({0, 0, 4}), 0, 3)), "incomparable or better entries are kept");
        SimpleDominanceChecker<VecState> e(&dv, 0);
        CHECK(none(e.is_dominated_or_insert(st({0, 0}), 0, 3)) && none(e.is_dominated_or_insert(st({0, 3}), 0, 0)) && none(e.is_dominated_or_insert(st({0, 1}), 0, 1)) &&
              none(e.is_dominated_or_insert(st({0, 0}), 0, 5)), "incomparable with value");
    }
    {   // pruning_threshold_when_value_is_used
        SimpleDominanceChecker<VecState> d(&dv, 0);
        CHECK(none(d.is_dominated_or_insert(st({0, 0}), 0, 3)), "first");
        auto a = d.is_dominated_or_insert(st({0, 0}), 0, 2), b = d.is_dominated_or_insert(st({0, 0}), 0, 1), c = d.is_dominated_or_insert(st({0, -1}), 0, 0);
        CHECK(a.dominated && a.threshold == std::optional<isize>(2) && b.dominated && b.threshold == std::optional<isize>(2) && c.dominated && c.threshold == std::optional<isize>(3),
              "thresholds 2 / 2 / 3");
    }
    {   // entry_is_added_only_when_dominant / entry_is_removed_when_dominated (sizes of the per-key entry list)
        SimpleDominanceChecker<VecState> d(&dv, 0);
        auto size = [&]() { size_t n = 0; for (auto& kv : d.data[0]->m) n += kv.second.size(); return n; };
        CHECK(none(d.is_dominated_or_insert(st({0, 0}), 0, 3)) && size() == 1, "one entry");
        auto r = d.is_dominated_or_insert(st({0, 0}), 0, 1);
        CHECK(r.dominated && r.threshold == std::optional<isize>(2) && size() == 1, "a dominated state is not stored");
        CHECK(none(d.is_dominated_or_insert(st({0, 1}), 0, 1)) && size() == 2, "incomparable: stored");
        CHECK(none(d.is_dominated_or_insert(st({0, -1}), 0, 5)) && size() == 3, "incomparable: stored (2)");
        SimpleDominanceChecker<VecState> e(&dv, 0);
        auto size_e = [&]() { size_t n = 0; for (auto& kv : e.data[0]->m) n += kv.second.size(); return n; };
        CHECK(none(e.is_dominated_or_insert(st({0, 0}), 0, 3)) && size_e() == 1, "one entry");
        CHECK(none(e.is_dominated_or_insert(st({0, 1}), 0, 5)) && size_e() == 1, "the dominated entry is replaced");
        CHECK(none(e.is_dominated_or_insert(st({0, 2}), 0, 7)) && size_e() == 1, "and again");
    }
}

static void test_dominance_cmp() {  // abstraction/dominance.rs:134-192
    DummyDominance dd; DummyDominanceWithValue dv;
    const VecState z{0, 0, 0};
    auto pc = [](const Dominance<VecState>& d, const VecState& a, isize va, const VecState& b, isize vb) { return d.partial_cmp(a, va, b, vb); };
    auto is = [](const std::optional<DominanceCmpResult>& r, int ord, bool only) { return r && (r->ordering < 0 ? -1 : (r->ordering > 0 ? 1 : 0)) == ord && r->only_val_diff == only; };
    CHECK(!dd.use_value(), "by_default_value_is_unused");
    CHECK(!pc(dd, z, 0, VecState{0, -1, 1}, 0) && !pc(dv, z, 0, VecState{0, 0, 1}, -1), "partial_cmp: None when coordinates (or the value) disagree");
    CHECK(is(pc(dd, z, 0, VecState{0, 0, 1}, 0), -1, false) && is(pc(dd, z, 0, VecState{0, 0, 1}, -1), -1, false) && is(pc(dd, z, 0, VecState{0, 0, -1}, 0), 1, false) &&
          is(pc(dd, z, 0, VecState{0, 0, -1}, 1), 1, false) && is(pc(dd, z, 0, z, 1), 0, false), "partial_cmp without value");
    CHECK(is(pc(dv, z, 0, z, 1), -1, true) && is(pc(dv, z, 0, VecState{1, 1, 1}, 1), -1, false) && is(pc(dv, z, 0, z, -1), 1, true) &&
          is(pc(dv, z, 0, VecState{-1, -1, -1}, -1), 1, false) && is(pc(dv, z, 0, z, 0), 0, false), "partial_cmp with value (only_val_diff)");
    auto sg = [](int c) { return c < 0 ? -1 : (c > 0 ? 1 : 0); };
    CHECK(sg(dd.cmp(z, 0, VecState{0, 0, 1}, 0)) == -1 && sg(dd.cmp(z, 0, VecState{0, 1, -1}, 0)) == -1 && sg(dd.cmp(z, 0, VecState{0, 0, 1}, -1)) == -1 &&
          sg(dd.cmp(z, 0, VecState{0, 0, -1}, 0)) == 1 && sg(dd.cmp(z, 0, VecState{0, -1, 1}, 0)) == 1 && sg(dd.cmp(z, 0, VecState{0, 0, -1}, 1)) == 1 &&
          sg(dd.cmp(z, 0, z, 1)) == 0, "cmp returns the first difference (coordinates)");
    CHECK(sg(dv.cmp(z, 0, VecState{0, 0, 1}, 0)) == -1 && sg(dv.cmp(z, 0, VecState{0, 1, -1}, 0)) == -1 && sg(dv.cmp(z, 0, VecState{0, 0, -1}, 1)) == -1 &&
          sg(dv.cmp(z, 0, VecState{0, 0, -1}, 0)) == 1 && sg(dv.cmp(z, 0, VecState{0, -1, 1}, 0)) == 1 && sg(dv.cmp(z, 0, VecState{0, 0, 1}, -1)) == 1 &&
          sg(dv.cmp(z, 0, z, 0)) == 0, "cmp returns the first difference (value first)");
}

int main() {
    test_flags();
    test_maxub_and_defaults();
    test_width();
    test_dominance();
    test_dominance_cmp();
    test_dummy_dd();
    test_locbounds();
    test_fringe();
    test_solvers_knapsack();
    test_solvers_knapsack7_and_primal();
    std::printf("%s: %d checks, %d failed\n", g_fail ? "SELFTEST FAILED" : "SELFTEST OK", g_checks, g_fail);
    return g_fail ? 1 : 0;
}

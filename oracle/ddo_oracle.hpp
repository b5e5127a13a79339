// ============================================================================
// ddo_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++17) of the reference algorithm for the hot path of
// xgillard/ddo (reference @ 3b39798): `Mdd::compile` and the solvers that
// drive it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg may build, load or call anything in oracle/.
//
// Parity status: PINNED.  The reference is Rust and cannot be compiled in this
// image (no cargo/rustc), so the oracle is pinned against the reference's own
// unit-test expectations (oracle/selftest.cpp restates clean.rs:1190-2398,
// node_flags.rs, no_duplicate.rs tests), the golden graphviz dumps in
// resources/visualisation_tests (tests/golden/*.dot) and the known optima of
// the DIMACS / knapsack / MAX2SAT / TSPTW instances asserted by examples/*/tests.rs
// (plus the MAX2SAT model unit vectors of examples/max2sat/model.rs:388-448).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/ddo/src unless stated).
//
// Canonical order rules (the reference leaves these to FxHashMap iteration
// order / unstable sorts).  Oracle and device engine
// share them so that a DD is a pure function of its CompilationInput:
//   C1. `next_l` is iterated / drained in node-creation order.
//   C2. the width-cut sort is stable w.r.t. C1 (matters only when the ranking
//       is not a strict total order).
//   C3. after a cut the surviving nodes are put back in node-creation order
//       (merged node, resp. the "saved" node of the recycled case, last).
//   C4. best terminal node = last maximum in C1 order (Rust max_by_key).
//   C5/C6 (MAX2SAT, models.hpp): stable variable order; ranking ties refined
//       by (depth, lexicographic benefits).
//   C7. a FRONTIER cutset is drained by (layer descending, position in the
//       layer ascending); the reference pushes its nodes in the order of the
//       bottom-up edge walk (clean.rs:586-606), which inherits the hash order.
// Everything else (e.g. the order in which relaxed edges are appended,
// clean.rs:851-866, and the `>=` "last tie wins" rule, clean.rs:215) is verbatim.
// ============================================================================
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace ddo_oracle {

using isize = int64_t;
constexpr isize ISIZE_MIN = std::numeric_limits<isize>::min();
constexpr isize ISIZE_MAX = std::numeric_limits<isize>::max();

// Rust isize::saturating_add / saturating_sub (used at clean.rs:208,364,426,427,466,504-511,526,747)
inline isize sat_add(isize a, isize b) {
    isize r;
    if (__builtin_add_overflow(a, b, &r)) return b > 0 ? ISIZE_MAX : ISIZE_MIN;
    return r;
}
inline isize sat_sub(isize a, isize b) {
    isize r;
    if (__builtin_sub_overflow(a, b, &r)) return b < 0 ? ISIZE_MAX : ISIZE_MIN;
    return r;
}

// ---------------------------------------------------------------------------
// common.rs:33-121
// ---------------------------------------------------------------------------
struct Variable { size_t id; };
struct Decision { size_t variable; isize value; };
inline bool operator==(const Decision& a, const Decision& b) { return a.variable == b.variable && a.value == b.value; }
using Solution = std::vector<Decision>;

template <class S>
struct SubProblem {  // common.rs:75-87
    std::shared_ptr<const S> state;
    isize value;
    std::vector<Decision> path;
    isize ub;
    size_t depth;
};
struct Threshold {  // common.rs:96 (derive Ord: value, then explored)
    isize value; bool explored;
    bool operator<(const Threshold& o) const { return value < o.value || (value == o.value && explored < o.explored); }
};
enum class Reason { CutoffOccurred };  // common.rs:108
struct Completion { bool is_exact; std::optional<isize> best_value; };  // common.rs:115

// ---------------------------------------------------------------------------
// abstraction/dp.rs:34-122
// ---------------------------------------------------------------------------
using DecisionCallback = std::function<void(Decision)>;
template <class S>
struct Problem {
    virtual ~Problem() = default;
    virtual size_t nb_variables() const = 0;
    virtual S initial_state() const = 0;
    virtual isize initial_value() const = 0;
    virtual S transition(const S& state, Decision d) const = 0;
    virtual isize transition_cost(const S& source, const S& dest, Decision d) const = 0;
    // next_layer is given in C1 (creation) order
    virtual std::optional<Variable> next_variable(size_t depth, const std::vector<const S*>& next_layer) const = 0;
    virtual void for_each_in_domain(Variable var, const S& state, const DecisionCallback& f) const = 0;
    virtual bool is_impacted_by(Variable, const S&) const { return true; }
};
template <class S>
struct Relaxation {
    virtual ~Relaxation() = default;
    virtual S merge(const std::vector<const S*>& states) const = 0;
    virtual isize relax(const S& source, const S& dest, const S& merged, Decision d, isize cost) const = 0;
    virtual isize fast_upper_bound(const S&) const { return ISIZE_MAX; }  // dp.rs:104-106
};
// abstraction/heuristics.rs:61-105.  compare: <0 Less, 0 Equal, >0 Greater
template <class S> struct StateRanking { virtual ~StateRanking() = default; virtual int compare(const S& a, const S& b) const = 0; };
template <class S> struct WidthHeuristic { virtual ~WidthHeuristic() = default; virtual size_t max_width(const SubProblem<S>&) const = 0; };
struct Cutoff { virtual ~Cutoff() = default; virtual bool must_stop() const = 0; };

// implementation/heuristics/width.rs:166-170,397-401,636-641,875-880
template <class S> struct FixedWidth : WidthHeuristic<S> {
    size_t w; explicit FixedWidth(size_t w_) : w(w_) {}
    size_t max_width(const SubProblem<S>&) const override { return w; }
};
template <class S> struct NbUnassignedWidth : WidthHeuristic<S> {
    size_t n; explicit NbUnassignedWidth(size_t n_) : n(n_) {}
    size_t max_width(const SubProblem<S>& x) const override { return n - x.path.size(); }
};
template <class S> struct Times : WidthHeuristic<S> {
    size_t k; const WidthHeuristic<S>* inner; Times(size_t k_, const WidthHeuristic<S>* i) : k(k_), inner(i) {}
    size_t max_width(const SubProblem<S>& x) const override { return std::max<size_t>(1, k * inner->max_width(x)); }
};
template <class S> struct DivBy : WidthHeuristic<S> {
    size_t k; const WidthHeuristic<S>* inner; DivBy(size_t k_, const WidthHeuristic<S>* i) : k(k_), inner(i) {}
    size_t max_width(const SubProblem<S>& x) const override { return std::max<size_t>(1, inner->max_width(x) / k); }
};
// implementation/heuristics/cutoff.rs:160-163,302-323
struct NoCutoff : Cutoff { bool must_stop() const override { return false; } };
struct TimeBudget : Cutoff {
    std::chrono::steady_clock::time_point deadline;
    explicit TimeBudget(double seconds)
        : deadline(std::chrono::steady_clock::now() + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(seconds))) {}
    bool must_stop() const override { return std::chrono::steady_clock::now() >= deadline; }
};
struct FlagCutoff : Cutoff {  // used by tests (CutoffAlways clean.rs:1318-1321)
    bool stop; explicit FlagCutoff(bool s) : stop(s) {}
    bool must_stop() const override { return stop; }
};

// ---------------------------------------------------------------------------
// abstraction/cache.rs:27-56 ; implementation/cache/{empty,simple}.rs
// ---------------------------------------------------------------------------
template <class S>
struct Cache {
    virtual ~Cache() = default;
    virtual bool must_explore(const SubProblem<S>& sp) const {  // cache.rs:32-39
        auto t = get_threshold(*sp.state, sp.depth);
        if (t) return sp.value > t->value || (sp.value == t->value && !t->explored);
        return true;
    }
    virtual void initialize(const Problem<S>&) = 0;
    virtual std::optional<Threshold> get_threshold(const S&, size_t depth) const = 0;
    virtual void update_threshold(std::shared_ptr<const S>, size_t depth, isize value, bool explored) = 0;
    virtual void clear_layer(size_t) = 0;
    virtual void clear() = 0;
};
template <class S>
struct EmptyCache : Cache<S> {  // cache/empty.rs:33-70
    bool must_explore(const SubProblem<S>&) const override { return true; }
    void initialize(const Problem<S>&) override {}
    std::optional<Threshold> get_threshold(const S&, size_t) const override { return std::nullopt; }
    void update_threshold(std::shared_ptr<const S>, size_t, isize, bool) override {}
    void clear_layer(size_t) override {}
    void clear() override {}
};
template <class S, class Hash, class Eq>
struct SimpleCache : Cache<S> {  // cache/simple.rs:36-74 (DashMap per depth; here one mutex per layer)
    struct PH { size_t operator()(const std::shared_ptr<const S>& p) const { return Hash()(*p); } };
    struct PE { bool operator()(const std::shared_ptr<const S>& a, const std::shared_ptr<const S>& b) const { return Eq()(*a, *b); } };
    struct LayerMap { std::unordered_map<std::shared_ptr<const S>, Threshold, PH, PE> m; mutable std::mutex mu; };
    std::vector<std::unique_ptr<LayerMap>> layers;
    void initialize(const Problem<S>& pb) override {
        for (size_t i = 0; i <= pb.nb_variables(); ++i) layers.emplace_back(new LayerMap());
    }
    std::optional<Threshold> get_threshold(const S& s, size_t depth) const override {
        auto& L = *layers[depth];
        std::lock_guard<std::mutex> g(L.mu);
        std::shared_ptr<const S> key(std::shared_ptr<const S>(), &s);  // aliasing, non-owning
        auto it = L.m.find(key);
        if (it == L.m.end()) return std::nullopt;
        return it->second;
    }
    void update_threshold(std::shared_ptr<const S> s, size_t depth, isize value, bool explored) override {  // simple.rs:62-66
        auto& L = *layers[depth];
        std::lock_guard<std::mutex> g(L.mu);
        Threshold t{value, explored};
        auto it = L.m.find(s);
        if (it == L.m.end()) L.m.emplace(std::move(s), t);
        else if (it->second < t) it->second = t;
    }
    void clear_layer(size_t d) override { std::lock_guard<std::mutex> g(layers[d]->mu); layers[d]->m.clear(); }
    void clear() override { for (auto& l : layers) { std::lock_guard<std::mutex> g(l->mu); l->m.clear(); } }
};

// ---------------------------------------------------------------------------
// abstraction/dominance.rs:37-126 ; implementation/dominance/{empty,simple}.rs
// ---------------------------------------------------------------------------
struct DominanceCmpResult { int ordering; bool only_val_diff; };
struct DominanceCheckResult { bool dominated; std::optional<isize> threshold; };
template <class S>
struct Dominance {
    virtual ~Dominance() = default;
    virtual std::optional<isize> get_key(const S&) const = 0;  // integer keys (knapsack); models with a structured key override get_key_bytes
    virtual std::optional<std::string> get_key_bytes(const S& s) const {  // Dominance::Key as an opaque byte string (dominance.rs:42-45)
        auto k = get_key(s);
        if (!k) return std::nullopt;
        return std::string(reinterpret_cast<const char*>(&*k), sizeof(isize));
    }
    virtual size_t nb_dimensions(const S&) const = 0;
    virtual isize get_coordinate(const S&, size_t i) const = 0;
    virtual bool use_value() const { return false; }
    static int cmp3(isize a, isize b) { return a < b ? -1 : (a > b ? 1 : 0); }
    std::optional<DominanceCmpResult> partial_cmp(const S& a, isize va, const S& b, isize vb) const {  // dominance.rs:57-79
        int ordering = 0;
        for (size_t i = 0; i < nb_dimensions(a); ++i) {
            int c = cmp3(get_coordinate(a, i), get_coordinate(b, i));
            if (ordering < 0 && c > 0) return std::nullopt;
            if (ordering > 0 && c < 0) return std::nullopt;
            if (ordering == 0 && c != 0) ordering = c;
        }
        if (use_value()) {
            int c = cmp3(va, vb);
            if (ordering < 0 && c > 0) return std::nullopt;
            if (ordering > 0 && c < 0) return std::nullopt;
            if (ordering == 0 && c > 0) return DominanceCmpResult{1, true};
            if (ordering == 0 && c < 0) return DominanceCmpResult{-1, true};
            return DominanceCmpResult{ordering, false};
        }
        return DominanceCmpResult{ordering, false};
    }
    int cmp(const S& a, isize va, const S& b, isize vb) const {  // dominance.rs:81-98
        if (use_value()) { int c = cmp3(va, vb); if (c) return c; }
        for (size_t i = 0; i < nb_dimensions(a); ++i) { int c = cmp3(get_coordinate(a, i), get_coordinate(b, i)); if (c) return c; }
        return 0;
    }
};
template <class S>
struct DominanceChecker {
    virtual ~DominanceChecker() = default;
    virtual void clear_layer(size_t) = 0;
    virtual DominanceCheckResult is_dominated_or_insert(std::shared_ptr<const S>, size_t depth, isize value) = 0;
    virtual int cmp(const S& a, isize va, const S& b, isize vb) const = 0;
};
template <class S>
struct EmptyDominanceChecker : DominanceChecker<S> {  // dominance/empty.rs:24-46
    void clear_layer(size_t) override {}
    DominanceCheckResult is_dominated_or_insert(std::shared_ptr<const S>, size_t, isize) override { return {false, std::nullopt}; }
    int cmp(const S&, isize, const S&, isize) const override { return 0; }
};
template <class S>
struct SimpleDominanceChecker : DominanceChecker<S> {  // dominance/simple.rs:37-116
    struct Entry { std::shared_ptr<const S> state; isize value; };
    struct LayerMap { std::unordered_map<std::string, std::vector<Entry>> m; std::mutex mu; };
    const Dominance<S>* dom;
    std::vector<std::unique_ptr<LayerMap>> data;
    SimpleDominanceChecker(const Dominance<S>* d, size_t nb_variables) : dom(d) {
        for (size_t i = 0; i <= nb_variables; ++i) data.emplace_back(new LayerMap());
    }
    void clear_layer(size_t d) override { std::lock_guard<std::mutex> g(data[d]->mu); data[d]->m.clear(); }
    DominanceCheckResult is_dominated_or_insert(std::shared_ptr<const S> state, size_t depth, isize value) override {  // simple.rs:71-111
        auto key = dom->get_key_bytes(*state);
        if (!key) return {false, std::nullopt};
        auto& L = *data[depth];
        std::lock_guard<std::mutex> g(L.mu);
        auto it = L.m.find(*key);
        if (it == L.m.end()) { L.m[*key].push_back({state, value}); return {false, std::nullopt}; }
        bool dominated = false;
        std::optional<isize> threshold = ISIZE_MAX;
        auto& vec = it->second;
        std::vector<Entry> kept;
        for (auto& other : vec) {
            auto c = dom->partial_cmp(*state, value, *other.state, other.value);
            bool keep = true;
            if (c) {
                if (c->ordering < 0) {
                    dominated = true;
                    if (dom->use_value()) {
                        isize cand = c->only_val_diff ? sat_sub(other.value, 1) : other.value;
                        threshold = std::min(*threshold, cand);
                    }
                } else keep = false;  // Equal or Greater: drop the old entry
            }
            if (keep) kept.push_back(other);
        }
        vec.swap(kept);
        if (!dominated) { threshold = std::nullopt; vec.push_back({state, value}); }
        return {dominated, threshold};
    }
    int cmp(const S& a, isize va, const S& b, isize vb) const override { return dom->cmp(a, va, b, vb); }
};

// ---------------------------------------------------------------------------
// implementation/mdd/node_flags.rs:48-185
// ---------------------------------------------------------------------------
struct NodeFlags {
    uint8_t bits;
    static constexpr uint8_t F_EXACT = 1, F_RELAXED = 2, F_MARKED = 4, F_CUTSET = 8, F_DELETED = 16, F_CACHE = 32, F_ABOVE_CUTSET = 64;
    static NodeFlags new_exact() { return NodeFlags{F_EXACT}; }
    static NodeFlags new_relaxed() { return NodeFlags{F_RELAXED}; }
    bool test(uint8_t m) const { return (bits & m) == m; }
    void set(uint8_t f, bool v) { if (v) bits |= f; else bits &= (uint8_t)~f; }
    void add(uint8_t f) { bits |= f; }
    bool is_exact() const { return test(F_EXACT) && !test(F_RELAXED); }  // node_flags.rs:88
    bool is_relaxed() const { return test(F_RELAXED); }
    bool is_marked() const { return test(F_MARKED); }
    bool is_cutset() const { return test(F_CUTSET); }
    bool is_above_cutset() const { return test(F_ABOVE_CUTSET); }
    bool is_deleted() const { return test(F_DELETED); }
    bool is_pruned_by_cache() const { return test(F_CACHE); }
    void set_exact(bool v) { set(F_EXACT, v); }
    void set_relaxed(bool v) { set(F_RELAXED, v); }
    void set_marked(bool v) { set(F_MARKED, v); }
    void set_cutset(bool v) { set(F_CUTSET, v); }
    void set_above_cutset(bool v) { set(F_ABOVE_CUTSET, v); }
    void set_deleted(bool v) { set(F_DELETED, v); }
    void set_pruned_by_cache(bool v) { set(F_CACHE, v); }
};

// ---------------------------------------------------------------------------
// abstraction/mdd.rs:24-114
// ---------------------------------------------------------------------------
enum class CompilationType { Exact, Relaxed, Restricted };
constexpr int LAST_EXACT_LAYER = 1;
constexpr int FRONTIER = 2;

template <class S>
struct CompilationInput {  // mdd.rs:51-71
    CompilationType comp_type;
    const Problem<S>* problem;
    const Relaxation<S>* relaxation;
    const StateRanking<S>* ranking;
    const Cutoff* cutoff;
    size_t max_width;
    const SubProblem<S>* residual;
    isize best_lb;
    Cache<S>* cache;
    DominanceChecker<S>* dominance;
};

// ---------------------------------------------------------------------------
// implementation/mdd/clean.rs -- Mdd<T, CUTSET_TYPE> (DefaultMDD)
// ---------------------------------------------------------------------------
template <class S, class Hash, class Eq>
class Mdd {
public:
    using NodeId = size_t;
    struct Node {  // clean.rs:37-69
        std::shared_ptr<const S> state;
        isize value_top, value_bot;
        int64_t best;     // edge id, -1 = None
        size_t inbound;   // edgelist id
        isize rub;
        std::optional<isize> theta;
        NodeFlags flags;
        size_t depth;
        size_t pos = 0;   // position in its layer after the cut (canonical order C1/C3); not a field of the reference's Node
    };
    struct Edge { NodeId from, to; Decision decision; isize cost; };  // clean.rs:74-85
    struct EdgesList { bool cons; size_t head, tail; };                // clean.rs:89-92
    struct Layer { size_t from, to; };                                 // clean.rs:96-99

    explicit Mdd(int cutset_type = LAST_EXACT_LAYER) : cutset_type_(cutset_type) {}

    // ---- DecisionDiagram trait (mdd.rs:75-114 / clean.rs:231-266) ----
    // returns false on Err(Reason::CutoffOccurred)
    bool compile(const CompilationInput<S>& input, Completion* out) { return _compile(input, out); }
    bool is_exact() const { return is_exact_ || has_exact_best_path_; }  // clean.rs:241-243
    std::optional<isize> best_value() const { if (best_node_ < 0) return std::nullopt; return nodes[best_node_].value_top; }
    std::optional<Solution> best_solution() const { if (best_node_ < 0) return std::nullopt; return _best_path(best_node_); }
    std::optional<isize> best_exact_value() const { if (best_exact_node_ < 0) return std::nullopt; return nodes[best_exact_node_].value_top; }
    std::optional<Solution> best_exact_solution() const { if (best_exact_node_ < 0) return std::nullopt; return _best_path(best_exact_node_); }

    template <class F>
    void drain_cutset(F func) {  // clean.rs:417-445
        auto bv = best_value();
        if (cutset_type_ == FRONTIER)
            // C7: the reference pushes frontier nodes in the order its bottom-up edge walk meets them (clean.rs:586-606), which inherits the
            // hash order of the layers; canonically they are drained by (layer descending, position in the layer ascending)
            std::stable_sort(cutset.begin(), cutset.end(), [&](NodeId a, NodeId b) {
                if (nodes[a].depth != nodes[b].depth) return nodes[a].depth > nodes[b].depth;
                return nodes[a].pos < nodes[b].pos;
            });
        if (bv) {
            for (NodeId id : cutset) {
                const Node& node = nodes[id];
                if (node.flags.is_marked()) {
                    isize rub = sat_add(node.value_top, node.rub);
                    isize locb = sat_add(node.value_top, node.value_bot);
                    isize ub = std::min(std::min(rub, locb), *bv);
                    func(SubProblem<S>{node.state, node.value_top, _best_path(id), ub, node.depth});
                }
            }
        }
        cutset.clear();
    }

    // ---- introspection for tests / parity checks ----
    std::vector<Layer> layers;
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    std::vector<EdgesList> edgelists;
    std::vector<NodeId> cutset;
    std::vector<NodeId> terminal_nodes() const { return next_order_; }
    std::optional<size_t> lel() const { return lel_; }
    int64_t best_node() const { return best_node_; }
    uint64_t expanded = 0;           // nodes passing the rub test at clean.rs:365 (metric definition)
    uint64_t transitions = 0;        // calls of _branch_on
    std::vector<size_t> layer_vars;  // variable chosen for each expanded layer
    std::vector<size_t> layer_widths;  // |curr_l| after filters and squash

    template <class F> void foreach_edge_of(NodeId id, F action) const {  // clean.rs:187-196 (newest first)
        size_t list = nodes[id].inbound;
        while (edgelists[list].cons) {
            Edge e = edges[edgelists[list].head];
            size_t tail = edgelists[list].tail;
            action(e);
            list = tail;
        }
    }

private:
    int cutset_type_;
    std::vector<NodeId> prev_l_;
    struct PH { size_t operator()(const S* p) const { return Hash()(*p); } };
    struct PE { bool operator()(const S* a, const S* b) const { return Eq()(*a, *b); } };
    std::unordered_map<const S*, NodeId, PH, PE> next_l_;
    std::vector<NodeId> next_order_;  // C1: creation order of next_l
    size_t curr_depth_ = 0;
    std::vector<Decision> path_to_root_;
    std::optional<size_t> lel_;
    int64_t best_node_ = -1, best_exact_node_ = -1;
    bool is_exact_ = true, has_exact_best_path_ = false;

    void append_edge_to(const Edge& edge) {  // clean.rs:199-220
        size_t new_eid = edges.size();
        size_t lst_id = edgelists.size();
        edges.push_back(edge);
        edgelists.push_back(EdgesList{true, new_eid, nodes[edge.to].inbound});
        const Node& parent = nodes[edge.from];
        bool parent_exact = parent.flags.is_exact();
        isize value = sat_add(parent.value_top, edge.cost);
        Node& node = nodes[edge.to];
        bool exact = parent_exact & node.flags.is_exact();
        node.flags.set_exact(exact);
        node.inbound = lst_id;
        if (value >= node.value_top) { node.best = (int64_t)new_eid; node.value_top = value; }
    }

    void _clear() {  // clean.rs:293-307
        layers.clear(); nodes.clear(); edges.clear(); edgelists.clear();
        prev_l_.clear(); next_l_.clear(); next_order_.clear(); path_to_root_.clear(); cutset.clear();
        lel_.reset(); best_node_ = -1; best_exact_node_ = -1; is_exact_ = true; has_exact_best_path_ = false;
        expanded = 0; transitions = 0; layer_vars.clear(); layer_widths.clear();
    }

    Solution _best_path(NodeId id) const {  // clean.rs:329-343
        Solution sol = path_to_root_;
        int64_t eid = nodes[id].best;
        while (eid >= 0) {
            const Edge& e = edges[eid];
            sol.push_back(e.decision);
            eid = nodes[e.from].best;
        }
        return sol;
    }

    void _initialize(const CompilationInput<S>& input) {  // clean.rs:383-405
        path_to_root_ = input.residual->path;
        edgelists.push_back(EdgesList{false, 0, 0});
        Node root{input.residual->state, input.residual->value, ISIZE_MIN, -1, 0, ISIZE_MAX, std::nullopt, NodeFlags::new_exact(), input.residual->depth};
        nodes.push_back(root);
        next_l_.emplace(nodes[0].state.get(), 0);
        next_order_.push_back(0);
        edgelists.push_back(EdgesList{false, 0, 0});
        curr_depth_ = input.residual->depth;
    }

    bool _compile(const CompilationInput<S>& input, Completion* out) {  // clean.rs:345-381
        _clear();
        _initialize(input);
        std::vector<NodeId> curr_l;
        std::vector<const S*> keys;
        for (;;) {
            keys.clear();
            for (NodeId id : next_order_) keys.push_back(nodes[id].state.get());
            auto var = input.problem->next_variable(curr_depth_, keys);
            if (!var) break;
            if (input.cutoff->must_stop()) return false;  // clean.rs:352-354
            if (!_move_to_next_layer(input, curr_l)) break;
            layer_vars.push_back(var->id);
            layer_widths.push_back(curr_l.size());
            for (NodeId node_id : curr_l) {  // clean.rs:360-370
                std::shared_ptr<const S> state = nodes[node_id].state;
                isize rub = input.relaxation->fast_upper_bound(*state);
                nodes[node_id].rub = rub;
                isize ub = sat_add(rub, nodes[node_id].value_top);
                if (ub > input.best_lb) {
                    ++expanded;
                    input.problem->for_each_in_domain(*var, *state, [&](Decision d) { _branch_on(node_id, d, *input.problem); });
                }
            }
            curr_depth_ += 1;
        }
        _finalize(input);
        out->is_exact = is_exact();
        out->best_value = best_value();
        return true;
    }

    void _finalize(const CompilationInput<S>& input) {  // clean.rs:407-414
        _finalize_layers();
        _find_best_node();
        _finalize_exact(input);
        _finalize_cutset(input);
        _compute_local_bounds(input);
        _compute_thresholds(input);
    }

    void _compute_local_bounds(const CompilationInput<S>& input) {  // clean.rs:448-475
        if (*lel_ < layers.size() && input.comp_type == CompilationType::Relaxed) {
            Layer last = layers.back();
            for (size_t i = last.from; i < last.to; ++i) { nodes[i].value_bot = 0; nodes[i].flags.set_marked(true); }
            for (size_t li = layers.size(); li-- > 0;) {
                Layer L = layers[li];
                for (size_t id = L.from; id < L.to; ++id) {
                    isize value = nodes[id].value_bot;
                    if (nodes[id].flags.is_marked()) {
                        foreach_edge_of(id, [&](const Edge& edge) {
                            isize using_edge = sat_add(value, edge.cost);
                            Node& parent = nodes[edge.from];
                            parent.flags.set_marked(true);
                            parent.value_bot = std::max(parent.value_bot, using_edge);
                        });
                    }
                }
            }
        }
    }

    void _compute_thresholds(const CompilationInput<S>& input) {  // clean.rs:478-532
        if (input.comp_type == CompilationType::Relaxed || is_exact_) {
            isize best_known = input.best_lb;
            if (best_exact_node_ >= 0) {
                isize bev = nodes[best_exact_node_].value_top;
                best_known = std::max(best_known, bev);
                for (NodeId id : next_order_) {
                    if ((cutset_type_ == LAST_EXACT_LAYER && is_exact_) || (cutset_type_ == FRONTIER && nodes[id].flags.is_exact()))
                        nodes[id].theta = best_known;
                }
            }
            for (size_t li = layers.size(); li-- > 0;) {
                Layer L = layers[li];
                for (size_t id = L.from; id < L.to; ++id) {
                    Node& node = nodes[id];
                    if (node.flags.is_deleted()) continue;
                    if (!node.flags.is_pruned_by_cache()) {
                        isize tot_rub = sat_add(node.value_top, node.rub);
                        if (tot_rub <= best_known) {
                            node.theta = sat_sub(best_known, node.rub);
                        } else if (node.flags.is_cutset()) {
                            isize tot_locb = sat_add(node.value_top, node.value_bot);
                            if (tot_locb <= best_known) {
                                isize theta = node.theta.value_or(ISIZE_MAX);
                                node.theta = std::min(theta, sat_sub(best_known, node.value_bot));
                            } else {
                                node.theta = node.value_top;
                            }
                        } else if (node.flags.is_exact() && !node.theta) {
                            node.theta = ISIZE_MAX;
                        }
                        _maybe_update_cache(node, input);
                    }
                    if (node.theta) {
                        isize my_theta = *node.theta;
                        foreach_edge_of(id, [&](const Edge& edge) {
                            Node& parent = nodes[edge.from];
                            isize theta = parent.theta.value_or(ISIZE_MAX);
                            parent.theta = std::min(theta, sat_sub(my_theta, edge.cost));
                        });
                    }
                }
            }
        }
    }

    void _maybe_update_cache(const Node& node, const CompilationInput<S>& input) {  // clean.rs:534-545
        if (node.theta && node.flags.is_above_cutset())
            input.cache->update_threshold(node.state, node.depth, *node.theta, !node.flags.is_cutset());
    }

    void _finalize_cutset(const CompilationInput<S>& input) {  // clean.rs:547-564
        if (!lel_) lel_ = layers.size();
        if (input.comp_type == CompilationType::Relaxed || is_exact_) {
            if (cutset_type_ == LAST_EXACT_LAYER) _compute_last_exact_layer_cutset(*lel_);
            else _compute_frontier_cutset();
        }
    }
    void _compute_last_exact_layer_cutset(size_t lel) {  // clean.rs:566-583
        if (lel < layers.size()) {
            Layer L = layers[lel];
            for (size_t id = L.from; id < L.to; ++id) {
                cutset.push_back(id);
                nodes[id].flags.add(NodeFlags::F_CUTSET | NodeFlags::F_ABOVE_CUTSET);
            }
        }
        for (size_t li = std::min(lel, layers.size()); li-- > 0;) {
            Layer L = layers[li];
            for (size_t id = L.from; id < L.to; ++id) nodes[id].flags.set_above_cutset(true);
        }
    }
    void _compute_frontier_cutset() {  // clean.rs:586-606
        for (size_t li = layers.size(); li-- > 0;) {
            Layer L = layers[li];
            for (size_t id = L.from; id < L.to; ++id) {
                if (nodes[id].flags.is_exact()) {
                    nodes[id].flags.set_above_cutset(true);
                } else {
                    foreach_edge_of(id, [&](const Edge& edge) {
                        Node& parent = nodes[edge.from];
                        if (parent.flags.is_exact() && !parent.flags.is_cutset()) {
                            cutset.push_back(edge.from);
                            parent.flags.set_cutset(true);
                        }
                    });
                }
            }
        }
    }

    void _finalize_layers() {  // clean.rs:608-618
        if (!next_order_.empty()) {
            if (layers.empty()) layers.push_back(Layer{0, nodes.size()});
            else layers.push_back(Layer{layers.back().to, nodes.size()});
        }
    }
    void _find_best_node() {  // clean.rs:620-632 (Rust max_by_key keeps the LAST maximum; C5)
        best_node_ = -1; best_exact_node_ = -1;
        for (NodeId id : next_order_) {
            if (best_node_ < 0 || nodes[id].value_top >= nodes[best_node_].value_top) best_node_ = (int64_t)id;
            if (nodes[id].flags.is_exact())
                if (best_exact_node_ < 0 || nodes[id].value_top >= nodes[best_exact_node_].value_top) best_exact_node_ = (int64_t)id;
        }
    }
    void _finalize_exact(const CompilationInput<S>& input) {  // clean.rs:634-641
        is_exact_ = !lel_.has_value();
        has_exact_best_path_ = input.comp_type == CompilationType::Relaxed && _has_exact_best_path(best_node_);
        if (has_exact_best_path_) best_exact_node_ = best_node_;
    }
    bool _has_exact_best_path(int64_t node) const {  // clean.rs:643-655
        while (node >= 0) {
            const Node& n = nodes[node];
            if (n.flags.is_exact()) return true;
            if (n.flags.is_relaxed()) return false;
            node = n.best >= 0 ? (int64_t)edges[n.best].from : -1;
        }
        return true;
    }

    bool _move_to_next_layer(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:657-687
        prev_l_.clear();
        prev_l_.swap(curr_l);
        curr_l = next_order_;  // C1: creation order instead of hash order
        next_order_.clear();
        next_l_.clear();
        if (curr_l.empty()) { layers.push_back(Layer{0, 0}); return false; }
        if (!layers.empty()) _filter_with_cache(input, curr_l);
        _filter_with_dominance(input, curr_l);
        _squash_if_needed(input, curr_l);
        for (size_t i = 0; i < curr_l.size(); ++i) nodes[curr_l[i]].pos = i;
        if (layers.empty()) layers.push_back(Layer{0, nodes.size()});
        else layers.push_back(Layer{layers.back().to, nodes.size()});
        return true;
    }

    void _filter_with_dominance(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:689-708
        std::stable_sort(curr_l.begin(), curr_l.end(), [&](NodeId a, NodeId b) {
            return input.dominance->cmp(*nodes[a].state, nodes[a].value_top, *nodes[b].state, nodes[b].value_top) > 0;  // reversed
        });
        std::vector<NodeId> kept;
        for (NodeId id : curr_l) {
            Node& node = nodes[id];
            if (node.flags.is_exact()) {
                auto r = input.dominance->is_dominated_or_insert(node.state, node.depth, node.value_top);
                if (r.dominated) { node.theta = r.threshold; continue; }
            }
            kept.push_back(id);
        }
        curr_l.swap(kept);
    }
    void _filter_with_cache(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:710-726
        std::vector<NodeId> kept;
        for (NodeId id : curr_l) {
            Node& node = nodes[id];
            auto t = input.cache->get_threshold(*node.state, node.depth);
            if (t && !(node.value_top > t->value)) {
                node.flags.set_pruned_by_cache(true);
                node.theta = t->value;
                continue;
            }
            kept.push_back(id);
        }
        curr_l.swap(kept);
    }

    void _branch_on(NodeId from_id, Decision decision, const Problem<S>& problem) {  // clean.rs:728-776
        ++transitions;
        std::shared_ptr<const S> src = nodes[from_id].state;
        auto next_state = std::make_shared<const S>(problem.transition(*src, decision));
        isize cost = problem.transition_cost(*src, *next_state, decision);
        auto it = next_l_.find(next_state.get());
        if (it == next_l_.end()) {
            const Node& parent = nodes[from_id];
            NodeId node_id = nodes.size();
            NodeFlags flags = NodeFlags::new_exact();
            flags.set_exact(parent.flags.is_exact());
            Node n{next_state, sat_add(parent.value_top, cost), ISIZE_MIN, -1, 0, ISIZE_MAX, std::nullopt, flags, parent.depth + 1};
            // NOTE (clean.rs:747 then :215): value_top is set here and then re-derived by append_edge_to with `>=`.
            nodes.push_back(std::move(n));
            append_edge_to(Edge{from_id, node_id, decision, cost});
            next_l_.emplace(nodes[node_id].state.get(), node_id);
            next_order_.push_back(node_id);
        } else {
            append_edge_to(Edge{from_id, it->second, decision, cost});
        }
    }

    void _squash_if_needed(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:779-795
        switch (input.comp_type) {
            case CompilationType::Exact: break;
            case CompilationType::Restricted:
                if (curr_l.size() > input.max_width) { _maybe_save_lel(); _restrict(input, curr_l); }
                break;
            case CompilationType::Relaxed:
                if (curr_l.size() > input.max_width && layers.size() > 1) { _maybe_save_lel(); _relax(input, curr_l); }
                break;
        }
    }
    void _maybe_save_lel() { if (!lel_) lel_ = layers.size() - 1; }  // clean.rs:796-800

    void _sort_for_cut(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:803-808 / 819-824 (C2: stable)
        std::stable_sort(curr_l.begin(), curr_l.end(), [&](NodeId a, NodeId b) {
            const Node& na = nodes[a]; const Node& nb = nodes[b];
            if (na.value_top != nb.value_top) return na.value_top > nb.value_top;
            return input.ranking->compare(*na.state, *nb.state) > 0;
        });
    }

    void _restrict(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:802-815
        _sort_for_cut(input, curr_l);
        for (size_t i = input.max_width; i < curr_l.size(); ++i) nodes[curr_l[i]].flags.set_deleted(true);
        curr_l.resize(input.max_width);
        std::sort(curr_l.begin(), curr_l.end());  // C3
    }

    void _relax(const CompilationInput<S>& input, std::vector<NodeId>& curr_l) {  // clean.rs:818-876
        _sort_for_cut(input, curr_l);
        size_t nkeep = input.max_width - 1;  // panics in the reference when max_width == 0 (clean.rs:827)
        std::vector<const S*> to_merge;
        for (size_t i = nkeep; i < curr_l.size(); ++i) to_merge.push_back(nodes[curr_l[i]].state.get());
        auto merged = std::make_shared<const S>(input.relaxation->merge(to_merge));

        int64_t recycled = -1;  // clean.rs:830
        for (size_t i = 0; i < nkeep; ++i)
            if (Eq()(*nodes[curr_l[i]].state, *merged)) { recycled = (int64_t)curr_l[i]; break; }

        NodeId merged_id;
        if (recycled >= 0) merged_id = (NodeId)recycled;
        else {
            merged_id = nodes.size();
            nodes.push_back(Node{merged, ISIZE_MIN, ISIZE_MIN, -1, 0, ISIZE_MAX, std::nullopt, NodeFlags::new_relaxed(), nodes[curr_l[nkeep]].depth});
        }
        nodes[merged_id].flags.set_relaxed(true);

        for (size_t i = nkeep; i < curr_l.size(); ++i) {  // clean.rs:851-866 (sorted order; inbound edges newest first)
            NodeId drop_id = curr_l[i];
            nodes[drop_id].flags.set_deleted(true);
            size_t list = nodes[drop_id].inbound;
            while (edgelists[list].cons) {
                Edge edge = edges[edgelists[list].head];
                size_t tail = edgelists[list].tail;
                isize rcost = input.relaxation->relax(*nodes[edge.from].state, *nodes[edge.to].state, *merged, edge.decision, edge.cost);
                append_edge_to(Edge{edge.from, merged_id, edge.decision, rcost});
                list = tail;
            }
        }

        if (recycled >= 0) {  // clean.rs:868-871
            curr_l.resize(input.max_width);
            NodeId saved_id = curr_l[input.max_width - 1];
            nodes[saved_id].flags.set_deleted(false);
            std::sort(curr_l.begin(), curr_l.end() - 1);  // C3 (saved node last)
        } else {  // clean.rs:872-875
            curr_l.resize(nkeep);
            std::sort(curr_l.begin(), curr_l.end());  // C3
            curr_l.push_back(merged_id);
        }
    }
};

// ---------------------------------------------------------------------------
// heuristics/subproblem_ranking.rs:76-92 (MaxUB) + fringe/no_duplicate.rs:52-323
// ---------------------------------------------------------------------------
template <class S>
struct MaxUB {
    const StateRanking<S>* ranking;
    int compare(const SubProblem<S>& l, const SubProblem<S>& r) const {  // subproblem_ranking.rs:86-90
        if (l.ub != r.ub) return l.ub < r.ub ? -1 : 1;
        if (l.value != r.value) return l.value < r.value ? -1 : 1;
        return ranking->compare(*l.state, *r.state);
    }
};

template <class S>
struct Fringe {
    virtual ~Fringe() = default;
    virtual void push(SubProblem<S> node) = 0;
    virtual std::optional<SubProblem<S>> pop() = 0;
    virtual void clear() = 0;
    virtual size_t len() const = 0;
    bool is_empty() const { return len() == 0; }
};

template <class S, class Hash, class Eq>
class NoDupFringe : public Fringe<S> {  // no_duplicate.rs:52-323
    MaxUB<S> cmp_;
    struct PH { size_t operator()(const S* p) const { return Hash()(*p); } };
    struct PE { bool operator()(const S* a, const S* b) const { return Eq()(*a, *b); } };
    std::unordered_map<const S*, size_t, PH, PE> states_;
    std::vector<SubProblem<S>> nodes_;
    std::vector<size_t> pos_, heap_, recycle_bin_;
public:
    explicit NoDupFringe(MaxUB<S> c) : cmp_(c) {}
    void push(SubProblem<S> node) override {  // no_duplicate.rs:88-140
        auto it = states_.find(node.state.get());
        if (it != states_.end()) {
            size_t id = it->second;
            isize old_lp = nodes_[id].value, old_ub = nodes_[id].ub;
            isize new_lp = node.value, new_ub = node.ub;
            node.ub = std::max(new_ub, old_ub);
            bool up = cmp_.compare(node, nodes_[id]) > 0;
            if (new_lp > old_lp) {
                // the map key points into the stored node's state: re-key
                states_.erase(it);
                nodes_[id] = std::move(node);
                states_.emplace(nodes_[id].state.get(), id);
            }
            if (new_ub > old_ub) nodes_[id].ub = new_ub;
            if (up) bubble_up(id);
        } else {
            size_t id;
            if (recycle_bin_.empty()) { id = nodes_.size(); nodes_.push_back(std::move(node)); pos_.push_back(0); }
            else { id = recycle_bin_.back(); recycle_bin_.pop_back(); nodes_[id] = std::move(node); }
            heap_.push_back(id);
            pos_[id] = heap_.size() - 1;
            states_.emplace(nodes_[id].state.get(), id);
            bubble_up(id);
        }
    }
    std::optional<SubProblem<S>> pop() override {  // no_duplicate.rs:144-164
        if (heap_.empty()) return std::nullopt;
        size_t id = heap_[0];
        heap_[0] = heap_.back(); heap_.pop_back();  // swap_remove(0)
        if (!heap_.empty()) { pos_[heap_[0]] = 0; bubble_down(heap_[0]); }
        recycle_bin_.push_back(id);
        states_.erase(nodes_[id].state.get());
        SubProblem<S> node = nodes_[id];
        return node;
    }
    void clear() override { states_.clear(); nodes_.clear(); pos_.clear(); heap_.clear(); recycle_bin_.clear(); }
    size_t len() const override { return heap_.size(); }
private:
    int compare_at_pos(size_t x, size_t y) const { return cmp_.compare(nodes_[heap_[x]], nodes_[heap_[y]]); }
    static size_t parent(size_t pos) { return pos == 0 ? 0 : (pos % 2 == 1 ? pos / 2 : pos / 2 - 1); }  // no_duplicate.rs:262-270
    void bubble_up(size_t id) {  // :227-242
        size_t me = pos_[id], par = parent(me);
        while (me != 0 && compare_at_pos(me, par) > 0) {
            size_t p_id = heap_[par];
            pos_[p_id] = me; pos_[id] = par; heap_[me] = p_id; heap_[par] = id;
            me = par; par = parent(me);
        }
    }
    size_t max_child_of(size_t pos) const {  // :279-295
        size_t size = heap_.size(), left = pos * 2 + 1, right = pos * 2 + 2;
        if (left >= size) return 0;
        if (right >= size) return left;
        return compare_at_pos(left, right) > 0 ? left : right;
    }
    void bubble_down(size_t id) {  // :244-259
        size_t me = pos_[id], kid = max_child_of(me);
        while (kid > 0 && compare_at_pos(me, kid) < 0) {
            size_t k_id = heap_[kid];
            pos_[k_id] = me; pos_[id] = kid; heap_[me] = k_id; heap_[kid] = id;
            me = kid; kid = max_child_of(me);
        }
    }
};

template <class S>
class SimpleFringe : public Fringe<S> {  // fringe/simple.rs:35-63 (binary heap on MaxUB)
    MaxUB<S> cmp_;
    std::vector<SubProblem<S>> heap_;
    struct Less { const MaxUB<S>* c; bool operator()(const SubProblem<S>& a, const SubProblem<S>& b) const { return c->compare(a, b) < 0; } };
public:
    explicit SimpleFringe(MaxUB<S> c) : cmp_(c) {}
    void push(SubProblem<S> n) override { heap_.push_back(std::move(n)); std::push_heap(heap_.begin(), heap_.end(), Less{&cmp_}); }
    std::optional<SubProblem<S>> pop() override {
        if (heap_.empty()) return std::nullopt;
        std::pop_heap(heap_.begin(), heap_.end(), Less{&cmp_});
        SubProblem<S> n = std::move(heap_.back()); heap_.pop_back(); return n;
    }
    void clear() override { heap_.clear(); }
    size_t len() const override { return heap_.size(); }
};

// ---------------------------------------------------------------------------
// Solvers.  Statistics common to all of them.
// ---------------------------------------------------------------------------
struct SolverStats {
    uint64_t explored = 0;     // Solver::explored(), solver.rs:96
    uint64_t expanded = 0;     // metric counter
    uint64_t transitions = 0;
    uint64_t compilations = 0;
    uint64_t waves = 0;
};

template <class S>
struct SolverConfig {
    const Problem<S>* problem;
    const Relaxation<S>* relaxation;
    const StateRanking<S>* ranking;
    const WidthHeuristic<S>* width;
    DominanceChecker<S>* dominance;
    const Cutoff* cutoff;
    Fringe<S>* fringe;
    Cache<S>* cache;
    int cutset_type = LAST_EXACT_LAYER;
};

// implementation/solver/sequential.rs:202-526
template <class S, class Hash, class Eq>
class SequentialSolver {
public:
    explicit SequentialSolver(SolverConfig<S> c) : c_(c), mdd_(c.cutset_type), open_by_layer_(c.problem->nb_variables() + 1, 0) {}
    SolverStats stats;
    isize best_lb = ISIZE_MIN, best_ub = ISIZE_MAX;
    std::optional<Solution> best_sol;
    bool aborted = false;

    void set_primal(isize value, Solution sol) { if (value > best_lb) { best_sol = std::move(sol); best_lb = value; } }  // :516-521

    Completion maximize() {  // sequential.rs:475-494
        initialize();
        for (;;) {
            // get_workload, sequential.rs:433-461
            while (first_active_layer_ < c_.problem->nb_variables() && open_by_layer_[first_active_layer_] == 0) {
                c_.cache->clear_layer(first_active_layer_);
                first_active_layer_ += 1;
            }
            if (c_.fringe->is_empty()) { best_ub = best_lb; break; }
            if (aborted) break;
            SubProblem<S> nn = *c_.fringe->pop();
            stats.explored += 1;
            open_by_layer_[nn.depth] -= 1;
            best_ub = nn.ub;
            if (!process_one_node(nn)) { abort_search(); break; }
        }
        if (best_sol) std::stable_sort(best_sol->begin(), best_sol->end(), [](const Decision& a, const Decision& b) { return a.variable < b.variable; });
        return Completion{!aborted, best_sol ? std::optional<isize>(best_lb) : std::nullopt};
    }
    double gap() const {  // abstraction/solver.rs:80-93: 1 while a bound is missing, else (u - l) / u over the absolute bounds, in f32
        if (best_ub == ISIZE_MAX || best_lb == ISIZE_MIN) return 1.0;
        isize aub = best_ub < 0 ? -best_ub : best_ub, alb = best_lb < 0 ? -best_lb : best_lb;
        isize u = std::max(aub, alb), l = std::min(aub, alb);
        return (double)((float)(u - l) / (float)u);  // 0 / 0 is NaN, as in the reference
    }
private:
    SolverConfig<S> c_;
    Mdd<S, Hash, Eq> mdd_;
    std::vector<size_t> open_by_layer_;
    size_t first_active_layer_ = 0;

    void initialize() {  // sequential.rs:308-313
        SubProblem<S> root{std::make_shared<const S>(c_.problem->initial_state()), c_.problem->initial_value(), {}, ISIZE_MAX, 0};
        c_.cache->initialize(*c_.problem);
        c_.fringe->push(root);
        open_by_layer_[0] += 1;
    }
    bool process_one_node(const SubProblem<S>& node) {  // sequential.rs:329-389
        isize node_ub = node.ub;
        isize lb = best_lb;
        if (node_ub <= lb) return true;
        if (!c_.cache->must_explore(node)) return true;
        size_t width = c_.width->max_width(node);
        CompilationInput<S> in{CompilationType::Restricted, c_.problem, c_.relaxation, c_.ranking, c_.cutoff, width, &node, lb, c_.cache, c_.dominance};
        Completion comp;
        if (!mdd_.compile(in, &comp)) return false;
        account();
        maybe_update_best();
        if (comp.is_exact) return true;
        in.comp_type = CompilationType::Relaxed;
        in.best_lb = best_lb;
        if (!mdd_.compile(in, &comp)) return false;
        account();
        maybe_update_best();
        if (!comp.is_exact) enqueue_cutset(node_ub);
        return true;
    }
    void account() { stats.compilations++; stats.expanded += mdd_.expanded; stats.transitions += mdd_.transitions; }
    void maybe_update_best() {  // sequential.rs:394-400
        isize v = mdd_.best_exact_value().value_or(ISIZE_MIN);
        if (v > best_lb) { best_lb = v; best_sol = mdd_.best_exact_solution(); }
    }
    void enqueue_cutset(isize ub) {  // sequential.rs:403-416
        isize lb = best_lb;
        mdd_.drain_cutset([&](SubProblem<S> n) {
            n.ub = std::min(ub, n.ub);
            if (n.ub > lb) {
                size_t depth = n.depth;
                size_t before = c_.fringe->len();
                c_.fringe->push(std::move(n));
                size_t after = c_.fringe->len();
                open_by_layer_[depth] += after - before;
            }
        });
    }
    void abort_search() { aborted = true; c_.fringe->clear(); c_.cache->clear(); }  // sequential.rs:418-422
};

// Wave-synchronous solver: the canonical batched schedule the device engine
// follows (K sub-problems popped per wave, restricted DDs compiled against one
// best_lb snapshot, incumbent updated in wave order, relaxed DDs compiled
// against the updated snapshot, cutsets enqueued in wave order).  K = 1 is
// exactly SequentialSolver with EmptyCache.  It is the parallel solver's
// schedule (parallel.rs:391-469,500-559) with K workers running in lock-step.
template <class S, class Hash, class Eq>
class WaveSolver {
public:
    WaveSolver(SolverConfig<S> c, size_t wave_size) : c_(c), K_(wave_size) {}
    SolverStats stats;
    isize best_lb = ISIZE_MIN, best_ub = ISIZE_MAX;
    std::optional<Solution> best_sol;
    isize sol_value = ISIZE_MIN;  // objective of best_sol (best_lb may be raised from outside: set_lower_bound)
    bool aborted = false;
    uint64_t max_waves = UINT64_MAX;
    // per-wave trace for parity tests: (wave size, best_lb after wave, fringe size after wave)
    struct WaveTrace { size_t popped; isize best_lb; size_t fringe_len; isize top_ub; };
    std::vector<WaveTrace> trace;

    // ---- stepwise form (what the fringe-sharded multi-GPU driver calls between its allreduce(max) steps) ----
    void init(bool push_root) {
        c_.fringe->clear();
        best_lb = ISIZE_MIN; best_ub = ISIZE_MAX; best_sol.reset(); sol_value = ISIZE_MIN; aborted = false; stats = SolverStats(); trace.clear();
        mdds_.clear();
        for (size_t i = 0; i < K_; ++i) mdds_.emplace_back(c_.cutset_type);
        if (push_root) c_.fringe->push(SubProblem<S>{std::make_shared<const S>(c_.problem->initial_state()), c_.problem->initial_value(), {}, ISIZE_MAX, 0});
    }
    // one wave; out3 = {best_lb, ub of the best open node before the wave (ISIZE_MIN if none), 1 if work remains}; false on cutoff
    bool wave(isize out3[3]) {
        std::vector<SubProblem<S>> wave;
        isize top_ub = ISIZE_MIN;
        while (wave.size() < K_ && !c_.fringe->is_empty()) {
            SubProblem<S> nn = *c_.fringe->pop();
            if (nn.ub <= best_lb) { c_.fringe->clear(); break; }  // parallel.rs:531-535
            if (wave.empty()) top_ub = nn.ub;
            wave.push_back(std::move(nn));
            stats.explored += 1;
        }
        out3[1] = top_ub;
        if (wave.empty()) { out3[0] = best_lb; out3[2] = 0; return true; }
        best_ub = top_ub;
        stats.waves += 1;
        // The K workers of the wave are emulated with ONE decision diagram (a wave of 2 048 sub-problems with a few dozen wide DDs would
        // otherwise keep gigabytes of arenas alive): what a worker's DD is asked for later -- its best exact value and solution, its
        // cutset -- is taken right after its compilation and applied in wave order, exactly as when every worker kept its own DD.
        Mdd<S, Hash, Eq>& mdd = mdds_[0];
        EmptyCache<S> cache;
        struct Found { size_t i; isize v; std::optional<Solution> sol; };
        auto remember = [&](std::vector<Found>& out, size_t i, isize floor_) {  // only a DD that beats the snapshot can become the incumbent
            const isize v = mdd.best_exact_value().value_or(ISIZE_MIN);
            if (v > floor_) out.push_back(Found{i, v, mdd.best_exact_solution()});
        };
        auto apply = [&](std::vector<Found>& found) {  // maybe_update_best in wave order: the first DD reaching a new maximum keeps its solution
            for (auto& f : found) if (f.v > best_lb) { best_lb = f.v; best_sol = std::move(f.sol); sol_value = f.v; }
        };
        // restricted
        isize lb = best_lb;
        std::vector<char> exact(wave.size(), 0);
        std::vector<size_t> widths(wave.size());
        std::vector<Found> found;
        for (size_t i = 0; i < wave.size(); ++i) {
            widths[i] = c_.width->max_width(wave[i]);
            CompilationInput<S> in{CompilationType::Restricted, c_.problem, c_.relaxation, c_.ranking, c_.cutoff, widths[i], &wave[i], lb, &cache, c_.dominance};
            Completion comp;
            if (!mdd.compile(in, &comp)) return false;
            exact[i] = comp.is_exact;
            account(mdd);
            remember(found, i, lb);
        }
        apply(found);
        // relaxed
        lb = best_lb;
        found.clear();
        std::vector<std::pair<size_t, std::vector<SubProblem<S>>>> cutsets;
        for (size_t i = 0; i < wave.size(); ++i) {
            if (exact[i]) continue;
            CompilationInput<S> in{CompilationType::Relaxed, c_.problem, c_.relaxation, c_.ranking, c_.cutoff, widths[i], &wave[i], lb, &cache, c_.dominance};
            Completion comp;
            if (!mdd.compile(in, &comp)) return false;
            exact[i] = comp.is_exact;
            account(mdd);
            remember(found, i, lb);
            if (!comp.is_exact) {
                cutsets.emplace_back(i, std::vector<SubProblem<S>>());
                auto& cs = cutsets.back().second;
                const isize ub = wave[i].ub;
                mdd.drain_cutset([&](SubProblem<S> n) {
                    n.ub = std::min(ub, n.ub);
                    if (n.ub > lb) cs.push_back(std::move(n));  // (lb <= the final incumbent of the wave: the filter below is the binding one)
                });
            }
        }
        apply(found);
        for (auto& pr : cutsets) {
            const isize blb = best_lb;
            for (auto& n : pr.second) if (n.ub > blb) c_.fringe->push(std::move(n));
        }
        trace.push_back({wave.size(), best_lb, c_.fringe->len(), top_ub});
        out3[0] = best_lb; out3[2] = c_.fringe->is_empty() ? 0 : 1;
        return true;
    }
    void set_lower_bound(isize lb) { if (lb > best_lb) best_lb = lb; }
    // keep every nranks-th open node of the common MaxUB order (SURVEY.md section 8e initial deal)
    void retain_share(size_t rank, size_t nranks) {
        if (nranks <= 1) return;
        std::vector<SubProblem<S>> keep;
        for (size_t idx = 0; !c_.fringe->is_empty(); ++idx) {
            SubProblem<S> n = *c_.fringe->pop();
            if ((idx + idx / nranks) % nranks == rank) keep.push_back(std::move(n));  // rotating deal (as Solver::retain_share of the device solver)
        }
        c_.fringe->clear();
        for (auto& n : keep) c_.fringe->push(std::move(n));
    }
    // work hand-off between ranks (mirrors Solver::export_open / import_open of the device solver): every other one of the best
    // 2 * max_nodes open nodes leaves, the rest is queued again
    std::vector<SubProblem<S>> export_open(size_t max_nodes) {
        std::vector<SubProblem<S>> out, keep;
        for (size_t idx = 0; out.size() < max_nodes && !c_.fringe->is_empty(); ++idx) {
            SubProblem<S> n = *c_.fringe->pop();
            if (idx & 1) out.push_back(std::move(n)); else keep.push_back(std::move(n));
        }
        for (auto& n : keep) c_.fringe->push(std::move(n));
        return out;
    }
    void import_open(std::vector<SubProblem<S>>&& nodes) {
        for (auto& n : nodes) if (n.ub > best_lb) c_.fringe->push(std::move(n));
    }
    void finish() {
        if (c_.fringe->is_empty() && !aborted) best_ub = best_lb;
        if (best_sol) std::stable_sort(best_sol->begin(), best_sol->end(), [](const Decision& a, const Decision& b) { return a.variable < b.variable; });
    }
    size_t fringe_len() const { return c_.fringe->len(); }

    Completion maximize() {
        init(true);
        for (;;) {
            if (c_.fringe->is_empty()) { best_ub = best_lb; break; }
            if (stats.waves >= max_waves) { aborted = true; break; }
            isize o3[3];
            if (!wave(o3)) { abort_search(); break; }
        }
        if (best_sol) std::stable_sort(best_sol->begin(), best_sol->end(), [](const Decision& a, const Decision& b) { return a.variable < b.variable; });
        return Completion{!aborted, best_sol ? std::optional<isize>(best_lb) : std::nullopt};
    }
private:
    SolverConfig<S> c_;
    size_t K_;
    std::vector<Mdd<S, Hash, Eq>> mdds_;
    void account(const Mdd<S, Hash, Eq>& m) { stats.compilations++; stats.expanded += m.expanded; stats.transitions += m.transitions; }
    void maybe_update_best(const Mdd<S, Hash, Eq>& m) {
        isize v = m.best_exact_value().value_or(ISIZE_MIN);
        if (v > best_lb) { best_lb = v; best_sol = m.best_exact_solution(); sol_value = v; }
    }
    void abort_search() { aborted = true; c_.fringe->clear(); }
};

// implementation/solver/parallel.rs:287-641 -- N worker threads over one mutex-protected fringe.
// This is the CPU baseline ("restated reference"): same data-structure choices as the reference.
template <class S, class Hash, class Eq>
class ParallelSolver {
public:
    ParallelSolver(SolverConfig<S> c, size_t nb_threads) : c_(c), nb_threads_(nb_threads) {
        size_t n = c.problem->nb_variables() + 1;
        open_by_layer_.assign(n, 0); ongoing_by_layer_.assign(n, 0);
    }
    SolverStats stats;
    isize best_lb = ISIZE_MIN, best_ub = ISIZE_MAX;
    std::optional<Solution> best_sol;
    bool aborted = false;

    Completion maximize() {  // parallel.rs:573-607
        {
            SubProblem<S> root{std::make_shared<const S>(c_.problem->initial_state()), c_.problem->initial_value(), {}, ISIZE_MAX, 0};
            c_.cache->initialize(*c_.problem);
            std::lock_guard<std::mutex> g(mu_);
            c_.fringe->push(root);
            open_by_layer_[0] += 1;
        }
        std::vector<std::thread> threads;
        for (size_t i = 0; i < nb_threads_; ++i) threads.emplace_back([this] { worker(); });
        for (auto& t : threads) t.join();
        if (best_sol) std::stable_sort(best_sol->begin(), best_sol->end(), [](const Decision& a, const Decision& b) { return a.variable < b.variable; });
        return Completion{!aborted, best_sol ? std::optional<isize>(best_lb) : std::nullopt};
    }
private:
    SolverConfig<S> c_;
    size_t nb_threads_;
    std::mutex mu_;
    std::condition_variable monitor_;
    size_t ongoing_ = 0, first_active_layer_ = 0;
    std::vector<size_t> open_by_layer_, ongoing_by_layer_;
    enum class WL { Complete, Aborted, Starvation, WorkItem };

    void worker() {
        Mdd<S, Hash, Eq> mdd(c_.cutset_type);
        for (;;) {
            SubProblem<S> node;
            WL w = get_workload(node);
            if (w == WL::Complete || w == WL::Aborted) break;
            if (w == WL::Starvation) continue;
            isize ub = node.ub; size_t depth = node.depth;
            bool ok = process_one_node(mdd, node);
            if (!ok) {
                std::lock_guard<std::mutex> g(mu_);  // abort_search parallel.rs:479-489
                aborted = true;
                best_ub = (best_ub == ISIZE_MAX) ? ub : std::max(ub, best_ub);
                c_.fringe->clear(); c_.cache->clear();
            }
            {
                std::lock_guard<std::mutex> g(mu_);  // notify_node_finished :471-477
                ongoing_ -= 1; ongoing_by_layer_[depth] -= 1;
            }
            monitor_.notify_all();
            if (!ok) break;
        }
    }
    WL get_workload(SubProblem<S>& out) {  // parallel.rs:500-559
        std::unique_lock<std::mutex> g(mu_);
        while (first_active_layer_ < c_.problem->nb_variables() && open_by_layer_[first_active_layer_] + ongoing_by_layer_[first_active_layer_] == 0) {
            c_.cache->clear_layer(first_active_layer_);
            first_active_layer_ += 1;
        }
        if (ongoing_ == 0 && c_.fringe->is_empty()) { best_ub = best_lb; monitor_.notify_all(); return WL::Complete; }
        if (aborted) return WL::Aborted;
        if (c_.fringe->is_empty()) { monitor_.wait(g); return WL::Starvation; }
        SubProblem<S> nn = *c_.fringe->pop();
        for (;;) {
            if (nn.ub <= best_lb) {
                c_.fringe->clear();
                std::fill(open_by_layer_.begin(), open_by_layer_.end(), 0);
                return WL::Starvation;
            }
            if (c_.cache->must_explore(nn)) {
                c_.cache->update_threshold(nn.state, nn.depth, nn.value, true);
                break;
            } else {
                open_by_layer_[nn.depth] -= 1;
                if (c_.fringe->is_empty()) return WL::Starvation;
                nn = *c_.fringe->pop();
            }
        }
        ongoing_ += 1; stats.explored += 1;
        open_by_layer_[nn.depth] -= 1; ongoing_by_layer_[nn.depth] += 1;
        out = std::move(nn);
        return WL::WorkItem;
    }
    isize read_lb() { std::lock_guard<std::mutex> g(mu_); return best_lb; }
    bool process_one_node(Mdd<S, Hash, Eq>& mdd, const SubProblem<S>& node) {  // parallel.rs:391-437
        isize node_ub = node.ub;
        isize lb = read_lb();
        if (node_ub <= lb) return true;
        size_t width = c_.width->max_width(node);
        CompilationInput<S> in{CompilationType::Restricted, c_.problem, c_.relaxation, c_.ranking, c_.cutoff, width, &node, lb, c_.cache, c_.dominance};
        Completion comp;
        if (!mdd.compile(in, &comp)) { account_aborted(mdd); return false; }
        maybe_update_best(mdd);
        if (comp.is_exact) return true;
        in.comp_type = CompilationType::Relaxed;
        in.best_lb = read_lb();
        if (!mdd.compile(in, &comp)) { account_aborted(mdd); return false; }
        maybe_update_best(mdd);
        if (!comp.is_exact) {  // enqueue_cutset :456-469
            std::lock_guard<std::mutex> g(mu_);
            isize blb = best_lb;
            mdd.drain_cutset([&](SubProblem<S> n) {
                n.ub = std::min(node_ub, n.ub);
                if (n.ub > blb) {
                    size_t depth = n.depth, before = c_.fringe->len();
                    c_.fringe->push(std::move(n));
                    open_by_layer_[depth] += c_.fringe->len() - before;
                }
            });
        }
        return true;
    }
    // metric counter only: the layers a DD expanded before Cutoff::must_stop fired (clean.rs:352) still count as work done
    void account_aborted(const Mdd<S, Hash, Eq>& mdd) { std::lock_guard<std::mutex> g(mu_); stats.expanded += mdd.expanded; stats.transitions += mdd.transitions; }
    void maybe_update_best(const Mdd<S, Hash, Eq>& mdd) {  // parallel.rs:446-453
        std::lock_guard<std::mutex> g(mu_);
        stats.compilations++; stats.expanded += mdd.expanded; stats.transitions += mdd.transitions;
        isize v = mdd.best_exact_value().value_or(ISIZE_MIN);
        if (v > best_lb) { best_lb = v; best_sol = mdd.best_exact_solution(); }
    }
};

}  // namespace ddo_oracle

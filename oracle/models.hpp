// ============================================================================
// models.hpp -- TEST INFRASTRUCTURE (see ddo_oracle.hpp).  CPU restatements of
// the reference's example models that sit on the hot path:
//   * MISP      examples/misp/main.rs:37-209, instance reader :258-317
//   * Knapsack  examples/knapsack/main.rs:36-218, reader :267-299
// and of the hand-written fixtures of the reference's unit tests
//   * DummyProblem & co      clean.rs:2552-2667
//   * LocBoundsAndThresholds clean.rs:2056-2181
// ============================================================================
#pragma once
#include "ddo_oracle.hpp"
#include <cmath>

namespace ddo_oracle {

// ---------------------------------------------------------------------------
// MISP.  State = bit set over vertices (reference: bit-set 0.5.3 `BitSet<u32>`,
// un-vendored; semantics restated: iter ascending, len = popcount, Ord =
// lexicographic over the ascending member sequence).
// ---------------------------------------------------------------------------
struct BitState {
    std::vector<uint64_t> w;
    bool contains(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    void remove(size_t i) { w[i >> 6] &= ~(1ull << (i & 63)); }
    void insert(size_t i) { w[i >> 6] |= (1ull << (i & 63)); }
    size_t len() const { size_t c = 0; for (uint64_t x : w) c += (size_t)__builtin_popcountll(x); return c; }
};
struct BitStateHash {
    size_t operator()(const BitState& s) const {  // FxHasher-style word mixing (fxhash 0.2.1 is un-vendored; only iteration order would leak, and C1 removes that)
        uint64_t h = 0;
        for (uint64_t x : s.w) h = ((h << 5 | h >> 59) ^ x) * 0x517cc1b727220a95ull;
        return (size_t)h;
    }
};
struct BitStateEq { bool operator()(const BitState& a, const BitState& b) const { return a.w == b.w; } };

// BitSet::cmp (bit-set 0.5.3: `self.iter().cmp(other.iter())`), word-parallel form
inline int bitset_lex_cmp(const BitState& a, const BitState& b) {
    size_t n = a.w.size();
    for (size_t j = 0; j < n; ++j) {
        uint64_t d = a.w[j] ^ b.w[j];
        if (!d) continue;
        int p = __builtin_ctzll(d);
        bool a_owns = (a.w[j] >> p) & 1;
        const BitState& other = a_owns ? b : a;
        // does `other` have a member greater than p ?
        uint64_t above = (p == 63) ? 0 : (other.w[j] >> (p + 1));
        bool has_more = above != 0;
        for (size_t k = j + 1; k < n && !has_more; ++k) has_more = other.w[k] != 0;
        int owner_cmp = has_more ? -1 : 1;  // owner of p is Less unless the other sequence ended
        return a_owns ? owner_cmp : -owner_cmp;
    }
    return 0;
}

constexpr isize MISP_YES = 1, MISP_NO = 0;

struct Misp : Problem<BitState> {  // misp/main.rs:37-148
    size_t nb_vars = 0;
    size_t words = 0;
    std::vector<BitState> neighbors;  // COMPLEMENT of the adjacency (misp/main.rs:40-45)
    std::vector<isize> weight;

    // misp/main.rs:280-309: full sets, then edges removed (1-based in files; here 0-based)
    Misp(size_t n, const isize* w, size_t m, const int32_t* src, const int32_t* dst) : nb_vars(n), words((n + 63) / 64) {
        BitState full; full.w.assign(words, 0);
        for (size_t i = 0; i < n; ++i) full.insert(i);
        neighbors.assign(n, full);
        weight.assign(n, 1);
        if (w) for (size_t i = 0; i < n; ++i) weight[i] = w[i];
        for (size_t e = 0; e < m; ++e) { neighbors[src[e]].remove(dst[e]); neighbors[dst[e]].remove(src[e]); }
    }
    size_t nb_variables() const override { return nb_vars; }
    BitState initial_state() const override {  // :69-71
        BitState s; s.w.assign(words, 0);
        for (size_t i = 0; i < nb_vars; ++i) s.insert(i);
        return s;
    }
    isize initial_value() const override { return 0; }
    BitState transition(const BitState& state, Decision d) const override {  // :77-85
        BitState res = state;
        res.remove(d.variable);
        if (d.value == MISP_YES) for (size_t j = 0; j < words; ++j) res.w[j] &= neighbors[d.variable].w[j];
        return res;
    }
    isize transition_cost(const BitState&, const BitState&, Decision d) const override {  // :87-93
        return d.value == MISP_NO ? 0 : weight[d.variable];
    }
    void for_each_in_domain(Variable v, const BitState& state, const DecisionCallback& f) const override {  // :95-102
        if (state.contains(v.id)) { f(Decision{v.id, MISP_YES}); f(Decision{v.id, MISP_NO}); }
        else f(Decision{v.id, MISP_NO});
    }
    std::optional<Variable> next_variable(size_t, const std::vector<const BitState*>& next_layer) const override {  // :109-143
        thread_local std::vector<size_t> heu;
        heu.assign(nb_vars, 0);
        for (const BitState* s : next_layer)
            for (size_t j = 0; j < words; ++j) {
                uint64_t x = s->w[j];
                while (x) { int b = __builtin_ctzll(x); heu[j * 64 + b] += 1; x &= x - 1; }
            }
        std::optional<Variable> best; size_t bestc = 0;
        for (size_t i = 0; i < nb_vars; ++i)
            if (heu[i] > 0 && (!best || heu[i] < bestc)) { best = Variable{i}; bestc = heu[i]; }  // min_by_key: first minimum
        return best;
    }
    bool is_impacted_by(Variable v, const BitState& s) const override { return s.contains(v.id); }  // :145-147
};
struct MispRelax : Relaxation<BitState> {  // misp/main.rs:168-194
    const Misp* pb;
    explicit MispRelax(const Misp* p) : pb(p) {}
    BitState merge(const std::vector<const BitState*>& states) const override {
        BitState s; s.w.assign(pb->words, 0);
        for (const BitState* x : states) for (size_t j = 0; j < pb->words; ++j) s.w[j] |= x->w[j];
        return s;
    }
    isize relax(const BitState&, const BitState&, const BitState&, Decision, isize cost) const override { return cost; }
    isize fast_upper_bound(const BitState& s) const override {
        isize sum = 0;
        for (size_t j = 0; j < pb->words; ++j) {
            uint64_t x = s.w[j];
            while (x) { int b = __builtin_ctzll(x); sum += pb->weight[j * 64 + b]; x &= x - 1; }
        }
        return sum;
    }
};
struct MispRanking : StateRanking<BitState> {  // misp/main.rs:201-209
    int compare(const BitState& a, const BitState& b) const override {
        size_t la = a.len(), lb = b.len();
        if (la != lb) return la < lb ? -1 : 1;
        return bitset_lex_cmp(a, b);
    }
};

// ---------------------------------------------------------------------------
// Knapsack (BASELINE config 1).  knapsack/main.rs:36-218
// ---------------------------------------------------------------------------
struct KnapsackState { size_t depth; size_t capacity; };
struct KnapsackHash { size_t operator()(const KnapsackState& s) const { return (size_t)((s.depth * 0x9E3779B97F4A7C15ull) ^ (s.capacity * 0xC2B2AE3D27D4EB4Full)); } };
struct KnapsackEq { bool operator()(const KnapsackState& a, const KnapsackState& b) const { return a.depth == b.depth && a.capacity == b.capacity; } };
constexpr isize TAKE_IT = 1, LEAVE_IT_OUT = 0;

struct Knapsack : Problem<KnapsackState> {
    size_t capacity;
    std::vector<isize> profit;
    std::vector<size_t> weight;
    std::vector<size_t> order;
    Knapsack(size_t cap, std::vector<isize> p, std::vector<size_t> w) : capacity(cap), profit(std::move(p)), weight(std::move(w)) {  // :62-69
        order.resize(profit.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
            return -(double)profit[a] / (double)weight[a] < -(double)profit[b] / (double)weight[b];
        });
    }
    size_t nb_variables() const override { return profit.size(); }
    void for_each_in_domain(Variable v, const KnapsackState& s, const DecisionCallback& f) const override {  // :93-99
        if (s.capacity >= weight[v.id]) f(Decision{v.id, TAKE_IT});
        f(Decision{v.id, LEAVE_IT_OUT});
    }
    KnapsackState initial_state() const override { return KnapsackState{0, capacity}; }
    isize initial_value() const override { return 0; }
    KnapsackState transition(const KnapsackState& s, Decision d) const override {  // :106-113
        KnapsackState r = s; r.depth += 1;
        if (d.value == TAKE_IT) r.capacity -= weight[d.variable];
        return r;
    }
    isize transition_cost(const KnapsackState&, const KnapsackState&, Decision d) const override { return profit[d.variable] * d.value; }
    std::optional<Variable> next_variable(size_t depth, const std::vector<const KnapsackState*>&) const override {  // :118-125
        if (depth < nb_variables()) return Variable{order[depth]};
        return std::nullopt;
    }
};
struct KPRelax : Relaxation<KnapsackState> {  // :147-181
    const Knapsack* pb;
    explicit KPRelax(const Knapsack* p) : pb(p) {}
    KnapsackState merge(const std::vector<const KnapsackState*>& states) const override {  // max_by_key: last maximum
        const KnapsackState* best = states[0];
        for (auto* s : states) if (s->capacity >= best->capacity) best = s;
        return *best;
    }
    isize relax(const KnapsackState&, const KnapsackState&, const KnapsackState&, Decision, isize cost) const override { return cost; }
    isize fast_upper_bound(const KnapsackState& state) const override {  // :158-180 (f64 floor)
        size_t depth = state.depth; isize max_profit = 0; size_t cap = state.capacity;
        while (cap > 0 && depth < pb->profit.size()) {
            size_t item = pb->order[depth];
            if (cap >= pb->weight[item]) { max_profit += pb->profit[item]; cap -= pb->weight[item]; }
            else {
                double ratio = (double)cap / (double)pb->weight[item];
                double ip = ratio * (double)pb->profit[item];
                max_profit += (isize)std::floor(ip);
                cap = 0;
            }
            depth += 1;
        }
        return max_profit;
    }
};
struct KPRanking : StateRanking<KnapsackState> {  // :184-191
    int compare(const KnapsackState& a, const KnapsackState& b) const override { return a.capacity < b.capacity ? -1 : (a.capacity > b.capacity ? 1 : 0); }
};
struct KPDominance : Dominance<KnapsackState> {  // :193-218
    std::optional<isize> get_key(const KnapsackState& s) const override { return (isize)s.depth; }
    size_t nb_dimensions(const KnapsackState&) const override { return 1; }
    isize get_coordinate(const KnapsackState& s, size_t) const override { return (isize)s.capacity; }
    bool use_value() const override { return true; }
};

// ---------------------------------------------------------------------------
// MAX2SAT (BASELINE config 3).  examples/max2sat/model.rs:29-348, relax.rs:43-89, heuristics.rs:30-37, data.rs:31-110
// State = depth + marginal benefit of assigning True to every variable (model.rs:59-62).
// Canonical choices where the reference leaves the outcome to an unstable sort / heap layout (SURVEY hard part 2):
//   * variable order: model.rs:149-151 sorts unstably by the sum of clause weights -> here stable (ties by variable id);
//   * ranking: heuristics.rs:33-37 compares rank() = sum |benefit| only; ties are refined by (depth, lexicographic benefits as
//     signed integers, ascending variable id), which makes the cut and the fringe order a function of the states alone.
// ---------------------------------------------------------------------------
struct M2State { size_t depth; std::vector<isize> sub; };
struct M2Hash {
    size_t operator()(const M2State& s) const {
        uint64_t h = (uint64_t)s.depth * 0x9E3779B97F4A7C15ull;
        for (isize x : s.sub) h = ((h << 5 | h >> 59) ^ (uint64_t)x) * 0x517cc1b727220a95ull;
        return (size_t)h;
    }
};
struct M2Eq { bool operator()(const M2State& a, const M2State& b) const { return a.depth == b.depth && a.sub == b.sub; } };
constexpr isize M2_T = 1, M2_F = -1;  // model.rs:30-32
struct M2Clause { isize w, x, y; };   // weight, literals (x == y: unit clause); literal of variable i (0-based) is +-(i+1)

struct Max2Sat : Problem<M2State> {  // model.rs:98-349
    size_t nb_vars;
    isize initial = 0;
    std::vector<isize> weights;  // (2n)^2 table, model.rs:141
    std::vector<isize> sum_of_clause_weights;
    std::vector<size_t> order;   // vars_by_sum_of_clause_weights
    std::vector<isize> nk, estimates;
    static size_t mk_lit(isize x) { size_t a = (size_t)((x < 0 ? -x : x) - 1); return a + a + (x > 0 ? 1 : 0); }  // model.rs:116-121
    size_t offset(isize x, isize y) const { isize a = std::min(x, y), b = std::max(x, y); return mk_lit(a) * 2 * nb_vars + mk_lit(b); }  // :165-170
    isize weight(isize x, isize y) const { return weights[offset(x, y)]; }
    static isize t(size_t v) { return (isize)v + 1; }
    static isize f(size_t v) { return -((isize)v + 1); }
    static isize pos(isize x) { return x > 0 ? x : 0; }

    // data.rs:99,106: `weights.insert(BinaryClause::new(x, y), w)` -- a repeated clause keeps the LAST weight.  The reference then iterates
    // the hash map (model.rs:136); every per-clause effect below is order independent once duplicates are resolved.
    Max2Sat(size_t n, const std::vector<M2Clause>& clauses) : nb_vars(n), weights(4 * n * n, 0), sum_of_clause_weights(n, 0), order(n) {
        std::vector<M2Clause> uniq;
        std::unordered_map<uint64_t, size_t> seen;
        for (const M2Clause& c : clauses) {
            isize a = std::min(c.x, c.y), b = std::max(c.x, c.y);
            uint64_t key = (uint64_t)mk_lit(a) * 2 * n + mk_lit(b);
            auto it = seen.find(key);
            if (it == seen.end()) { seen.emplace(key, uniq.size()); uniq.push_back(M2Clause{c.w, a, b}); }
            else uniq[it->second].w = c.w;
        }
        for (const M2Clause& c : uniq) {  // model.rs:136-147
            weights[offset(c.x, c.y)] = c.w;
            sum_of_clause_weights[(size_t)((c.x < 0 ? -c.x : c.x) - 1)] += c.w;
            if (c.x != c.y) sum_of_clause_weights[(size_t)((c.y < 0 ? -c.y : c.y) - 1)] += c.w;
            if (c.x == -c.y) initial += c.w;
        }
        for (size_t i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return sum_of_clause_weights[a] < sum_of_clause_weights[b]; });  // :149-151
        estimates.assign(n, 0); nk.assign(n, 0);
        // precompute_estimate(k), model.rs:204-238, as a suffix sum: estimate(k) = estimate(k+1) + terms of i = k
        isize acc = 0;
        for (size_t i = n; i-- > 0;) {
            size_t vi = order[i];
            for (size_t j = i + 1; j < n; ++j) {
                size_t vj = order[j];
                isize tt = weight(t(vi), t(vj)), tf = weight(t(vi), f(vj)), ft = weight(f(vi), t(vj)), ff = weight(f(vi), f(vj));
                isize wtt = tt + tf + ft, wtf = tt + tf + ff, wft = tt + ft + ff, wff = tf + ft + ff;
                acc += std::max(std::max(wtt, wtf), std::max(wft, wff));
            }
            acc += weight(t(vi), f(vi)) + std::max(weight(t(vi), t(vi)), weight(f(vi), f(vi)));
            estimates[i] = acc;
        }
        for (size_t k = 0; k < n; ++k) {  // precompute_nk, model.rs:190-197
            isize sum = 0;
            for (size_t i = 0; i < k; ++i) sum += weight(t(order[i]), f(order[i]));
            nk[k] = sum;
        }
    }
    isize fast_upper_bound(const M2State& state) const {  // model.rs:240-249
        isize mb = 0;
        for (isize b : state.sub) mb += b < 0 ? -b : b;
        return mb + estimates[state.depth] - initial + nk[state.depth];
    }
    size_t nb_variables() const override { return nb_vars; }
    M2State initial_state() const override { return M2State{0, std::vector<isize>(nb_vars, 0)}; }
    isize initial_value() const override { return initial; }
    void for_each_in_domain(Variable v, const M2State&, const DecisionCallback& cb) const override {  // :270-273
        cb(Decision{v.id, M2_T}); cb(Decision{v.id, M2_F});
    }
    M2State transition(const M2State& state, Decision d) const override {  // :275-292
        size_t k = d.variable;
        M2State ret = state;
        ret.depth += 1;
        ret.sub[k] = 0;
        size_t nrem = nb_vars - (state.depth + 1);  // varset(), :173-181
        if (d.value == M2_F) for (size_t i = 0; i < nrem; ++i) { size_t l = order[i]; ret.sub[l] += weight(t(k), t(l)) - weight(t(k), f(l)); }
        else for (size_t i = 0; i < nrem; ++i) { size_t l = order[i]; ret.sub[l] += weight(f(k), t(l)) - weight(f(k), f(l)); }
        return ret;
    }
    isize transition_cost(const M2State& state, const M2State&, Decision d) const override {  // :294-328
        size_t k = d.variable;
        size_t nrem = nb_vars - (state.depth + 1);
        if (d.value == M2_F) {
            isize res = pos(-state.sub[k]);
            isize sum = weight(f(k), f(k));
            for (size_t i = 0; i < nrem; ++i) {
                size_t l = order[i];
                isize wff = weight(f(k), f(l)), wft = weight(f(k), t(l)), wtt = weight(t(k), t(l)), wtf = weight(t(k), f(l));
                sum += (wff + wft) + std::min(pos(state.sub[l]) + wtt, pos(-state.sub[l]) + wtf);
            }
            return res + sum;
        }
        isize res = pos(state.sub[k]);
        isize sum = weight(t(k), t(k));
        for (size_t i = 0; i < nrem; ++i) {
            size_t l = order[i];
            isize wtt = weight(t(k), t(l)), wtf = weight(t(k), f(l)), wff = weight(f(k), f(l)), wft = weight(f(k), t(l));
            sum += (wtf + wtt) + std::min(pos(state.sub[l]) + wft, pos(-state.sub[l]) + wff);
        }
        return res + sum;
    }
    std::optional<Variable> next_variable(size_t, const std::vector<const M2State*>& next_layer) const override {  // :330-348
        if (next_layer.empty()) return std::nullopt;
        size_t depth = next_layer[0]->depth;
        if (depth < nb_vars) return Variable{order[nb_vars - depth - 1]};
        return std::nullopt;
    }
};
struct Max2SatRelax : Relaxation<M2State> {  // relax.rs:43-89
    const Max2Sat* pb;
    explicit Max2SatRelax(const Max2Sat* p) : pb(p) {}
    M2State merge(const std::vector<const M2State*>& states) const override {  // :46-77
        std::vector<isize> benefits(pb->nb_vars, 0);
        for (size_t v = 0; v < pb->nb_vars; ++v) {
            isize sign = 0, min_benef = ISIZE_MAX; bool same = true;
            for (const M2State* st : states) {
                isize x = st->sub[v], ax = x < 0 ? -x : x;
                min_benef = std::min(min_benef, ax);
                if (sign == 0 && x != 0) sign = ax / x;
                else if (sign * x < 0) { same = false; break; }
            }
            if (same) benefits[v] = sign * min_benef;
        }
        return M2State{states[0]->depth, std::move(benefits)};
    }
    isize relax(const M2State&, const M2State& dst, const M2State& relaxed, Decision, isize cost) const override {  // :78-84
        isize rc = cost;
        for (size_t v = 0; v < pb->nb_vars; ++v) {
            isize a = dst.sub[v], b = relaxed.sub[v];
            rc += (a < 0 ? -a : a) - (b < 0 ? -b : b);
        }
        return rc;
    }
    isize fast_upper_bound(const M2State& s) const override { return pb->fast_upper_bound(s); }  // :86-88
};
struct Max2SatRanking : StateRanking<M2State> {  // heuristics.rs:30-37 (+ canonical refinement of ties, see above)
    static isize rank(const M2State& s) { isize r = 0; for (isize x : s.sub) r += x < 0 ? -x : x; return r; }  // model.rs:77-81
    int compare(const M2State& a, const M2State& b) const override {
        isize ra = rank(a), rb = rank(b);
        if (ra != rb) return ra < rb ? -1 : 1;
        if (a.depth != b.depth) return a.depth < b.depth ? -1 : 1;
        for (size_t v = 0; v < a.sub.size(); ++v) if (a.sub[v] != b.sub[v]) return a.sub[v] < b.sub[v] ? -1 : 1;
        return 0;
    }
};

// ---------------------------------------------------------------------------
// Fixtures of the reference's unit tests (clean.rs:2552-2667)
// ---------------------------------------------------------------------------
struct DummyState { isize value; size_t depth; };
struct DummyHash { size_t operator()(const DummyState& s) const { return (size_t)(s.value * 1000003 + (isize)s.depth); } };
struct DummyEq { bool operator()(const DummyState& a, const DummyState& b) const { return a.value == b.value && a.depth == b.depth; } };
struct DummyProblem : Problem<DummyState> {  // clean.rs:2558-2597
    bool infeasible = false;  // DummyInfeasibleProblem clean.rs:2599-2636
    size_t nb_variables() const override { return 3; }
    isize initial_value() const override { return 0; }
    DummyState initial_state() const override { return DummyState{0, 0}; }
    DummyState transition(const DummyState& s, Decision d) const override { return DummyState{s.value + d.value, 1 + s.depth}; }
    isize transition_cost(const DummyState&, const DummyState&, Decision d) const override { return d.value; }
    std::optional<Variable> next_variable(size_t depth, const std::vector<const DummyState*>&) const override {
        if (depth < nb_variables()) return Variable{depth};
        return std::nullopt;
    }
    void for_each_in_domain(Variable var, const DummyState&, const DecisionCallback& f) const override {
        if (infeasible) return;
        for (isize d = 0; d <= 2; ++d) f(Decision{var.id, d});
    }
};
struct DummyRelax : Relaxation<DummyState> {  // clean.rs:2638-2657
    DummyState merge(const std::vector<const DummyState*>& s) const override { return DummyState{100, s[0]->depth}; }
    isize relax(const DummyState&, const DummyState&, const DummyState&, Decision, isize) const override { return 20; }
    isize fast_upper_bound(const DummyState& s) const override { return (isize)(3 - s.depth) * 10; }
};
struct DummyRanking : StateRanking<DummyState> {  // clean.rs:2659-2667
    int compare(const DummyState& a, const DummyState& b) const override { return a.value < b.value ? 1 : (a.value > b.value ? -1 : 0); }
};

// clean.rs:2056-2181
struct CharHash { size_t operator()(char c) const { return (size_t)c; } };
struct CharEq { bool operator()(char a, char b) const { return a == b; } };
struct LocBoundsPb : Problem<char> {
    size_t nb_variables() const override { return 4; }
    char initial_state() const override { return 'r'; }
    isize initial_value() const override { return 0; }
    std::optional<Variable> next_variable(size_t, const std::vector<const char*>& next_layer) const override {
        char c = next_layer.empty() ? 'z' : *next_layer[0];
        switch (c) {
            case 'r': return Variable{0};
            case 'a': case 'b': return Variable{1};
            case 'c': case 'd': case 'M': case 'e': case 'f': return Variable{2};
            case 'g': case 'h': case 'i': return Variable{0};
            default: return std::nullopt;
        }
    }
    void for_each_in_domain(Variable v, const char& s, const DecisionCallback& f) const override {
        std::vector<isize> dom;
        switch (s) {
            case 'r': dom = {10, 7}; break;
            case 'a': dom = {2}; break;
            case 'b': dom = {3, 6, 5}; break;
            case 'M': dom = {4}; break;
            case 'e': dom = {0}; break;
            case 'f': dom = {1, 2}; break;
            case 'g': case 'h': case 'i': dom = {0}; break;
            default: break;
        }
        for (isize x : dom) f(Decision{v.id, x});
    }
    char transition(const char& s, Decision d) const override {
        if (s == 'r' && d.value == 10) return 'a';
        if (s == 'r' && d.value == 7) return 'b';
        if (s == 'a' && d.value == 2) return 'c';
        if (s == 'b' && d.value == 3) return 'd';
        if (s == 'b' && d.value == 6) return 'e';
        if (s == 'b' && d.value == 5) return 'f';
        if (s == 'M' && d.value == 4) return 'g';
        if (s == 'e' && d.value == 0) return 'h';
        if (s == 'f' && d.value == 1) return 'h';
        if (s == 'f' && d.value == 2) return 'i';
        return 't';
    }
    isize transition_cost(const char&, const char&, Decision d) const override { return d.value; }
};
struct LocBoundsRelax : Relaxation<char> {
    char merge(const std::vector<const char*>&) const override { return 'M'; }
    isize relax(const char&, const char&, const char&, Decision, isize cost) const override { return cost; }
    isize fast_upper_bound(const char& s) const override {
        switch (s) {
            case 'r': return 30;
            case 'a': case 'b': return 20;
            case 'M': case 'e': case 'f': return 10;
            default: return 0;
        }
    }
};
struct CmpChar : StateRanking<char> {
    int compare(const char& a, const char& b) const override { return a < b ? -1 : (a > b ? 1 : 0); }
};

}  // namespace ddo_oracle

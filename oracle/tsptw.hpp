// tsptw.hpp -- CPU restatement of ddo's TSPTW example (BASELINE config 4), TEST INFRASTRUCTURE like the rest of oracle/.
// Follows, under /root/reference/ddo/examples/tsptw/:
//   state.rs:34-101      TsptwState, Position, ElapsedTime
//   instance.rs:37-108   TimeWindow, TsptwInstance (the f32 x 10000 -> usize conversion is done by the caller, tests/oracle_lib.py)
//   model.rs:29-217      Tsptw: Problem
//   relax.rs:32-264      TsptwRelax: merge / relax / fast_upper_bound
//   heuristics.rs:29-51  TsptwRanking, TsptwWidth
//   dominance.rs:26-60   TsptwDominance, TsptwKey
// Third-party arithmetic: smallbitset 0.7.1 `Set256` (Cargo.lock) -- a 256-bit set; only add / remove / union / inter / diff / flip /
// len / ascending iteration are used (state.rs:40-55, relax.rs:100-160) and restated here over four 64-bit words.
// Pinned by tests/test_oracle_golden.py against the optima asserted in examples/tsptw/tests.rs (Langevin / SolomonPotvinBengio files).
#pragma once
#include <array>

#include "ddo_oracle.hpp"

namespace ddo_oracle {

struct Set256 {
    std::array<uint64_t, 4> w{0, 0, 0, 0};
    void add(size_t i) { w[i >> 6] |= 1ull << (i & 63); }
    void remove(size_t i) { w[i >> 6] &= ~(1ull << (i & 63)); }
    bool contains(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1ull; }
    void union_with(const Set256& o) { for (int j = 0; j < 4; ++j) w[j] |= o.w[j]; }
    void inter_with(const Set256& o) { for (int j = 0; j < 4; ++j) w[j] &= o.w[j]; }
    void diff_with(const Set256& o) { for (int j = 0; j < 4; ++j) w[j] &= ~o.w[j]; }
    Set256 flip() const { Set256 r; for (int j = 0; j < 4; ++j) r.w[j] = ~w[j]; return r; }
    size_t len() const { size_t c = 0; for (int j = 0; j < 4; ++j) c += (size_t)__builtin_popcountll(w[j]); return c; }
    bool operator==(const Set256& o) const { return w == o.w; }
    template <class F> void for_each(F f) const {  // ascending
        for (int j = 0; j < 4; ++j) { uint64_t x = w[j]; while (x) { f((size_t)(64 * j + __builtin_ctzll(x))); x &= x - 1; } }
    }
};

struct TimeWindow { size_t earliest, latest; };  // instance.rs:37-48
struct TsptwInstance { size_t nb_nodes = 0; std::vector<std::vector<size_t>> distances; std::vector<TimeWindow> timewindows; };  // instance.rs:51-59

// state.rs:34-69
struct TsptwState {
    bool virtual_pos = false; uint16_t node = 0; Set256 pool;     // Position::Node(node) | Position::Virtual(pool)
    bool fuzzy = false; size_t earliest = 0, latest = 0;          // ElapsedTime::FixedAmount{duration = earliest} | FuzzyAmount{earliest, latest}
    Set256 must_visit;
    bool has_maybe = false; Set256 maybe_visit;                   // Option<Set256>
    uint16_t depth = 0;
    size_t elapsed_earliest() const { return earliest; }          // state.rs:90-95
};
struct TsptwEq {  // #[derive(PartialEq, Eq)]: variants and payloads
    bool operator()(const TsptwState& a, const TsptwState& b) const {
        if (a.virtual_pos != b.virtual_pos || (a.virtual_pos ? !(a.pool == b.pool) : a.node != b.node)) return false;
        if (a.fuzzy != b.fuzzy || a.earliest != b.earliest || (a.fuzzy && a.latest != b.latest)) return false;
        if (!(a.must_visit == b.must_visit) || a.has_maybe != b.has_maybe || (a.has_maybe && !(a.maybe_visit == b.maybe_visit))) return false;
        return a.depth == b.depth;
    }
};
struct TsptwHash {
    size_t operator()(const TsptwState& s) const {
        uint64_t h = 0x9E3779B97F4A7C15ull * (uint64_t)(s.depth + 1);
        auto mix = [&](uint64_t x) { h = ((h << 5 | h >> 59) ^ x) * 0x517cc1b727220a95ull; };
        mix(s.virtual_pos); if (s.virtual_pos) for (uint64_t x : s.pool.w) mix(x); else mix(s.node);
        mix(s.fuzzy); mix(s.earliest); if (s.fuzzy) mix(s.latest);
        for (uint64_t x : s.must_visit.w) mix(x);
        mix(s.has_maybe); if (s.has_maybe) for (uint64_t x : s.maybe_visit.w) mix(x);
        return (size_t)h;
    }
};

struct Tsptw : Problem<TsptwState> {  // model.rs:29-217
    TsptwInstance instance;
    TsptwState initial;
    explicit Tsptw(TsptwInstance inst) : instance(std::move(inst)) {  // model.rs:35-47
        for (size_t i = 1; i < instance.nb_nodes; ++i) initial.must_visit.add(i);
    }
    size_t nb_variables() const override { return instance.nb_nodes; }
    TsptwState initial_state() const override { return initial; }
    isize initial_value() const override { return 0; }

    size_t min_distance_to(const TsptwState& s, size_t j) const {  // model.rs:196-205
        if (!s.virtual_pos) return instance.distances[s.node][j];
        size_t m = SIZE_MAX;
        s.pool.for_each([&](size_t i) { m = std::min(m, instance.distances[i][j]); });
        return m;
    }
    size_t max_distance_to(const TsptwState& s, size_t j) const {  // model.rs:206-216
        if (!s.virtual_pos) return instance.distances[s.node][j];
        size_t m = 0;
        s.pool.for_each([&](size_t i) { m = std::max(m, instance.distances[i][j]); });
        return m;
    }
    bool can_move_to(const TsptwState& s, size_t j) const {  // model.rs:150-157: the earliest arrival must not exceed the window's end
        return s.earliest + min_distance_to(s, j) <= instance.timewindows[j].latest;
    }
    void for_each_in_domain(Variable var, const TsptwState& s, const DecisionCallback& f) const override {  // model.rs:65-93
        if ((size_t)s.depth == nb_variables() - 1) {
            if (can_move_to(s, 0)) f(Decision{var.id, 0});
            return;
        }
        bool ok = true;
        s.must_visit.for_each([&](size_t i) { if (ok && !can_move_to(s, i)) ok = false; });
        if (!ok) return;
        s.must_visit.for_each([&](size_t i) { f(Decision{var.id, (isize)i}); });
        if (s.has_maybe) s.maybe_visit.for_each([&](size_t i) { if (can_move_to(s, i)) f(Decision{var.id, (isize)i}); });
    }
    TsptwState transition(const TsptwState& s, Decision d) const override {  // model.rs:95-114 + arrival_time :158-195
        TsptwState r;
        const size_t j = (size_t)d.value;
        r.must_visit = s.must_visit; r.must_visit.remove(j);
        r.has_maybe = s.has_maybe; r.maybe_visit = s.maybe_visit;
        if (r.has_maybe) r.maybe_visit.remove(j);
        size_t min_arrival = s.earliest + min_distance_to(s, j);                        // FixedAmount: duration; FuzzyAmount: earliest
        size_t max_arrival = (s.fuzzy ? s.latest : s.earliest) + max_distance_to(s, j);  // FixedAmount: duration; FuzzyAmount: latest
        const TimeWindow tw = instance.timewindows[j];
        if (min_arrival == max_arrival) { r.fuzzy = false; r.earliest = std::max(min_arrival, tw.earliest); r.latest = r.earliest; }
        else {
            size_t e = std::max(min_arrival, tw.earliest), l = std::min(max_arrival, tw.latest);
            if (e == l) { r.fuzzy = false; r.earliest = e; r.latest = e; } else { r.fuzzy = true; r.earliest = e; r.latest = l; }
        }
        r.virtual_pos = false; r.node = (uint16_t)j;
        r.depth = (uint16_t)(s.depth + 1);
        return r;
    }
    isize transition_cost(const TsptwState& s, const TsptwState&, Decision d) const override {  // model.rs:116-138
        const size_t j = (size_t)d.value;
        const TimeWindow tw = instance.timewindows[j];
        const size_t travel = min_distance_to(s, j);
        const size_t waiting = (s.earliest + travel) < tw.earliest ? tw.earliest - (s.earliest + travel) : 0;
        return -(isize)(travel + waiting);
    }
    std::optional<Variable> next_variable(size_t depth, const std::vector<const TsptwState*>&) const override {  // model.rs:140-147
        if (depth == nb_variables()) return std::nullopt;
        return Variable{depth};
    }
};

struct TsptwRelax : Relaxation<TsptwState> {  // relax.rs:32-264
    const Tsptw* pb;
    std::vector<size_t> cheapest_edge;
    explicit TsptwRelax(const Tsptw* p) : pb(p) {  // relax.rs:50-64
        const size_t n = pb->nb_variables();
        for (size_t i = 0; i < n; ++i) {
            size_t m = SIZE_MAX;
            for (size_t j = 0; j < n; ++j) if (i != j) m = std::min(m, pb->instance.distances[j][i]);
            cheapest_edge.push_back(m);
        }
    }
    TsptwState merge(const std::vector<const TsptwState*>& states) const override {  // relax.rs:169-190 with RelaxHelper :66-163
        uint16_t depth = 0; Set256 position; size_t earliest = SIZE_MAX, latest = 0;
        Set256 all_must, all_agree = Set256().flip(), all_maybe;
        for (const TsptwState* s : states) {
            depth = std::max(depth, s->depth);
            if (s->virtual_pos) position.union_with(s->pool); else position.add(s->node);
            earliest = std::min(earliest, s->earliest);
            latest = std::max(latest, s->fuzzy ? s->latest : s->earliest);
            all_agree.inter_with(s->must_visit);
            all_must.union_with(s->must_visit);
            if (s->has_maybe) all_maybe.union_with(s->maybe_visit);
        }
        TsptwState r;
        r.depth = depth;
        r.virtual_pos = true; r.pool = position;
        if (earliest == latest) { r.fuzzy = false; r.earliest = earliest; r.latest = earliest; } else { r.fuzzy = true; r.earliest = earliest; r.latest = latest; }
        r.must_visit = all_agree;
        Set256 maybe = all_maybe; maybe.union_with(all_must); maybe.diff_with(all_agree);
        r.has_maybe = maybe.len() > 0; if (r.has_maybe) r.maybe_visit = maybe;
        return r;
    }
    isize relax(const TsptwState&, const TsptwState&, const TsptwState&, Decision, isize cost) const override { return cost; }  // relax.rs:192-194
    isize fast_upper_bound(const TsptwState& s) const override {  // relax.rs:196-264
        size_t complete_tour = pb->nb_variables() - (size_t)s.depth;
        std::vector<size_t> tmp;
        size_t mandatory = 0, back_to_depot = SIZE_MAX;
        bool infeasible = false;
        s.must_visit.for_each([&](size_t i) {
            if (infeasible) return;
            complete_tour -= 1;
            mandatory += cheapest_edge[i];
            back_to_depot = std::min(back_to_depot, pb->instance.distances[i][0]);
            if (s.earliest + cheapest_edge[i] > pb->instance.timewindows[i].latest) infeasible = true;
        });
        if (infeasible) return ISIZE_MIN;
        if (s.has_maybe) {
            size_t violations = 0;
            s.maybe_visit.for_each([&](size_t i) {
                tmp.push_back(cheapest_edge[i]);
                back_to_depot = std::min(back_to_depot, pb->instance.distances[i][0]);
                if (s.earliest + cheapest_edge[i] > pb->instance.timewindows[i].latest) violations += 1;
            });
            if (tmp.size() - violations < complete_tour) return ISIZE_MIN;
            std::sort(tmp.begin(), tmp.end());
            for (size_t q = 0; q < complete_tour && q < tmp.size(); ++q) mandatory += tmp[q];
        }
        if (mandatory == 0) {
            size_t here = SIZE_MAX;
            if (!s.virtual_pos) here = pb->instance.distances[s.node][0];
            else s.pool.for_each([&](size_t x) { here = std::min(here, pb->instance.distances[x][0]); });
            back_to_depot = std::min(back_to_depot, here);
        }
        const size_t total = mandatory + back_to_depot;
        if (s.earliest + total > pb->instance.timewindows[0].latest) return ISIZE_MIN;
        return -(isize)total;
    }
};

struct TsptwRanking : StateRanking<TsptwState> {  // heuristics.rs:29-37
    int compare(const TsptwState& a, const TsptwState& b) const override { return a.depth < b.depth ? -1 : (a.depth > b.depth ? 1 : 0); }
};
struct TsptwWidth : WidthHeuristic<TsptwState> {  // heuristics.rs:39-51
    size_t nb_vars, factor;
    TsptwWidth(size_t n, size_t f) : nb_vars(n), factor(f) {}
    size_t max_width(const SubProblem<TsptwState>& s) const override { return nb_vars * (s.depth + 1) * factor; }
};
struct TsptwDominance : Dominance<TsptwState> {  // dominance.rs:26-60: key = (position, must_visit), no coordinates, value compared
    std::optional<isize> get_key(const TsptwState&) const override { return 0; }
    std::optional<std::string> get_key_bytes(const TsptwState& s) const override {
        std::string k;
        k.push_back(s.virtual_pos ? 'V' : 'N');
        if (s.virtual_pos) k.append(reinterpret_cast<const char*>(s.pool.w.data()), 32); else k.append(reinterpret_cast<const char*>(&s.node), 2);
        k.append(reinterpret_cast<const char*>(s.must_visit.w.data()), 32);
        return k;
    }
    size_t nb_dimensions(const TsptwState&) const override { return 0; }
    isize get_coordinate(const TsptwState&, size_t) const override { return 0; }
    bool use_value() const override { return true; }
};

}  // namespace ddo_oracle

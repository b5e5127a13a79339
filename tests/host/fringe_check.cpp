// fringe_check.cpp -- host-logic test of ddo::NoDupFringe (ddo_b200/csrc/solver.cu; no device code is touched): a burst pushed with
// push_many (several host threads over the sharded state index) must leave the fringe in exactly the state the sequential push() loop
// leaves it in (no_duplicate.rs:88-140: one entry per state, the longer path wins, ub = max) and pop in the same MaxUB order
// (no_duplicate.rs:144-164, subproblem_ranking.rs:86-90).  Built and run by tests/test_host_fringe.py against libddo_b200.so.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <tuple>
#include <vector>

#include "../../ddo_b200/csrc/solver.hpp"

using ddo::NoDupFringe;

struct Burst {
    std::vector<uint64_t> states, bits;
    std::vector<int32_t> value, ub, depth, rec, nbw;
};

static Burst make_burst(std::mt19937_64& rng, int kind, int W, int PW, size_t n, size_t distinct, int max_depth) {
    // `distinct` base states; records pick one at random, so a burst holds many duplicates with different values / bounds / paths
    std::vector<uint64_t> base(distinct * W);
    std::vector<int32_t> base_depth(distinct);
    for (size_t i = 0; i < distinct; ++i) {
        for (int j = 0; j < W; ++j) {
            uint64_t x = rng();
            if (kind == DDO_MODEL_MAX2SAT) x = (uint64_t)(uint32_t)(int32_t)((int)(rng() % 41) - 20) | ((uint64_t)(uint32_t)(int32_t)((int)(rng() % 41) - 20) << 32);
            else x &= rng();  // sparser bitsets
            base[i * W + j] = x;
        }
        base_depth[i] = 1 + (int)(rng() % max_depth);
    }
    Burst b;
    b.states.resize(n * W); b.bits.resize(n * PW); b.value.resize(n); b.ub.resize(n); b.depth.resize(n); b.rec.resize(n); b.nbw.resize(n);
    for (size_t r = 0; r < n; ++r) {
        const size_t s = rng() % distinct;
        std::memcpy(&b.states[r * W], &base[s * W], (size_t)W * 8);
        // MAX2SAT: the depth is part of the state; MISP: equal bitsets may arrive with different depths
        b.depth[r] = kind == DDO_MODEL_MAX2SAT ? base_depth[s] : 1 + (int)(rng() % max_depth);
        b.value[r] = (int32_t)(rng() % 7);      // few distinct values and bounds: plenty of ties for the dedup rules and the order
        b.ub[r] = 10 + (int32_t)(rng() % 9);
        b.rec[r] = (int32_t)(rng() % 1000);
        b.nbw[r] = (b.depth[r] + 63) / 64;
        for (int q = 0; q < PW; ++q) b.bits[r * PW + q] = q < b.nbw[r] ? rng() : 0;
    }
    return b;
}

// Independent model of no_duplicate.rs:88-140: one entry per state (MAX2SAT: per state and depth), the longer path wins, ub = max.
struct ModelItem { int32_t value, ub, depth, rec; std::vector<uint64_t> bits; };
using Model = std::map<std::vector<uint64_t>, ModelItem>;
static std::vector<uint64_t> model_key(int kind, const uint64_t* st, int W, int32_t depth) {
    std::vector<uint64_t> k(st, st + W);
    if (kind == DDO_MODEL_MAX2SAT) k.push_back((uint64_t)depth);
    return k;
}
static void push_model(Model& m, int kind, const Burst& b, int W, int PW) {
    for (size_t r = 0; r < b.value.size(); ++r) {
        auto key = model_key(kind, &b.states[r * W], W, b.depth[r]);
        std::vector<uint64_t> bits(PW, 0);
        for (int q = 0; q < b.nbw[r]; ++q) bits[q] = b.bits[r * PW + q];
        auto it = m.find(key);
        if (it == m.end()) { m.emplace(key, ModelItem{b.value[r], b.ub[r], b.depth[r], b.rec[r], bits}); continue; }
        ModelItem& o = it->second;
        const int32_t merged = std::max(o.ub, b.ub[r]);
        if (b.value[r] > o.value) o = ModelItem{b.value[r], merged, b.depth[r], b.rec[r], bits};
        o.ub = merged;
    }
}
// MaxUB (subproblem_ranking.rs:86-90) over the model's ranking, written from the reference's definitions (not from solver.cu):
//   MISP    MispRanking = (len, BitSet::cmp) (misp/main.rs:203-208); BitSet::cmp compares the ascending member lists lexicographically;
//   MAX2SAT Max2SatRanking = rank = sum |benefit| (heuristics.rs:33-37), refined canonically by (depth, lexicographic signed benefits).
static int ref_cmp(int kind, int W, const uint64_t* sa, const NoDupFringe::Item& a, const uint64_t* sb, const NoDupFringe::Item& b) {
    if (a.ub != b.ub) return a.ub < b.ub ? -1 : 1;
    if (a.value != b.value) return a.value < b.value ? -1 : 1;
    if (kind == DDO_MODEL_MAX2SAT) {
        const int32_t* x = reinterpret_cast<const int32_t*>(sa); const int32_t* y = reinterpret_cast<const int32_t*>(sb);
        long long rx = 0, ry = 0;
        for (int j = 0; j < 2 * W; ++j) { rx += std::llabs((long long)x[j]); ry += std::llabs((long long)y[j]); }
        if (rx != ry) return rx < ry ? -1 : 1;
        if (a.depth != b.depth) return a.depth < b.depth ? -1 : 1;
        for (int j = 0; j < 2 * W; ++j) if (x[j] != y[j]) return x[j] < y[j] ? -1 : 1;
        return 0;
    }
    std::vector<int> ma, mb;
    for (int v = 0; v < 64 * W; ++v) { if ((sa[v >> 6] >> (v & 63)) & 1) ma.push_back(v); if ((sb[v >> 6] >> (v & 63)) & 1) mb.push_back(v); }
    if (ma.size() != mb.size()) return ma.size() < mb.size() ? -1 : 1;
    if (ma == mb) return 0;
    return std::lexicographical_compare(ma.begin(), ma.end(), mb.begin(), mb.end()) ? -1 : 1;
}

// pops `count` nodes of f: each must be the model's entry for its state, and the (ub, value) keys must never increase (MaxUB order,
// subproblem_ranking.rs:86-90)
static int check_against_model(NoDupFringe& f, Model& m, int kind, int W, int PW, size_t count, const char* what) {
    std::tuple<int32_t, int32_t> prev{INT32_MAX, INT32_MAX};
    std::vector<uint64_t> prev_state; NoDupFringe::Item prev_item{};
    for (size_t i = 0; i < count && !f.empty(); ++i) {
        const int x = f.pop();
        const NoDupFringe::Item it = f.item(x);
        if (!prev_state.empty() && ref_cmp(kind, W, prev_state.data(), prev_item, f.state(x), it) <= 0) {
            std::printf("FAIL %s: pop %zu is not below its predecessor in the reference's MaxUB order\n", what, i);
            return 1;
        }
        prev_state.assign(f.state(x), f.state(x) + W); prev_item = it;
        auto key = model_key(kind, f.state(x), W, it.depth);
        auto mi = m.find(key);
        if (mi == m.end()) { std::printf("FAIL %s: popped a state the model does not hold\n", what); return 1; }
        const ModelItem& o = mi->second;
        if (o.value != it.value || o.ub != it.ub || o.depth != it.depth || o.rec != it.rec || std::memcmp(o.bits.data(), f.bits(x), (size_t)PW * 8)) {
            std::printf("FAIL %s: popped item differs from the model (value %d/%d ub %d/%d depth %d/%d rec %d/%d)\n", what, it.value, o.value, it.ub, o.ub, it.depth, o.depth, it.rec, o.rec);
            return 1;
        }
        std::tuple<int32_t, int32_t> cur{it.ub, it.value};
        if (cur > prev) { std::printf("FAIL %s: pop order not MaxUB\n", what); return 1; }
        prev = cur;
        m.erase(mi);
    }
    return 0;
}

static void push_seq(NoDupFringe& f, const Burst& b, int W, int PW) {
    for (size_t r = 0; r < b.value.size(); ++r) f.push(&b.states[r * W], b.value[r], b.ub[r], b.depth[r], b.rec[r], &b.bits[r * PW], b.nbw[r]);
}
static void push_par(NoDupFringe& f, const Burst& b, int W, int PW) {
    std::vector<NoDupFringe::PushRec> recs;
    for (size_t r = 0; r < b.value.size(); ++r) recs.push_back(NoDupFringe::PushRec{&b.states[r * W], &b.bits[r * PW], b.value[r], b.ub[r], b.depth[r], b.rec[r], b.nbw[r]});
    f.push_many(recs);
}

// `a` is drained with pop(), `b` with pop_many() in chunks: same nodes in the same order
static int compare_pops(NoDupFringe& a, NoDupFringe& b, int W, int PW, size_t count, const char* what) {
    std::vector<int> ids;
    size_t have = 0, done = 0;
    for (size_t i = 0; i < count; ++i) {
        if (have == done) {
            if (a.empty() != b.empty()) { std::printf("FAIL %s: emptiness differs after %zu pops\n", what, i); return 1; }
            if (a.empty()) break;
            const size_t want = std::min<size_t>(777, count - i);
            b.pop_many((int)want, ids);
            have = ids.size(); done = 0;
            if (have != std::min(want, a.len())) { std::printf("FAIL %s: pop_many returned %zu of %zu\n", what, have, std::min(want, a.len())); return 1; }
        }
        const int x = a.pop(), y = ids[done++];
        const NoDupFringe::Item ia = a.item(x), ib = b.item(y);
        if (ia.value != ib.value || ia.ub != ib.ub || ia.depth != ib.depth || ia.rec != ib.rec || std::memcmp(a.state(x), b.state(y), (size_t)W * 8) ||
            std::memcmp(a.bits(x), b.bits(y), (size_t)PW * 8)) {
            std::printf("FAIL %s: pop %zu differs (value %d/%d ub %d/%d depth %d/%d rec %d/%d)\n", what, i, ia.value, ib.value, ia.ub, ib.ub, ia.depth, ib.depth, ia.rec, ib.rec);
            return 1;
        }
    }
    if (a.len() != b.len()) { std::printf("FAIL %s: lengths differ after the drain (%zu vs %zu)\n", what, a.len(), b.len()); return 1; }
    return 0;
}

int main() {
    int fails = 0, checks = 0;
    for (int kind : {DDO_MODEL_MISP, DDO_MODEL_MAX2SAT}) {
        for (int W : {1, 8, 16}) {
            const int PW = 8;
            std::mt19937_64 rng(1234 + 17 * W + kind);
            NoDupFringe seq(W, PW, kind), par(W, PW, kind), third(W, PW, kind);
            Model model;
            // rounds of (burst, partial drain): small bursts take the sequential path of push_many, large ones the threaded path; later
            // bursts hit nodes already in the fringe and slots recycled by the pops
            // the 150 000 / 120 000 bursts are large enough for the split flush (best entries sorted at once, the rest by a background thread):
            // the short drains after them pop from the sorted part while that thread runs, the small bursts that follow push while it runs
            // (no slot is recycled meanwhile), the long drains run into the background-sorted part
            const size_t sizes[] = {100, 20000, 50, 9000, 60000, 3000, 30000, 150000, 700, 4000, 120000, 64};
            const size_t short_drain[] = {0, 0, 0, 0, 0, 0, 0, 3000, 2500, 0, 6000, 0};
            for (size_t round = 0; round < sizeof(sizes) / sizeof(sizes[0]); ++round) {
                const size_t n = sizes[round];
                Burst b = make_burst(rng, kind, W, PW, n, std::max<size_t>(8, n / 3), 400);
                push_seq(seq, b, W, PW);
                push_par(par, b, W, PW);
                push_par(third, b, W, PW);
                push_model(model, kind, b, W, PW);
                ++checks;
                if (third.len() != model.size()) { std::printf("FAIL kind %d W %d round %zu: len %zu vs model %zu\n", kind, W, round, third.len(), model.size()); ++fails; break; }
                const size_t drain = short_drain[round] ? short_drain[round] : third.len() / 2 + 1;
                if (check_against_model(third, model, kind, W, PW, drain, "model, partial drain")) { ++fails; break; }
                if (seq.len() != par.len()) { std::printf("FAIL kind %d W %d round %zu: len %zu vs %zu\n", kind, W, round, seq.len(), par.len()); ++fails; break; }
                if (compare_pops(seq, par, W, PW, short_drain[round] ? short_drain[round] : seq.len() / 2 + 1, "partial drain")) { ++fails; break; }
            }
            if (compare_pops(seq, par, W, PW, (size_t)-1, "final drain")) ++fails;
            if (check_against_model(third, model, kind, W, PW, (size_t)-1, "model, final drain") || !model.empty()) ++fails;
            if (!seq.empty() || !par.empty()) { std::printf("FAIL: fringe not empty after the final drain\n"); ++fails; }
            // clear() then reuse
            Burst b = make_burst(rng, kind, W, PW, 12000, 2000, 100);
            seq.clear(); par.clear();
            push_seq(seq, b, W, PW); push_par(par, b, W, PW);
            ++checks;
            if (seq.len() != par.len() || compare_pops(seq, par, W, PW, (size_t)-1, "after clear")) ++fails;
        }
    }
    std::printf("%s: %d bursts checked, %d failures\n", fails ? "FAILED" : "OK", checks, fails);
    return fails ? 1 : 0;
}

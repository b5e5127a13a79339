// solver.hpp -- host-side branch-and-bound driver over the device engine.
//
// Restates, natively, the callers of the hot path: `ParallelSolver` (ddo/src/implementation/solver/
// parallel.rs:287-641) with its K workers running in lock-step as one device batch ("wave"), the `NoDupFringe`
// (fringe/no_duplicate.rs:52-323) ordered by `MaxUB` (heuristics/subproblem_ranking.rs:86-90) with `MispRanking`
// (examples/misp/main.rs:201-209), `FixedWidth` / `NbUnassignedWidth` (heuristics/width.rs:166-170,397-401) and `TimeBudget`
// (heuristics/cutoff.rs:302-323).  Paths are kept as a tree of per-DD records so that a sub-problem costs O(state) host memory.
#pragma once
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "engine.hpp"

namespace ddo {

// One record per compiled relaxed DD that produced open sub-problems: the decisions from that DD's root to its cutset layer share the
// variable sequence; each sub-problem only stores one bit per layer.
struct PathRec {
    int32_t parent_rec;              // record of the DD root's own path, -1 for the problem root
    std::vector<uint64_t> parent_bits;  // the DD root's decision bits inside parent_rec
    std::vector<int32_t> vars;       // branching variable of each layer between the DD root and its deepest cutset node
    int32_t base_depth = 0;          // depth of the DD root: a sub-problem of depth d uses the first d - base_depth layers (all of them for
                                     // a LAST_EXACT_LAYER cutset; the nodes of a FRONTIER cutset sit in different layers)
};

class NoDupFringe {
public:
    // kind: DDO_MODEL_MISP -- MispRanking (popcount, BitSet::cmp), states compared without their depth (BitSet alone is the state);
    //       DDO_MODEL_MAX2SAT -- Max2SatRanking (rank = sum |benefit|, heuristics.rs:33-37) refined canonically by (depth, lexicographic
    //       benefits); the depth is part of the state (model.rs:59-62 derives Hash / Eq over both fields).
    NoDupFringe(int words, int pw, int kind = 0) : W(words), PW(pw), kind_(kind) {
        nodes_.init(HW + words + pw);
        if (const char* e = std::getenv("DDO_FRINGE_ASYNC")) async_sort_ = std::atoi(e) != 0;
    }
    ~NoDupFringe() { drop_cold(); }
    NoDupFringe(const NoDupFringe&) = delete;
    NoDupFringe& operator=(const NoDupFringe&) = delete;
    struct Item { int32_t value, ub, depth, rec; };
    size_t len() const { return live_; }
    bool empty() const { return live_ == 0; }
    void clear();
    // no_duplicate.rs:88-140
    void push(const uint64_t* state, int32_t value, int32_t ub, int32_t depth, int32_t rec, const uint64_t* bits, int nbits_words);
    // A burst of pushes (the cutsets of a wide wave: hundreds of thousands of sub-problems), equivalent to push() of every record in order.
    // Large bursts run on several host threads: the state index is split into NS shards by the top bits of the state hash, equal states
    // fall into one shard and are handled by one thread in record order, so duplicates resolve exactly as in the sequential loop.
    struct PushRec { const uint64_t* state; const uint64_t* bits; int32_t value, ub, depth, rec, nbits_words; };
    void push_many(const std::vector<PushRec>& recs);
    // no_duplicate.rs:144-164; returns node id (valid until the next push)
    int pop();
    // the next min(k, len()) nodes in pop() order; ids valid until the next push.  With `out`, the records of the popped nodes (item, packed
    // state, the first bits_words words of the path bits) are appended to the caller's vectors while their cache lines arrive.
    struct PopOut { std::vector<Item>* items; std::vector<uint64_t>* states; std::vector<uint64_t>* bits; int bits_words; };
    int pop_many(int k, std::vector<int>& ids, const PopOut* out = nullptr);
    void prefetch(int id) const { const char* p = reinterpret_cast<const char*>(nodes_.at(id)); for (int b = 0; b < (HW + W + PW) * 8; b += 64) __builtin_prefetch(p + b); }
    const uint64_t* state(int id) const { return nodes_.at(id) + HW; }
    const uint64_t* bits(int id) const { return nodes_.at(id) + HW + W; }
    const Item& item(int id) const { return hdr(id).it; }
private:
    // Priority structure: the reference's updatable binary heap (no_duplicate.rs:206-323) pops in the MaxUB order, a strict total order
    // for MISP (ub, value, then MispRanking on distinct states).  Pushes arrive in bursts (the cutsets of a wave) and pops in bursts (the
    // next wave), so the same order is served by sorted runs: a burst of pushes is sorted once into a new run, a pop takes the largest of
    // the runs' tails.  An entry is stale when its node was popped or re-keyed since (version mismatch) and is skipped.
    struct Ent { uint64_t k1, k2; int id; uint32_t ver; };
    int W, PW, kind_;
    // states live in fixed blocks (no reallocation copies: a MAX2SAT fringe holds gigabytes of 2 KB states)
    struct Arena {
        int W = 1; int shift = 0; size_t per_block = 1; std::vector<std::unique_ptr<uint64_t[]>> blocks; size_t count = 0;
        void init(int w) {  // ~32 MB blocks, a power of two rows; the block list never reallocates (a background sort reads it while the owner grows it)
            W = w; shift = 0; while (((size_t)2 << shift) * (size_t)w <= ((size_t)1 << 22)) ++shift; per_block = (size_t)1 << shift; blocks.reserve(1 << 15);
        }
        uint64_t* at(size_t id) const { return blocks[id >> shift].get() + (id & (per_block - 1)) * (size_t)W; }
        void grow() { if (count == blocks.size() * per_block) blocks.emplace_back(new uint64_t[per_block * (size_t)W]); ++count; }
        void clear() { count = 0; }  // the blocks stay (no page faults when the next search refills them)
    } nodes_;
    // One record per node: [Hdr][packed state, W words][path bits, PW words], contiguous (160 bytes for MISP at n = 500): a pop touches one
    // run of adjacent cache lines instead of one line in each of five arrays.
    struct Hdr { Item it; uint64_t hash; uint32_t ver; int32_t popc; };  // popc = ranking key of the state: popcount (MISP) / sum |benefit| (MAX2SAT)
    static constexpr int HW = 4;  // header words
    static_assert(sizeof(Hdr) == HW * 8, "node header layout");
    Hdr& hdr(int id) const { return *reinterpret_cast<Hdr*>(nodes_.at(id)); }
    uint64_t* state_w(int id) const { return nodes_.at(id) + HW; }
    uint64_t* bits_w(int id) const { return nodes_.at(id) + HW + W; }
    void new_node() { nodes_.grow(); Hdr& h = hdr((int)nodes_.count - 1); h.it = Item{}; h.hash = 0; h.ver = 0; h.popc = 0; }
    std::vector<int> recycle_;
    std::vector<Ent> pending_;
    // A run is sorted ascending (pops take the back).  A large burst is split at flush time: its best few thousand entries are sorted at once (`v`),
    // the rest (`cold`, every entry below every entry of `v`) is sorted by a background thread while the device compiles the next wave and
    // joins the run when `v` is used up.  While such a sort is in flight no node slot is recycled (its comparator reads node states).
    struct Cold { std::vector<Ent> ents; std::thread th; std::atomic<bool> done{false}; };
    struct Run { std::vector<Ent> v; std::unique_ptr<Cold> cold; };
    std::vector<Run> runs_;
    int cold_open_ = 0;
    std::vector<std::vector<uint32_t>> cand_; std::vector<size_t> head_, scan_;  // scratch of pop_many
    bool async_sort_ = true;  // DDO_FRINGE_ASYNC=0: sort every burst at once (A/B runs)
    static constexpr size_t kHotLarge = 16384, kHotSmall = 4096;
    void sort_ents(std::vector<Ent>& v) const;
    void join_cold(Run& run);
    bool cold_busy();   // true while a background sort is still running (finished ones are folded into their runs)
    void drop_cold();
    std::vector<Ent>& tail_run(size_t r);  // run r with stale tail entries dropped and, if its sorted part is used up, its cold part joined
    size_t live_ = 0;
    // state index: NS open-addressing tables (node id, -1 empty, -2 tombstone), shard = top bits of the state hash
    static constexpr int NS_BITS = 6, NS = 1 << NS_BITS;
    struct alignas(128) Shard { std::vector<int> tab; size_t used = 0, live = 0; };  // used = occupied + tombstones; one cache-line pair per shard (threads own shards)
    Shard shards_[NS];
    static int shard_of(uint64_t h) { return (int)(h >> (64 - NS_BITS)); }
    uint64_t key_hash(const uint64_t* st, int32_t depth) const;
    int push_one(const uint64_t* st, uint64_t h, int32_t value, int32_t ub, int32_t depth, int32_t rec, const uint64_t* bits, int nbits_words, int new_id,
                 std::vector<Ent>& pending);  // returns 1 when new_id was consumed (vacant entry), 0 when an existing node was updated
    Ent make_ent(int id) const;
    bool ent_less(const Ent& a, const Ent& b) const;  // a strictly below b in the MaxUB order
    int state_cmp(int a, int b) const;  // final tie-break of the ranking between two stored nodes
    void flush_pending();
    void table_insert(int id);
    int table_find(const uint64_t* st, uint64_t h, int32_t depth) const;
    void table_erase(int id);
    void rehash(Shard& sh, size_t min_cap);
    static uint64_t hash_state(const uint64_t* st, int W);
};

// the collectives a sharded search needs, as plain function pointers (ddo_comm_* in production; anything else in tests)
struct ShardComm {
    void* ctx; int nranks, rank;
    int (*allgather)(void* ctx, const int64_t* values, int32_t count, int64_t* recv);
    int (*send)(void* ctx, const void* buf, int64_t bytes, int32_t peer);
    int (*recv)(void* ctx, void* buf, int64_t bytes, int32_t peer);
};
constexpr int64_t kMaxHandoff = 8192, kMinDonor = 64;  // open nodes per hand-off; what a donor keeps for itself

struct Solver {
    Engine* eng; int model_kind; int n_vars, words;
    std::vector<uint64_t> root_state; int64_t root_value;  // Problem::initial_state / initial_value
    int width_kind; uint64_t width; int wave_size;
    NoDupFringe fringe;
    std::vector<PathRec> recs;
    int64_t best_lb = INT64_MIN, best_ub = INT64_MAX;
    bool has_sol = false; std::vector<ddo_decision> best_sol;
    int64_t sol_value = INT64_MIN;   // objective of best_sol: best_lb may be raised from outside (another rank's incumbent) without a solution
    double deadline_ms = 0;          // TimeBudget of the running maximize() (0 = none), polled before every device batch
    bool aborted = false;
    uint64_t explored = 0, expanded = 0, transitions = 0, compilations = 0, waves = 0;
    double device_ms = 0, fringe_ms = 0;
    FILE* trace_file = nullptr;  // DDO_WAVE_TRACE=<path>: one line per wave (wave, popped, general DDs, inexact, small ms, general ms, layer steps, expanded, fringe)
    // Speculative pre-pop (maximize() only): while the device compiles the fast-path DDs of a wave, the host pops the nodes of the NEXT wave.
    // If the wave turns out to push nothing and leaves best_lb alone (the common case: every DD of the fast path is exact), those are exactly
    // the nodes the next wave would pop; otherwise they are pushed back before anything else happens.
    bool pipeline = false, pre_valid = false;
    std::vector<uint64_t> pre_states, pre_bits; std::vector<NoDupFringe::Item> pre_items;
    void prepop(); void unpop();
    ~Solver();
    Solver(const Solver&) = delete;
    Solver& operator=(const Solver&) = delete;
    // scratch of one wave (kept to avoid reallocations)
    std::vector<uint64_t> w_states, w_bits, p_states, p_bits; std::vector<NoDupFringe::Item> w_items; std::vector<int32_t> p_val, p_ub, p_vars, p_tt; std::vector<NoDupFringe::PushRec> push_recs; std::vector<int> pop_ids;

    Solver(Engine* e, int model_kind, const uint64_t* root_state, int64_t root_value, int wk, uint64_t w, int ws);
    int init(bool push_root);
    int wave(const volatile int32_t* cutoff_flag, int64_t out3[3]);
    int wave_body(const volatile int32_t* cutoff_flag, int64_t out3[3]);
    int maximize(double time_budget_s, uint64_t max_waves, int32_t* is_exact, int32_t* has_value, int64_t* best_value);
    void finish();
    void full_path(int32_t rec, const uint64_t* bits, int32_t depth, std::vector<ddo_decision>& out) const;
    int retain_share(int rank, int nranks);
    // Work hand-off between ranks (the reference's workers share ONE fringe, parallel.rs:500-559; here a rank whose fringe runs dry is
    // refilled by a loaded one).  export_open pops up to 2 * max_nodes of the best open nodes, gives away every other one (so donor and
    // receiver keep nodes of the same quality) and re-queues the rest; a node travels as packed state, value, upper bound, depth and its
    // FULL decision path.  import_open queues such nodes (each gets a path record of its own).
    // packed node: int64 words [value, ub (INT64_MAX: none), depth, state words ..., decisions four per word: 16 bits each = variable | bit << 15]
    int node_words() const { return 3 + words + (n_vars + 3) / 4; }
    int export_open(int max_nodes, int64_t* rows, int32_t* count);
    int import_open(int count, const int64_t* rows);
    int maximize_sharded(const ShardComm& cm, double time_budget_s, uint64_t max_waves, bool rebalance, int64_t out[8]);
    size_t open_len() const { return fringe.len() + (pre_valid ? pre_items.size() : 0); }  // nodes popped ahead of time are still open
};

}  // namespace ddo

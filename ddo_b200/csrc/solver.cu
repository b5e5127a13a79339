// solver.cu -- host-side branch-and-bound driver (see solver.hpp).  No device code in this file.
#include "solver.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

namespace ddo {

static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ---------------------------------------------------------------------------------------------------------------
// NoDupFringe
// ---------------------------------------------------------------------------------------------------------------
uint64_t NoDupFringe::hash_state(const uint64_t* st, int W) {
    // four independent multiply-xor lanes (a MAX2SAT state is 250 words: one dependent chain would cost a multiply latency per word)
    uint64_t h[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
    int j = 0;
    for (; j + 4 <= W; j += 4)
        for (int q = 0; q < 4; ++q) { h[q] = (h[q] ^ st[j + q]) * 0xff51afd7ed558ccdULL; h[q] ^= h[q] >> 29; }
    for (; j < W; ++j) { h[0] = (h[0] ^ st[j]) * 0xff51afd7ed558ccdULL; h[0] ^= h[0] >> 29; }
    uint64_t x = h[0] ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7);
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 32;
    return x;
}
// BitSet::cmp of bit-set 0.5.3 (lexicographic over ascending members), word-parallel
static int lex_cmp(const uint64_t* a, const uint64_t* b, int W) {
    for (int j = 0; j < W; ++j) {
        const uint64_t d = a[j] ^ b[j];
        if (!d) continue;
        const int p = __builtin_ctzll(d);
        const bool a_owns = (a[j] >> p) & 1;
        const uint64_t* other = a_owns ? b : a;
        bool has_more = p < 63 && (other[j] >> (p + 1)) != 0;
        for (int q = j + 1; q < W && !has_more; ++q) has_more = other[q] != 0;
        const int owner_cmp = has_more ? -1 : 1;
        return a_owns ? owner_cmp : -owner_cmp;
    }
    return 0;
}
// canonical lexicographic order of two MAX2SAT states: benefits as signed integers, ascending variable id (oracle/models.hpp Max2SatRanking)
static int lex_cmp_i32(const uint64_t* a, const uint64_t* b, int W) {
    const int32_t* x = reinterpret_cast<const int32_t*>(a); const int32_t* y = reinterpret_cast<const int32_t*>(b);
    for (int j = 0; j < 2 * W; ++j) if (x[j] != y[j]) return x[j] < y[j] ? -1 : 1;
    return 0;
}
int NoDupFringe::state_cmp(int a, int b) const {
    if (kind_ == DDO_MODEL_MAX2SAT) return lex_cmp_i32(state(a), state(b), W);
    return lex_cmp(state(a), state(b), W);
}

static inline uint64_t lex_word_host(uint64_t w) {  // ~bitreverse: larger = Greater in BitSet::cmp among equal popcounts
    w = ((w >> 1) & 0x5555555555555555ull) | ((w & 0x5555555555555555ull) << 1);
    w = ((w >> 2) & 0x3333333333333333ull) | ((w & 0x3333333333333333ull) << 2);
    w = ((w >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((w & 0x0F0F0F0F0F0F0F0Full) << 4);
    return ~__builtin_bswap64(w);
}
NoDupFringe::Ent NoDupFringe::make_ent(int id) const {
    const Item& it = hdr(id).it;
    Ent e;
    e.k1 = ((uint64_t)((uint32_t)it.ub ^ 0x80000000u) << 32) | ((uint32_t)it.value ^ 0x80000000u);
    if (kind_ == DDO_MODEL_MAX2SAT) e.k2 = ((uint64_t)(uint32_t)hdr(id).popc << 32) | ((uint64_t)(uint32_t)it.depth << 8);  // (rank, depth)
    else e.k2 = ((uint64_t)(uint16_t)hdr(id).popc << 48) | (lex_word_host(state(id)[0]) >> 16);
    e.id = id; e.ver = hdr(id).ver;
    return e;
}
bool NoDupFringe::ent_less(const Ent& a, const Ent& b) const {
    if (a.k1 != b.k1) return a.k1 < b.k1;
    if (a.k2 != b.k2) return a.k2 < b.k2;
    if (a.id == b.id) return a.ver < b.ver;
    return state_cmp(a.id, b.id) < 0;  // same ub, value and ranking key prefix
}

void NoDupFringe::clear() {  // no_duplicate.rs:168-174
    drop_cold();
    nodes_.clear(); recycle_.clear();
    pending_.clear(); runs_.clear(); live_ = 0;
    for (Shard& sh : shards_) { std::fill(sh.tab.begin(), sh.tab.end(), -1); sh.used = 0; sh.live = 0; }  // the tables keep their size for the next search
}
void NoDupFringe::rehash(Shard& sh, size_t min_cap) {
    size_t cap = 256;
    while (cap < min_cap) cap <<= 1;
    std::vector<int> old;
    old.swap(sh.tab);
    sh.tab.assign(cap, -1); sh.used = 0;
    const size_t mask = cap - 1;
    for (int id : old)
        if (id >= 0) {
            size_t s = hdr(id).hash & mask;
            while (sh.tab[s] >= 0) s = (s + 1) & mask;
            sh.tab[s] = id; ++sh.used;
        }
}
void NoDupFringe::table_insert(int id) {
    Shard& sh = shards_[shard_of(hdr(id).hash)];
    if ((sh.used + 1) * 2 > sh.tab.size()) rehash(sh, (sh.live + 1) * 4);
    const size_t mask = sh.tab.size() - 1;
    size_t s = hdr(id).hash & mask;
    while (sh.tab[s] >= 0) s = (s + 1) & mask;
    if (sh.tab[s] == -1) ++sh.used;
    sh.tab[s] = id; ++sh.live;
}
int NoDupFringe::table_find(const uint64_t* st, uint64_t h, int32_t depth) const {
    const Shard& sh = shards_[shard_of(h)];
    if (sh.tab.empty()) return -1;
    const size_t mask = sh.tab.size() - 1;
    size_t s = h & mask;
    while (sh.tab[s] != -1) {
        const int id = sh.tab[s];
        if (id >= 0 && hdr(id).hash == h && std::memcmp(state(id), st, (size_t)W * 8) == 0 && (kind_ != DDO_MODEL_MAX2SAT || hdr(id).it.depth == depth)) return id;
        s = (s + 1) & mask;
    }
    return -1;
}
void NoDupFringe::table_erase(int id) {
    Shard& sh = shards_[shard_of(hdr(id).hash)];
    const size_t mask = sh.tab.size() - 1;
    size_t s = hdr(id).hash & mask;
    while (sh.tab[s] != id) s = (s + 1) & mask;
    sh.tab[s] = -2; --sh.live;
}
uint64_t NoDupFringe::key_hash(const uint64_t* st, int32_t depth) const {
    return hash_state(st, W) ^ (kind_ == DDO_MODEL_MAX2SAT ? 0x9E3779B97F4A7C15ull * (uint64_t)(depth + 1) : 0ull);
}
// One push (no_duplicate.rs:88-140) against the shard of `h`.  Touches only that shard's table and the per-node slots of the node it finds
// or of `new_id`, so pushes into different shards may run concurrently.
int NoDupFringe::push_one(const uint64_t* st, uint64_t h, int32_t value, int32_t ub, int32_t depth, int32_t rec, const uint64_t* bits, int nbits_words,
                          int new_id, std::vector<Ent>& pending) {
    const int found = table_find(st, h, depth);
    if (found >= 0) {  // Occupied, no_duplicate.rs:92-118: keep the longer path, ub = max of the known ubs
        const int id = found;
        const int32_t old_lp = hdr(id).it.value, old_ub = hdr(id).it.ub;
        const int32_t merged_ub = std::max(ub, old_ub);
        bool changed = false;
        if (value > old_lp) {
            hdr(id).it = Item{value, merged_ub, depth, rec};
            std::memset(bits_w(id), 0, (size_t)PW * 8);
            std::memcpy(bits_w(id), bits, (size_t)nbits_words * 8);
            changed = true;
        }
        if (ub > old_ub) { hdr(id).it.ub = ub; changed = true; }
        if (changed) { ++hdr(id).ver; pending.push_back(make_ent(id)); }  // re-keyed: the old entry goes stale
        return 0;
    }
    const int id = new_id;  // Vacant, no_duplicate.rs:119-135
    hdr(id).it = Item{value, ub, depth, rec};
    std::memcpy(state_w(id), st, (size_t)W * 8);
    std::memset(bits_w(id), 0, (size_t)PW * 8);
    std::memcpy(bits_w(id), bits, (size_t)nbits_words * 8);
    int pc = 0;
    if (kind_ == DDO_MODEL_MAX2SAT) { const int32_t* x = reinterpret_cast<const int32_t*>(st); for (int j = 0; j < 2 * W; ++j) pc += x[j] < 0 ? -x[j] : x[j]; }
    else for (int j = 0; j < W; ++j) pc += __builtin_popcountll(st[j]);
    hdr(id).popc = pc; hdr(id).hash = h;
    ++hdr(id).ver;
    table_insert(id);
    pending.push_back(make_ent(id));
    return 1;
}
void NoDupFringe::push(const uint64_t* st, int32_t value, int32_t ub, int32_t depth, int32_t rec, const uint64_t* bits, int nbits_words) {
    int id;
    const bool fresh = recycle_.empty() || cold_busy();  // no slot is reused while a background sort may still compare its old state
    if (fresh) {
        id = (int)nodes_.count;
        new_node();
    } else id = recycle_.back();
    if (push_one(st, key_hash(st, depth), value, ub, depth, rec, bits, nbits_words, id, pending_)) {
        if (!fresh) recycle_.pop_back();
        ++live_;
    } else if (fresh) recycle_.push_back(id);  // the slot was not needed: keep it for the next push
}
void NoDupFringe::push_many(const std::vector<PushRec>& recs) {
    const size_t n = recs.size();
    const int T = (int)std::min(16u, std::max(2u, std::thread::hardware_concurrency()));
    if (n < 8192) { for (const PushRec& r : recs) push(r.state, r.value, r.ub, r.depth, r.rec, r.bits, r.nbits_words); return; }
    for (Run& r : runs_) if (r.cold) join_cold(r);  // the burst reuses recycled slots
    // hashes, then the records of every shard in record order (stable counting sort)
    static const bool prof = std::getenv("DDO_FRINGE_PROF") != nullptr; double tp0 = now_ms(), tp1 = 0, tp2 = 0, tp3 = 0, tp4 = 0;
    std::vector<uint64_t> h(n);
    {
        std::vector<std::thread> ts;
        for (int w = 0; w < T; ++w)
            ts.emplace_back([&, w] { for (size_t i = n * w / T; i < n * (w + 1) / T; ++i) h[i] = key_hash(recs[i].state, recs[i].depth); });
        for (auto& t : ts) t.join();
    }
    tp1 = now_ms();
    std::vector<uint32_t> start(NS + 1, 0), order(n);
    for (size_t i = 0; i < n; ++i) ++start[shard_of(h[i]) + 1];
    for (int s = 0; s < NS; ++s) start[s + 1] += start[s];
    {
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for (size_t i = 0; i < n; ++i) order[fill[shard_of(h[i])]++] = (uint32_t)i;
    }
    // tentative node slot of every record, handed out in SHARD order so that the slots one thread writes are contiguous (no cache lines
    // shared between threads): recycled slots first, then fresh ones; a record that hits an existing node leaves its slot unused
    tp2 = now_ms();
    std::vector<int> slot(n);
    const size_t nrec = std::min(n, recycle_.size());
    const size_t base = nodes_.count, fresh = n - nrec;
    for (size_t q = 0; q < n; ++q) slot[order[q]] = q < fresh ? (int)(base + q) : recycle_[recycle_.size() - 1 - (q - fresh)];
    recycle_.resize(recycle_.size() - nrec);
    for (size_t i = 0; i < fresh; ++i) nodes_.grow();  // (their headers are initialised by the thread that fills them: no pass over cold memory here)
    tp3 = now_ms();
    std::vector<std::vector<Ent>> pend(T);
    std::vector<std::vector<int>> unused(T);
    std::vector<size_t> added(T, 0);
    {
        std::vector<std::thread> ts;
        for (int w = 0; w < T; ++w)
            ts.emplace_back([&, w] {
                pend[w].reserve(n / T + 64);
                size_t add = 0;  // thread-local: the per-thread result slots share cache lines
                for (int s = w; s < NS; s += T) {
                    Shard& sh = shards_[s];  // one growth step for the whole burst instead of a chain of doublings
                    const size_t want = sh.used + (start[s + 1] - start[s]) + 1;
                    if (want * 2 > sh.tab.size()) rehash(sh, std::max<size_t>(want * 2, (sh.live + (start[s + 1] - start[s]) + 1) * 4));
                    for (uint32_t q = start[s]; q < start[s + 1]; ++q) {
                        const uint32_t i = order[q];
                        const PushRec& r = recs[i];
                        if ((size_t)slot[i] >= base) hdr(slot[i]).ver = 0;
                        if (push_one(r.state, h[i], r.value, r.ub, r.depth, r.rec, r.bits, r.nbits_words, slot[i], pend[w])) ++add;
                        else unused[w].push_back(slot[i]);
                    }
                }
                added[w] = add;
            });
        for (auto& t : ts) t.join();
    }
    tp4 = now_ms();
    for (int w = 0; w < T; ++w) {
        pending_.insert(pending_.end(), pend[w].begin(), pend[w].end());
        recycle_.insert(recycle_.end(), unused[w].begin(), unused[w].end());
        live_ += added[w];
    }
    if (prof) std::fprintf(stderr, "[push_many %zu] hash %.1f  order %.1f  slots %.1f  insert %.1f  gather %.1f ms\n", n, tp1 - tp0, tp2 - tp1, tp3 - tp2, tp4 - tp3, now_ms() - tp4);
}
void NoDupFringe::flush_pending() {
    if (pending_.empty()) return;
    static const bool prof = std::getenv("DDO_FRINGE_PROF") != nullptr; const double tf0 = now_ms(); double tf1 = 0, tf2 = 0; const size_t nf = pending_.size();
    // entries re-keyed or popped since they were queued are dropped first: a stale entry's slot may hold another state by now, and the
    // comparator must never dereference it (the runs below stay sorted by the keys of LIVE nodes only)
    pending_.erase(std::remove_if(pending_.begin(), pending_.end(), [this](const Ent& e) { return e.ver != hdr(e.id).ver; }), pending_.end());
    if (pending_.empty()) return;
    auto less = [this](const Ent& a, const Ent& b) { return ent_less(a, b); };
    tf1 = now_ms();
    Run nr;
    // a burst is split when it holds more than the next two waves can pop from it: 4096 entries of a mid-sized burst (the cutsets of a
    // narrow wave, ~10 000 nodes: 0.5 ms instead of 1.7 ms of sorting on the critical path, a dozen times per solve), 16 384 of a large one
    const size_t kHot = pending_.size() >= (1u << 16) ? kHotLarge : kHotSmall;
    if (pending_.size() < 2 * kHotSmall || !async_sort_) { sort_ents(pending_); nr.v.swap(pending_); }
    else {
        // a wide wave's cutsets (hundreds of thousands of nodes): the next waves only need the best few thousand of them -- select those,
        // sort them now, and let a background thread sort the rest while the device works
        std::nth_element(pending_.begin(), pending_.end() - kHot, pending_.end(), less);
        nr.v.assign(pending_.end() - kHot, pending_.end());
        std::sort(nr.v.begin(), nr.v.end(), less);
        pending_.resize(pending_.size() - kHot);
        nr.cold.reset(new Cold());
        nr.cold->ents.swap(pending_);
        Cold* c = nr.cold.get();
        c->th = std::thread([this, c] { sort_ents(c->ents); c->done.store(true, std::memory_order_release); });
        ++cold_open_;
    }
    pending_.clear();
    tf2 = now_ms();
    if (prof && nf > 10000) std::fprintf(stderr, "[flush %zu] stale filter %.1f  sort %.1f ms\n", nf, tf1 - tf0, tf2 - tf1);
    runs_.push_back(std::move(nr));
    // Keep the runs few (a pop compares the tails of at most eight runs; merging half a million entries more often costs more than that
    // saves): beyond eight, the last two runs that are not being sorted in the background are merged.  The order of the runs is irrelevant
    // (the MaxUB order is strict over live entries).
    while (runs_.size() > 8) {
        size_t k = runs_.size() - 1;
        while (k >= 1 && (runs_[k].cold || runs_[k - 1].cold)) --k;
        if (k < 1) { k = runs_.size() - 1; if (runs_[k].cold) join_cold(runs_[k]); if (runs_[k - 1].cold) join_cold(runs_[k - 1]); }
        std::vector<Ent>& x = runs_[k - 1].v;
        std::vector<Ent>& y = runs_[k].v;
        std::vector<Ent> m;
        m.reserve(x.size() + y.size());
        size_t i = 0, j = 0;
        auto skip_stale = [this](const std::vector<Ent>& v, size_t& q) { while (q < v.size() && v[q].ver != hdr(v[q].id).ver) ++q; };
        for (;;) {  // stale entries are dropped BEFORE they are compared (their slot may have been recycled for another state)
            skip_stale(x, i); skip_stale(y, j);
            if (i == x.size() && j == y.size()) break;
            const bool take_x = j == y.size() || (i < x.size() && !ent_less(y[j], x[i]));
            m.push_back(take_x ? x[i++] : y[j++]);
        }
        runs_[k - 1].v.swap(m);
        runs_.erase(runs_.begin() + (long)k);
    }
}
void NoDupFringe::sort_ents(std::vector<Ent>& v) const {
    auto less = [this](const Ent& a, const Ent& b) { return ent_less(a, b); };
    if (v.size() < (1u << 15)) { std::sort(v.begin(), v.end(), less); return; }
    constexpr int T = 16;  // sixteen slices on as many threads, then pairwise merges
    const size_t n = v.size();
    size_t cut[T + 1];
    for (int i = 0; i <= T; ++i) cut[i] = n * (size_t)i / T;
    {
        std::vector<std::thread> ts;
        for (int i = 0; i < T; ++i) ts.emplace_back([&, i] { std::sort(v.begin() + cut[i], v.begin() + cut[i + 1], less); });
        for (auto& t : ts) t.join();
    }
    for (int step = 1; step < T; step *= 2) {
        std::vector<std::thread> ts;
        for (int i = 0; i + step < T; i += 2 * step)
            ts.emplace_back([&, i, step] { std::inplace_merge(v.begin() + cut[i], v.begin() + cut[i + step], v.begin() + cut[std::min(i + 2 * step, T)], less); });
        for (auto& t : ts) t.join();
    }
}
void NoDupFringe::join_cold(Run& run) {
    Cold& c = *run.cold;
    c.th.join();
    --cold_open_;
    c.ents.insert(c.ents.end(), run.v.begin(), run.v.end());  // every cold entry is below every entry of the sorted part
    run.v.swap(c.ents);
    run.cold.reset();
}
bool NoDupFringe::cold_busy() {
    if (cold_open_ == 0) return false;
    for (Run& r : runs_) if (r.cold && r.cold->done.load(std::memory_order_acquire)) join_cold(r);
    return cold_open_ > 0;
}
void NoDupFringe::drop_cold() {
    for (Run& r : runs_) if (r.cold) { r.cold->th.join(); r.cold.reset(); }
    cold_open_ = 0;
}
std::vector<NoDupFringe::Ent>& NoDupFringe::tail_run(size_t r) {
    Run& run = runs_[r];
    for (;;) {
        while (!run.v.empty() && run.v.back().ver != hdr(run.v.back().id).ver) run.v.pop_back();  // stale
        if (!run.v.empty() || !run.cold) return run.v;
        join_cold(run);
    }
}
int NoDupFringe::pop() {
    if (live_ == 0) return -1;
    flush_pending();
    int best_run = -1;
    for (size_t r = 0; r < runs_.size(); ++r) {
        auto& run = tail_run(r);
        if (run.empty()) continue;
        if (best_run < 0 || ent_less(runs_[best_run].v.back(), run.back())) best_run = (int)r;
    }
    auto& brun = runs_[best_run].v;
    const int id = brun.back().id;
    brun.pop_back();
    if (brun.size() >= 8) {  // the next pops most likely come from the same run: start fetching their node records and states
        const int nid = brun[brun.size() - 8].id;
        prefetch(nid);
    }
    ++hdr(id).ver;  // any other entry of this node is now stale
    recycle_.push_back(id);
    table_erase(id);
    --live_;
    return id;
}
// The next `k` nodes in pop() order at once (the workload of a wave, parallel.rs:500-559).  Knowing the ids ahead lets every random access
// -- version words during the selection, then hash, index slot, node record, state and path bits -- be prefetched a few nodes ahead
// instead of being paid as a chain of cache misses per pop.
int NoDupFringe::pop_many(int k, std::vector<int>& ids, const PopOut* out) {
    ids.clear();
    if (out) { out->items->clear(); out->states->clear(); out->bits->clear(); }
    if (live_ == 0 || k <= 0) return 0;
    flush_pending();
    static double acc1 = 0, acc2 = 0; static bool reg = false;
    if (!reg && std::getenv("DDO_FRINGE_PROF")) { reg = true; std::atexit([] { std::fprintf(stderr, "[pop_many] select %.1f ms, erase + copy %.1f ms\n", acc1, acc2); }); }
    const double tpm0 = now_ms();
    const size_t want = std::min<size_t>((size_t)k, live_);
    const size_t R = runs_.size();
    // 1. selection = a merge of the runs' live entries from their tails down (what pop() picks one by one).  Every run keeps a short
    //    look-ahead of validated (live) positions, refilled a chunk at a time: the version loads of a chunk are independent, so their
    //    cache misses overlap (a pop()-style loop pays one miss per pop behind an unpredictable branch).
    constexpr size_t CHUNK = 48;
    if (cand_.size() < R) cand_.resize(R);
    head_.assign(R, 0); scan_.resize(R);
    for (size_t r = 0; r < R; ++r) { cand_[r].clear(); scan_[r] = runs_[r].v.size(); }
    auto refill = [&](size_t r) -> bool {  // false: the run has no live entry left
        std::vector<uint32_t>& c = cand_[r];
        for (;;) {
            const std::vector<Ent>& v = runs_[r].v;
            size_t q = scan_[r];
            const size_t stop = c.size() + CHUNK;
            while (q > 0 && c.size() < stop) {
                --q;
                if (q >= 16) __builtin_prefetch(nodes_.at(v[q - 16].id));
                if (v[q].ver == hdr(v[q].id).ver) c.push_back((uint32_t)q);
            }
            scan_[r] = q;
            if (head_[r] < c.size()) return true;
            if (!runs_[r].cold) return false;
            const size_t shift = runs_[r].cold->ents.size();  // the sorted part is used up: the background-sorted part joins below it
            join_cold(runs_[r]);
            for (uint32_t& x : c) x += (uint32_t)shift;
            scan_[r] = shift;
        }
    };
    // runs ordered by their heads, best first; after a pop only the run it came from moves (usually it stays in front)
    int order[64]; int no = 0;
    auto head_ent = [&](int r) -> const Ent& { return runs_[r].v[cand_[r][head_[r]]]; };
    for (size_t r = 0; r < R && r < 64; ++r) {
        if (!refill(r)) continue;
        int p = no++;
        while (p > 0 && ent_less(head_ent(order[p - 1]), head_ent((int)r))) { order[p] = order[p - 1]; --p; }
        order[p] = (int)r;
    }
    ids.reserve(want);
    while (ids.size() < want && no > 0) {
        const int r = order[0];
        ids.push_back(head_ent(r).id);
        ++head_[r];
        if (head_[r] >= cand_[r].size() && !refill((size_t)r)) { for (int p = 1; p < no; ++p) order[p - 1] = order[p]; --no; continue; }
        int p = 0;
        while (p + 1 < no && ent_less(head_ent(r), head_ent(order[p + 1]))) { order[p] = order[p + 1]; ++p; }
        order[p] = r;
    }
    for (size_t r = 0; r < R; ++r) if (head_[r] > 0) runs_[r].v.resize(cand_[r][head_[r] - 1]);  // the taken entries and the stale ones above them
    const size_t n = ids.size();
    const double tpm1 = now_ms(); acc1 += tpm1 - tpm0;
    struct Fin { double& a; double t; ~Fin() { a += now_ms() - t; } } fin{acc2, tpm1};
    // 3. per node: version bump (any other entry of the node is stale from here on), index erase, slot recycling, and the copy of its record
    //    for the caller -- every random access prefetched a few nodes ahead
    if (out) { out->items->reserve(n); out->states->reserve(n * (size_t)W); out->bits->reserve(n * (size_t)out->bits_words); }
    constexpr size_t D1 = 24, D2 = 12;
    for (size_t i = 0; i < n; ++i) {
        if (i + D1 < n) __builtin_prefetch(nodes_.at(ids[i + D1]), 1);  // header: item, hash, version
        if (i + D2 < n) {
            const uint64_t h = hdr(ids[i + D2]).hash;
            const Shard& sh = shards_[shard_of(h)];
            __builtin_prefetch(&sh.tab[h & (sh.tab.size() - 1)], 1);
            prefetch(ids[i + D2]);
        }
        const int id = ids[i];
        ++hdr(id).ver;
        recycle_.push_back(id);
        table_erase(id);
        if (out) {
            out->items->push_back(hdr(id).it);
            out->states->insert(out->states->end(), state_w(id), state_w(id) + W);
            out->bits->insert(out->bits->end(), bits_w(id), bits_w(id) + out->bits_words);
        }
    }
    live_ -= n;
    return (int)n;
}
// DDO_FRINGE_PROF: wall-clock accounting of a solve by host phase (diagnostics; printed at the end of maximize())
namespace {
struct PhaseProf { double f_stage = 0, f_launch = 0, f_prepop = 0, f_wait = 0, f_res = 0, f_prep = 0; double pop = 0, small_wall = 0, small_dev = 0, stage = 0, gen_wall = 0, gen_dev = 0, fetch = 0, collect = 0, enqueue = 0, waves_wall = 0; };
PhaseProf g_prof;
struct PhaseTimer { double& acc; double t0; explicit PhaseTimer(double& a) : acc(a), t0(now_ms()) {} ~PhaseTimer() { acc += now_ms() - t0; } };
}
// ---------------------------------------------------------------------------------------------------------------
// Solver
// ---------------------------------------------------------------------------------------------------------------
Solver::Solver(Engine* e, int kind, const uint64_t* rs, int64_t rv, int wk, uint64_t w, int ws)
    : eng(e), model_kind(kind), n_vars(e->n_vars), words(e->abi_words), root_state(rs, rs + e->abi_words), root_value(rv), width_kind(wk), width(w),
      wave_size(ws), fringe(e->abi_words, (e->n_vars + 63) / 64, kind) {
    if (const char* p = std::getenv("DDO_WAVE_TRACE")) {  // one file per rank under torchrun
        std::string path(p);
        if (const char* r = std::getenv("RANK")) path += std::string(".rank") + r;
        trace_file = std::fopen(path.c_str(), "w");
    }
}
Solver::~Solver() {
    if (trace_file) std::fclose(trace_file);
}

int Solver::init(bool push_root) {  // parallel.rs:368-385
    fringe.clear(); recs.clear(); pre_valid = false; pre_items.clear();
    best_lb = INT64_MIN; best_ub = INT64_MAX; has_sol = false; sol_value = INT64_MIN; best_sol.clear(); aborted = false;
    explored = expanded = transitions = compilations = waves = 0; device_ms = fringe_ms = 0;
    if (push_root) {
        std::vector<uint64_t> bits(1, 0);
        fringe.push(root_state.data(), (int32_t)root_value, INT32_MAX, 0, -1, bits.data(), 0);  // parallel.rs:368-385
    }
    return DDO_OK;
}

void Solver::full_path(int32_t rec, const uint64_t* bits, int32_t depth, std::vector<ddo_decision>& out) const {
    if (rec == -1) return;
    const PathRec& r = recs[rec];
    full_path(r.parent_rec, r.parent_bits.data(), r.base_depth, out);
    const size_t len = std::min<size_t>(r.vars.size(), (size_t)std::max(0, depth - r.base_depth));  // decisions taken inside this DD
    for (size_t i = 0; i < len; ++i) out.push_back(ddo_decision{r.vars[i], eng->bit_value[(bits[i >> 6] >> (i & 63)) & 1]});
}

int Solver::wave(const volatile int32_t* cutoff_flag, int64_t out3[3]) {
    const int rc = wave_body(cutoff_flag, out3);
    // Err(CutoffOccurred) (clean.rs:352-354 -> parallel.rs:479-489 abort_search): the nodes of this wave are gone from the fringe, so the
    // search can no longer prove optimality -- finish() must not close the gap
    if (rc == DDO_CUTOFF) aborted = true;
    return rc;
}

int Solver::wave_body(const volatile int32_t* cutoff_flag, int64_t out3[3]) {
    const int W = words, PWN = (n_vars + 63) / 64;
    // ---- get_workload (parallel.rs:500-559): pop up to wave_size open sub-problems ---------------------------------------------
    double t0 = now_ms();
    const double tr_wave0 = t0;
    int64_t top_ub = INT64_MIN;
    w_states.clear(); w_bits.clear(); w_items.clear();
    if (!pre_valid) {  // (otherwise: the nodes were popped ahead of time, while the device compiled the previous wave)
        const NoDupFringe::PopOut po{&pre_items, &pre_states, &pre_bits, PWN};
        fringe.pop_many(wave_size, pop_ids, &po);
    }
    pre_valid = false;
    for (size_t i = 0; i < pre_items.size(); ++i) {  // re-checked against the current incumbent in pop order
        const NoDupFringe::Item it = pre_items[i];
        const int64_t ub = it.ub == INT32_MAX ? INT64_MAX : it.ub;
        if (ub <= best_lb) { fringe.clear(); break; }  // parallel.rs:531-535
        if (w_items.empty()) top_ub = ub;
        w_states.insert(w_states.end(), &pre_states[i * W], &pre_states[i * W] + W);
        w_bits.insert(w_bits.end(), &pre_bits[i * PWN], &pre_bits[i * PWN] + PWN);
        w_items.push_back(it);
        ++explored;
    }
    pre_items.clear();
    fringe_ms += now_ms() - t0;
    const double tr_pop = now_ms() - t0;
    g_prof.pop += tr_pop;
    PhaseTimer wave_timer(g_prof.waves_wall);
    const double t_prep0 = now_ms();
    out3[1] = top_ub;
    const int cnt = (int)w_items.size();
    if (cnt == 0) { out3[0] = best_lb; out3[2] = 0; return DDO_OK; }
    if (top_ub != INT64_MIN) best_ub = top_ub;
    ++waves;
    std::vector<uint64_t> widths(cnt);
    std::vector<int64_t> values(cnt);
    std::vector<int32_t> depths(cnt);
    for (int i = 0; i < cnt; ++i) {
        const uint64_t unassigned = (uint64_t)(n_vars - w_items[i].depth);  // NbUnassignedWidth, width.rs:397-401 (path.len() == depth)
        switch (width_kind) {
            case DDO_WIDTH_FIXED: widths[i] = width; break;                                                            // width.rs:166-170
            case DDO_WIDTH_TIMES_NB_UNASSIGNED: widths[i] = std::max<uint64_t>(1, width * unassigned); break;            // width.rs:636-641
            case DDO_WIDTH_DIVBY_NB_UNASSIGNED: widths[i] = std::max<uint64_t>(1, unassigned / std::max<uint64_t>(width, 1)); break;  // width.rs:875-880
            default: widths[i] = unassigned; break;
        }
        values[i] = w_items[i].value; depths[i] = w_items[i].depth;
    }
    g_prof.f_prep += now_ms() - t_prep0;
    auto past_deadline = [&]() { return (deadline_ms > 0 && now_ms() >= deadline_ms) || (cutoff_flag && *cutoff_flag); };  // TimeBudget polled before every device batch
    struct Res { bool exact = false, has = false; int32_t best = 0; };
    std::vector<Res> res(cnt);
    const int64_t lb0 = best_lb;  // every restricted DD of the wave is compiled against this snapshot
    float ms = 0;
    int rc;
    double tr_small = 0, tr_general = 0; const uint64_t tr_steps0 = eng->layer_steps, tr_exp0 = expanded;  // DDO_WAVE_TRACE (diagnostics only)
    int cap = eng->K;  // DDs the general engine compiles in lock-step (set below from the depth of the sub-problems that need it)
    std::vector<uint64_t> w2, s2; std::vector<int64_t> v2; std::vector<int32_t> d2;
    auto stage_subset = [&](const int* idx, int oc) -> int {
        w2.resize(oc); s2.resize((size_t)oc * W); v2.resize(oc); d2.resize(oc);
        for (int j = 0; j < oc; ++j) {
            const int i = idx[j];
            w2[j] = widths[i]; v2[j] = values[i]; d2[j] = depths[i];
            std::memcpy(&s2[(size_t)j * W], &w_states[(size_t)i * W], (size_t)W * 8);
        }
        return eng->stage_roots(oc, w2.data(), s2.data(), v2.data(), d2.data());
    };
    // the DD that improves the incumbent is recompiled alone by the general engine to read its best path (a handful of times per solve)
    auto take_solution = [&](int wave_index, int comp_type, int64_t lb) -> int {
        int r2 = stage_subset(&wave_index, 1);
        if (r2 != DDO_OK) return r2;
        r2 = eng->compile_staged(1, comp_type, lb, cutoff_flag, &ms);
        if (r2 != DDO_OK) return r2;
        device_ms += ms;
        std::vector<ddo_decision> dd(n_vars + 1);
        int32_t len = (int32_t)dd.size();
        r2 = eng->best_solution(0, 1, dd.data(), &len);
        if (r2 != DDO_OK) return r2;
        best_sol.clear();
        full_path(w_items[wave_index].rec, &w_bits[(size_t)wave_index * PWN], w_items[wave_index].depth, best_sol);
        best_sol.insert(best_sol.end(), dd.begin(), dd.begin() + len);
        has_sol = true; sol_value = best_lb;
        return DDO_OK;
    };

    // A general-engine DD that improves on everything seen so far in the wave has its best exact path read while its batch is still
    // resident on the device; only an improver of the fast path (which keeps no paths) has to be recompiled by take_solution.
    std::vector<std::pair<int, std::vector<ddo_decision>>> captured;  // (wave index, decisions from that DD's root), restricted / relaxed
    int32_t run_best_restricted = INT32_MIN, run_best_relaxed = INT32_MIN;
    auto capture = [&](int slot, int wave_index) -> int {
        std::vector<ddo_decision> dd(n_vars + 1);
        int32_t len = (int32_t)dd.size();
        const int r2 = eng->best_solution(slot, 1, dd.data(), &len);
        if (r2 != DDO_OK) return r2;
        dd.resize(len);
        captured.emplace_back(wave_index, std::move(dd));
        return DDO_OK;
    };
    auto use_captured = [&](int wave_index) -> bool {
        for (auto it = captured.rbegin(); it != captured.rend(); ++it)
            if (it->first == wave_index) {
                best_sol.clear();
                full_path(w_items[wave_index].rec, &w_bits[(size_t)wave_index * PWN], w_items[wave_index].depth, best_sol);
                best_sol.insert(best_sol.end(), it->second.begin(), it->second.end());
                has_sol = true; sol_value = best_lb;
                return true;
            }
        return false;
    };

    // ---- 1a. shared-memory fast path: one CTA per sub-problem; DDs that never need a cut are exact and finish here ----------------
    std::vector<int> ov;  // sub-problems that need the general engine (a layer outgrew the fast path)
    if (eng->small_ws > 0) {
        PhaseTimer small_timer(g_prof.small_wall);
        { PhaseTimer pt(g_prof.f_stage); rc = eng->stage_roots(cnt, widths.data(), w_states.data(), values.data(), depths.data()); }
        if (rc != DDO_OK) return rc;
        { PhaseTimer pt(g_prof.f_launch); rc = eng->compile_small_launch(cnt, lb0, eng->small_ws_first > 0 && eng->small_ws_first < eng->small_ws ? eng->small_ws_first : eng->small_ws); }
        if (rc != DDO_OK) return rc;
        if (pipeline) { PhaseTimer pt(g_prof.f_prepop); const double tp = now_ms(); prepop(); fringe_ms += now_ms() - tp; }  // overlaps the device
        { PhaseTimer pt(g_prof.f_wait); rc = eng->compile_small_wait(&ms); }
        if (rc != DDO_OK) return rc;
        PhaseTimer res_timer(g_prof.f_res);
        device_ms += ms; tr_small += ms; g_prof.small_dev += ms;
        for (int i = 0; i < cnt; ++i) {
            const SmallOut& o = eng->h_small[i];
            if (o.status != 0) { ov.push_back(i); continue; }
            res[i].exact = true; res[i].has = o.has_best != 0; res[i].best = o.best_value;
            expanded += o.expanded; transitions += o.transitions; ++compilations;
        }
        // second tier: the DDs that outgrew the first (small, many CTAs per SM) capacity get the full fast-path capacity before they
        // fall back to the layer-by-layer engine
        if (!ov.empty() && eng->small_ws_first > 0 && eng->small_ws_first < eng->small_ws) {
            std::vector<int> ov1;
            ov1.swap(ov);
            rc = stage_subset(ov1.data(), (int)ov1.size());
            if (rc != DDO_OK) return rc;
            rc = eng->compile_small_launch((int)ov1.size(), lb0, eng->small_ws);
            if (rc != DDO_OK) return rc;
            rc = eng->compile_small_wait(&ms);
            if (rc != DDO_OK) return rc;
            device_ms += ms; tr_small += ms; g_prof.small_dev += ms;
            for (size_t j = 0; j < ov1.size(); ++j) {
                const SmallOut& o = eng->h_small[j];
                const int i = ov1[j];
                if (o.status != 0) { ov.push_back(i); continue; }
                res[i].exact = true; res[i].has = o.has_best != 0; res[i].best = o.best_value;
                expanded += o.expanded; transitions += o.transitions; ++compilations;
            }
        }
    } else {
        for (int i = 0; i < cnt; ++i) ov.push_back(i);
    }
    if (pre_valid) {  // keep the speculation only if this wave cannot change the fringe or the incumbent
        bool keep = ov.empty();
        for (int i = 0; i < cnt && keep; ++i) if (res[i].has && (int64_t)res[i].best > best_lb) keep = false;
        if (!keep) { const double tp = now_ms(); unpop(); fringe_ms += now_ms() - tp; }
    }
    // ---- 1b. restriction with the general engine (parallel.rs:396-423) --------------------------------------------------------------
    // Dual mode: the relaxed twin of a restricted DD forks on the device at the first width cut and advances in the same launches, so an
    // inexact sub-problem costs one pass of ~n layers instead of two.  Both twins see lb0; the relaxed results are kept only if the
    // restricted DDs of this wave did not improve the incumbent (lb1 == lb0, the common case) -- otherwise they are recompiled against lb1.
    struct Pending { int wave_index, lel, first, count; };
    struct TwinRes { int wave_index; unsigned long long expanded, transitions; bool has; int32_t best; };
    std::vector<Pending> pend;
    std::vector<TwinRes> twins;
    p_states.clear(); p_bits.clear(); p_val.clear(); p_ub.clear(); p_vars.clear(); p_tt.clear();
    const bool frontier = eng->cutset_type == DDO_FRONTIER;  // cutset nodes of one DD sit in different layers (clean.rs:586-606)
    std::vector<int64_t> caps, lbs;
    std::vector<int32_t> vars;
    // collects the cutset records of the last relaxed batch: slot -> wave index through `slot_wave`
    static const bool fringe_prof = std::getenv("DDO_FRINGE_PROF") != nullptr;
    bool p_direct = false;  // the wave's only relaxed batch: its records are pushed straight from the engine's pinned drain buffer (no staging copy)
    auto collect_drain = [&](int slots, const std::vector<int>& slot_wave) -> int {
        int pw = 1;
        PhaseTimer collect_timer(g_prof.collect);
        const double tc0 = now_ms();
        const int total = eng->drain_all(slots, caps.data(), lbs.data(), &pw);
        if (total < 0) return total;
        const double tc1 = now_ms();
        if (total > 0) { const int r2 = eng->fetch_vars_all(slots, vars); if (r2 != DDO_OK) return r2; }
        const size_t base = p_val.size();
        p_val.resize(base + total); p_ub.resize(base + total); p_tt.resize(base + total);
        p_bits.resize((base + total) * (size_t)PWN, 0ull);
        if (!p_direct) p_states.resize((base + total) * (size_t)W);
        // the DD boundaries (records arrive grouped by DD slot), sequentially: one Pending per DD that drained something
        int cur_dd = -1;
        for (int r = 0; r < total; ++r) {
            const int j = eng->h_out_dd[r];
            if (j != cur_dd) {
                cur_dd = j;
                const int lel = frontier ? std::max(0, eng->h_ctl[j].t_term) : eng->h_ctl[j].lel;  // layers whose variables the paths may use
                pend.push_back(Pending{slot_wave[j], lel, (int)(base + r), 0});
                p_vars.insert(p_vars.end(), vars.begin() + (size_t)j * eng->Lcur, vars.begin() + (size_t)j * eng->Lcur + lel);
            }
            pend.back().count++;
        }
        // the records themselves: independent copies, on a few threads when the batch drained a wide wave's cutsets
        auto fill = [&](int r0, int r1) {
            for (int r = r0; r < r1; ++r) {
                const size_t q = base + r;
                if (!p_direct) std::memcpy(&p_states[q * W], &eng->h_out_state[(size_t)r * eng->S], (size_t)W * 8);
                std::memcpy(&p_bits[q * PWN], &eng->h_out_path[(size_t)r * pw], (size_t)std::min(pw, PWN) * 8);
                p_val[q] = eng->h_out_val[r]; p_ub[q] = eng->h_out_ub[r];
                p_tt[q] = frontier ? eng->h_out_tt[r] : eng->h_ctl[eng->h_out_dd[r]].lel;
            }
        };
        if (total < 32768) fill(0, total);
        else {
            constexpr int T = 8;
            std::vector<std::thread> ts;
            for (int w = 0; w < T; ++w) ts.emplace_back(fill, (int)((int64_t)total * w / T), (int)((int64_t)total * (w + 1) / T));
            for (auto& t : ts) t.join();
        }
        if (fringe_prof && total > 10000) std::fprintf(stderr, "[collect_drain %d] drain_all %.1f  collect %.1f ms\n", total, tc1 - tc0, now_ms() - tc1);
        return DDO_OK;
    };
    {   // the deeper the sub-problems, the fewer layers their DDs log, the more of them fit the log pool (Engine::slots_for)
        int lneed = 1;
        for (int i : ov) lneed = std::max(lneed, eng->layers_bound(&w_states[(size_t)i * W], depths[i]));
        cap = std::max(1, eng->slots_for(lneed));
    }
    bool dual = eng->dual_enabled && cap >= 2;
    for (int i : ov) if (widths[i] < 1) dual = false;
    const int chunk = dual ? cap / 2 : cap;
    std::vector<int> slot_wave;
    for (size_t s0 = 0; s0 < ov.size(); s0 += (size_t)chunk) {
        const int oc = (int)std::min<size_t>((size_t)chunk, ov.size() - s0);
        if (past_deadline()) return DDO_CUTOFF;
        { PhaseTimer pt(g_prof.stage); rc = stage_subset(&ov[s0], oc); }
        if (rc != DDO_OK) return rc;
        { PhaseTimer pt(g_prof.gen_wall); rc = dual ? eng->compile_dual(oc, lb0, cutoff_flag, &ms) : eng->compile_staged(oc, DDO_RESTRICTED, lb0, cutoff_flag, &ms); }
        if (rc != DDO_OK) return rc;
        device_ms += ms; tr_general += ms; g_prof.gen_dev += ms;
        PhaseTimer fetch_timer(g_prof.fetch);
        rc = eng->fetch_ctl(dual ? 2 * oc : oc);
        if (rc != DDO_OK) return rc;
        for (int j = 0; j < oc; ++j) {
            const DDCtl& c = eng->h_ctl[j];
            Res& r = res[ov[s0 + j]];
            r.exact = c.lel < 0; r.has = c.has_best_exact != 0; r.best = c.best_exact_value;
            expanded += c.expanded; transitions += c.transitions; ++compilations;
            if (r.has && (int64_t)r.best > lb0 && r.best > run_best_restricted) {
                run_best_restricted = r.best;
                rc = capture(j, ov[s0 + j]);
                if (rc != DDO_OK) return rc;
            }
        }
        if (dual) {
            caps.assign(2 * oc, 0); lbs.assign(2 * oc, INT64_MAX);
            slot_wave.assign(2 * oc, -1);
            bool any = false;
            for (int j = 0; j < oc; ++j) {
                const DDCtl& c = eng->h_ctl[oc + j];
                if (c.status == ST_WAITING) continue;  // the restricted DD never needed a cut: exact, no relaxation (parallel.rs:421-423)
                const int wi = ov[s0 + j];
                twins.push_back(TwinRes{wi, c.expanded, c.transitions, c.has_best_exact != 0, c.best_exact_value});
                if (c.has_best_exact && (int64_t)c.best_exact_value > lb0 && c.best_exact_value > run_best_relaxed) {
                    run_best_relaxed = c.best_exact_value;
                    rc = capture(oc + j, -1 - wi);  // relaxed twins are filed under -1 - wave index
                    if (rc != DDO_OK) return rc;
                }
                const bool exact = (c.lel < 0) || c.ebpo;
                const int32_t rub = w_items[wi].ub;
                caps[oc + j] = rub == INT32_MAX ? INT64_MAX : rub;
                lbs[oc + j] = exact ? INT64_MAX : lb0;
                slot_wave[oc + j] = wi;
                any = any || !exact;
            }
            if (any) { rc = collect_drain(2 * oc, slot_wave); if (rc != DDO_OK) return rc; }
        }
    }
    {   // maybe_update_best in wave order (parallel.rs:446-453): the first DD reaching the new maximum keeps its solution
        int last = -1;
        for (int i = 0; i < cnt; ++i) if (res[i].has && (int64_t)res[i].best > best_lb) { best_lb = res[i].best; last = i; }
        if (last >= 0 && !use_captured(last)) { rc = take_solution(last, DDO_RESTRICTED, lb0); if (rc != DDO_OK) return rc; }
    }
    std::vector<int> open;  // sub-problems whose restricted DD is not exact
    for (int i : ov) if (!res[i].exact) open.push_back(i);

    // ---- 2. relaxation (parallel.rs:425-434) + enqueue_cutset (parallel.rs:456-469) ----------------------------------------------------
    if (!open.empty()) {
        const int64_t lb1 = best_lb;  // every relaxed DD of the wave is compiled against this snapshot
        int improver = -1;
        if (dual && lb1 == lb0) {
            // the twins were compiled against the right bound: adopt them, in wave order
            for (const TwinRes& tw : twins) {
                expanded += tw.expanded; transitions += tw.transitions; ++compilations;
                if (tw.has && (int64_t)tw.best > best_lb) { best_lb = tw.best; improver = tw.wave_index; }
            }
        } else {
            pend.clear(); p_states.clear(); p_bits.clear(); p_val.clear(); p_ub.clear(); p_vars.clear(); p_tt.clear();
            p_direct = open.size() <= (size_t)cap;
            for (size_t s0 = 0; s0 < open.size(); s0 += (size_t)cap) {
                const int oc = (int)std::min<size_t>((size_t)cap, open.size() - s0);
                if (past_deadline()) return DDO_CUTOFF;
                rc = stage_subset(&open[s0], oc);
                if (rc != DDO_OK) return rc;
                rc = eng->compile_staged(oc, DDO_RELAXED, lb1, cutoff_flag, &ms);
                if (rc != DDO_OK) return rc;
                device_ms += ms; tr_general += ms;
                rc = eng->fetch_ctl(oc);
                if (rc != DDO_OK) return rc;
                caps.assign(oc, 0); lbs.assign(oc, 0); slot_wave.assign(oc, -1);
                for (int j = 0; j < oc; ++j) {
                    const DDCtl& c = eng->h_ctl[j];
                    expanded += c.expanded; transitions += c.transitions; ++compilations;
                    if (c.has_best_exact && (int64_t)c.best_exact_value > best_lb) {
                        best_lb = c.best_exact_value; improver = open[s0 + j];
                        rc = capture(j, -1 - improver);
                        if (rc != DDO_OK) return rc;
                    }
                    const bool exact = (c.lel < 0) || c.ebpo;
                    const int32_t rub = w_items[open[s0 + j]].ub;
                    caps[j] = rub == INT32_MAX ? INT64_MAX : rub;
                    lbs[j] = exact ? INT64_MAX : best_lb;  // a lower bound on the final filter; re-applied below once the wave is complete
                    slot_wave[j] = open[s0 + j];
                }
                rc = collect_drain(oc, slot_wave);
                if (rc != DDO_OK) return rc;
            }
        }
        if (improver >= 0) {
            const int wi = improver;
            bool got = false;
            for (auto it = captured.rbegin(); it != captured.rend() && !got; ++it)
                if (it->first == -1 - wi) {
                    best_sol.clear();
                    full_path(w_items[wi].rec, &w_bits[(size_t)wi * PWN], w_items[wi].depth, best_sol);
                    best_sol.insert(best_sol.end(), it->second.begin(), it->second.end());
                    has_sol = true; sol_value = best_lb; got = true;
                }
            if (!got) { rc = take_solution(improver, DDO_RELAXED, lb1); if (rc != DDO_OK) return rc; }
        }
        t0 = now_ms();
        size_t var_off = 0;
        push_recs.clear();
        for (const Pending& pd : pend) {
            PathRec pr;
            pr.parent_rec = w_items[pd.wave_index].rec;
            pr.parent_bits.assign(&w_bits[(size_t)pd.wave_index * PWN], &w_bits[(size_t)pd.wave_index * PWN] + PWN);
            pr.vars.assign(p_vars.begin() + var_off, p_vars.begin() + var_off + pd.lel);
            pr.base_depth = w_items[pd.wave_index].depth;
            var_off += pd.lel;
            int rec_id = -1;
            for (int q = 0; q < pd.count; ++q) {
                const size_t r = (size_t)pd.first + q;
                if ((int64_t)p_ub[r] <= best_lb) continue;  // parallel.rs:461 with the final incumbent of the wave
                if (rec_id < 0) { recs.push_back(pr); rec_id = (int)recs.size() - 1; }
                push_recs.push_back(NoDupFringe::PushRec{p_direct ? &eng->h_out_state[r * (size_t)eng->S] : &p_states[r * W], &p_bits[r * PWN], p_val[r], p_ub[r],
                                                         w_items[pd.wave_index].depth + p_tt[r], rec_id, (p_tt[r] + 63) / 64});
            }
        }
        const double tq = now_ms();
        PhaseTimer enq_timer(g_prof.enqueue);
        fringe.push_many(push_recs);  // enqueue_cutset (parallel.rs:456-469) in wave order
        fringe_ms += now_ms() - t0;
        if (fringe_prof && push_recs.size() > 10000) std::fprintf(stderr, "[enqueue %zu] build %.1f  push_many %.1f ms\n", push_recs.size(), tq - t0, now_ms() - tq);
    }
    if (trace_file)
        std::fprintf(trace_file, "%llu %d %zu %zu %.3f %.3f %llu %llu %zu %.3f %.3f\n", (unsigned long long)waves, cnt, ov.size(), open.size(), tr_small, tr_general,
                     (unsigned long long)(eng->layer_steps - tr_steps0), (unsigned long long)(expanded - tr_exp0), fringe.len(), tr_pop, now_ms() - tr_wave0);
    out3[0] = best_lb;
    out3[2] = (fringe.empty() && !pre_valid) ? 0 : 1;
    return DDO_OK;
}

void Solver::prepop() {
    const int PWN = (n_vars + 63) / 64;
    const NoDupFringe::PopOut po{&pre_items, &pre_states, &pre_bits, PWN};
    fringe.pop_many(wave_size, pop_ids, &po);
    for (size_t i = 0; i < pre_items.size(); ++i) {
        const int64_t ub = pre_items[i].ub == INT32_MAX ? INT64_MAX : pre_items[i].ub;
        if (ub <= best_lb) {  // nothing left can improve on the incumbent (it only grows): safe ahead of time too
            fringe.clear();
            pre_items.resize(i); pre_states.resize(i * (size_t)words); pre_bits.resize(i * (size_t)PWN);
            break;
        }
    }
    pre_valid = !pre_items.empty();
}
void Solver::unpop() {
    const int W = words, PWN = (n_vars + 63) / 64;
    for (size_t i = 0; i < pre_items.size(); ++i) {
        const NoDupFringe::Item& it = pre_items[i];
        fringe.push(&pre_states[i * W], it.value, it.ub, it.depth, it.rec, &pre_bits[i * PWN], PWN);
    }
    pre_items.clear(); pre_valid = false;
}

void Solver::finish() { if (fringe.empty() && !aborted) best_ub = best_lb; }  // parallel.rs:512-515

int Solver::maximize(double time_budget_s, uint64_t max_waves, int32_t* is_exact, int32_t* has_value, int64_t* best_value) {  // parallel.rs:573-607
    int rc = init(true);
    if (rc != DDO_OK) return rc;
    const double t_end = time_budget_s > 0 ? now_ms() + time_budget_s * 1000.0 : 0;
    deadline_ms = t_end;
    volatile int32_t cutoff = 0;
    pipeline = true; pre_valid = false;
    for (;;) {
        if (fringe.empty() && !pre_valid) break;
        if ((max_waves && waves >= max_waves) || (t_end > 0 && now_ms() >= t_end)) { aborted = true; break; }  // TimeBudget, cutoff.rs:302-323
        int64_t o3[3];
        rc = wave(&cutoff, o3);
        if (rc == DDO_CUTOFF) { aborted = true; break; }
        if (rc != DDO_OK) return rc;
    }
    pipeline = false; pre_valid = false; deadline_ms = 0;
    if (std::getenv("DDO_FRINGE_PROF")) {
        const PhaseProf& q = g_prof;
        std::fprintf(stderr, "[solve] waves wall %.1f (+pop %.1f) | fast path wall %.1f (device %.1f) | general: stage %.1f, compile wall %.1f (device %.1f), fetch+capture(+collect) %.1f, "
                     "collect %.1f, enqueue (push_many only) %.1f ms | fast path: widths %.1f stage %.1f launch %.1f prepop %.1f wait %.1f results(+tier 2) %.1f\n", q.waves_wall, q.pop, q.small_wall, q.small_dev, q.stage, q.gen_wall, q.gen_dev, q.fetch, q.collect, q.enqueue, q.f_prep, q.f_stage, q.f_launch, q.f_prepop, q.f_wait, q.f_res);
        g_prof = PhaseProf{};
    }
    if (aborted) fringe.clear();  // abort_search, parallel.rs:479-489
    else best_ub = best_lb;
    std::stable_sort(best_sol.begin(), best_sol.end(), [](const ddo_decision& a, const ddo_decision& b) { return a.variable < b.variable; });  // parallel.rs:605
    if (is_exact) *is_exact = !aborted;
    if (has_value) *has_value = has_sol;
    if (best_value) *best_value = has_sol ? sol_value : 0;
    return DDO_OK;
}

// Initial deal of the open sub-problems across ranks: every rank compiled the same root DD, so each one simply
// keeps its share of the common MaxUB order -- no data-path collective.  The deal rotates (node idx goes to rank (idx + idx / nranks)
// mod nranks): with a plain round-robin rank 0 receives the best node of every group of nranks, which in MaxUB order is the one with the
// largest remaining graph, and carries twice the work of the others (profiles/r02_rank_timeline_8gpu.txt).
int Solver::retain_share(int rank, int nranks) {
    if (nranks <= 1) return DDO_OK;
    if (rank < 0 || rank >= nranks) { set_error("retain_share: bad rank"); return DDO_ERR_INVALID; }
    struct Keep { std::vector<uint64_t> state, bits; NoDupFringe::Item it; };
    std::vector<Keep> keep;
    const int W = words, PWN = (n_vars + 63) / 64;
    for (size_t idx = 0; !fringe.empty(); ++idx) {
        const int id = fringe.pop();
        if ((int)((idx + idx / (size_t)nranks) % (size_t)nranks) != rank) continue;  // rotating: a plain idx % nranks hands rank 0 the best node of every group
        Keep k; k.state.assign(fringe.state(id), fringe.state(id) + W); k.bits.assign(fringe.bits(id), fringe.bits(id) + PWN); k.it = fringe.item(id);
        keep.push_back(std::move(k));
    }
    fringe.clear();
    for (auto& k : keep) fringe.push(k.state.data(), k.it.value, k.it.ub, k.it.depth, k.it.rec, k.bits.data(), PWN);
    return DDO_OK;
}


int Solver::export_open(int max_nodes, int64_t* rows, int32_t* count) {
    if (max_nodes < 0 || !rows || !count) { set_error("export_open: invalid argument"); return DDO_ERR_INVALID; }
    if (pre_valid) unpop();
    const int W = words, PWN = (n_vars + 63) / 64, RW = node_words();
    struct Keep { std::vector<uint64_t> state, bits; NoDupFringe::Item it; };
    std::vector<Keep> keep;
    int out = 0;
    std::vector<ddo_decision> path;
    for (int idx = 0; out < max_nodes && !fringe.empty(); ++idx) {
        const int id = fringe.pop();
        const NoDupFringe::Item it = fringe.item(id);
        if (idx & 1) {  // every other node of the MaxUB order leaves
            int64_t* r = rows + (size_t)out * RW;
            std::memset(r, 0, (size_t)RW * 8);
            r[0] = it.value; r[1] = it.ub == INT32_MAX ? INT64_MAX : it.ub; r[2] = it.depth;
            std::memcpy(r + 3, fringe.state(id), (size_t)W * 8);
            path.clear();
            full_path(it.rec, fringe.bits(id), it.depth, path);
            if ((int)path.size() != it.depth) { set_error("export_open: inconsistent path"); return DDO_ERR_INVALID; }
            uint16_t* d16 = reinterpret_cast<uint16_t*>(r + 3 + W);
            for (int j = 0; j < it.depth; ++j) d16[j] = (uint16_t)(path[j].variable | (path[j].value == eng->bit_value[1] ? 0x8000 : 0));
            ++out;
        } else {
            Keep k; k.state.assign(fringe.state(id), fringe.state(id) + W); k.bits.assign(fringe.bits(id), fringe.bits(id) + PWN); k.it = it;
            keep.push_back(std::move(k));
        }
    }
    for (auto& k : keep) fringe.push(k.state.data(), k.it.value, k.it.ub, k.it.depth, k.it.rec, k.bits.data(), PWN);
    *count = out;
    return DDO_OK;
}

int Solver::import_open(int count, const int64_t* rows) {
    if (count < 0 || (count > 0 && !rows)) { set_error("import_open: invalid argument"); return DDO_ERR_INVALID; }
    if (pre_valid) unpop();  // the imported nodes compete with the ones popped ahead of time
    const int W = words, PWN = (n_vars + 63) / 64, RW = node_words();
    std::vector<uint64_t> bits(PWN);
    for (int i = 0; i < count; ++i) {
        const int64_t* r = rows + (size_t)i * RW;
        const int d = (int)r[2];
        if (d < 0 || d > n_vars) { set_error("import_open: bad depth"); return DDO_ERR_INVALID; }
        if (r[1] <= best_lb) continue;  // parallel.rs:460-461: nothing to gain from this node any more
        PathRec pr; pr.parent_rec = -1; pr.base_depth = 0; pr.vars.resize(d);
        std::fill(bits.begin(), bits.end(), 0ull);
        const uint16_t* d16 = reinterpret_cast<const uint16_t*>(r + 3 + W);
        for (int j = 0; j < d; ++j) {
            pr.vars[j] = d16[j] & 0x7FFF;
            if (d16[j] & 0x8000) bits[j >> 6] |= 1ull << (j & 63);
        }
        recs.push_back(std::move(pr));
        fringe.push(reinterpret_cast<const uint64_t*>(r + 3), (int32_t)r[0], r[1] >= INT32_MAX ? INT32_MAX : (int32_t)r[1], d, (int)recs.size() - 1, bits.data(), PWN);
    }
    return DDO_OK;
}

// The fringe-sharded search as ONE native call (the protocol of ddo_b200/sharded.py, which drives CPU stand-ins in
// the tests): root DD on every rank, deterministic deal, then per wave one all-gather of {best_lb, bound of the best open node, open nodes,
// objective of the held solution} and -- only when the fringes are out of balance -- point-to-point hand-offs of packed open nodes.  At the
// end the solution travels from the lowest rank that holds one of the optimal value to the others.
// out[8] = best_lb, best_ub, is_exact, waves, collectives, hand-offs, nodes sent, nodes received.
int Solver::maximize_sharded(const ShardComm& cm, double time_budget_s, uint64_t max_waves, bool rebalance, int64_t out[8]) {
    const int world = cm.nranks, rank = cm.rank;
    int rc = init(true);
    if (rc != DDO_OK) return rc;
    const double t_end = time_budget_s > 0 ? now_ms() + time_budget_s * 1000.0 : 0;
    deadline_ms = t_end;
    pipeline = true; pre_valid = false;
    volatile int32_t cutoff = 0;
    int64_t o3[3];
    rc = wave(&cutoff, o3);  // the root DD: identical on every rank
    if (rc != DDO_OK && rc != DDO_CUTOFF) return rc;
    if (pre_valid) unpop();
    rc = retain_share(rank, world);
    if (rc != DDO_OK) return rc;
    int64_t lb = best_lb, top = INT64_MIN, glob_ub = INT64_MAX;
    uint64_t nwaves = 1, colls = 0, handoffs = 0, sent = 0, received = 0;
    std::vector<int64_t> mine(4), all((size_t)4 * world), rows;
    const int RW = node_words();
    bool cut = rc == DDO_CUTOFF;
    for (;;) {
        mine[0] = best_lb; mine[1] = top; mine[2] = (int64_t)open_len(); mine[3] = has_sol ? sol_value : INT64_MIN;
        if (cut) mine[2] = -1;  // a rank that ran out of time stops everybody
        const double tg0 = now_ms();
        rc = cm.allgather(cm.ctx, mine.data(), 4, all.data());  // ---- the ONE collective of the wave
        if (rc != DDO_OK) return rc;
        const double tg1 = now_ms();
        ++colls;
        int64_t g_lb = INT64_MIN, g_top = INT64_MIN, total = 0; bool any_cut = false;
        std::vector<int64_t> lens(world);
        for (int r = 0; r < world; ++r) {
            g_lb = std::max(g_lb, all[4 * r]); g_top = std::max(g_top, all[4 * r + 1]);
            if (all[4 * r + 2] < 0) any_cut = true;
            lens[r] = std::max<int64_t>(0, all[4 * r + 2]); total += lens[r];
        }
        if (g_lb > best_lb) best_lb = g_lb;
        lb = best_lb;
        if (g_top != INT64_MIN) glob_ub = g_top;
        if (any_cut) { aborted = true; break; }
        if (total == 0) break;
        if (rebalance) {
            // the same plan on every rank: the emptiest ranks (below a quarter of the mean) are refilled by the fullest ones, a donor serves
            // one receiver per wave and gives half of what it holds above the mean
            const double mean = (double)total / world;
            std::vector<int> lo(world), hi(world);
            for (int r = 0; r < world; ++r) lo[r] = hi[r] = r;
            std::stable_sort(lo.begin(), lo.end(), [&](int a, int b) { return lens[a] < lens[b]; });
            std::stable_sort(hi.begin(), hi.end(), [&](int a, int b) { return lens[a] > lens[b]; });
            std::vector<char> used(world, 0);
            if (world >= 2 && total >= 2 * world)
                for (int dst : lo) {
                    if ((double)lens[dst] * 4 >= mean) break;
                    for (int src : hi) {
                        if (src == dst || used[src] || (double)lens[src] <= mean || lens[src] < 2 * kMinDonor) continue;
                        const int64_t cnt = (int64_t)std::min<double>({(double)kMaxHandoff, ((double)lens[src] - mean) / 2 + 1, (double)((lens[src] - kMinDonor) / 2),
                                                                       std::max<double>(mean - (double)lens[dst], 1.0)});
                        if (cnt > 0) {
                            used[src] = 1; lens[src] -= cnt; lens[dst] += cnt;
                            if (rank == src) {
                                rows.resize((size_t)cnt * RW);
                                int32_t k = 0;
                                rc = export_open((int)cnt, rows.data(), &k);
                                if (rc != DDO_OK) return rc;
                                int64_t kk = k;
                                rc = cm.send(cm.ctx, &kk, 8, dst);
                                if (rc == DDO_OK && k) rc = cm.send(cm.ctx, rows.data(), (int64_t)k * RW * 8, dst);
                                if (rc != DDO_OK) return rc;
                                ++handoffs; sent += (uint64_t)k;
                            } else if (rank == dst) {
                                int64_t kk = 0;
                                rc = cm.recv(cm.ctx, &kk, 8, src);
                                if (rc != DDO_OK) return rc;
                                if (kk) {
                                    rows.resize((size_t)kk * RW);
                                    rc = cm.recv(cm.ctx, rows.data(), kk * RW * 8, src);
                                    if (rc == DDO_OK) rc = import_open((int)kk, rows.data());
                                    if (rc != DDO_OK) return rc;
                                }
                                ++handoffs; received += (uint64_t)kk;
                            }
                        }
                        break;
                    }
                }
        }
        if (trace_file)  // "G": wave about to run, ms spent in the gather (= waiting for the slowest rank), ms in hand-offs, own open nodes, all open nodes
            std::fprintf(trace_file, "G %llu %.3f %.3f %lld %lld\n", (unsigned long long)nwaves + 1, tg1 - tg0, now_ms() - tg1, (long long)mine[2], (long long)total);
        rc = wave(&cutoff, o3);  // a rank with an empty fringe returns immediately (top = INT64_MIN)
        if (rc == DDO_CUTOFF) { cut = true; top = INT64_MIN; continue; }
        if (rc != DDO_OK) return rc;
        top = o3[1];
        ++nwaves;
        if ((max_waves && nwaves >= max_waves) || (t_end > 0 && now_ms() >= t_end)) cut = true;  // the others learn it from the next gather
    }
    pipeline = false; deadline_ms = 0;
    if (pre_valid) unpop();
    if (aborted) fringe.clear(); else best_ub = best_lb;
    if (aborted && glob_ub != INT64_MAX) best_ub = std::max(glob_ub, best_lb);
    // the solution travels once, from the lowest rank that holds one of the optimal value
    int owner = -1;
    for (int r = 0; r < world && owner < 0; ++r) if (all[4 * r + 3] == best_lb) owner = r;
    if (owner >= 0 && world > 1) {
        std::stable_sort(best_sol.begin(), best_sol.end(), [](const ddo_decision& a, const ddo_decision& b) { return a.variable < b.variable; });
        if (rank == owner) {
            int64_t k = (int64_t)best_sol.size();
            for (int r = 0; r < world; ++r) if (r != owner) {
                rc = cm.send(cm.ctx, &k, 8, r);
                if (rc == DDO_OK && k) rc = cm.send(cm.ctx, best_sol.data(), k * (int64_t)sizeof(ddo_decision), r);
                if (rc != DDO_OK) return rc;
            }
        } else {
            int64_t k = 0;
            rc = cm.recv(cm.ctx, &k, 8, owner);
            if (rc != DDO_OK) return rc;
            best_sol.resize((size_t)k);
            if (k) { rc = cm.recv(cm.ctx, best_sol.data(), k * (int64_t)sizeof(ddo_decision), owner); if (rc != DDO_OK) return rc; }
            has_sol = true; sol_value = best_lb;
        }
    }
    std::stable_sort(best_sol.begin(), best_sol.end(), [](const ddo_decision& a, const ddo_decision& b) { return a.variable < b.variable; });
    out[0] = best_lb; out[1] = best_ub; out[2] = !aborted; out[3] = (int64_t)nwaves; out[4] = (int64_t)colls; out[5] = (int64_t)handoffs;
    out[6] = (int64_t)sent; out[7] = (int64_t)received;
    return DDO_OK;
}

}  // namespace ddo

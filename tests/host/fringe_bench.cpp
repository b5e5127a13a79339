// Host micro-benchmark of the fringe under the burst pattern of a config-2 solve (waves 5-7: the cutsets of a wide wave arrive as one burst
// of 250 000 / 290 000 open nodes, the next wave pops 2048 of them).  Not a test: prints the time of every phase.
//   g++ -O2 -std=c++17 -pthread -I/usr/local/cuda/include tests/host/fringe_bench.cpp -L ddo_b200 -l:libddo_b200.so ... && ./a.out
#include <chrono>
#include <cstdio>
#include <random>
#include <thread>
#include <vector>

#include "../../ddo_b200/csrc/solver.hpp"

using namespace ddo;
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
    const int W = 8, PW = 8;
    NoDupFringe f(W, PW, 0);
    std::mt19937_64 rng(7);
    for (int rep = 0; rep < 3; ++rep) {
        f.clear();
        for (int burst = 0; burst < 2; ++burst) {
            const size_t n = burst == 0 ? 250000 : 290000;
            std::vector<uint64_t> st(n * W), bits(n * PW, 0);
            std::vector<NoDupFringe::PushRec> recs;
            for (size_t i = 0; i < n; ++i) {
                for (int j = 0; j < W; ++j) st[i * W + j] = rng() & rng() & rng() & rng();  // ~31 members of 512
                recs.push_back(NoDupFringe::PushRec{&st[i * W], &bits[i * PW], (int32_t)(rng() % 6), (int32_t)(10 + rng() % 5), 20, 0, 1});
            }
            double t0 = now();
            f.push_many(recs);
            double t1 = now();
            std::vector<int> ids;
            f.pop_many(2048, ids);
            double t2 = now();
            f.pop_many(2048, ids);
            double t3 = now();
            std::printf("rep %d burst %zu: push_many %.1f ms, first pop_many %.1f ms (sort of the burst), second pop_many %.2f ms, len %zu\n", rep, n, t1 - t0, t2 - t1, t3 - t2, f.len());
        }
        // the steady state of a solve: ~250 waves pop 2048 nodes each and copy item, state and path bits (Solver::prepop)
        std::vector<int> ids; std::vector<uint64_t> cs, cb; std::vector<NoDupFringe::Item> ci;
        std::this_thread::sleep_for(std::chrono::milliseconds(400));  // let the background sorts of the bursts finish
        size_t popped = 0; double t0 = now();
        while (!f.empty()) {
            const NoDupFringe::PopOut po{&ci, &cs, &cb, PW};
            f.pop_many(2048, ids, &po);
            popped += ids.size();
        }
        std::printf("rep %d drain: %zu nodes in waves of 2048: %.1f ms = %.0f ns per node\n", rep, popped, now() - t0, (now() - t0) * 1e6 / popped);
    }
    return 0;
}

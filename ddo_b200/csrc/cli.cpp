// cli.cpp -- the reference's example executables on the device engine, over the C ABI only (include/ddo_b200.h).
//
//   misp    <fname> [-t|--threads N] [-d|--duration SECONDS] [-w|--width W]        ddo/examples/misp/main.rs:222-398
//   max2sat -f|--file FILE [-w|--width W] [-t|--timeout SECONDS]                   ddo/examples/max2sat/main.rs:36-110
//
// Same instance formats (DIMACS `p edge` / `n` / `e` lines, main.rs:258-317; `p wcnf` with binary and unit clauses, data.rs:66-110), same
// width / cutoff policy (FixedWidth(w) or NbUnassignedWidth, TimeBudget or NoCutoff), same report on stdout.  `--threads` is accepted and
// ignored (the workers are the DDs of one device batch); device-side knobs are extra long options: --wave-size, --batch-cap, --device,
// --cutset lel|frontier.  There is no CPU fallback: without a CUDA device the program exits with the library's error text.
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/ddo_b200.h"

namespace {

[[noreturn]] void die(const std::string& msg) { std::fprintf(stderr, "error: %s\n", msg.c_str()); std::exit(2); }
void check(int rc, const char* what) { if (rc < 0) die(std::string(what) + ": " + ddo_last_error()); }

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
// tokens of a line; `ok` is false when a numeric field does not parse (the reference's ParseIntError)
std::vector<std::string> split(const std::string& s) { std::istringstream is(s); std::vector<std::string> t; std::string x; while (is >> x) t.push_back(x); return t; }
bool to_i64(const std::string& s, long long* out, bool allow_neg) {
    if (s.empty()) return false;
    size_t i = 0;
    if (s[0] == '-') { if (!allow_neg || s.size() == 1) return false; i = 1; }
    for (size_t j = i; j < s.size(); ++j) if (!std::isdigit((unsigned char)s[j])) return false;
    errno = 0;
    *out = std::strtoll(s.c_str(), nullptr, 10);
    return errno == 0;
}

struct Options {
    std::string file;
    bool has_width = false; unsigned long long width = 0;
    bool has_time = false; unsigned long long seconds = 0;
    int wave_size = 0, batch_cap = 0, device = 0, cutset = DDO_LAST_EXACT_LAYER;
};

// ---- misp/main.rs:258-317 ----------------------------------------------------------------------------------------------------------
struct MispFile { int n = 0; std::vector<int64_t> weight; std::vector<int32_t> src, dst; };
MispFile read_misp(const std::string& fname) {
    std::ifstream f(fname);
    if (!f) die("io error " + fname + ": " + std::strerror(errno));
    MispFile g;
    std::string raw;
    while (std::getline(f, raw)) {
        const std::string line = trim(raw);
        if (line.empty()) continue;
        if (line[0] == 'c' && line.size() > 1 && std::isspace((unsigned char)line[1])) continue;  // ^c\s.*$
        const std::vector<std::string> t = split(line);
        long long a, b;
        if (t[0] == "p" && t.size() == 4 && t[1] == "edge" && to_i64(t[2], &a, false) && to_i64(t[3], &b, false)) {  // ^p\s+edge\s+(\d+)\s+(\d+)$
            g.n = (int)a; g.weight.assign((size_t)a, 1);
            continue;
        }
        if (t[0] == "n" && t.size() >= 3 && to_i64(t[1], &a, false) && to_i64(t[2], &b, true)) {  // ^n\s+(\d+)\s+(-?\d+)
            if (a < 1 || a > g.n) die("ill formed instance");
            g.weight[(size_t)a - 1] = b;
            continue;
        }
        if (t[0] == "e" && t.size() >= 3 && to_i64(t[1], &a, false) && to_i64(t[2], &b, false)) {  // ^e\s+(\d+)\s+(\d+)
            if (a < 1 || b < 1 || a > g.n || b > g.n) die("ill formed instance");
            g.src.push_back((int32_t)a - 1); g.dst.push_back((int32_t)b - 1);
            continue;
        }
        die("ill formed instance");  // Error::Format
    }
    return g;
}

// ---- max2sat/data.rs:66-110 ----------------------------------------------------------------------------------------------------------
struct WcnfFile { int n = 0; std::vector<int64_t> clauses; };  // (weight, x, y) triples; x == y encodes a unit clause
WcnfFile read_wcnf(const std::string& fname) {
    std::ifstream f(fname);
    if (!f) die("io error " + fname + ": " + std::strerror(errno));
    WcnfFile w;
    std::string raw;
    while (std::getline(f, raw)) {
        const std::string line = trim(raw);
        if (line.empty()) continue;
        if (line[0] == 'c' && line.size() > 1 && std::isspace((unsigned char)line[1])) continue;
        const std::vector<std::string> t = split(line);
        long long a, b, c, z;
        if (t[0] == "p" && t.size() >= 4 && t[1] == "wcnf" && to_i64(t[2], &a, false)) { w.n = (int)a; continue; }          // ^p\s+wcnf\s+(\d+)\s+(\d+)
        if (t.size() >= 4 && to_i64(t[0], &a, true) && to_i64(t[1], &b, true) && to_i64(t[2], &c, true) && to_i64(t[3], &z, false) && t[3][0] == '0') {
            w.clauses.insert(w.clauses.end(), {a, b, c});                                                                      // ^(-?\d+)\s+(-?\d+)\s+(-?\d+)\s+0
            continue;
        }
        if (t.size() >= 3 && to_i64(t[0], &a, true) && to_i64(t[1], &b, true) && to_i64(t[2], &z, false) && t[2][0] == '0') {
            w.clauses.insert(w.clauses.end(), {a, b, b});                                                                      // ^(-?\d+)\s+(-?\d+)-?\s+0
            continue;
        }
        // anything else is skipped (data.rs has no Format error)
    }
    return w;
}

void usage(bool misp) {
    if (misp) std::fprintf(stderr, "Usage: misp [OPTIONS] <FNAME>\n  -t, --threads <THREADS>    accepted, ignored (default 8)\n  -d, --duration <DURATION>  time budget in seconds\n"
                                   "  -w, --width <WIDTH>        maximum number of nodes per layer\n");
    else std::fprintf(stderr, "Usage: max2sat [OPTIONS] --file <FILE>\n  -f, --file <FILE>\n  -w, --width <WIDTH>\n  -t, --timeout <TIMEOUT>\n");
    std::fprintf(stderr, "      --wave-size <N>  --batch-cap <N>  --device <ID>  --cutset <lel|frontier>\n");
    std::exit(2);
}

Options parse_args(int argc, char** argv, bool misp) {
    Options o;
    auto num = [&](int& i) -> unsigned long long {
        long long v;
        if (i + 1 >= argc || !to_i64(argv[i + 1], &v, false)) usage(misp);
        ++i;
        return (unsigned long long)v;
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-w" || a == "--width") { o.has_width = true; o.width = num(i); }
        else if (misp && (a == "-t" || a == "--threads")) { (void)num(i); }
        else if (misp && (a == "-d" || a == "--duration")) { o.has_time = true; o.seconds = num(i); }
        else if (!misp && (a == "-t" || a == "--timeout")) { o.has_time = true; o.seconds = num(i); }
        else if (!misp && (a == "-f" || a == "--file")) { if (i + 1 >= argc) usage(misp); o.file = argv[++i]; }
        else if (a == "--wave-size") o.wave_size = (int)num(i);
        else if (a == "--batch-cap") o.batch_cap = (int)num(i);
        else if (a == "--device") o.device = (int)num(i);
        else if (a == "--cutset") {
            if (i + 1 >= argc) usage(misp);
            const std::string c = argv[++i];
            if (c == "lel") o.cutset = DDO_LAST_EXACT_LAYER; else if (c == "frontier") o.cutset = DDO_FRONTIER; else usage(misp);
        }
        else if (a == "-h" || a == "--help") usage(misp);
        else if (misp && a[0] != '-' && o.file.empty()) o.file = a;
        else usage(misp);
    }
    if (o.file.empty()) usage(misp);
    return o;
}

// Solver::gap, src/abstraction/solver.rs:80-93 (f32 arithmetic; 0 / 0 prints NaN like Rust does)
std::string gap_text(int64_t lb, int64_t ub) {
    float g;
    if (ub == INT64_MAX || lb == INT64_MIN) g = 1.0f;
    else {
        const int64_t aub = ub < 0 ? -ub : ub, alb = lb < 0 ? -lb : lb;
        const int64_t u = aub > alb ? aub : alb, l = aub > alb ? alb : aub;
        g = (float)(u - l) / (float)u;
    }
    if (std::isnan(g)) return "NaN";
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.3f", g);
    return buf;
}

}  // namespace

int main(int argc, char** argv) {
    std::string prog = argv[0];
    const size_t slash = prog.find_last_of('/');
    if (slash != std::string::npos) prog = prog.substr(slash + 1);
    const bool misp = prog.find("max2sat") == std::string::npos;
    const Options o = parse_args(argc, argv, misp);

    ddo_model* model = nullptr;
    WcnfFile wcnf;
    int n = 0;
    if (misp) {
        const MispFile g = read_misp(o.file);
        n = g.n;
        check(ddo_model_create_misp(g.n, g.weight.data(), (int64_t)g.src.size(), g.src.data(), g.dst.data(), o.device, &model), "ddo_model_create_misp");
    } else {
        wcnf = read_wcnf(o.file);
        n = wcnf.n;
        check(ddo_model_create_max2sat(wcnf.n, (int64_t)(wcnf.clauses.size() / 3), wcnf.clauses.data(), o.device, &model), "ddo_model_create_max2sat");
    }
    // max_width (main.rs:322-328): FixedWidth(w) or NbUnassignedWidth(nb_variables)
    const int width_kind = o.has_width ? DDO_WIDTH_FIXED : DDO_WIDTH_NB_UNASSIGNED;
    const uint64_t cap = o.has_width ? o.width : (uint64_t)n;
    const int wave = o.wave_size > 0 ? o.wave_size : (misp ? 2048 : 148);
    const int batch = o.batch_cap > 0 ? o.batch_cap : (misp ? std::min(wave, 512) : std::min(wave, 148));
    ddo_mdd* mdd = nullptr;
    check(ddo_mdd_create(model, o.device, cap, batch, o.cutset, &mdd), "ddo_mdd_create");
    ddo_solver* solver = nullptr;
    check(ddo_solver_create(model, mdd, width_kind, o.has_width ? o.width : 0, wave, &solver), "ddo_solver_create");

    const auto start = std::chrono::steady_clock::now();
    int32_t is_exact = 0, has_value = 0;
    int64_t best_value = 0;
    check(ddo_solver_maximize(solver, o.has_time ? (double)o.seconds : 0.0, 0, &is_exact, &has_value, &best_value), "ddo_solver_maximize");
    const double duration = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    const int64_t ub = ddo_solver_best_upper_bound(solver), lb = ddo_solver_best_lower_bound(solver);

    std::vector<ddo_decision> sol((size_t)n + 1);
    int32_t len = (int32_t)sol.size();
    if (has_value) check(ddo_solver_best_solution(solver, sol.data(), &len), "ddo_solver_best_solution"); else len = 0;
    sol.resize((size_t)len);  // already sorted by variable (parallel.rs:605)

    std::printf("Duration:   %.3f seconds\n", duration);
    std::printf("Objective:  %lld\n", has_value ? (long long)best_value : -1LL);
    std::printf("Upper Bnd:  %lld\n", (long long)ub);
    std::printf("Lower Bnd:  %lld\n", (long long)lb);
    std::printf("Gap:        %s\n", gap_text(lb, ub).c_str());
    std::printf("Aborted:    %s\n", is_exact ? "false" : "true");
    std::string out = "[";
    bool first = true;
    if (misp) {  // ids of the vertices taken (main.rs:374-381)
        for (const ddo_decision& d : sol) if (d.value == 1) { out += (first ? "" : ", ") + std::to_string(d.variable); first = false; }
    } else {     // signed literals v(variable) * value, and the weight of the violated clauses (max2sat/main.rs:73-110)
        std::vector<int> val((size_t)n + 1, 0);
        for (const ddo_decision& d : sol) { val[(size_t)d.variable + 1] = d.value; out += (first ? "" : ", ") + std::to_string((long long)(d.variable + 1) * d.value); first = false; }
        long long cost = 0;
        if (has_value) {
            // a repeated clause keeps its last weight (data.rs:99,106)
            std::vector<int64_t>& c = wcnf.clauses;
            const size_t m = c.size() / 3;
            for (size_t i = 0; i < m; ++i) {
                const int64_t x = std::min(c[3 * i + 1], c[3 * i + 2]), y = std::max(c[3 * i + 1], c[3 * i + 2]);
                bool last = true;
                for (size_t j = i + 1; j < m && last; ++j) last = !(std::min(c[3 * j + 1], c[3 * j + 2]) == x && std::max(c[3 * j + 1], c[3 * j + 2]) == y);
                if (!last) continue;
                const bool sx = val[(size_t)std::llabs(x)] * x > 0, sy = val[(size_t)std::llabs(y)] * y > 0;
                if (!sx && !sy) cost += c[3 * i];
            }
        }
        std::printf("Cost:       %lld\n", cost);
    }
    out += "]";
    std::printf("Solution:   %s\n", out.c_str());

    ddo_solver_destroy(solver);
    ddo_mdd_destroy(mdd);
    ddo_model_destroy(model);
    return 0;
}

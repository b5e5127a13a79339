/* ============================================================================
 * ddo_b200.h -- C ABI of the B200-native MDD-compilation engine.
 *
 * Drop-in boundary for the ONE hot path of xgillard/ddo: `DecisionDiagram::compile`
 * (`Mdd::_compile`, ddo/src/implementation/mdd/clean.rs:345-381) plus the solver loop that
 * drives it (`Solver::maximize`, ddo/src/implementation/solver/parallel.rs:573-607).
 * Plain C: opaque handles, plain pointers and sizes, `int` status codes, caller-allocated
 * outputs; no C++/torch types, no unwinding across the boundary.  All pointers are HOST
 * pointers unless the name ends in `_dev`.  The reference itself has no FFI for this path
 * (it is pure Rust); each entry point cites the Rust item a binding would replace -- the
 * Rust-side stub a maintainer would add is shown in INTEGRATION.md.
 *
 * Reference paths are relative to /root/reference/ddo/.
 * ========================================================================== */
#ifndef DDO_B200_H
#define DDO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------- */
#define DDO_OK 0
#define DDO_CUTOFF 1            /* Err(Reason::CutoffOccurred), src/common.rs:108-111; raised at clean.rs:352-354 */
#define DDO_ERR_INVALID (-1)    /* bad argument (the reference panics: e.g. max_width == 0 underflows at clean.rs:827) */
#define DDO_ERR_CUDA (-2)       /* CUDA runtime failure; ddo_last_error() has the text */
#define DDO_ERR_CAPACITY (-3)   /* a layer (Exact compilation) or a batch outgrew the arenas given to ddo_mdd_create */
#define DDO_ERR_UNSUPPORTED (-4)/* feature outside the device model (e.g. value range > 31 bits, n > 1024) */
#define DDO_ERR_NO_DEVICE (-5)  /* no CUDA device: there is NO CPU fallback */

/* ---- enums (values follow the reference) --------------------------------- */
/* CompilationType, src/abstraction/mdd.rs:40-47 */
#define DDO_EXACT 0
#define DDO_RELAXED 1
#define DDO_RESTRICTED 2
/* CutsetType const generic, src/abstraction/mdd.rs:24-28 */
#define DDO_LAST_EXACT_LAYER 1
#define DDO_FRONTIER 2
/* WidthHeuristic, src/implementation/heuristics/width.rs:166-170 (FixedWidth), :397-401 (NbUnassignedWidth) */
#define DDO_MODEL_MISP 0      /* examples/misp/main.rs */
#define DDO_MODEL_MAX2SAT 1   /* examples/max2sat/{model,relax,heuristics}.rs */

#define DDO_WIDTH_FIXED 0
#define DDO_WIDTH_NB_UNASSIGNED 1
#define DDO_WIDTH_TIMES_NB_UNASSIGNED 2   /* Times(width, NbUnassignedWidth(n)), src/implementation/heuristics/width.rs:636-641: max(1, width * (n - depth)) */
#define DDO_WIDTH_DIVBY_NB_UNASSIGNED 3   /* DivBy(width, NbUnassignedWidth(n)), width.rs:875-880: max(1, (n - depth) / width) */

typedef struct ddo_model ddo_model;   /* an immutable problem instance resident in HBM (Problem + Relaxation + StateRanking) */
typedef struct ddo_mdd ddo_mdd;       /* D::default(): one reusable batch of DD workspaces on one GPU (parallel.rs:580)   */
typedef struct ddo_solver ddo_solver; /* ParallelSolver state: fringe, incumbent, bounds (parallel.rs:32-81)              */

/* Decision, src/common.rs:58-64 */
typedef struct ddo_decision {
    int32_t variable;
    int32_t value;
} ddo_decision;

/* Completion, src/common.rs:115-121, plus what DecisionDiagram's getters return (mdd.rs:84-101) */
typedef struct ddo_completion {
    int32_t is_exact;          /* Completion::is_exact == DecisionDiagram::is_exact()  (clean.rs:241-243,377-380) */
    int32_t has_best_value;    /* best_value().is_some()                                                         */
    int64_t best_value;        /* DecisionDiagram::best_value()        (clean.rs:309-311)                        */
    int32_t has_best_exact;    /* best_exact_value().is_some()                                                   */
    int32_t cutset_size;       /* number of sub-problems drain_cutset would emit (MARKED cutset nodes)           */
    int64_t best_exact_value;  /* DecisionDiagram::best_exact_value()  (clean.rs:317-319)                        */
    int32_t lel_depth;         /* depth (root_depth + layer index) of the last exact layer, -1 when never squashed */
    int32_t n_layers;          /* layers compiled, terminal layer included                                       */
    uint64_t expanded;         /* nodes that passed the rough-upper-bound test of clean.rs:365 (the bench metric) */
    uint64_t transitions;      /* calls of Mdd::_branch_on (clean.rs:728)                                         */
} ddo_completion;

/* ---- library ------------------------------------------------------------- */
const char* ddo_last_error(void);           /* thread-local text of the last failure */
int ddo_device_count(void);                 /* number of visible CUDA devices (0 => every create call fails) */
uint64_t ddo_kernel_launches(void);         /* count of this library's kernel launches since load (bench "gpu_launches") */

/* ---- model: examples/misp/main.rs:37-209 (Misp, MispRelax, MispRanking) --- */
/* n vertices, weights[n] (NULL => all 1, main.rs:286), m undirected edges given 0-based (the reader at
 * main.rs:299-307 converts the file's 1-based ids).  Uploads the complement-adjacency rows once (main.rs:40-45). */
int ddo_model_create_misp(int32_t n, const int64_t* weights, int64_t m, const int32_t* edge_src, const int32_t* edge_dst,
                          int device, ddo_model** out);
/* ---- model: examples/max2sat/model.rs:98-349 (Max2Sat), relax.rs:43-89 (Max2SatRelax), heuristics.rs:30-37 (Max2SatRanking) ---- */
/* n variables, m clauses as int64 triples (weight, literal x, literal y); the literal of variable i (0-based) is +-(i+1), x == y encodes a
 * unit clause (data.rs:31-58).  A repeated clause keeps its LAST weight (data.rs:99,106 insert into a hash map).  Uploads the clause-weight
 * rows of every branching variable and the fast_upper_bound tables (model.rs:183-238) once.  A state is n int32 marginal benefits
 * (model.rs:59-62), packed two per uint64 word at this ABI; its depth travels as root_depth.  Decision values: T = 1, F = -1 (model.rs:30-32).
 * Canonical refinements where the reference leaves ties to unstable sorts: variable order ties by variable id (model.rs:149-151); ranking
 * ties (equal sum |benefit|) by depth, then lexicographic signed benefits. */
int ddo_model_create_max2sat(int32_t n, int64_t m, const int64_t* clauses, int device, ddo_model** out);
int32_t ddo_model_kind(const ddo_model*);           /* DDO_MODEL_* */
void ddo_model_destroy(ddo_model*);
int32_t ddo_model_nb_variables(const ddo_model*);   /* Problem::nb_variables, src/abstraction/dp.rs:39 */
int32_t ddo_model_state_words(const ddo_model*);    /* uint64 words of one packed state (bit v of the BitSet = bit v%64 of word v/64) */
int ddo_model_initial_state(const ddo_model*, uint64_t* state_out, int64_t* value_out); /* dp.rs:41-43 / misp main.rs:69-75 */

/* ---- DecisionDiagram: src/abstraction/mdd.rs:75-114 ---------------------- */
/* max_width_cap: largest max_width any compile will ask for; batch_cap: DDs compiled per call (>= 1). */
int ddo_mdd_create(const ddo_model*, int device, uint64_t max_width_cap, int32_t batch_cap, int32_t cutset_type, ddo_mdd** out);
void ddo_mdd_destroy(ddo_mdd*);

/* DecisionDiagram::compile(&CompilationInput) (mdd.rs:81, clean.rs:237-239).  CompilationInput fields (mdd.rs:51-71):
 * comp_type, max_width, residual = (root_state, root_value, root_depth), best_lb; cutoff = *cutoff_flag != 0 polled between
 * layers (Cutoff::must_stop, clean.rs:352); cache / dominance are the Empty* implementations.  The root path stays with the caller. */
int ddo_mdd_compile(ddo_mdd*, int32_t comp_type, uint64_t max_width, const uint64_t* root_state, int64_t root_value, int32_t root_depth,
                    int64_t best_lb, const volatile int32_t* cutoff_flag, ddo_completion* out);
/* The same for `count` independent sub-problems in lock-step (what N worker threads do at parallel.rs:576-602).
 * root_states: count x state_words; max_widths / root_values / root_depths / completions: count entries. */
int ddo_mdd_compile_batch(ddo_mdd*, int32_t count, int32_t comp_type, const uint64_t* max_widths, const uint64_t* root_states,
                          const int64_t* root_values, const int32_t* root_depths, int64_t best_lb, const volatile int32_t* cutoff_flag,
                          ddo_completion* out);
/* best_solution()/best_exact_solution() (mdd.rs:88-101, clean.rs:313-343) of DD `index` of the last batch: decisions from the DD root
 * to the terminal node, in the reference's order (terminal -> root, clean.rs:337-341).  *len in: capacity, out: length.  DDO_ERR_INVALID
 * when the DD has no such solution. */
int ddo_mdd_best_solution(ddo_mdd*, int32_t index, int32_t exact, ddo_decision* out, int32_t* len);
/* drain_cutset (mdd.rs:107-113, clean.rs:417-445) of DD `index` of the last RELAXED batch, SoA instead of a per-node callback.
 * Emits, in cutset order, every MARKED cutset node whose ub' = min(ub, ub_cap) is > lb_filter (the filter of parallel.rs:460-461; pass
 * INT64_MAX / INT64_MIN to disable): states (count x words), values, ubs, and the decisions from the DD root to the node (path_len each,
 * same for all nodes of a LEL cutset; terminal->root order like clean.rs:329-343).  *count in: capacity, out: number emitted. */
/* (LAST_EXACT_LAYER engines; a FRONTIER engine answers DDO_ERR_UNSUPPORTED: its nodes have different depths, see the batch form.) */
int ddo_mdd_drain_cutset(ddo_mdd*, int32_t index, int64_t ub_cap, int64_t lb_filter, uint64_t* states, int64_t* values, int64_t* ubs,
                         int32_t* depth_out, int32_t* path_len_out, ddo_decision* paths, int32_t* count);
/* drain_cutset for DDs 0..count-1 of the last RELAXED batch at once (what the wave solver uses): DD i emits its MARKED cutset nodes with
 * min(ub, ub_caps[i]) > lb_filters[i], concatenated in DD order.  Outputs (caller-allocated, *total in: capacity in records, out: records
 * emitted): states (words each), values, ubs, dd_index (owning DD), path_bits (*path_words uint64 each: bit t = decision taken in layer t
 * of the owning DD; its variables come from ddo_mdd_layer_trace).  Any output pointer may be NULL. */
int ddo_mdd_drain_cutset_batch(ddo_mdd*, int32_t count, const int64_t* ub_caps, const int64_t* lb_filters, uint64_t* states, int64_t* values,
                               int64_t* ubs, int32_t* dd_index, uint64_t* path_bits, int32_t* path_words, int64_t* total);
/* Layer (inside its owning DD) of every record of the last ddo_mdd_drain_cutset_batch: the depth of record r is root_depth[dd_index[r]] +
 * layer_index[r] and its path has layer_index[r] decisions.  A LAST_EXACT_LAYER cutset has one layer per DD (clean.rs:566-583); the nodes of
 * a FRONTIER cutset (clean.rs:586-606: every exact node with an edge into an inexact node) come from different layers and are emitted in the
 * canonical order (layer descending, position in the layer ascending).  Returns the number of records (< 0: error). */
int ddo_mdd_drain_layer_index(ddo_mdd*, int32_t* layer_index, int64_t cap);
/* per-kernel device time (CUDA events around every launch; slows the launch loop, off by default).  kernel ids: 0 k_expand, 1 k_finish,
 * 2 k_compact, 3 k_finalize+k_bottomup, 4 drain kernels, 5 k_small.  Times accumulate until reset (on != 0 also resets). */
int ddo_mdd_set_profiling(ddo_mdd*, int32_t on);
int ddo_mdd_kernel_times(ddo_mdd*, double ms[6], uint64_t launches[6]);
/* per-layer trace of DD `index`: branching variable (Problem::next_variable) and layer width after the cut; returns #layers expanded */
int ddo_mdd_layer_trace(ddo_mdd*, int32_t index, int32_t* vars, int32_t* widths, int32_t cap);

/* Device-resident variant for benchmarking the kernels alone: roots already in HBM (uploaded by ddo_mdd_stage_roots), results left on
 * the device; returns after the stream is idle.  compile_staged(count) == compile_batch without the H2D / D2H copies. */
int ddo_mdd_stage_roots(ddo_mdd*, int32_t count, const uint64_t* max_widths, const uint64_t* root_states, const int64_t* root_values,
                        const int32_t* root_depths);
int ddo_mdd_compile_staged(ddo_mdd*, int32_t count, int32_t comp_type, int64_t best_lb, float* device_ms);
int ddo_mdd_fetch_completions(ddo_mdd*, int32_t count, ddo_completion* out);

/* ---- Solver: src/abstraction/solver.rs:32-97, implementation/solver/parallel.rs:287-641 ----
 * Branch-and-bound over a NoDupFringe (fringe/no_duplicate.rs) ordered by MaxUB (heuristics/subproblem_ranking.rs:86-90); each wave
 * pops up to wave_size open sub-problems and compiles their restricted then relaxed DDs on the device.  wave_size may exceed the mdd's
 * batch_cap: narrow sub-problems are compiled by the shared-memory fast path (one CTA each), the others batch_cap at a time.
 * width_kind / width = the WidthHeuristic (heuristics/width.rs), evaluated per sub-problem on the host: DDO_WIDTH_FIXED (FixedWidth(width)),
 * DDO_WIDTH_NB_UNASSIGNED (width ignored), DDO_WIDTH_TIMES_NB_UNASSIGNED / DDO_WIDTH_DIVBY_NB_UNASSIGNED (Times / DivBy(width, NbUnassignedWidth)). */
int ddo_solver_create(const ddo_model*, ddo_mdd*, int32_t width_kind, uint64_t width, int32_t wave_size, ddo_solver** out);
void ddo_solver_destroy(ddo_solver*);
/* Solver::maximize (solver.rs:56).  time_budget_s <= 0: NoCutoff; else TimeBudget (heuristics/cutoff.rs:302-323). max_waves 0: unlimited */
int ddo_solver_maximize(ddo_solver*, double time_budget_s, uint64_t max_waves, int32_t* is_exact, int32_t* has_value, int64_t* best_value);
/* stepwise form used by the multi-GPU driver (one allreduce(max) between waves): */
int ddo_solver_init(ddo_solver*, int32_t push_root);                   /* parallel.rs:368-374 (root pushed only where push_root != 0) */
/* one wave; out3 = {best_lb, ub of the best open node before the wave (INT64_MIN if none), 1 if work remains after it} */
int ddo_solver_wave(ddo_solver*, const volatile int32_t* cutoff_flag, int64_t out3[3]);
int ddo_solver_set_lower_bound(ddo_solver*, int64_t best_lb);          /* adopt a bound found elsewhere (cf. set_primal, solver.rs:77) */
/* initial deal of the open sub-problems over `nranks` processes (one per GPU): every rank compiled the same root DD; this keeps this
 * rank's share of the common MaxUB order (node i goes to rank (i + i / nranks) mod nranks) and drops the others -- the path shards with no
 * data-path collective. */
int ddo_solver_retain_share(ddo_solver*, int32_t rank, int32_t nranks);
/* Work hand-off between ranks.  The reference's workers pull from ONE shared fringe (parallel.rs:500-559), so its load balances itself;
 * here a rank whose fringe runs dry is refilled by a loaded one.  export_open pops up to 2 * max_nodes of the best open nodes, hands out
 * every other one (donor and receiver keep nodes of the same quality) and re-queues the rest.  A node travels as ddo_solver_node_words()
 * int64 words: value, upper bound (INT64_MAX: none yet), depth, the packed state [state_words], then its full decision path, four
 * decisions per word, 16 bits each (variable | path bit << 15; path bit 1 = YES / T).  import_open queues such nodes (those whose bound no
 * longer beats the incumbent are dropped, parallel.rs:461). */
int32_t ddo_solver_node_words(const ddo_solver*);
int ddo_solver_export_open(ddo_solver*, int32_t max_nodes, int64_t* rows, int32_t* count);
int ddo_solver_import_open(ddo_solver*, int32_t count, const int64_t* rows);
int ddo_solver_finish(ddo_solver*);                                    /* best_ub = best_lb when the fringe is empty (parallel.rs:512-515) */
int64_t ddo_solver_best_lower_bound(const ddo_solver*);                /* solver.rs:83 */
int64_t ddo_solver_best_upper_bound(const ddo_solver*);                /* solver.rs:86 */
int ddo_solver_best_value(const ddo_solver*, int32_t* has, int64_t* value);      /* solver.rs:74 */
int ddo_solver_best_solution(const ddo_solver*, ddo_decision* out, int32_t* len);/* solver.rs:71; sorted by variable (parallel.rs:605) */
uint64_t ddo_solver_explored(const ddo_solver*);                       /* solver.rs:96 */
uint64_t ddo_solver_fringe_len(const ddo_solver*);
/* stats[0..7] = expanded nodes, transitions, compilations, waves, device ms in compile (CUDA events), host ms in fringe,
 * bytes copied host->device and device->host by the engine since it was created */
int ddo_solver_stats(const ddo_solver*, double stats[8]);

/* ---- Collectives between the ranks of a fringe-sharded search (one process per GPU): replaces `Mutex<Critical>` of
 * implementation/solver/parallel.rs:32-81 (best_lb read at :398 / :426, written at :446-453; termination test at :512) across processes.
 * NCCL over NVLink, loaded at run time (libnccl.so.2); buffers are HOST memory, staged through the communicator's device scratch.
 * Bootstrap: rank 0 calls ddo_comm_unique_id and hands the 128 bytes to the other ranks by any out-of-band channel. */
typedef struct ddo_comm ddo_comm;
int32_t ddo_comm_size(const ddo_comm*);
int32_t ddo_comm_rank(const ddo_comm*);
int ddo_comm_unique_id(void* id128);
int ddo_comm_init(int32_t nranks, int32_t rank, const void* id128, int device, ddo_comm** out);
void ddo_comm_destroy(ddo_comm*);
/* one collective per wave: in place max over the ranks of `count` int64 ({best_lb, ub of the best open node, has_work}) ... */
int ddo_comm_allreduce_max(ddo_comm*, int64_t* values, int32_t count);
/* ... or, when the driver also balances the fringes, the same words of EVERY rank: recv[r * count + i] = rank r's values[i] */
int ddo_comm_allgather(ddo_comm*, const int64_t* values, int32_t count, int64_t* recv);
/* point-to-point hand-off of packed open nodes (ddo_solver_export_open / import_open) and of the final solution */
int ddo_comm_send(ddo_comm*, const void* buf, int64_t bytes, int32_t peer);
int ddo_comm_recv(ddo_comm*, void* buf, int64_t bytes, int32_t peer);
/* Solver::maximize of a fringe-sharded search as ONE call per rank (what a Rust host binds): root DD on every rank, deterministic deal,
 * one ddo_comm_allgather per wave, hand-offs of open nodes when `rebalance` != 0, solution gathered from the rank that holds it
 * (ddo_solver_best_value / best_solution afterwards).  out[8] = best_lb, best_ub, is_exact, waves, collectives, hand-offs, nodes sent,
 * nodes received.  Every rank must call it with the same arguments. */
int ddo_solver_maximize_sharded(ddo_solver*, ddo_comm*, double time_budget_s, uint64_t max_waves, int32_t rebalance, int64_t out[8]);

#ifdef __cplusplus
}
#endif
#endif /* DDO_B200_H */

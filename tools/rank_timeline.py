"""Per-rank wave timeline of a fringe-sharded solve from the DDO_WAVE_TRACE files of its ranks (tools/gpu_trace_ranks.sh):
usage: python tools/rank_timeline.py gpurun_out/r02_trace_8gpu 8 > profiles/r02_rank_timeline_8gpu.txt"""
import sys

prefix, N = sys.argv[1], int(sys.argv[2])
ranks = []
for r in range(N):
    gl, wl = [], []
    for line in open(f"{prefix}.rank{r}"):
        p = line.split()
        if p[0] == "G":
            gl.append((int(p[1]), float(p[2]), float(p[3]), int(p[4]), int(p[5])))
        else:
            wl.append((int(p[0]), int(p[1]), int(p[2]), int(p[3]), float(p[4]), float(p[5]), int(p[6]), int(p[7]), int(p[8]), float(p[9]), float(p[10])))
    starts = [i for i, w in enumerate(wl) if w[0] == 1] + [len(wl)]  # wave numbers restart at 1 with every solve
    segs = [wl[a:b] for a, b in zip(starts, starts[1:]) if b - a > 1]   # (bench.py ends with a one-wave solve that sizes the root DD pair)
    wl = segs[-1]
    gl = gl[[i for i, g in enumerate(gl) if g[0] == 2][-1]:]
    ranks.append((wl, gl))
print(f"# {N} ranks, one config-2 solve (MISP G(500,0.5), W = 10 000, waves of 2048 per rank); times in ms")
print("# per rank: waves run, wall clock inside the waves (root DD alone), device time fast path / general engine, time blocked in the per-wave all-gather")
print("#           (= waiting for the slowest rank), time in hand-offs, nodes expanded")
for r, (wl, gl) in enumerate(ranks):
    print(f"rank {r}: waves {len(wl):3d}  wave wall {sum(w[10] for w in wl):6.1f} (root {wl[0][10]:.1f})  device {sum(w[4] for w in wl):5.1f} / {sum(w[5] for w in wl):6.1f}"
          f"  gather wait {sum(g[1] for g in gl):6.1f}  hand-offs {sum(g[2] for g in gl):5.1f}  expanded {sum(w[7] for w in wl) / 1e6:6.1f} M")
print("# global waves (all ranks run wave w between two gathers): wall clock of wave w per rank, '-' = the rank had no open node")
nw = max(len(wl) for wl, _ in ranks)
tot = 0.0
for w in range(1, nw + 1):
    cells, mx = [], 0.0
    for wl, _ in ranks:
        row = [x for x in wl if x[0] == w]
        if row:
            cells.append(f"{row[0][10]:6.1f}{'*' if row[0][2] else ' '}"); mx = max(mx, row[0][10])
        else:
            cells.append("     - ")
    tot += mx
    if mx >= 2.0:
        print(f"wave {w:3d}: " + " ".join(cells) + f"   max {mx:6.1f}")
print(f"# sum over the global waves of the slowest rank's wave: {tot:.1f} ms  (waves under 2 ms are not listed; * = the wave held general-engine DDs)")

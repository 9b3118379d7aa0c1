"""Timing model of one directional sweep of k_sweep_patch (no GPU needed): which term of the makespan is what.

The kernel is a dependency network: tile (U,V) = NU warps x `rows` march steps; warp w of a tile does step a after
  * its own step a-1                                  (+ t_step: one march step of a warp)
  * warp w-1's step a of the same tile                (+ t_hand: tagged shared-memory ring, poll)
  * for warp 0: the last warp's step a of tile (U-1,V) (+ t_mail: global mailbox round trip through the importer)
  * the last lane of tile (U,V-1): its row a + dmf     (+ t_mail)
and a tile starts when a CTA is free (tickets in the kernel's order, `ctas` persistent CTAs).
Every quantity is in microseconds.  The model has no jitter and no back-pressure (ring depth); it is the optimistic
schedule for given constants, to be compared with the measured sweep (1.16 ms at 512^3, 6.4 ms at 1024^3).

usage: pipeline_model.py [N] [--step us] [--stepdep us] [--hand us] [--mail us] [--ctas n] [--nu planes] [--lanes n]
                         [--lagu k] [--lagv k] [--overlap]
  --overlap: the same network for two consecutive sweeps that differ in the i direction only, the second one following
             the first with the dependency of DESIGN.md section 8 item 1 (its tile (U,V) at row a needs the first sweep's
             tiles (nU-1-U, V) and (nU-1-U, V+1) to be `gate` rows further), one ticket list, same CTAs.
"""
import heapq
import sys

import numpy as np


def arg(name, default, cast=float):
    return cast(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


N = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 512
T_STEP = arg("--step", 0.46)      # one march step of a compute warp (900 cycles at 1.965 GHz)
T_HAND = arg("--hand", 0.06)      # ring hand-off between two warps (visibility + poll)
T_MAIL = arg("--mail", 2.0)       # tile to tile through the global mailbox and the importer
CTAS = arg("--ctas", 148, int)
NU = arg("--nu", 8, int)
LANES = arg("--lanes", 128, int)   # lanes of v per tile (128 = 4 per thread; 64 = the 2-lanes-per-thread patch of DESIGN section 8)
LAG_U = arg("--lagu", NU + 4)      # ticket key = U * LAG_U + V * LAG_V (the kernel: NU + 4 and 128)
LAG_V = arg("--lagv", LANES)
T_STEP_DEP = arg("--stepdep", T_STEP)   # step of a tile that imports words (has a U or V predecessor)
GATE = arg("--gate", 24, int)     # rows the first sweep must be ahead of the second one's loader (TMA look-ahead + publish period)
OVERLAP = "--overlap" in sys.argv

nU = (N + NU - 1) // NU
nV = (N + LANES - 1) // LANES
rows = N + LANES - 1              # march steps of a tile (sheared rows that hold a node of the tile)


def tile_schedule(start, up_last, left_last, gate=None, t_step=None):
    """finish times of warp 0 and of the last warp of a tile for all rows, given when the CTA is free (`start`), the last
    warp's times of tile U-1 (`up_last`), of tile V-1 (`left_last`, already shifted to this tile's rows) and an optional
    gate (earliest time per row)."""
    t_prev = np.full(rows, start)           # constraint from the previous warp in the chain (or the inputs of warp 0)
    if up_last is not None:
        t_prev = np.maximum(t_prev, up_last + T_MAIL)
    if left_last is not None:
        t_prev = np.maximum(t_prev, left_last + T_MAIL)
    if gate is not None:
        t_prev = np.maximum(t_prev, gate)
    first = None
    T_STEP = t_step if t_step is not None else globals()["T_STEP"]
    k = np.arange(rows)
    for w in range(NU):
        # t[a] = max(t[a-1], t_prev[a]) + T_STEP  ==  (a+1) T + max(start, max_{k<=a} (t_prev[k] - k T))
        t = (k + 1) * T_STEP + np.maximum(start, np.maximum.accumulate(t_prev - k * T_STEP))
        if w == 0:
            first = t
        t_prev = t + T_HAND
    return first, t


def run(n_sweeps):
    # ticket order of the kernel: estimated start time U * (NU + 4) + first lane, per sweep; the second sweep's tiles are
    # keyed behind the first sweep's tile they wait for
    tickets = []
    for s in range(n_sweeps):
        for U in range(nU):
            for V in range(nV):
                key = U * LAG_U + V * LAG_V
                if s == 1:
                    key += (nU - 1 - U) * LAG_U + GATE + NU + 8   # after first-sweep tile (nU-1-U, V) has got that far
                tickets.append((key, s, U, V))
    tickets.sort()
    free = [0.0] * min(CTAS, len(tickets))
    heapq.heapify(free)
    last = {}
    busy = 0.0
    for key, s, U, V in tickets:
        start = heapq.heappop(free)
        up = last.get((s, U - 1, V))
        left = last.get((s, U, V - 1))
        if left is not None:
            # local row a of tile V is local row a + LANES of tile V-1; it needs that tile's lane 127 of the row before
            sh = np.full(rows, -np.inf)
            sh[:rows - (LANES - 1)] = left[LANES - 1:]
            left = sh
        gate = None
        if s == 1:
            a = last[(0, nU - 1 - U, V)]
            gate = np.concatenate([a[GATE:], np.full(GATE, a[-1])])
            b = last.get((0, nU - 1 - U, V + 1))
            if b is not None:      # lane v0+128 of the row after: local row a - (LANES - 1) of the first sweep's tile V+1
                sh = np.full(rows, -np.inf)
                idx = np.arange(rows) - (LANES - 1) + GATE
                ok = idx >= 0
                sh[ok] = b[np.minimum(idx[ok], rows - 1)]
                gate = np.maximum(gate, sh)
        first, t_last = tile_schedule(start, up, left, gate, T_STEP if (up is None and left is None) else T_STEP_DEP)
        last[(s, U, V)] = t_last
        heapq.heappush(free, t_last[-1])
        busy += t_last[-1] - start
    end = max(v[-1] for v in last.values())
    return end, busy


end1, busy1 = run(1)
print(f"N={N}: {nU} x {nV} tiles of {NU} planes x {LANES} lanes, {rows} steps each, {CTAS} CTAs; "
      f"t_step {T_STEP} us, t_hand {T_HAND} us, t_mail {T_MAIL} us")
print(f"  chain of steps alone : {(N + rows + (nV - 1) * LANES) * T_STEP / 1e3:.3f} ms  (u planes + rows of a tile + lane offsets of the tiles)")
print(f"  tile work / CTAs     : {nU * nV * rows * T_STEP / CTAS / 1e3:.3f} ms  (every CTA busy all the time)")
print(f"  one sweep, modelled  : {end1 / 1e3:.3f} ms   CTA occupancy {busy1 / (CTAS * end1):.2f}")
if OVERLAP:
    end2, busy2 = run(2)
    print(f"  two sweeps overlapped: {end2 / 1e3:.3f} ms = {end2 / 2e3:.3f} ms per sweep ({2 * end1 / end2:.2f}x)   CTA occupancy {busy2 / (CTAS * end2):.2f}")

"""Schedule model of one directional sweep of k_sweep_march (no GPU needed): tiles of PUT planes x TW lanes march `steps` rows at
`tau` us per step; tile (U,V) needs tile (U-1,V) to be PUT-1 steps and tile (U,V-1) to be TW-1 steps further, plus `mail` us per hop;
tickets in the kernel's order on `slots` persistent CTAs.  With the measured constants (tau 0.232, mail 3, 148 slots, 16 x 32 tiles)
it gives 0.66 ms for 512^3; the measured sweep is 0.686 ms.
    python tools/schedule_model.py [N] [PUT] [TW] [slots] [tau] [mail]"""
import heapq, sys, numpy as np
def sim(N, PUT, TW, slots, tau, mail, lag_u=None, tile_ovh=3.0, verbose=False):
    nU = -(-N // PUT); nV = -(-N // TW)
    steps = N + TW + PUT
    if lag_u is None: lag_u = PUT + 2
    tickets = sorted((U * lag_u + V * TW + V, U, V) for U in range(nU) for V in range(nV))
    free = [0.0] * min(slots, len(tickets)); heapq.heapify(free)
    last = {}; busy = 0; k = np.arange(steps)
    for key, U, V in tickets:
        start = heapq.heappop(free) + tile_ovh
        dep = np.full(steps, start)
        up = last.get((U - 1, V)); left = last.get((U, V - 1))
        if up is not None:
            sh = np.full(steps, up[-1]); sh[:steps - (PUT - 1)] = up[PUT - 1:]; dep = np.maximum(dep, sh + mail)
        if left is not None:
            sh = np.full(steps, left[-1]); sh[:steps - (TW - 1)] = left[TW - 1:]; dep = np.maximum(dep, sh + mail)
        # t[s] = finish of step s: max(t[s-1], dep[s]) + tau
        t = (k + 1) * tau + np.maximum.accumulate(dep - k * tau)
        last[(U, V)] = t
        heapq.heappush(free, t[-1]); busy += t[-1] - start
    end = max(v[-1] for v in last.values())
    work = nU * nV * steps * tau / slots
    return end, work, busy / (slots * end), nU * nV, steps
if __name__ == "__main__":
    N = 512
if __name__ == "__main__":
    a = sys.argv[1:]
    N = int(a[0]) if len(a) > 0 else 512
    PUT = int(a[1]) if len(a) > 1 else 16
    TW = int(a[2]) if len(a) > 2 else 32
    slots = int(a[3]) if len(a) > 3 else 148
    tau = float(a[4]) if len(a) > 4 else 0.232
    mail = float(a[5]) if len(a) > 5 else 3.0
    end, work, occ, nt, steps = sim(N, PUT, TW, slots, tau, mail)
    print(f"N={N}: {nt} tiles of {PUT} x {TW}, {steps} steps each, {slots} CTA slots, step {tau} us, hop {mail} us")
    print(f"  chain alone      : {(N + steps + (-(-N // TW) - 1) * TW) * tau / 1e3 + ((-(-N // PUT)) + (-(-N // TW)) - 2) * mail / 1e3:.3f} ms")
    print(f"  tile work / slots: {work / 1e3:.3f} ms")
    print(f"  modelled sweep   : {end / 1e3:.3f} ms (CTA occupancy {occ:.2f})")

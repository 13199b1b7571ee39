"""Randomised model check of the neighbour-synchronisation protocol of fused_kernel<FLUX> (gcmf_fused.cuh) and of its
opt-in variants, without a GPU.

The kernel orders its shared-memory traffic with per-warp mbarriers instead of CTA barriers (racecheck cannot see
that), so the protocol is restated here as a discrete-event model and executed under random schedules:

  * 16 warps in a 2 x 8 grid; per level: wait for the TMA landing, [wait for the neighbours' last phase], extract
    (reads the landing tiles, writes its rows of S0), publish, count itself as drained; k steps, each: wait for the
    neighbours' previous phase, read its own and its neighbours' rows of S[(s-1)&1], write its rows of S[s&1],
    publish; store.  The landing tiles are re-armed for the next level by the warp that drains them last (default)
    or by warp 0, which polls the drain counter between its steps (EDGEREFILL).
  * mbarriers: two per warp, expected arrivals = number of neighbours, waits by phase parity exactly as the kernel
    computes them (barrier need & 1, parity ((need - 1) >> 1) & 1); a wait passes when the barrier's current phase
    parity differs from the one waited for -- so a protocol that falls two phases behind IS caught.
  * every tile row group carries the (level, phase) of its last write; a read checks, at the start and at the end of
    the reading phase, that it sees exactly the version the algorithm needs (catches missing waits and overwrites
    while a neighbour still reads); the landing tiles carry the level they hold.

    python tests/tools/sync_model.py [--runs 300] [--levels 4]

Exit code 1 on a violation or a deadlock.  `--latewait` additionally models a variant that was considered and
REJECTED by this checker: no neighbour wait in front of extract when k is even (extract only writes the warp's own
rows of S0, which nobody reads after step k-1).  The data argument holds, but a neighbour may then arrive for phase
g+2 on a barrier whose phase g is still open; the counting barrier completes with the wrong mix of arrivals and a
warp starts a step before a slow neighbour has published its rows.
"""
import argparse
import random
import sys

WX, WY = 2, 8
NW = WX * WY


def neighbours(w):
    wy, wx = divmod(w, WX)
    out = []
    if wx > 0:
        out.append(w - 1)
    if wx < WX - 1:
        out.append(w + 1)
    if wy > 0:
        out.append(w - WX)
    if wy < WY - 1:
        out.append(w + WX)
    return out


class Violation(Exception):
    pass


class Model:
    def __init__(self, k, levels, edgerefill, latewait, skiplast, rng):
        self.k, self.levels = k, levels
        self.edgerefill, self.latewait, self.skiplast = edgerefill, latewait, skiplast
        self.rng = rng
        self.nb = [neighbours(w) for w in range(NW)]
        # mbarriers: [warp][2] -> (phase counter, pending arrivals)
        self.bar_phase = [[0, 0] for _ in range(NW)]
        self.bar_pending = [[len(self.nb[w]), len(self.nb[w])] for w in range(NW)]
        # S tiles: version[tile][warp] = (level, phase) of the last write of that warp's rows; phase 0 = extract
        self.ver = [[None] * NW for _ in range(2)]
        self.landed = 0          # level held by the landing tiles (level 0 is loaded in the prologue)
        self.inflight = None     # level being copied in by the TMA engine
        self.xcount = 0
        # per-run speeds: a few warps (and sometimes the copy engine) are much slower than the rest, so that one-sided
        # races get a real chance instead of the near-lockstep of a uniform random schedule
        self.weight = [rng.choice([1.0, 1.0, 1.0, 0.2, 0.03]) for _ in range(NW)] + [rng.choice([1.0, 0.1, 0.01])]
        self.prog = [self.program(w) for w in range(NW)]
        self.blocked = [None] * NW
        self.done = [False] * NW

    # ---- primitives -------------------------------------------------------------------------------------
    def arrive(self, w, done_phase):
        for n in self.nb[w]:
            b = done_phase & 1
            self.bar_pending[n][b] -= 1
            if self.bar_pending[n][b] == 0:
                self.bar_phase[n][b] += 1
                self.bar_pending[n][b] = len(self.nb[n])
            elif self.bar_pending[n][b] < 0:
                raise Violation(f"warp {n}: more arrivals than neighbours on barrier {b}")

    def wait_ok(self, w, need):
        if need == 0:
            return True
        b, parity = need & 1, ((need - 1) >> 1) & 1
        return (self.bar_phase[w][b] & 1) != parity

    def check_read(self, w, tile, level, phase, who):
        for n in [w] + self.nb[w]:
            if self.ver[tile][n] != (level, phase):
                raise Violation(f"warp {w} ({who}, level {level}) reads S{tile} rows of warp {n}: holds "
                                f"{self.ver[tile][n]}, needs {(level, phase)}")

    def refill(self, level):
        if self.inflight is not None:
            raise Violation("two refills in flight")
        if self.xcount < NW * level:
            raise Violation(f"refill for level {level} before all warps drained level {level - 1}")
        self.inflight = level

    # ---- one warp's program (a generator: every yield is a scheduling point; yielding a callable = blocked) ----
    def program(self, w):
        k = self.k
        for it in range(self.levels):
            g0 = it * (k + 1)
            yield lambda it=it: self.landed == it                      # mbar_wait(&mb[1], it & 1)
            skip = self.latewait and k % 2 == 0
            if not skip:
                yield lambda g0=g0: self.wait_ok(w, g0)
            # extract: reads the landing tiles, writes its rows of S0
            if self.landed != it:
                raise Violation(f"warp {w} extracts level {it}, landing tiles hold {self.landed}")
            yield None
            if self.landed != it:
                raise Violation(f"landing tiles overwritten while warp {w} extracts level {it}")
            self.ver[0][w] = (it, 0)
            self.arrive(w, g0 + 1)
            self.xcount += 1
            refill_due = False
            if it + 1 < self.levels:
                if self.edgerefill:
                    refill_due = w == 0
                elif self.xcount == NW * (it + 1):
                    self.refill(it + 1)
            for s in range(1, k + 1):
                if refill_due and self.xcount >= NW * (it + 1):          # try_refill(false)
                    self.refill(it + 1)
                    refill_due = False
                yield lambda need=g0 + s: self.wait_ok(w, need)
                src, dst = (s - 1) & 1, s & 1
                self.check_read(w, src, it, s - 1, f"step {s} begin")
                yield None
                self.check_read(w, src, it, s - 1, f"step {s} end")
                if not (self.skiplast and s == k):
                    self.ver[dst][w] = (it, s)
                self.arrive(w, g0 + s + 1)
                yield None
            if refill_due:                                                 # try_refill(true)
                yield lambda it=it: self.xcount >= NW * (it + 1)
                self.refill(it + 1)
            yield None                                                     # store
        self.done[w] = True

    # ---- scheduler --------------------------------------------------------------------------------------
    def run(self):
        steps = 0
        while not all(self.done):
            steps += 1
            if steps > 2000000:
                raise Violation("no progress (livelock)")
            choices = list(range(NW)) + ([NW] if self.inflight is not None else [])
            # weighted random order (Efraimidis-Spirakis keys)
            choices.sort(key=lambda c: -(self.rng.random() ** (1.0 / self.weight[c])))
            progressed = False
            for c in choices:
                if c == NW:  # the TMA engine lands the copy
                    self.landed, self.inflight = self.inflight, None
                    progressed = True
                    break
                if self.done[c]:
                    continue
                cond = self.blocked[c]
                if cond is not None and not cond():
                    continue
                self.blocked[c] = None
                try:
                    nxt = next(self.prog[c])
                except StopIteration:
                    self.done[c] = True
                    nxt = None
                self.blocked[c] = nxt
                progressed = True
                break
            if not progressed:
                raise Violation("deadlock: " + ", ".join(f"w{w}" for w in range(NW) if not self.done[w]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--runs", type=int, default=300)
    ap.add_argument("--levels", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--latewait", action="store_true", help="also model the rejected LATEWAIT variant (violations expected)")
    args = ap.parse_args()
    rng = random.Random(args.seed)
    bad = 0
    for edge in (0, 1):
        for late in ((0, 1) if args.latewait else (0,)):
            for skip in (0, 1):
                for k in (1, 2, 3, 4):
                    fails = 0
                    for _ in range(args.runs):
                        try:
                            Model(k, args.levels, edge, late, skip, rng).run()
                        except Violation as exc:
                            fails += 1
                            if fails == 1:
                                print(f"EDGEREFILL={edge} LATEWAIT={late} SKIPLAST={skip} k={k}: {exc}")
                    bad += fails
    print("violations:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

"""Crude timing model of one CTA of fused_kernel<FLUX> (16 warps, neighbour barriers): what does the re-arming of the
landing tiles cost when it lands on the critical path of whichever inner warp drains them last, and how much of it
comes back when an edge warp does it (GCMF_OPT_EDGEREFILL)?  Units: one `extract` = 1; an inner warp's step = 4, an
edge warp's 3, 2, 1, 0 rows of 4; store = 1; re-arm = 3 (~300 of the ~1730 instructions of a warp-level); durations
jittered.  The model reproduces the measured cost of the neighbour waits (10.1 ms vs 8.3 ms without them: +22 %;
model: 21.4-23.3 vs 18 units per level) and predicts 10-15 % from EDGEREFILL.  An estimate to be checked on the GPU.

    python tests/tools/sync_timing_model.py
"""
import random, statistics
WX,WY=2,8; NW=16
def nbrs(w):
    wy,wx=divmod(w,WX); o=[]
    if wx>0:o.append(w-1)
    if wx<WX-1:o.append(w+1)
    if wy>0:o.append(w-WX)
    if wy<WY-1:o.append(w+WX)
    return o
NB=[nbrs(w) for w in range(NW)]
def run(edge, levels=40, k=4, jitter=0.1, d_refill=3.0, d_tma=6.0, seed=0, d_inner=4.0, d_edge=1.5):
    rng=random.Random(seed)
    def J(x): return x*(1+jitter*(2*rng.random()-1))
    edgew=[w for w in range(NW) if divmod(w,WX)[0] in (0,WY-1)]
    free=[0.0]*NW                      # time warp becomes free
    done_prev=[0.0]*NW                 # completion time of last phase of previous level
    landed=0.0
    t_level=[]
    for it in range(levels):
        # extract
        st=[max(free[w], landed, max(done_prev[n] for n in NB[w])) for w in range(NW)]
        ex=[st[w]+J(1.0) for w in range(NW)]
        free=ex[:]
        alldrained=max(ex)
        refill_end=None
        if not edge:
            last=max(range(NW), key=lambda w: ex[w])
            free[last]+=J(d_refill); refill_end=free[last]
        prev=ex[:]   # completion time of previous phase per warp
        for s in range(1,k+1):
            if edge and refill_end is None and free[0]>=alldrained:
                free[0]+=J(d_refill); refill_end=free[0]
            start=[max(free[w], max(prev[n] for n in NB[w])) for w in range(NW)]
            # work: edge warps shrink: rows in region: s=1:3,2:2,3:1,4:0 of 4
            end=[]
            for w in range(NW):
                if w in edgew: d=d_inner*max(0,4-s)/4.0+0.2
                else: d=d_inner
                end.append(start[w]+J(d))
            free=end[:]; prev=end[:]
        if edge and refill_end is None:
            free[0]=max(free[0],alldrained)+J(d_refill); refill_end=free[0]
        landed=refill_end+J(d_tma)
        # store
        free=[free[w]+J(1.0) for w in range(NW)]
        done_prev=prev[:]
        t_level.append(max(free))
    per=[(t_level[i]-t_level[i-1]) for i in range(5,levels)]
    return statistics.mean(per)
for jit in (0.05,0.15,0.3):
    a=statistics.mean(run(0,jitter=jit,seed=s) for s in range(20)); b=statistics.mean(run(1,jitter=jit,seed=s) for s in range(20))
    nosync=1+4*4+1
    print(f"jitter {jit}: default {a:.2f}  edgerefill {b:.2f}  ({(a/b-1)*100:.1f}% faster); ideal (no waits, no refill) {nosync}")

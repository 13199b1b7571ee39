// gcmf_march.cuh -- temporally blocked Chebyshev steps of the FLUX family as a ROW-STREAMING pipeline.
//
// The tile form (gcmf_fused.cuh) keeps a 32 x 128 tile of one level in shared memory and sweeps it k times; its
// state lives in registers (48 per thread), it re-reads three coefficient tiles from shared memory on every step,
// 4 of its 16 warps own halo rows only, and ncu shows 12 busy warps of long dependent fp64 chains that cannot hide
// their latencies (profiles/variants_r02.md).  This form turns the tile on its side:
//
//   * a CTA owns a strip of 120 output columns (+ 4 halo columns per side = 128 threads) of MARCH_LV levels at once
//     and marches north through a band of rows, one row per iteration;
//   * the K steps run as a software pipeline along the march, two rows apart: in iteration t a thread (one column of
//     one level) performs step 1 at row t, step 2 at row t-2, ... step K at row t-2(K-1).  Every input of every step
//     was produced in an earlier iteration, so the K steps of an iteration are INDEPENDENT (a first version with the
//     steps one row apart fed each step's result to the next one as its north neighbour: four dependent fp64 chains per
//     iteration, 1.3x slower than the tile form).  The N / S neighbours of a point are registers of the same thread
//     (three rows per stage), there is NO redundant row work inside a band (the tile form recomputes 4 halo rows per 24);
//   * only the W / E neighbours go through shared memory: at the top of an iteration every thread publishes the K
//     values that become stage centres in the next iteration (sanitized) in a double-buffered exchange row -- one named
//     barrier per iteration and level (128 threads);
//   * rows of T_{i-1}, T_{i-2}, bar of the LV levels (8-slot ring) and of the three coefficient planes, shared by the
//     levels (16-slot ring: a coefficient row stays in use for 2K iterations), are streamed by the TMA engine (one lane
//     per array, cp.async.bulk -> UBLKCP, mbarrier complete_tx) several rows ahead by a dedicated producer warp.
//     (MARCH_LV = 3: with four levels the producer would be a 17th warp, five warps on one of the four register-file
//     sub-partitions, which caps every thread at 96 registers and spills; folding the producer duty into consumer
//     warp 0 instead made the whole CTA run at that warp's pace -- measured 1.3x to 2x slower than the tile form.)
//
// HBM traffic per grid-point step: (6w + 3w/LV) / K * (128/120) = 14.4 B (fp64, K = 4) plus 3(K-1) priming iterations
// per band, against 13.1 B of the tile form -- but all 16 warps do the same work, the per-step shared-memory traffic is
// 7 LDS.64 + 1 STS.64 per point instead of a full tile sweep.
//
// The arithmetic of a point is the same inline code as everywhere else (flux_lap, shifted_flux, cheb_next,
// bar_update): results are bit-identical to the tile form and to the one-step kernels.
// Device-only: the host emulator keeps the tile form (tests/cabi/gpu_vs_emu.c compares the GPU with it bit for bit).
#pragma once
#include "gcmf_fused.cuh"

namespace gcmf {

#ifndef GCMF_MARCH_LV
#define GCMF_MARCH_LV 3  // 12 consumer warps + the producer warp = 13 warps: at most 4 on a sub-partition, 128 registers
#endif
constexpr int MARCH_LV = GCMF_MARCH_LV;  // levels per CTA (they share the coefficient rows)
constexpr int MARCH_W = 128;             // threads per level = staged columns per row
constexpr int MARCH_DS = 8;              // state ring slots (rows of T1, T2, bar of every level)
constexpr int MARCH_DC = 16;             // coefficient ring slots (rows of ce, cn, ra)

template <typename T> struct MarchGeom {
    static constexpr int H = FUSED_H;
    static constexpr int SW = MARCH_W - 2 * H;                   // output columns per strip: 120
    static constexpr int NS = 3 * MARCH_LV;                      // state arrays of a row: T1, T2, bar per level
    static constexpr int SSLOT = NS * MARCH_W;                   // elements per state slot
    static constexpr int CSLOT = 3 * MARCH_W;                    // elements per coefficient slot
    static constexpr int XB = 2 * MARCH_LV * FUSED_H * MARCH_W;  // exchange rows: [parity][level][stage][column]
    static constexpr int PAD = 16;                               // elements in front of the exchange rows (column -1 reads)
    static constexpr int NTHREADS = MARCH_W * MARCH_LV + 32;
    static constexpr int NBAR = 2 * (MARCH_DS + MARCH_DC);
    static constexpr size_t smem_bytes() {
        return ((size_t)PAD + XB + (size_t)MARCH_DC * CSLOT + (size_t)MARCH_DS * SSLOT + PAD) * sizeof(T) + NBAR * sizeof(uint64_t) + 128;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// EDGE as in fused_kernel: bit 0 = the block starts at recurrence step 1, bit 1 = it ends at step n_steps.
// K = steps of the block (compile time: the step loop is straight-line code).
template <typename T, int EDGE, int K>
__global__ void __launch_bounds__(MarchGeom<T>::NTHREADS, 1)
    march_kernel(const __grid_constant__ FusedParams<T> P, int nstrips, int nlg, int ry) {
    using G = MarchGeom<T>;
    constexpr bool FIRST = (EDGE & 1) != 0, LAST = (EDGE & 2) != 0;
    constexpr int H = G::H;
    // shared memory: [pad][exchange rows][coefficient ring][state ring][pad][barriers].  West / east neighbours are read
    // at column +-1 without clamping: column -1 of the first array / +1 of the last one fall into a pad or the
    // neighbouring array -- defined memory whose value only reaches cells outside the dependency cone.
    extern __shared__ __align__(128) unsigned char march_smem[];
    T* xb = reinterpret_cast<T*>(march_smem) + G::PAD;
    T* cring = xb + G::XB;
    T* sring = cring + (size_t)MARCH_DC * G::CSLOT;
    uint64_t* fullS = reinterpret_cast<uint64_t*>(sring + (size_t)MARCH_DS * G::SSLOT + G::PAD);
    uint64_t* emptyS = fullS + MARCH_DS;
    uint64_t* fullC = emptyS + MARCH_DS;
    uint64_t* emptyC = fullC + MARCH_DC;

    // block id -> (level group fastest: the groups of one strip / band re-read the same coefficient rows, strip, band)
    unsigned bid = blockIdx.x;
    const int lg = (int)(bid % (unsigned)nlg);
    bid /= (unsigned)nlg;
    const int cx = (int)(bid % (unsigned)nstrips);
    const int band = (int)(bid / (unsigned)nstrips);
    const int ny = P.g.ny, nx = P.g.nx;
    const int j0 = band * ry, j1 = j0 + ry < ny ? j0 + ry : ny;
    const int lev0 = lg * MARCH_LV;
    const int nlev = (int)(P.nb - lev0 < MARCH_LV ? P.nb - lev0 : MARCH_LV);  // active levels of this CTA
    const int tid = threadIdx.x;
    const bool wrap = (P.g.flags & FL_WRAP_Y) != 0;
    // row r of the arrays: periodic, or a latitude band with FUSED_H ghost rows physically present on either side
    auto rowidx = [&](int r) { return wrap ? (r < 0 ? r + ny : (r >= ny ? r - ny : r)) : r; };
    const int R0 = j0 - K;                 // first staged row
    const int nrows = (j1 - j0) + 2 * K;   // staged rows R0 .. j1 + K - 1 (both rings)
    if (tid == 0) {
        const unsigned nw = (unsigned)(nlev * (MARCH_W / 32));
        for (int s = 0; s < MARCH_DS; ++s) { mbar_init(&fullS[s], 1); mbar_init(&emptyS[s], nw); }
        for (int s = 0; s < MARCH_DC; ++s) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], nw); }
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= MARCH_W * MARCH_LV) {
        // ---- producer warp: lanes 0..NS-1 own the state arrays (level-major T1, T2, bar), NS..NS+2 own ce, cn, ra
        const int lane = tid & 31;
        const T* pbase = nullptr;
        int64_t ppitch = 0;
        bool pbar = false;
        if (lane < G::NS) {
            const int pl_ = lane / 3, kind = lane % 3;
            if (pl_ < nlev) {
                const int64_t plev = lev0 + pl_;
                if (kind == 0) { pbase = P.t1_in.p + plev * P.t1_in.bstride; ppitch = P.t1_in.pitch; }
                if (kind == 1 && !FIRST) { pbase = P.t2_in.p + plev * P.t2_in.bstride; ppitch = P.t2_in.pitch; }
                if (kind == 2 && !FIRST) { pbase = P.bar.p + plev * P.bar.bstride; ppitch = P.bar.pitch; pbar = true; }
            }
        } else if (lane < G::NS + 3) {
            const int pc = lane - G::NS;
            pbase = reinterpret_cast<const T*>(pc == 0 ? P.plane[0].p : (pc == 1 ? P.plane[1].p : P.plane[2].p));
            ppitch = pc == 0 ? P.plane[0].pitch : (pc == 1 ? P.plane[1].pitch : P.plane[2].pitch);
        }
        const int pcol0 = cx * G::SW - H;
        const int pgx = pcol0 < 0 ? pcol0 + nx : pcol0;
        const int pn1 = (nx - pgx) < MARCH_W ? (nx - pgx) : MARCH_W;
        const unsigned per = (unsigned)(MARCH_W * sizeof(T));
        const unsigned tx_halo = per * (unsigned)(nlev * (FIRST ? 1 : 2));  // bar only exists for the owned rows
        const unsigned tx_own = per * (unsigned)(nlev * (FIRST ? 1 : 3));
        T* const dstS = sring + lane * MARCH_W;
        T* const dstC = cring + (lane - G::NS) * MARCH_W;
        for (int i = 0; i < nrows; ++i) {  // one row of either ring per iteration, as soon as its slot is free
            const int r = R0 + i;
            const bool own = r >= j0 && r < j1;
            const int ss = i & (MARCH_DS - 1), sc = i & (MARCH_DC - 1);
            if (i >= MARCH_DS) mbar_wait(&emptyS[ss], (unsigned)(((i / MARCH_DS) - 1) & 1));
            if (i >= MARCH_DC) mbar_wait(&emptyC[sc], (unsigned)(((i / MARCH_DC) - 1) & 1));
            if (i >= MARCH_DS) fence_proxy_async();
            if (lane == 0) {
                mbar_expect_tx(&fullS[ss], own ? tx_own : tx_halo);
                mbar_expect_tx(&fullC[sc], 3u * per);
            }
            __syncwarp();
            if (pbase != nullptr && (!pbar || own)) {
                const T* row = pbase + (int64_t)rowidx(r) * ppitch;
                const bool state = lane < G::NS;
                T* dst = state ? dstS + (size_t)ss * G::SSLOT : dstC + (size_t)sc * G::CSLOT;
                uint64_t* fb = state ? &fullS[ss] : &fullC[sc];
                bulk_copy_g2s(dst, row + pgx, (unsigned)(pn1 * sizeof(T)), fb);
                if (pn1 < MARCH_W) bulk_copy_g2s(dst + pn1, row, (unsigned)((MARCH_W - pn1) * sizeof(T)), fb);
            }
        }
        return;
    }

    // ---- consumers: thread = (level l, column c)
    const int l = tid / MARCH_W, c = tid % MARCH_W;
    if (l >= nlev) return;
    const int64_t lev = lev0 + l;
    const int gc = cx * G::SW + c - H;  // global column (unwrapped)
    const bool emit_col = c >= H && c < MARCH_W - H && gc < nx;
    const T cc = (T)P.c;
    const int oT1 = 3 * l * MARCH_W, oT2 = oT1 + MARCH_W, oBar = oT2 + MARCH_W;  // arrays inside a state slot
    constexpr int oCe = 0, oCn = MARCH_W, oRa = 2 * MARCH_W;                      // arrays inside a coefficient slot
    const T* const sring_c = sring + c;
    const T* const cring_c = cring + c;
    // ring row of staged index idx (= row - R0)
    auto srow = [&](int idx) { return sring_c + (size_t)(idx & (MARCH_DS - 1)) * G::SSLOT; };
    auto crow = [&](int idx) { return cring_c + (size_t)(idx & (MARCH_DC - 1)) * G::CSLOT; };
    constexpr int XBP = MARCH_LV * FUSED_H * MARCH_W;
    T* xr = xb + (size_t)l * FUSED_H * MARCH_W + c;  // exchange rows read in this iteration ...
    T* xw = xr + XBP;                                // ... and written for the next one
    auto wait_full = [&](uint64_t* fullb, int D, int idx) { mbar_wait(&fullb[idx & (D - 1)], (unsigned)((idx / D) & 1)); };

    // ---- per-thread pipeline state.  Stage q (0..K-1) is the input of step q+1, which runs at row t-2q in iteration t:
    //   san{S,C,N}[q]  sanitized stage-q values at rows t-2q-1, t-2q, t-2q+1
    //   raw{SS,S,C,N}[q]  raw values at rows t-2q-2 .. t-2q+1 (C: the "-x" term of step q+1; SS: T_{i-2} of step q+2)
    //   accA / accB[q]  running bar of the rows that finished step q+1 one / two iterations ago
    T sanS[K], sanC[K], sanN[K], rawSS[K], rawS[K], rawC[K], rawN[K], accA[K], accB[K];
#pragma unroll
    for (int q = 0; q < K; ++q)
        sanS[q] = sanC[q] = sanN[q] = rawSS[q] = rawS[q] = rawC[q] = rawN[q] = accA[q] = accB[q] = T(0);
    const int t0 = j0 - (K - 1), t1 = j1 - 1 + 2 * (K - 1);
    const int last_idx = nrows - 1;
    {   // prologue: stage 0 holds rows t0-1, t0, t0+1 (staged indices 0, 1, 2); coefficient rows 0 and 1
        wait_full(fullS, MARCH_DS, 0);
        wait_full(fullS, MARCH_DS, 1);
        wait_full(fullS, MARCH_DS, 2);
        wait_full(fullC, MARCH_DC, 0);
        rawS[0] = srow(0)[oT1];
        rawC[0] = srow(1)[oT1];
        rawN[0] = srow(2)[oT1];
        sanS[0] = nan2num(rawS[0]);
        sanC[0] = nan2num(rawC[0]);
        sanN[0] = nan2num(rawN[0]);
        xr[0] = sanC[0];  // what the neighbours read as stage-0 centre in the first iteration
#pragma unroll
        for (int q = 1; q < FUSED_H; ++q) xr[q * MARCH_W] = T(0);
        // state row 0 (row t0-1) only contributes its T1, lifted just now: release its slot (the loop releases row t in
        // iteration t, starting with staged index 1)
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&emptyS[0]);
        named_barrier(1 + l, MARCH_W);
    }
    // running element offsets of the output rows (row rk = t - 2(K-1))
    int64_t ob = lev * P.bar.bstride + (int64_t)(t0 - 2 * (K - 1)) * P.bar.pitch + gc;
    int64_t o1 = 0, o2 = 0;
    if (!LAST) {
        o1 = lev * P.t1_out.bstride + (int64_t)(t0 - 2 * (K - 1)) * P.t1_out.pitch + gc;
        o2 = lev * P.t2_out.bstride + (int64_t)(t0 - 2 * (K - 1)) * P.t2_out.pitch + gc;
    }
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
        const int ti = t - R0;  // staged index of row t
        // publish what becomes the stage centres of the next iteration: the north rows held now
#pragma unroll
        for (int q = 0; q < K; ++q) xw[q * MARCH_W] = sanN[q];
        // new north row of stage 0 for the next iteration: input row t+2
        T rawN0 = T(0);
        if (ti + 2 <= last_idx) {
            wait_full(fullS, MARCH_DS, ti + 2);
            rawN0 = srow(ti + 2)[oT1];
        }
        if (ti <= last_idx) wait_full(fullC, MARCH_DC, ti);
        T t2in = T(0), barin = T(0);
        if (!FIRST && ti <= last_idx) {
            t2in = srow(ti)[oT2];
            if (t >= j0 && t < j1) barin = srow(ti)[oBar];
        }
        T tn[K], an[K];
#pragma unroll
        for (int s = 1; s <= K; ++s) {
            const int q = s - 1;
            const int r = t - 2 * q;
            // step s at row r feeds an owned row only inside the dependency cone of the band (uniform over the CTA)
            const bool active = r >= j0 - (K - s) && r <= j1 - 1 + (K - s);
            tn[q] = T(0);
            an[q] = T(0);
            if (active) {
                const T* cr = crow(ti - 2 * q);       // coefficient row r
                const T* crs = crow(ti - 2 * q - 1);  // row r-1: its north faces are the south faces of row r
                const T lap = flux_lap<T>(sanC[q], xr[q * MARCH_W - 1], xr[q * MARCH_W + 1], sanN[q], sanS[q],
                                          cr[oCe], cr[oCe - 1], cr[oCn], crs[oCn], cr[oRa]);
                const T a = shifted_flux<T>(rawC[q], cc, lap);                           // filter.py:171
                const bool start = FIRST && s == 1;
                const T tm2 = s == 1 ? t2in : rawSS[q >= 1 ? q - 1 : 0];
                tn[q] = start ? a : cheb_next<T>(a, tm2);                                 // filter.py:192-194 / 197-203
                const double b0 = s == 1 ? (start ? P.p0 * (double)rawC[0] : (double)barin) : (double)accB[q >= 1 ? q - 1 : 0];
                an[q] = (T)bar_update(b0, P.p[q], (double)tn[q]);                        // filter.py:195 / 204
            }
        }
        // outputs of row rk = t-2(K-1): T_{i+K-1} = the value step K produced, T_{i+K-2} = the centre of stage K-1
        const int rk = t - 2 * (K - 1);
        if (emit_col && rk >= j0 && rk < j1) {
            if (!LAST) {
                P.t1_out.p[o1] = tn[K - 1];
                P.t2_out.p[o2] = rawC[K - 1];
            }
            P.bar.p[ob] = an[K - 1];
        }
        ob += P.bar.pitch;
        if (!LAST) {
            o1 += P.t1_out.pitch;
            o2 += P.t2_out.pitch;
        }
        // shift every window one row north; the value step q produced enters stage q as its new north row
#pragma unroll
        for (int q = 0; q < K; ++q) {
            const T nraw = q == 0 ? rawN0 : tn[q >= 1 ? q - 1 : 0];
            rawSS[q] = rawS[q]; rawS[q] = rawC[q]; rawC[q] = rawN[q]; rawN[q] = nraw;
            sanS[q] = sanC[q]; sanC[q] = sanN[q]; sanN[q] = nan2num(nraw);
            accB[q] = accA[q]; accA[q] = an[q];
        }
        // releases: state row t is done (T2 / bar read above; its T1 was lifted two iterations ago); coefficient row
        // t-2K+1 was last used as the south faces of step K's row
        __syncwarp();
        if ((tid & 31) == 0) {
            if (ti >= 0 && ti <= last_idx) mbar_arrive(&emptyS[ti & (MARCH_DS - 1)]);
            const int ci = ti - 2 * K + 1;
            if (ci >= 0 && ci <= last_idx) mbar_arrive(&emptyC[ci & (MARCH_DC - 1)]);
        }
        {   // swap the exchange rows
            T* tmp = xr;
            xr = xw;
            xw = tmp;
        }
        named_barrier(1 + l, MARCH_W);
    }
}
#endif

}  // namespace gcmf

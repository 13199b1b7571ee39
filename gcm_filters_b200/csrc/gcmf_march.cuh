// gcmf_march.cuh -- temporally blocked Chebyshev steps of the FLUX family as a ROW-STREAMING pipeline.
//
// The tile form (gcmf_fused.cuh) keeps a 32 x 128 tile of one level in shared memory and sweeps it k times; its
// state lives in registers (48 per thread), it re-reads three coefficient tiles from shared memory on every step,
// 4 of its 16 warps own halo rows only, and ncu shows 12 busy warps of long dependent fp64 chains that cannot hide
// their latencies (profiles/variants_r02.md).  This form turns the tile on its side:
//
//   * a CTA owns a strip of 120 output columns (+ 4 halo columns per side = 128 threads) of MARCH_LV levels at once
//     and marches north through a band of rows, one row per iteration;
//   * the k steps run as a software pipeline along the march: in iteration t a thread (one column of one level)
//     performs step 1 at row t, step 2 at row t-1, ... step k at row t-k+1 -- each step's north neighbour is the
//     value the previous step of the same iteration has just produced, its own and south values are two registers
//     per stage, so N / S neighbours never touch shared memory and there is NO redundant row work inside a band
//     (the tile form recomputes 4 halo rows per 24);
//   * only the W / E neighbours go through shared memory: every thread publishes the k values it produced (sanitized)
//     in a double-buffered exchange row and reads its two neighbours' in the next iteration -- one named barrier per
//     iteration and level (128 threads);
//   * rows of T_{i-1}, T_{i-2}, bar of the LV levels and of the three coefficient planes (shared by the levels) are
//     streamed through a 10-slot shared-memory ring (one lane per array, cp.async.bulk -> UBLKCP, mbarrier
//     complete_tx), up to four rows ahead; consumers release a row k iterations after they first used it.  Warp 0
//     doubles as the producer at the top of its iterations: a 17th warp would cap every thread at 96 registers (the
//     register file is split over four sub-partitions: 5 warps on one of them), which spills.
//
// HBM traffic per grid-point step: (6w + 3w/LV) / k * (128/120) = 14.4 B (fp64, k = 4) plus 2(k-1) priming rows per
// band, against 13.1 B of the tile form -- but all 16 warps do the same work, the per-step shared-memory traffic is
// 6 LDS.64 + 1 STS.64 per point instead of a full tile sweep, and the per-thread state is 2 rows x k stages.
//
// The arithmetic of a point is the same inline code as everywhere else (flux_lap, shifted_flux, cheb_next,
// bar_update): results are bit-identical to the tile form and to the one-step kernels.
// Device-only: the host emulator keeps the tile form (tests/cabi/gpu_vs_emu.c compares the GPU with it bit for bit).
#pragma once
#include "gcmf_fused.cuh"

namespace gcmf {

#ifndef GCMF_MARCH_LV
#define GCMF_MARCH_LV 4
#endif
constexpr int MARCH_LV = GCMF_MARCH_LV;  // levels per CTA (they share the coefficient rows)
constexpr int MARCH_W = 128;             // threads per level = staged columns per row
constexpr int MARCH_D = 10;              // ring slots (rows)

template <typename T> struct MarchGeom {
    static constexpr int H = FUSED_H;
    static constexpr int SW = MARCH_W - 2 * H;                   // output columns per strip: 120
    static constexpr int NARR = 3 * MARCH_LV + 3;                // T1, T2, bar per level + ce, cn, ra
    static constexpr int SLOT = NARR * MARCH_W;                  // elements per ring slot
    static constexpr int XB = 2 * MARCH_LV * FUSED_H * MARCH_W;  // exchange rows: [parity][level][stage][column]
    static constexpr int PAD = 16;                               // elements in front of the exchange rows (column -1 reads)
    static constexpr int NTHREADS = MARCH_W * MARCH_LV;
    static constexpr size_t smem_bytes() {
        return ((size_t)PAD + XB + (size_t)MARCH_D * SLOT) * sizeof(T) + 2 * MARCH_D * sizeof(uint64_t) + 128;
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
    // shared memory: [pad][exchange rows][ring][barriers].  West / east neighbours are read at column +-1 without
    // clamping: column -1 of the first array / +1 of the last one fall into the pad, the neighbouring array or the
    // barrier words -- defined memory whose value only reaches cells outside the dependency cone.
    extern __shared__ __align__(128) unsigned char march_smem[];
    T* xb = reinterpret_cast<T*>(march_smem) + G::PAD;
    T* ring = xb + G::XB;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)MARCH_D * G::SLOT);
    uint64_t* empty = full + MARCH_D;

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
    const int nrows = (j1 - j0) + 2 * K;   // staged rows R0 .. j1 + K - 1
    if (tid == 0) {
        for (int s = 0; s < MARCH_D; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (unsigned)(nlev * (MARCH_W / 32)));
        }
        fence_mbar_init();
    }
    __syncthreads();

    // ---- thread = (level l, column c)
    const int l = tid / MARCH_W, c = tid % MARCH_W;
    if (l >= nlev) return;
    const int64_t lev = lev0 + l;
    const int gc = cx * G::SW + c - H;  // global column (unwrapped)
    const bool emit_col = c >= H && c < MARCH_W - H && gc < nx;
    const T cc = (T)P.c;
    // element offsets of the arrays inside a slot, relative to this thread's column of array 0
    const int oT1 = 3 * l * MARCH_W, oT2 = oT1 + MARCH_W, oBar = oT2 + MARCH_W;
    constexpr int oCe = 3 * MARCH_LV * MARCH_W, oCn = oCe + MARCH_W, oRa = oCn + MARCH_W;
    const T* const ring_c = ring + c;
    const T* const ring_end = ring_c + (size_t)MARCH_D * G::SLOT;
    constexpr int XBP = MARCH_LV * FUSED_H * MARCH_W;
    T* xr = xb + (size_t)l * FUSED_H * MARCH_W + c;  // exchange rows read in this iteration ...
    T* xw = xr + XBP;                                // ... and written for the next one

    // ---- producer duty of warp 0: lane a owns array a of a slot (level-major T1, T2, bar; then ce, cn, ra)
    const bool producer = tid < 32;
    const T* pbase = nullptr;
    int64_t ppitch = 0;
    bool pbar = false;
    if (producer) {
        const int lane = tid;
        if (lane < 3 * MARCH_LV) {
            const int pl_ = lane / 3, kind = lane % 3;
            if (pl_ < nlev) {
                const int64_t plev = lev0 + pl_;
                if (kind == 0) { pbase = P.t1_in.p + plev * P.t1_in.bstride; ppitch = P.t1_in.pitch; }
                if (kind == 1 && !FIRST) { pbase = P.t2_in.p + plev * P.t2_in.bstride; ppitch = P.t2_in.pitch; }
                if (kind == 2 && !FIRST) { pbase = P.bar.p + plev * P.bar.bstride; ppitch = P.bar.pitch; pbar = true; }
            }
        } else if (lane < G::NARR) {
            const int pc = lane - 3 * MARCH_LV;
            pbase = reinterpret_cast<const T*>(pc == 0 ? P.plane[0].p : (pc == 1 ? P.plane[1].p : P.plane[2].p));
            ppitch = pc == 0 ? P.plane[0].pitch : (pc == 1 ? P.plane[1].pitch : P.plane[2].pitch);
        }
    }
    const int pcol0 = cx * G::SW - H;
    const int pgx = pcol0 < 0 ? pcol0 + nx : pcol0;
    const int pn1 = (nx - pgx) < MARCH_W ? (nx - pgx) : MARCH_W;
    const unsigned per = (unsigned)(MARCH_W * sizeof(T));
    const unsigned tx_halo = per * (unsigned)(3 + nlev * (FIRST ? 1 : 2));  // bar only exists for the owned rows
    const unsigned tx_own = per * (unsigned)(3 + nlev * (FIRST ? 1 : 3));
    int next_s = 0;  // next row of the band to stage (producer warp)
    // stage rows up to `upto`; rows up to `must` are waited for (they are needed next), later ones only if their slot is free
    auto produce = [&](int upto, int must) {
        while (next_s < nrows && R0 + next_s <= upto) {
            const int slot = next_s % MARCH_D;
            if (next_s >= MARCH_D) {
                const unsigned par = (unsigned)(((next_s / MARCH_D) - 1) & 1);
                if (R0 + next_s <= must) {
                    mbar_wait(&empty[slot], par);
                } else {
                    unsigned ok = 0;
                    if (tid == 0) ok = mbar_test(&empty[slot], par);
                    ok = __shfl_sync(0xffffffffu, ok, 0);
                    if (!ok) break;
                }
                fence_proxy_async();
            }
            const int r = R0 + next_s;
            const bool own = r >= j0 && r < j1;
            if (tid == 0) mbar_expect_tx(&full[slot], own ? tx_own : tx_halo);
            __syncwarp();
            if (pbase != nullptr && (!pbar || own)) {
                const T* row = pbase + (int64_t)rowidx(r) * ppitch;
                T* dst = ring + (size_t)slot * G::SLOT + tid * MARCH_W;
                bulk_copy_g2s(dst, row + pgx, (unsigned)(pn1 * sizeof(T)), &full[slot]);
                if (pn1 < MARCH_W) bulk_copy_g2s(dst + pn1, row, (unsigned)((MARCH_W - pn1) * sizeof(T)), &full[slot]);
            }
            ++next_s;
        }
    };

    // per-stage state: stage s = input of step s+1; S / C = rows (t-s-1, t-s) at the start of iteration t
    T sanS[K], sanC[K], rawS[K], rawC[K], acc[K];
#pragma unroll
    for (int s = 0; s < K; ++s) sanS[s] = sanC[s] = rawS[s] = rawC[s] = acc[s] = T(0);
    const int t0 = j0 - (K - 1), t1 = j1 - 1 + (K - 1);
    if (producer) produce(R0 + MARCH_D - 1, R0 + MARCH_D - 1);  // fill the ring
    // ring rows in use in iteration t: row[i] = this thread's column of array 0 of row t+1-i, i = 0 .. K+1
    const T* row[K + 2];
    // staged index of row t0-1 is 0 (t0 - 1 = R0), of row t0 is 1: the loop starts with row t0+1 = index 2
    int fslot = 2 % MARCH_D;   // slot of the row waited for in the next iteration
    unsigned fpar = 0;
    int eslot = 0;             // slot released next (row t - K)
    {   // prologue: rows t0-1 and t0 of the input
        mbar_wait(&full[0], 0);
        mbar_wait(&full[1], 0);
        // state "at the end of iteration t0-1": row[0] = row t0 (slot 1), row[1] = row t0-1 (slot 0); the rows below R0
        // do not exist and are never dereferenced (their steps are inactive)
#pragma unroll
        for (int i = 0; i < K + 2; ++i) row[i] = ring_c;
        row[0] = ring_c + G::SLOT;
        rawS[0] = ring_c[oT1];
        rawC[0] = ring_c[G::SLOT + oT1];
        sanS[0] = nan2num(rawS[0]);
        sanC[0] = nan2num(rawC[0]);
        xr[0] = sanC[0];
#pragma unroll
        for (int s = 1; s < FUSED_H; ++s) xr[s * MARCH_W] = T(0);
        named_barrier(1 + l, MARCH_W);
    }
    // running element offsets of the output rows (row rk = t - (K-1))
    int64_t ob = lev * P.bar.bstride + (int64_t)(t0 - (K - 1)) * P.bar.pitch + gc;
    int64_t o1 = 0, o2 = 0;
    if (!LAST) {
        o1 = lev * P.t1_out.bstride + (int64_t)(t0 - (K - 1)) * P.t1_out.pitch + gc;
        o2 = lev * P.t2_out.bstride + (int64_t)(t0 - (K - 1)) * P.t2_out.pitch + gc;
    }
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
        if (producer) produce(t + 1 + MARCH_D, t + 2);  // as far ahead as released slots allow; row t+2 at the latest
        // rotate the row pointers: row[0] becomes row t+1
#pragma unroll
        for (int i = K + 1; i > 0; --i) row[i] = row[i - 1];
        {
            const T* nx_ = row[1] + G::SLOT;
            row[0] = nx_ == ring_end ? ring_c : nx_;
        }
        mbar_wait(&full[fslot], fpar);
        if (++fslot == MARCH_D) { fslot = 0; fpar ^= 1u; }
        const T rawN0 = row[0][oT1];
        const T sanN0 = nan2num(rawN0);
        xw[0] = sanN0;
        T t2in = T(0), barin = T(0);
        if (!FIRST) {
            t2in = row[1][oT2];
            if (t >= j0 && t < j1) barin = row[1][oBar];
        }
        T on = sanN0, rawN = rawN0;  // north input of the current step = the value entering stage s-1 in this iteration
        T newsan[K], newraw[K], newacc[K];
#pragma unroll
        for (int s = 1; s <= K; ++s) {
            const int r = t - (s - 1);
            newraw[s - 1] = rawN;   // becomes the centre of stage s-1 in the next iteration
            newsan[s - 1] = on;
            // step s at row r feeds an owned row only inside the dependency cone of the band (uniform over the CTA);
            // outside it nothing is computed and nothing staged is touched (rows below R0 + 1 do not exist in the ring)
            const bool active = r >= j0 - (K - s) && r <= j1 - 1 + (K - s);
            T tn = T(0);
            newacc[s - 1] = T(0);
            if (active) {
                const T* cr = row[s];        // ring row r
                const T* crs = row[s + 1];   // ring row r-1: its north faces are the south faces of row r
                const T lap = flux_lap<T>(sanC[s - 1], xr[(s - 1) * MARCH_W - 1], xr[(s - 1) * MARCH_W + 1], on, sanS[s - 1],
                                          cr[oCe], cr[oCe - 1], cr[oCn], crs[oCn], cr[oRa]);
                const T a = shifted_flux<T>(rawC[s - 1], cc, lap);                       // filter.py:171
                const bool start = FIRST && s == 1;
                const T tm2 = s == 1 ? t2in : rawS[s >= 2 ? s - 2 : 0];
                tn = start ? a : cheb_next<T>(a, tm2);                                    // filter.py:192-194 / 197-203
                // acc of row r after step s-1 = what step s-1 left in the previous iteration
                const double b0 = s == 1 ? (start ? P.p0 * (double)rawC[0] : (double)barin) : (double)acc[s >= 2 ? s - 2 : 0];
                newacc[s - 1] = (T)bar_update(b0, P.p[s - 1], (double)tn);               // filter.py:195 / 204
            }
            // the value just produced is the north input of the next step (one row further south)
            rawN = tn;
            if (s < K) {
                on = nan2num(tn);
                if (active) xw[s * MARCH_W] = on;
            }
        }
        // outputs of row rk = t-(K-1): T_{i+K-1} = the last value produced, T_{i+K-2} = centre of stage K-1
        const int rk = t - (K - 1);
        if (emit_col && rk >= j0 && rk < j1) {
            if (!LAST) {
                P.t1_out.p[o1] = rawN;
                P.t2_out.p[o2] = rawC[K - 1];
            }
            P.bar.p[ob] = newacc[K - 1];
        }
        ob += P.bar.pitch;
        if (!LAST) {
            o1 += P.t1_out.pitch;
            o2 += P.t2_out.pitch;
        }
        // shift the windows one row north
#pragma unroll
        for (int s = 0; s < K; ++s) {
            sanS[s] = sanC[s]; sanC[s] = newsan[s];
            rawS[s] = rawC[s]; rawC[s] = newraw[s];
            acc[s] = newacc[s];
        }
        // row t-K is not needed any more (its north faces were the south faces of row t-K+1 in step K just now)
        if (t - K >= R0) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[eslot]);
            if (++eslot == MARCH_D) eslot = 0;
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

// gcmf_vec2.cuh -- TWO Chebyshev steps of a VECTOR operator per HBM round trip (temporal blocking of the vector
// recurrence, filter.py:225-283, for the C-grid viscous Laplacian kernels.py:647-696 and the B-grid Laplacian
// kernels.py:740-837).
//
// The one-step kernels already stream every array once per step and sit at 0.8-0.9 of the HBM roofline (C-grid: 20 rows
// in -- u, v, 14 coefficient planes, T_{i-2}, bar -- and 4 rows out per row of points = 192 B per grid-point step at
// nb = 1).  The only way below that is to do more steps per byte.  This kernel keeps the row-streaming shape of
// cgrid_tma_kernel -- a CTA owns a strip of columns and marches north through a band of rows -- and runs step i+1
// ONE ROW behind step i inside the same march:
//
//   iteration s (row j = R0 + s):   step i   consumes input row j+1 from the ring and produces T_i(j)          (registers)
//                                   step i+1 consumes T_i(j) as its "row j+1" and produces T_{i+1}(j-1)         (registers)
//
// (A two-row lag -- step i+1 consuming T_i one iteration later, so that the two chains of an iteration are independent --
// was measured and is gone: cfg5 0.409 vs 0.400 ms per launch; it costs a coefficient-ring slot and the consumers are not
// what the launch waits for.)
//
// Both steps are the same marching recurrence (an operator policy: the state of "row j" lives in registers, W / E
// neighbours come from the adjacent lanes by warp shuffle), fed from shared memory (step i) or from the registers step i
// wrote (step i+1).  The coefficient rows are staged ONCE for both steps (a row lives for three iterations in its ring
// instead of two), T_i never touches HBM on its way into step i+1:
//
//   C-grid, per 2 steps and row: 20 rows read + 6 rows written (T_{i+1}, T_i, bar) = 104 B per grid-point step (192: x 1.85)
//   B-grid, per 2 steps and row: 14 rows read + 6 rows written                     =  80 B per grid-point step (144: x 1.80)
//
// Halo: a warp's 32 lanes see 30 valid columns after step i and 28 after step i+1 (lanes 2..29 emit); a band primes with
// two extra rows on either side.  Two rings (TMA bulk copies, cp.async.bulk -> UBLKCP, one lane per array, mbarrier
// complete_tx / consumer-release "empty" barriers as in cgrid_tma_kernel): the 6 field rows are released after one
// iteration, the coefficient rows after two.
//
// Whole periodic grids, or a latitude band of a decomposed domain (no GCMF_FLAG_WRAP_Y): rows -2, -1 and ny, ny+1 of
// the fields and of every coefficient plane are then ghost rows present in memory, which the caller refreshes once per
// two-step block (scheduler.FusedBandedFilter).
//
// Same expressions in the same order as the one-step kernels: results are bit-identical to two gcmf_cheb_step calls.
// Device-only (mbarriers + TMA); the host emulator takes the one-step path, which the GPU tests compare against.
#pragma once
#include "gcmf_fused.cuh"

namespace gcmf {

#ifndef GCMF_CG2_WARPS
#define GCMF_CG2_WARPS 8
#endif
// Ring depths, measured on cfg5 / the same geometry for the B-grid (ms per two-step launch): 5 field + 6 coefficient
// slots 0.3356 / 0.2513 (B-grid: 8 coefficient slots), 4 + 7: 0.3386 / 0.2706; consumer loop unrolled x1 0.346, x2 0.3386,
// x4 0.3348 (166 registers).
#ifndef GCMF_CG2_FS
#define GCMF_CG2_FS 5   // field ring slots (a row is live for 2 iterations)
#endif
#ifndef GCMF_CG2_CS
#define GCMF_CG2_CS 6   // coefficient ring slots, 14 arrays (a row is live for 3 iterations)
#endif
#ifndef GCMF_CG2_UNROLL
#define GCMF_CG2_UNROLL 2  // consumer loop unrolled by two: the row-state rotation becomes register renaming
#endif
#ifndef GCMF_CG2_UNROLL
#define GCMF_CG2_UNROLL 2  // consumer loop unrolled by two: the row-state rotation becomes register renaming
#endif
constexpr int CG2_UNROLL = GCMF_CG2_UNROLL;
constexpr int CG2_WARPS = GCMF_CG2_WARPS;  // consumer warps per CTA (+ 2 producer warps)
constexpr int CG2_COLS = 28;               // at most 28 output columns per warp (lanes 2..29)

template <typename T, int NC_> struct Vec2Geom {
    static constexpr int AV = 16 / (int)sizeof(T);
    static constexpr int HALO = AV < 2 ? 2 : AV;             // staged halo columns per side (two are needed), 16-byte aligned
    static constexpr int LW = CG2_WARPS * CG2_COLS + 2 * HALO;  // row pitch of the rings: 228 (f64) / 232 (f32)
    static constexpr int NC = NC_, NF = 6;                   // coefficient arrays; field arrays: u, v, t2u, t2v, bar_u, bar_v
    static constexpr int FS = GCMF_CG2_FS;
    static constexpr int CS = NC_ > 8 ? GCMF_CG2_CS : GCMF_CG2_CS + 2;  // fewer planes: a deeper ring fits
    static constexpr int NTHREADS = 32 * (CG2_WARPS + 2);  // consumers + the two producer warps
    static constexpr size_t SMEM = ((size_t)FS * NF + (size_t)CS * NC) * LW * sizeof(T) + 2 * (FS + CS) * sizeof(uint64_t) + 128;
    static_assert(SMEM <= 232448, "rings exceed the 227 KB of shared memory a CTA may use");
    static_assert(NC_ + 6 <= 32, "one producer lane per array");
    static constexpr size_t smem_bytes() { return SMEM; }
};
template <typename T> using Cg2Geom = Vec2Geom<T, 14>;

template <typename T> struct Cg2Params {
    Geo g;
    PlaneRef plane[14];
    FieldRef<const T> t1[2];   // T_{i-1} (u, v): the field step i acts on
    FieldRef<const T> t2[2];   // T_{i-2}                         (not FIRST)
    FieldRef<T> t1o[2];        // T_{i+1}                         (not LAST)
    FieldRef<T> t2o[2];        // T_i                             (not LAST)
    FieldRef<T> bar[2];        // running filtered field, updated in place (FIRST: written only)
    double c;                  // 2/s_max or 2/(s_max dx_min^2)   (filter.py:232-236)
    double p0, pa, pb;         // p[0] (FIRST), p[i], p[i+1]
    int64_t nb;
    int32_t cpw;               // output columns per warp, <= CG2_COLS (strips of equal width)
    int32_t lw;                // staged columns per row: CG2_WARPS*cpw + 2*HALO
    // latitude bands with the ghost-row exchange fused in (gcmf_cheb_fused_halo): where the first / last two rows of
    // t1o (halo[0]) and t2o (halo[1]) go in the neighbouring GPUs' arrays; the flags and counters of halo[0] are used
    HaloRef<T> halo[2];
};

#ifdef __CUDACC__
// ---- operator policies: the marching recurrence of one row.
//   init(S, f0, c0)                       row R0 (this lane's column of field slot 0 / coefficient slot 0) becomes "row j"
//   row(S, un, vn, nxt, cur, lap, x)      (un, vn): raw field values of row j+1; nxt / cur: this lane's column of the
//                                         coefficient rows j+1 / j (array k at offset k*LW).  Returns the Laplacian at
//                                         (j, i) and the raw (u, v) of row j, then shifts the state one row north.
// All 32 lanes of the warp call row() together (shuffles).  The output of row() is valid from the second call on.

// VECTOR_C (kernels.py:647-696), the recurrence of cgrid_march_kernel: point products and weighted stresses in registers
template <typename T, int LW> struct CgOp {
    static constexpr int NC = 14;
    struct Row {
        T u, v;            // raw field values of row j
        T b, c, e;         // v/dxCv, v/dyCv, u/dxCu of row j (nan_to_num'ed field)
        T k0, k1, k2, k3;  // reciprocal spacings of row j
        T p1, p2;          // dyT^2 * str_xx, dxT^2 * str_xx at T point (j, i)
        T p3m;             // dxBu^2 * str_xy at q point (j-1, i)
    };
    static __device__ __forceinline__ void zero(Row& S) {
        S.u = S.v = S.b = S.c = S.e = S.k0 = S.k1 = S.k2 = S.k3 = S.p1 = S.p2 = S.p3m = T(0);
    }
    static __device__ __forceinline__ void init(Row& S, const T* f0, const T* c0) {
        S.u = f0[0]; S.v = f0[LW];
        S.k0 = c0[0]; S.k1 = c0[LW]; S.k2 = c0[2 * LW]; S.k3 = c0[3 * LW];
        const T zu = nan2num(S.u), zv = nan2num(S.v);
        S.b = zv * S.k1; S.c = zv * S.k2; S.e = zu * S.k3;
        S.p1 = S.p2 = S.p3m = T(0);
    }
    static __device__ __forceinline__ void row(Row& S, T un, T vn, const T* nxt, const T* cur, T (&lap)[2], T (&x)[2]) {
        const T k0_n = nxt[0], k1_n = nxt[LW], k2_n = nxt[2 * LW], k3_n = nxt[3 * LW];
        const T k4 = nxt[4 * LW], k5 = nxt[5 * LW], k8 = nxt[8 * LW], k9 = nxt[9 * LW];
        const T k6 = cur[6 * LW], k7 = cur[7 * LW], k10 = cur[10 * LW], k11 = cur[11 * LW];
        const T zu = nan2num(un), zv = nan2num(vn);
        const T a_n = zu * k0_n, b_n = zv * k1_n, c_n = zv * k2_n, e_n = zu * k3_n;
        const T a_w = __shfl_up_sync(0xffffffffu, a_n, 1);
        const T sxx = -(k4 * (a_n - a_w) - k5 * (b_n - S.b));      // kernels.py:653-661
        const T p1_n = k8 * sxx, p2_n = k9 * sxx;
        const T c_e = __shfl_down_sync(0xffffffffu, S.c, 1);
        const T sxy = -(k6 * (c_e - S.c) + k7 * (e_n - S.e));      // kernels.py:663-670
        const T p3_j = k10 * sxy, p4_j = k11 * sxy;
        const T p1_e = __shfl_down_sync(0xffffffffu, S.p1, 1);
        const T p4_w = __shfl_up_sync(0xffffffffu, p4_j, 1);
        T uc = S.k0 * (S.p1 - p1_e);                               // kernels.py:672-694
        uc = uc + S.k3 * (S.p3m - p3_j);
        lap[0] = uc * cur[12 * LW];
        T vc = S.k2 * (p4_w - p4_j);
        vc = vc - S.k1 * (S.p2 - p2_n);
        lap[1] = vc * cur[13 * LW];
        x[0] = S.u;
        x[1] = S.v;
        S.u = un; S.v = vn; S.b = b_n; S.c = c_n; S.e = e_n;
        S.k0 = k0_n; S.k1 = k1_n; S.k2 = k2_n; S.k3 = k3_n;
        S.p1 = p1_n; S.p2 = p2_n; S.p3m = p3_j;
    }
};

// VECTOR_B (kernels.py:740-837): the 10-term sums of OpVectorB, left to right, on the 8 precombined planes
// (cc, dun, dus, due, duw, dmc, dmn, dme; dms = -dmn, dmw = -dme, kernels.py:804-805)
template <typename T, int LW> struct BgOp {
    static constexpr int NC = 8;
    struct Row {
        T u, v;      // raw field values of row j
        T zu, zv;    // nan_to_num'ed row j
        T su, sv;    // nan_to_num'ed row j-1
    };
    static __device__ __forceinline__ void zero(Row& S) { S.u = S.v = S.zu = S.zv = S.su = S.sv = T(0); }
    static __device__ __forceinline__ void init(Row& S, const T* f0, const T*) {
        S.u = f0[0]; S.v = f0[LW];
        S.zu = nan2num(S.u); S.zv = nan2num(S.v);
        S.su = S.sv = T(0);
    }
    static __device__ __forceinline__ void row(Row& S, T un, T vn, const T*, const T* cur, T (&lap)[2], T (&x)[2]) {
        const T nu = nan2num(un), nv = nan2num(vn);                 // kernels.py:743-744
        const T ue = __shfl_down_sync(0xffffffffu, S.zu, 1), uw = __shfl_up_sync(0xffffffffu, S.zu, 1);
        const T we = __shfl_down_sync(0xffffffffu, S.zv, 1), ww = __shfl_up_sync(0xffffffffu, S.zv, 1);
        const T cc = cur[0], dun = cur[LW], dus = cur[2 * LW], due = cur[3 * LW], duw = cur[4 * LW];
        const T dmc = cur[5 * LW], dmn = cur[6 * LW], dme = cur[7 * LW];
        const T dms = -dmn, dmw = -dme;
        lap[0] = ((((((((cc * S.zu + dun * nu) + dus * S.su) + due * ue) + duw * uw) + dmc * S.zv) + dmn * nv) + dms * S.sv) +
                  dme * we) + dmw * ww;
        lap[1] = ((((((((cc * S.zv + dun * nv) + dus * S.sv) + due * we) + duw * ww) + dmc * S.zu) + dmn * nu) + dms * S.su) +
                  dme * ue) + dmw * uw;
        x[0] = S.u;
        x[1] = S.v;
        S.su = S.zu; S.sv = S.zv; S.zu = nu; S.zv = nv; S.u = un; S.v = vn;
    }
};

// mbarrier helpers on 32-bit shared-memory addresses (the consumers keep the barrier addresses in registers)
__device__ __forceinline__ void mbar_wait_a(uint32_t mb, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(mb), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t mb) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mb) : "memory");
}
// position of a staged row in a ring of N slots, advanced without division: slot index and phase parity of its barriers
template <int N> struct RingPos {
    int slot;
    unsigned phase;
    __device__ __forceinline__ void next() {
        if (++slot == N) {
            slot = 0;
            phase ^= 1u;
        }
    }
};

// EDGE bit 0: the block starts at recurrence step 1 (t1 = prepared input, no T_{i-2}, no bar yet);
// EDGE bit 1: the block ends at step n_steps (no T is stored).
// Warps 0..CG2_WARPS-1 consume; warp CG2_WARPS streams the field ring, warp CG2_WARPS+1 the coefficient ring (each at
// its own pace: the field rows are released one iteration earlier than the coefficient rows).
// HALO (latitude band): border row-bands wait for the neighbours' flags before touching the ghost rows, store their
// first / last two rows of T_{i+1} and T_i straight into the neighbours' ghost rows (peer memory, NVLink) as they emit
// them, and the last border CTA to finish raises the neighbours' flags -- halo_wait / halo_signal of gcmf.cu, the
// protocol of gcmf_cheb_step_halo with two rows and two arrays per block.  Border bands are scheduled first.
template <typename T, template <typename, int> class OPT, int EDGE, bool HALO>
__global__ void __launch_bounds__(32 * (CG2_WARPS + 2), 1) vec2_kernel(const __grid_constant__ Cg2Params<T> P, unsigned ctas_x, int ry) {
    using OP = OPT<T, Vec2Geom<T, 14>::LW>;
    using G = Vec2Geom<T, OP::NC>;
    constexpr bool FIRST = (EDGE & 1) != 0, LAST = (EDGE & 2) != 0;
    constexpr int LW = G::LW, FS = G::FS, CS = G::CS;
    constexpr int FSLOT = G::NF * LW, CSLOT = G::NC * LW;
    extern __shared__ __align__(128) unsigned char cg2_smem[];
    T* fring = reinterpret_cast<T*>(cg2_smem);
    T* cring = fring + (size_t)FS * FSLOT;
    uint64_t* fullF = reinterpret_cast<uint64_t*>(cring + (size_t)CS * CSLOT);
    uint64_t* emptyF = fullF + FS;
    uint64_t* fullC = emptyF + FS;
    uint64_t* emptyC = fullC + CS;

    unsigned bid = blockIdx.x;
    const int cx = (int)(bid % ctas_x);
    bid /= ctas_x;
    const unsigned nbu = (unsigned)P.nb;
    const int b = (int)(bid % nbu);
    const int ny = P.g.ny, nx = P.g.nx;
    const int band = HALO ? halo_band_order((int)(bid / nbu), (ny + ry - 1) / ry) : (int)(bid / nbu);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = band * ry, j1 = j0 + ry < ny ? j0 + ry : ny;
    const bool bottom = j0 == 0, top = j1 >= ny;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FS; ++s) { mbar_init(&fullF[s], 1); mbar_init(&emptyF[s], CG2_WARPS); }
        for (int s = 0; s < CS; ++s) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], CG2_WARPS); }
        fence_mbar_init();
    }
    __syncthreads();
    if (HALO) halo_wait<T>(P.halo[0], bottom, top);
    // staged rows R0 .. j1+1: step i+1 emits rows j0 .. j1-1, needs T_i on j0-1 .. j1, which needs the input on j0-2 .. j1+1
    const int R0 = j0 - 2;
    const int nstage = (j1 - j0) + 4;
    const int strip = CG2_WARPS * P.cpw;  // output columns per CTA

    if (warp >= CG2_WARPS) {
        // ---- producer warps: lane a owns one array and issues that array's row copy
        //   field warp:        lanes 0, 1: u, v (every row)   2, 3: T_{i-2} (rows j0-1 .. j1)   4, 5: bar (rows j0 .. j1-1)
        //   coefficient warp:  lanes 0 .. NC-1: the coefficient planes
        const bool coef = warp == CG2_WARPS + 1;
        // row r of the arrays: periodic y (ny >= 4), or a latitude band whose two ghost rows per side are present in memory
        const bool wrap = (P.g.flags & FL_WRAP_Y) != 0;
        auto rowidx = [&](int r) { return wrap ? (r < 0 ? r + ny : (r >= ny ? r - ny : r)) : r; };
        const T* base = nullptr;
        int64_t pitch = 0;
        int kind = -1;  // 0: every row, 2: T_{i-2}, 3: bar
        if (coef) {
#pragma unroll
            for (int k = 0; k < G::NC; ++k)
                if (lane == k) {
                    base = plane_base<T>(P.plane[k], b);
                    pitch = P.plane[k].pitch;
                    kind = 0;
                }
        } else if (lane < 2) {
            base = P.t1[lane].p + (int64_t)b * P.t1[lane].bstride;
            pitch = P.t1[lane].pitch;
            kind = 0;
        } else if (!FIRST && lane < 4) {
            base = P.t2[lane - 2].p + (int64_t)b * P.t2[lane - 2].bstride;
            pitch = P.t2[lane - 2].pitch;
            kind = 2;
        } else if (!FIRST && lane < 6) {
            base = P.bar[lane - 4].p + (int64_t)b * P.bar[lane - 4].bstride;
            pitch = P.bar[lane - 4].pitch;
            kind = 3;
        }
        T* const ring = coef ? cring : fring;
        uint64_t* const full = coef ? fullC : fullF;
        uint64_t* const empty = coef ? emptyC : emptyF;
        const int nslot = coef ? CS : FS, slot_elems = coef ? CSLOT : FSLOT;
        T* const dst0 = ring + lane * LW;
        const int col0 = cx * strip - G::HALO;
        const int gx = col0 < 0 ? col0 + nx : col0;
        const int lw = P.lw;
        const int n1 = (nx - gx) < lw ? (nx - gx) : lw;
        const unsigned rowb = (unsigned)(lw * sizeof(T));
        if (HALO) asm volatile("fence.proxy.async;" ::: "memory");  // ghost rows written by the peers, acquired above
        int slot = 0;
        unsigned phase = 0;  // parity of the "empty" phase the producer waits for (second lap: 0, third: 1, ...)
        for (int q = 0; q < nstage; ++q) {
            if (q >= nslot) {
                mbar_wait(&empty[slot], phase);
                fence_proxy_async();
            }
            const bool has_t2 = !FIRST && q >= 1 && q <= nstage - 2;
            const bool has_bar = !FIRST && q >= 2 && q <= nstage - 3;
            if (lane == 0)
                mbar_expect_tx(&full[slot], coef ? rowb * (unsigned)G::NC : rowb * (2u + (has_t2 ? 2u : 0u) + (has_bar ? 2u : 0u)));
            __syncwarp();
            if (kind == 0 || (kind == 2 && has_t2) || (kind == 3 && has_bar)) {
                const T* row = base + (int64_t)rowidx(R0 + q) * pitch;
                T* dst = dst0 + (size_t)slot * slot_elems;
                bulk_copy_g2s(dst, row + gx, (unsigned)(n1 * sizeof(T)), &full[slot]);
                if (n1 < lw) bulk_copy_g2s(dst + n1, row, (unsigned)((lw - n1) * sizeof(T)), &full[slot]);
            }
            if (++slot == nslot) {
                slot = 0;
                if (q >= nslot) phase ^= 1u;
            }
        }
    } else {
    // ---- consumers
    const int lc = G::HALO - 2 + warp * P.cpw + lane;              // this lane's column in a staged row
    const int i = cx * strip + warp * P.cpw + lane - 2;            // global column (unwrapped; < 0 or >= nx: never emitted)
    const bool emit = lane >= 2 && lane < 2 + P.cpw && i >= 0 && i < nx;
    const T cc = (T)P.c;
    const T* const fcol = fring + lc;
    const T* const ccol = cring + lc;
    const uint32_t fullF_a = smem_u32(fullF), emptyF_a = smem_u32(emptyF), fullC_a = smem_u32(fullC), emptyC_a = smem_u32(emptyC);
    typename OP::Row S1, S2;
    OP::zero(S2);
    {   // prologue: row R0 becomes the state of step i
        mbar_wait_a(fullF_a, 0);
        mbar_wait_a(fullC_a, 0);
        OP::init(S1, fcol, ccol);
    }
    // between the two steps: T_{i-1} and bar-after-step-i of the row step i+1 emits in the next iteration
    T xd[2] = {T(0), T(0)}, bd[2] = {T(0), T(0)};
    T* pb[2];
    T* p1[2] = {nullptr, nullptr};
    T* p2[2] = {nullptr, nullptr};
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // output row pointers (row j0 at the first emission, s = 3)
        pb[k] = P.bar[k].p + ((int64_t)b * P.bar[k].bstride + (int64_t)j0 * P.bar[k].pitch + i);
        if (!LAST) {
            p1[k] = P.t1o[k].p + ((int64_t)b * P.t1o[k].bstride + (int64_t)j0 * P.t1o[k].pitch + i);
            p2[k] = P.t2o[k].p + ((int64_t)b * P.t2o[k].bstride + (int64_t)j0 * P.t2o[k].pitch + i);
        }
    }
    // ring positions of staged rows s+1 ("next") in both rings; element offsets of rows s-1, s, s+1
    RingPos<FS> fn{0, 0};
    RingPos<CS> cn{0, 0};
    int fs_cur = 0, cs_cur = 0, cs_prev = 0;  // slots of rows s (both rings) and s-1 (coefficient ring)
#pragma unroll CG2_UNROLL
    for (int s = 0; s + 1 < nstage; ++s) {
        fn.next();
        cn.next();
        const int f_nxt = fn.slot * FSLOT, c_nxt = cn.slot * CSLOT;
        const int f_cur = fs_cur * FSLOT, c_cur = cs_cur * CSLOT, c_prev = cs_prev * CSLOT;
        mbar_wait_a(fullF_a + 8u * (unsigned)fn.slot, fn.phase);
        mbar_wait_a(fullC_a + 8u * (unsigned)cn.slot, cn.phase);
        // ---- step i at row j = R0 + s (valid from s = 1 on: rows j0-1 .. j1)
        const T* fnx = fcol + f_nxt;
        const T* fc = fcol + f_cur;
        T lap[2], x[2], tn[2], b1[2];
        OP::row(S1, fnx[0], fnx[LW], ccol + c_nxt, ccol + c_cur, lap, x);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const T a = -x[k] - cc * lap[k];                                   // filter.py:232-236
            if (FIRST) {
                tn[k] = a;
                b1[k] = (T)bar_update(P.p0 * (double)x[k], P.pa, (double)a);   // filter.py:253-254
            } else {
                tn[k] = cheb_next<T>(a, fc[(2 + k) * LW]);                      // filter.py:263-264
                b1[k] = (T)bar_update((double)fc[(4 + k) * LW], P.pa, (double)tn[k]);  // filter.py:265-266
            }
        }
        // ---- step i+1 at row j-1: its "row j+1" is the T_i(j) just produced (valid from s = 3 on: rows j0 .. j1-1)
        if (s >= 1) {
            T lap2[2], x2[2];
            OP::row(S2, tn[0], tn[1], ccol + c_cur, ccol + c_prev, lap2, x2);
            if (s >= 3) {
                if (emit) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const T a2 = -x2[k] - cc * lap2[k];
                        const T t = cheb_next<T>(a2, xd[k]);
                        if (!LAST) {
                            *p1[k] = t;
                            *p2[k] = x2[k];
                            if (HALO) {  // rows 0, 1 -> the south neighbour's north ghost rows; ny-2, ny-1 -> the north neighbour's south ghosts
                                const int rj = j0 + s - 3;
                                if (rj < 2 && P.halo[0].south[k]) {
                                    const int64_t o = (int64_t)b * P.halo[0].sbs + (int64_t)rj * P.t1o[k].pitch + i;
                                    P.halo[0].south[k][o] = t;
                                    P.halo[1].south[k][o] = x2[k];
                                }
                                if (rj >= ny - 2 && P.halo[0].north[k]) {
                                    const int64_t o = (int64_t)b * P.halo[0].nbs + (int64_t)(rj - (ny - 2)) * P.t1o[k].pitch + i;
                                    P.halo[0].north[k][o] = t;
                                    P.halo[1].north[k][o] = x2[k];
                                }
                            }
                        }
                        *pb[k] = (T)bar_update((double)bd[k], P.pb, (double)t);
                    }
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    pb[k] += P.bar[k].pitch;
                    if (!LAST) { p1[k] += P.t1o[k].pitch; p2[k] += P.t2o[k].pitch; }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            xd[k] = x[k];
            bd[k] = b1[k];
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_a(emptyF_a + 8u * (unsigned)fs_cur);                // field row s: read for the last time just now
            if (s >= 1) mbar_arrive_a(emptyC_a + 8u * (unsigned)cs_prev);   // coefficient row s-1: step i+1 is done with it
        }
        cs_prev = cs_cur;
        cs_cur = cn.slot;
        fs_cur = fn.slot;
    }
    }  // consumers
    if (HALO) halo_signal<T>(P.halo[0], bottom, top, ctas_x * nbu);
}
#endif  // __CUDACC__

}  // namespace gcmf

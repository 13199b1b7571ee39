// gcmf_vec2.cuh -- TWO Chebyshev steps of a VECTOR operator per HBM round trip (temporal blocking of the vector
// recurrence, filter.py:225-283, for the C-grid viscous Laplacian kernels.py:647-696 and the B-grid Laplacian
// kernels.py:740-837).
//
// The one-step kernels already stream every array once per step and sit at 0.8-0.9 of the HBM roofline (C-grid: 20 rows
// in -- u, v, 14 coefficient planes, T_{i-2}, bar -- and 4 rows out per row of points = 192 B per grid-point step at
// nb = 1).  The only way below that is to do more steps per byte.  This kernel keeps the row-streaming shape of
// cgrid_tma_kernel -- a CTA owns a strip of columns and marches north through a band of rows -- and runs step i+1
// LAG rows behind step i inside the same march:
//
//   iteration s (row j = R0 + s):   step i   consumes input row j+1 from the ring and produces T_i(j)          (registers)
//                                   step i+1 consumes T_i(j) [LAG 1] or T_i(j-1) [LAG 2] as its "row j+1" and
//                                            produces T_{i+1}(j-LAG)                                            (registers)
//
// Both steps are the same marching recurrence (an operator policy: the state of "row j" lives in registers, W / E
// neighbours come from the adjacent lanes by warp shuffle), fed from shared memory (step i) or from the registers step i
// wrote (step i+1).  The coefficient rows are staged ONCE for both steps (a row lives for 2 + LAG iterations in its ring
// instead of two), T_i never touches HBM on its way into step i+1:
//
//   C-grid, per 2 steps and row: 20 rows read + 6 rows written (T_{i+1}, T_i, bar) = 104 B per grid-point step (192: x 1.85)
//   B-grid, per 2 steps and row: 14 rows read + 6 rows written                     =  80 B per grid-point step (144: x 1.80)
//
// Halo: a warp's 32 lanes see 30 valid columns after step i and 28 after step i+1 (lanes 2..29 emit); a band primes with
// two extra rows on either side.  Two rings (TMA bulk copies, cp.async.bulk -> UBLKCP, one lane per array, mbarrier
// complete_tx / consumer-release "empty" barriers as in cgrid_tma_kernel): the 6 field rows are released after one
// iteration, the coefficient rows after 1 + LAG.
//
// Same expressions in the same order as the one-step kernels: results are bit-identical to two gcmf_cheb_step calls.
// Device-only (mbarriers + TMA); the host emulator takes the one-step path, which the GPU tests compare against.
#pragma once
#include "gcmf_fused.cuh"

namespace gcmf {

#ifndef GCMF_CG2_WARPS
#define GCMF_CG2_WARPS 8
#endif
#ifndef GCMF_CG2_LAG
#define GCMF_CG2_LAG 1  // rows between step i and step i+1: 1 = step i+1 consumes T_i(j) in the iteration that made it;
#endif                  // 2 = one iteration later (two independent chains per iteration, one more coefficient row live)
#ifndef GCMF_CG2_FS
#define GCMF_CG2_FS (GCMF_CG2_LAG == 1 ? 5 : 4)   // field ring slots (a row is live for 2 iterations)
#endif
#ifndef GCMF_CG2_CS
#define GCMF_CG2_CS (GCMF_CG2_LAG == 1 ? 6 : 7)   // coefficient ring slots, 14 arrays (a row is live for 2 + LAG iterations)
#endif
constexpr int CG2_WARPS = GCMF_CG2_WARPS;  // consumer warps per CTA (+ 1 producer warp)
constexpr int CG2_COLS = 28;               // at most 28 output columns per warp (lanes 2..29)

template <typename T, int NC_> struct Vec2Geom {
    static constexpr int AV = 16 / (int)sizeof(T);
    static constexpr int HALO = AV < 2 ? 2 : AV;             // staged halo columns per side (two are needed), 16-byte aligned
    static constexpr int LW = CG2_WARPS * CG2_COLS + 2 * HALO;  // row pitch of the rings: 228 (f64) / 232 (f32)
    static constexpr int NC = NC_, NF = 6;                   // coefficient arrays; field arrays: u, v, t2u, t2v, bar_u, bar_v
    static constexpr int FS = GCMF_CG2_FS;
    static constexpr int CS = NC_ > 8 ? GCMF_CG2_CS : GCMF_CG2_CS + 2;  // fewer planes: a deeper ring fits
    static constexpr int NTHREADS = 32 * (CG2_WARPS + 1);
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

// EDGE bit 0: the block starts at recurrence step 1 (t1 = prepared input, no T_{i-2}, no bar yet);
// EDGE bit 1: the block ends at step n_steps (no T is stored).
template <typename T, template <typename, int> class OPT, int EDGE>
__global__ void __launch_bounds__(32 * (CG2_WARPS + 1), 1) vec2_kernel(const __grid_constant__ Cg2Params<T> P, unsigned ctas_x, int ry) {
    using OP = OPT<T, Vec2Geom<T, 14>::LW>;
    using G = Vec2Geom<T, OP::NC>;
    constexpr bool FIRST = (EDGE & 1) != 0, LAST = (EDGE & 2) != 0;
    constexpr int LW = G::LW, FS = G::FS, CS = G::CS, LAG = GCMF_CG2_LAG;
    static_assert(LAG == 1 || LAG == 2, "GCMF_CG2_LAG");
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
    const int band = (int)(bid / nbu);
    const int ny = P.g.ny, nx = P.g.nx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = band * ry, j1 = j0 + ry < ny ? j0 + ry : ny;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FS; ++s) { mbar_init(&fullF[s], 1); mbar_init(&emptyF[s], CG2_WARPS); }
        for (int s = 0; s < CS; ++s) { mbar_init(&fullC[s], 1); mbar_init(&emptyC[s], CG2_WARPS); }
        fence_mbar_init();
    }
    __syncthreads();
    // staged rows R0 .. j1+1: step i+1 emits rows j0 .. j1-1, needs T_i on j0-1 .. j1, which needs the input on j0-2 .. j1+1
    const int R0 = j0 - 2;
    const int nstage = (j1 - j0) + 4;
    auto rowidx = [&](int r) { return r < 0 ? r + ny : (r >= ny ? r - ny : r); };  // periodic y (ny >= 4)
    const int strip = CG2_WARPS * P.cpw;  // output columns per CTA

    if (warp == CG2_WARPS) {
        // ---- producer warp: lane a owns one array and issues that array's row copy.
        //   lanes 0, 1: u, v     lanes 2..NC+1: the coefficient planes     NC+2, NC+3: T_{i-2}     NC+4, NC+5: bar
        const T* base = nullptr;
        int64_t pitch = 0;
        int kind = -1;  // 0: field (every row), 1: coefficient, 2: T_{i-2} (rows j0-1 .. j1), 3: bar (rows j0 .. j1-1)
        T* dst0 = nullptr;
        if (lane < 2) {
            base = P.t1[lane].p + (int64_t)b * P.t1[lane].bstride;
            pitch = P.t1[lane].pitch;
            kind = 0;
            dst0 = fring + lane * LW;
        }
#pragma unroll
        for (int k = 0; k < G::NC; ++k)
            if (lane == 2 + k) {
                base = plane_base<T>(P.plane[k], b);
                pitch = P.plane[k].pitch;
                kind = 1;
                dst0 = cring + k * LW;
            }
        if (!FIRST) {
            constexpr int L2 = G::NC + 2, LB = G::NC + 4;
            if (lane == L2 || lane == L2 + 1) {
                base = P.t2[lane - L2].p + (int64_t)b * P.t2[lane - L2].bstride;
                pitch = P.t2[lane - L2].pitch;
                kind = 2;
                dst0 = fring + (2 + lane - L2) * LW;
            }
            if (lane == LB || lane == LB + 1) {
                base = P.bar[lane - LB].p + (int64_t)b * P.bar[lane - LB].bstride;
                pitch = P.bar[lane - LB].pitch;
                kind = 3;
                dst0 = fring + (4 + lane - LB) * LW;
            }
        }
        const int col0 = cx * strip - G::HALO;
        const int gx = col0 < 0 ? col0 + nx : col0;
        const int lw = P.lw;
        const int n1 = (nx - gx) < lw ? (nx - gx) : lw;
        const unsigned rowb = (unsigned)(lw * sizeof(T));
        for (int q = 0; q < nstage; ++q) {
            const int sf = q % FS, sc = q % CS;
            if (q >= FS) mbar_wait(&emptyF[sf], (unsigned)(((q / FS) - 1) & 1));
            if (q >= CS) mbar_wait(&emptyC[sc], (unsigned)(((q / CS) - 1) & 1));
            if (q >= FS || q >= CS) fence_proxy_async();
            const int r = R0 + q;
            const bool has_t2 = !FIRST && q >= 1 && q <= nstage - 2;
            const bool has_bar = !FIRST && q >= 2 && q <= nstage - 3;
            if (lane == 0) {
                mbar_expect_tx(&fullF[sf], rowb * (2u + (has_t2 ? 2u : 0u) + (has_bar ? 2u : 0u)));
                mbar_expect_tx(&fullC[sc], rowb * (unsigned)G::NC);
            }
            __syncwarp();
            const bool go = kind == 0 || kind == 1 || (kind == 2 && has_t2) || (kind == 3 && has_bar);
            if (go) {
                const T* row = base + (int64_t)rowidx(r) * pitch;
                const bool coef = kind == 1;
                T* dst = dst0 + (coef ? (size_t)sc * CSLOT : (size_t)sf * FSLOT);
                uint64_t* fb = coef ? &fullC[sc] : &fullF[sf];
                bulk_copy_g2s(dst, row + gx, (unsigned)(n1 * sizeof(T)), fb);
                if (n1 < lw) bulk_copy_g2s(dst + n1, row, (unsigned)((lw - n1) * sizeof(T)), fb);
            }
        }
        return;
    }

    // ---- consumers
    const int lc = G::HALO - 2 + warp * P.cpw + lane;              // this lane's column in a staged row
    const int i = cx * strip + warp * P.cpw + lane - 2;            // global column (unwrapped; < 0 or >= nx: never emitted)
    const bool emit = lane >= 2 && lane < 2 + P.cpw && i >= 0 && i < nx;
    const T cc = (T)P.c;
    auto frow = [&](int q) { return fring + (size_t)(q % FS) * FSLOT + lc; };
    auto crow = [&](int q) { return cring + (size_t)(q % CS) * CSLOT + lc; };
    typename OP::Row S1, S2;
    OP::zero(S2);
    {   // prologue: row R0 becomes the state of step i
        mbar_wait(&fullF[0], 0);
        mbar_wait(&fullC[0], 0);
        OP::init(S1, frow(0), crow(0));
    }
    // delay lines between the two steps: T_{i-1} and bar-after-step-i of the rows step i+1 has not emitted yet, and (LAG 2)
    // the T_i row step i+1 consumes one iteration after step i produced it
    T xd[LAG][2], bd[LAG][2], tnd[2] = {T(0), T(0)};
#pragma unroll
    for (int d = 0; d < LAG; ++d) xd[d][0] = xd[d][1] = bd[d][0] = bd[d][1] = T(0);
    int64_t o1[2] = {0, 0}, o2[2] = {0, 0}, ob[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // running element offsets of the output row (starts at row j0 with the first emission)
        ob[k] = (int64_t)b * P.bar[k].bstride + (int64_t)j0 * P.bar[k].pitch + i;
        if (!LAST) {
            o1[k] = (int64_t)b * P.t1o[k].bstride + (int64_t)j0 * P.t1o[k].pitch + i;
            o2[k] = (int64_t)b * P.t2o[k].bstride + (int64_t)j0 * P.t2o[k].pitch + i;
        }
    }
    // step i+1 on the T_i row (un, vn): its state row is LAG rows behind step i's; emits one output row from iteration
    // 2 + LAG on (row j0)
    auto step2 = [&](int s, T un, T vn) {
        T lap2[2], x2[2];
        OP::row(S2, un, vn, crow(s - LAG + 1), crow(s - LAG), lap2, x2);
        if (s >= 2 + LAG) {
            if (emit) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const T a2 = -x2[k] - cc * lap2[k];
                    const T t = cheb_next<T>(a2, xd[LAG - 1][k]);
                    if (!LAST) {
                        P.t1o[k].p[o1[k]] = t;
                        P.t2o[k].p[o2[k]] = x2[k];
                    }
                    P.bar[k].p[ob[k]] = (T)bar_update((double)bd[LAG - 1][k], P.pb, (double)t);
                }
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                ob[k] += P.bar[k].pitch;
                if (!LAST) { o1[k] += P.t1o[k].pitch; o2[k] += P.t2o[k].pitch; }
            }
        }
    };
    const int niter = nstage - 2 + LAG;
    for (int s = 0; s < niter; ++s) {
        const int sn = s + 1;
        const bool has_next = sn < nstage;  // (LAG 2: the last iteration only drains step i+1)
        if (has_next) {
            mbar_wait(&fullF[sn % FS], (unsigned)((sn / FS) & 1));
            mbar_wait(&fullC[sn % CS], (unsigned)((sn / CS) & 1));
        }
        // LAG 2: step i+1 consumes the T_i row of the PREVIOUS iteration -- independent of this iteration's step i, so the
        // two dependent fp64 chains of an iteration overlap
        if (LAG == 2 && s >= 2) step2(s, tnd[0], tnd[1]);
        T x[2] = {T(0), T(0)}, tn[2] = {T(0), T(0)}, b1[2] = {T(0), T(0)};
        if (has_next) {
            // ---- step i at row j = R0 + s (valid from s = 1 on: rows j0-1 .. j1)
            const T* fn = frow(sn);
            const T* fc = frow(s);
            T lap[2];
            OP::row(S1, fn[0], fn[LW], crow(sn), crow(s), lap, x);
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
        }
        // LAG 1: step i+1 at row j-1 consumes the T_i(j) just produced
        if (LAG == 1 && s >= 1) step2(s, tn[0], tn[1]);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (LAG == 2) { xd[LAG - 1][k] = xd[0][k]; bd[LAG - 1][k] = bd[0][k]; tnd[k] = tn[k]; }
            xd[0][k] = x[k];
            bd[0][k] = b1[k];
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&emptyF[s % FS]);                       // field row s: u, v read last iteration, T_{i-2} / bar just now
            if (s >= LAG) mbar_arrive(&emptyC[(s - LAG) % CS]);  // coefficient row s-LAG: step i+1 is done with it
        }
    }
}
#endif  // __CUDACC__

}  // namespace gcmf

// gcmf_fused.cuh -- temporally blocked Chebyshev steps: k <= 4 recurrence steps per HBM round trip.
//
// One CTA owns one spatial tile (core CH x CW plus a 4-cell halo) and loops over a slab of batch
// slices ("levels").  Per level:
//   * the T_{i-1} tile (+halo) is staged in shared memory by the TMA engine: one bulk async copy
//     (cp.async.bulk, SASS UBLKCP) per tile row, split in two where the row crosses the periodic
//     x boundary, completing on an mbarrier; the copy for level l+1 is in flight while level l is
//     computed (double buffer);
//   * T_{i-2} is only ever needed at the point itself: it goes straight from HBM into registers;
//   * k steps run out of shared memory on a region that shrinks by one cell per step (overlapped /
//     ghost-zone tiling), ping-ponging between two tiles: step s reads its neighbours from tile S and
//     overwrites, in place, the T_{i-2} value held in tile D;
//   * the running filtered field `bar` stays in registers for the k steps and is read / written once;
//   * the coefficient tiles (ce, cn, ra) are staged once per CTA and reused for every level.
// HBM traffic per grid-point step: (2*rho + 4) * w / k bytes instead of 5 * w  (rho = tile/core area).
//
// The arithmetic of a point is the same inline code as the one-step kernels (flux_lap, -fmad=false),
// and `bar` is accumulated in the same order, so fused and un-fused results are bit-identical.
//
// Every phase below is a __host__ __device__ function of (thread id, per-thread state): the device
// kernel separates the phases with __syncthreads()/mbarrier waits, the test-only host emulator runs
// each phase for all thread ids in turn.
#pragma once
#include "gcmf_stencils.cuh"

namespace gcmf {

template <typename T> struct FusedGeom {
    static constexpr int VX = 16 / (int)sizeof(T);  // one 16-byte vector per thread and row
    static constexpr int H = 4;                     // halo width = max fused steps
    static constexpr int NTX = 64;                  // threads along x
    static constexpr int TW = NTX * VX;             // tile width  incl. halo: 128 (f64) / 256 (f32)
    static constexpr int R = 4;                     // consecutive rows per thread
    static constexpr int NTY = 9;                   // threads along y
    static constexpr int TH = R * NTY;              // tile height incl. halo: 36
    static constexpr int CW = TW - 2 * H;           // core width  120 / 248
    static constexpr int CH = TH - 2 * H;           // core height 28
    static constexpr int NTHREADS = NTX * NTY;      // 576
    static constexpr int PLANE = TH * TW;           // elements per shared-memory tile (36 KiB)
    static constexpr int NPLANES = 6;               // P0, P1 (double-buffered T1), Q, ce, cn, ra
    static constexpr size_t SMEM_BYTES = (size_t)NPLANES * PLANE * sizeof(T) + 64;
};

template <typename T> struct FusedParams {
    Geo g;
    PlaneRef plane[3];          // ce, cn, ra (shared 2-D planes: nb == 1)
    FieldRef<const T> t1_in;    // T_{i-1}
    FieldRef<const T> t2_in;    // T_{i-2}
    FieldRef<T> t1_out;         // T_{i+k-1}
    FieldRef<T> t2_out;         // T_{i+k-2}
    FieldRef<T> bar;            // bar += sum_s p[s] T_{i+s}
    double c;
    double p[FusedGeom<T>::H];  // Chebyshev coefficients of the k steps
    int32_t k;                  // fused steps, 1..H
    int32_t ncx, ncy;           // core tiles along x / y
    int64_t nb;                 // batch slices
    int32_t levels_per_cta;
};

template <typename T> struct FusedThread {  // per-thread registers that live across phases
    T t2[FusedGeom<T>::R][FusedGeom<T>::VX];
    T acc[FusedGeom<T>::R][FusedGeom<T>::VX];
};

GCMF_HD int wrap_index(int v, int n) {
    v %= n;
    return v < 0 ? v + n : v;
}

// One flux-form Laplacian value (shared with OpFlux::apply; kernels.py:297-315, 564-585).
template <typename T>
GCMF_HD T flux_lap(T oc, T ow, T oe, T on, T os, T ce, T cew, T cn, T cs, T ra) {
    const T fe = (oe - oc) * ce;
    const T fw = (oc - ow) * cew;
    const T fn = (on - oc) * cn;
    const T fs = (oc - os) * cs;
    return (((fe - fw) + fn) - fs) * ra;
}

// ---- bulk async copy global -> shared, completing on an mbarrier (device) / memcpy (host emulator) ----
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* mb, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mb, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mb)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mb, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(mb)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif

GCMF_HD void bulk_copy_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* mb) {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(mb))
                 : "memory");
#else
    (void)mb;
    const char* s = (const char*)src_gmem;
    char* d = (char*)dst_smem;
    for (unsigned i = 0; i < bytes; ++i) d[i] = s[i];
#endif
}

template <typename T> struct FusedTile {
    using G = FusedGeom<T>;
    const FusedParams<T>& P;
    int gy0, gx0;      // global coordinates of tile element (0,0) (may be negative: wraps)
    int cy0, cx0;      // global coordinates of the first core element
    T* smem;           // NPLANES tiles

    GCMF_HD FusedTile(const FusedParams<T>& P_, int tile, T* smem_) : P(P_), smem(smem_) {
        const int cx = tile % P.ncx, cy = tile / P.ncx;
        cy0 = cy * G::CH;
        cx0 = cx * G::CW;
        gy0 = cy0 - G::H;
        gx0 = cx0 - G::H;
    }
    GCMF_HD T* tileP(int buf) const { return smem + (size_t)buf * G::PLANE; }
    GCMF_HD T* tileQ() const { return smem + (size_t)2 * G::PLANE; }
    GCMF_HD T* tileC(int which) const { return smem + (size_t)(3 + which) * G::PLANE; }

    // Stage tile row r of a (2-D slice of a) global array: <= 2 bulk copies (split at the x wrap).
    GCMF_HD void copy_row(T* dst_tile, const T* src_slice, int64_t pitch, int r, uint64_t* mb) const {
        const int gy = wrap_index(gy0 + r, P.g.ny);
        const int gx = wrap_index(gx0, P.g.nx);
        const T* row = src_slice + (int64_t)gy * pitch;
        const int n1 = (P.g.nx - gx) < G::TW ? (P.g.nx - gx) : G::TW;
        bulk_copy_g2s(dst_tile + r * G::TW, row + gx, (unsigned)(n1 * sizeof(T)), mb);
        if (n1 < G::TW) bulk_copy_g2s(dst_tile + r * G::TW + n1, row, (unsigned)((G::TW - n1) * sizeof(T)), mb);
    }
    // phase: thread r < TH issues the coefficient rows
    GCMF_HD void issue_coef_row(int r, uint64_t* mb) const {
        for (int s = 0; s < 3; ++s)
            copy_row(tileC(s), reinterpret_cast<const T*>(P.plane[s].p), P.plane[s].pitch, r, mb);
    }
    // phase: thread r < TH issues row r of T1(level) into P[buf]
    GCMF_HD void issue_t1_row(int r, int64_t level, int buf, uint64_t* mb) const {
        copy_row(tileP(buf), P.t1_in.p + level * P.t1_in.bstride, P.t1_in.pitch, r, mb);
    }

    // does this thread's row q / column group belong to the core AND lie inside the domain?
    GCMF_HD bool owns_cols(int tx) const {
        const int lc0 = tx * G::VX;
        return lc0 >= G::H && lc0 < G::TW - G::H && (cx0 + lc0 - G::H) < P.g.nx;
    }
    GCMF_HD bool owns_row(int lr) const { return lr >= G::H && lr < G::TH - G::H && (cy0 + lr - G::H) < P.g.ny; }

    // phase: T2 (whole tile) and bar (owned points) from HBM into registers
    GCMF_HD void load_regs(int tid, int64_t level, FusedThread<T>& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const int lc0 = tx * G::VX;
        const int gx = wrap_index(gx0 + lc0, P.g.nx);
        const T* t2 = P.t2_in.p + level * P.t2_in.bstride;
        const bool oc = owns_cols(tx);
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int lr = ty * G::R + q;
            const int gy = wrap_index(gy0 + lr, P.g.ny);
            Ld<T, G::VX>::go(t2 + (int64_t)gy * P.t2_in.pitch + gx, st.t2[q]);
            if (oc && owns_row(lr)) {
                Ld<T, G::VX>::go(P.bar.p + level * P.bar.bstride + (int64_t)(cy0 + lr - G::H) * P.bar.pitch +
                                     (cx0 + lc0 - G::H), st.acc[q]);
            } else {
#pragma unroll
                for (int v = 0; v < G::VX; ++v) st.acc[q][v] = T(0);
            }
        }
    }

    // phase: recurrence step s (1-based) on the region [s, TH-s) x [s, TW-s); S = source tile, D = in-place tile
    template <bool FIRST>
    GCMF_HD void step(int tid, int s, const T* __restrict__ S, T* D, FusedThread<T>& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const int lc0 = tx * G::VX;
        if (lc0 + G::VX <= s || lc0 >= G::TW - s) return;  // column group outside the region
        const int iw = lc0 > 0 ? lc0 - 1 : 0;
        const int ie = lc0 + G::VX < G::TW ? lc0 + G::VX : G::TW - 1;
        const T* CE = tileC(0);
        const T* CN = tileC(1);
        const T* RA = tileC(2);
        const T c = (T)P.c;
        const double pk = P.p[s - 1];
        bool have = false;
        T xs[G::VX], xc[G::VX], xn[G::VX];  // sanitized rows lr-1, lr, lr+1
        T raw[G::VX], rawn[G::VX];          // raw rows lr, lr+1
        T cs[G::VX], cn[G::VX];             // north-face coefficients of rows lr-1, lr
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int lr = ty * G::R + q;
            if (lr < s || lr >= G::TH - s) {
                have = false;
                continue;
            }
            if (!have) {
                Ld<T, G::VX>::go(S + (lr - 1) * G::TW + lc0, xs);
                Ld<T, G::VX>::go(S + lr * G::TW + lc0, raw);
                Ld<T, G::VX>::go(CN + (lr - 1) * G::TW + lc0, cs);
#pragma unroll
                for (int v = 0; v < G::VX; ++v) {
                    xs[v] = nan2num(xs[v]);
                    xc[v] = nan2num(raw[v]);
                }
            }
            Ld<T, G::VX>::go(S + (lr + 1) * G::TW + lc0, rawn);
#pragma unroll
            for (int v = 0; v < G::VX; ++v) xn[v] = nan2num(rawn[v]);
            const T ow = nan2num(S[lr * G::TW + iw]);
            const T oe = nan2num(S[lr * G::TW + ie]);
            T ce[G::VX], ra[G::VX], t2[G::VX], t0[G::VX];
            Ld<T, G::VX>::go(CE + lr * G::TW + lc0, ce);
            const T cew = CE[lr * G::TW + iw];
            Ld<T, G::VX>::go(CN + lr * G::TW + lc0, cn);
            Ld<T, G::VX>::go(RA + lr * G::TW + lc0, ra);
            if (FIRST) {
#pragma unroll
                for (int v = 0; v < G::VX; ++v) t2[v] = st.t2[q][v];
            } else {
                Ld<T, G::VX>::go(D + lr * G::TW + lc0, t2);
            }
#pragma unroll
            for (int v = 0; v < G::VX; ++v) {
                const T o_e = v == G::VX - 1 ? oe : xc[v + 1];
                const T o_w = v == 0 ? ow : xc[v - 1];
                const T lap = flux_lap<T>(xc[v], o_w, o_e, xn[v], xs[v], ce[v], v == 0 ? cew : ce[v - 1], cn[v], cs[v],
                                          ra[v]);
                const T a = -raw[v] - c * lap;                                  // filter.py:171
                t0[v] = T(2) * a - t2[v];                                       // filter.py:197-203
                st.acc[q][v] = (T)((double)st.acc[q][v] + pk * (double)t0[v]);  // filter.py:204
            }
            St<T, G::VX>::go(D + lr * G::TW + lc0, t0);
#pragma unroll
            for (int v = 0; v < G::VX; ++v) {  // slide the window one row north
                xs[v] = xc[v];
                xc[v] = xn[v];
                raw[v] = rawn[v];
                cs[v] = cn[v];
            }
            have = true;
        }
    }

    // phase: write the owned core points of T1', T2' and bar back to HBM
    GCMF_HD void store(int tid, int64_t level, const T* T1n, const T* T2n, const FusedThread<T>& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        if (!owns_cols(tx)) return;
        const int lc0 = tx * G::VX;
        const int gx = cx0 + lc0 - G::H;
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int lr = ty * G::R + q;
            if (!owns_row(lr)) continue;
            const int gy = cy0 + lr - G::H;
            T a[G::VX], b[G::VX];
            Ld<T, G::VX>::go(T1n + lr * G::TW + lc0, a);
            Ld<T, G::VX>::go(T2n + lr * G::TW + lc0, b);
            St<T, G::VX>::go(P.t1_out.p + level * P.t1_out.bstride + (int64_t)gy * P.t1_out.pitch + gx, a);
            St<T, G::VX>::go(P.t2_out.p + level * P.t2_out.bstride + (int64_t)gy * P.t2_out.pitch + gx, b);
            St<T, G::VX>::go(P.bar.p + level * P.bar.bstride + (int64_t)gy * P.bar.pitch + gx, st.acc[q]);
        }
    }
};

#ifdef __CUDACC__
template <typename T>
__global__ void __launch_bounds__(FusedGeom<T>::NTHREADS, 1) fused_flux_kernel(const __grid_constant__ FusedParams<T> P) {
    using G = FusedGeom<T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    uint64_t* mb = reinterpret_cast<uint64_t*>(smem_raw + (size_t)G::NPLANES * G::PLANE * sizeof(T));  // [coef, L0, L1]
    const int tid = threadIdx.x;
    const int ntiles = P.ncx * P.ncy;
    const int tile = blockIdx.x % ntiles;
    const int grp = blockIdx.x / ntiles;
    const int64_t l0 = (int64_t)grp * P.levels_per_cta;
    const int64_t l1 = l0 + P.levels_per_cta < P.nb ? l0 + P.levels_per_cta : P.nb;
    if (l0 >= l1) return;
    FusedTile<T> tl(P, tile, smem);
    FusedThread<T> st;
    constexpr unsigned ROW_BYTES = G::TW * sizeof(T);
    if (tid == 0) {
        mbar_init(&mb[0], G::TH);
        mbar_init(&mb[1], G::TH);
        mbar_init(&mb[2], G::TH);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid < G::TH) {
        mbar_expect_tx(&mb[0], 3 * ROW_BYTES);
        tl.issue_coef_row(tid, &mb[0]);
        mbar_expect_tx(&mb[1], ROW_BYTES);
        tl.issue_t1_row(tid, l0, 0, &mb[1]);
    }
    int it = 0;
    for (int64_t l = l0; l < l1; ++l, ++it) {
        const int buf = it & 1;
        if (l + 1 < l1 && tid < G::TH) {  // prefetch the next level (tile P[buf^1] was released by the last barrier)
            fence_proxy_async();
            mbar_expect_tx(&mb[1 + (buf ^ 1)], ROW_BYTES);
            tl.issue_t1_row(tid, l + 1, buf ^ 1, &mb[1 + (buf ^ 1)]);
        }
        tl.load_regs(tid, l, st);
        if (it == 0) mbar_wait(&mb[0], 0);
        mbar_wait(&mb[1 + buf], (unsigned)((it >> 1) & 1));
        T* Pb = tl.tileP(buf);
        T* Q = tl.tileQ();
        tl.template step<true>(tid, 1, Pb, Q, st);
        __syncthreads();
#pragma unroll 1
        for (int s = 2; s <= P.k; ++s) {
            if (s & 1) tl.template step<false>(tid, s, Pb, Q, st);
            else tl.template step<false>(tid, s, Q, Pb, st);
            __syncthreads();
        }
        // k odd: newest T in Q, previous in P[buf]; k even: newest in P[buf], previous in Q
        if (P.k & 1) tl.store(tid, l, Q, Pb, st);
        else tl.store(tid, l, Pb, Q, st);
        __syncthreads();
    }
}
#endif

}  // namespace gcmf

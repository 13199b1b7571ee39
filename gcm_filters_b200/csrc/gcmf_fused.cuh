// gcmf_fused.cuh -- temporally blocked Chebyshev steps: k <= 4 recurrence steps per HBM round trip.
//
// One CTA owns one spatial tile (core CH x CW plus a 4-cell halo) and loops over a slab of batch
// slices ("levels").  Every thread owns R x VX fixed points of the tile for the CTA's whole life.
// Per level:
//   * the T_{i-1} and T_{i-2} tiles (+halo) are staged in shared memory by the TMA engine: one bulk
//     async copy (cp.async.bulk, SASS UBLKCP) per tile row and array, split in two where the row
//     crosses the periodic x boundary, completing on an mbarrier.  As soon as a level has been pulled
//     out of the landing tiles the copies for the next level are issued, so they fly during the k steps;
//   * each thread lifts the RAW values of its own points into registers (t1, t2) -- the recurrence
//     "-x" and "- T_{i-2}" terms are point-wise -- and publishes only the SANITIZED value
//     (nan_to_num, times the wet mask for REGULAR5) of T_{i-1} in a work tile: that is all a neighbour
//     ever needs, so nan_to_num runs once per produced value instead of once per read;
//   * k steps ping-pong between two work tiles on a region that shrinks by one cell per step
//     (overlapped / ghost-zone tiling); the thread's own rows and the rows above / below and the W/E
//     columns are read from the work tile;
//   * the running filtered field `bar` stays in registers for the k steps; T_{i+k-1}, T_{i+k-2} and
//     bar are written back from registers;
//   * FLUX: the coefficient tiles (ce, cn, ra) are staged once per CTA and reused for every level; warps
//     synchronise with their neighbour warps only, through per-warp mbarriers (no CTA barrier per step);
//     REGULAR5: the wet mask and wet_fac of the own points are packed into two registers once per CTA.
// HBM traffic per grid-point step: (2*rho + 4) * w / k bytes instead of 5 * w  (rho = tile/core area).
//
// The arithmetic of a point is the same inline code as the one-step kernels (-fmad=false) and `bar`
// is accumulated in the same order, so fused and un-fused results are bit-identical.
//
// Every phase below is a __host__ __device__ function of (thread id, per-thread state): the device
// kernel separates the phases with __syncthreads()/mbarrier waits, the test-only host emulator runs
// each phase for all thread ids in turn.
#pragma once
#include "gcmf_stencils.cuh"

namespace gcmf {

enum : int { FK_FLUX = 0, FK_REG5 = 1 };

constexpr int FUSED_H = 4;  // halo width = max fused steps

// Compile-time knobs of the fused kernels.  Every one was A/B-measured on a B200 with tests/tools/build_variant.py +
// variant_bench.py (cfg3 shape, 62 x 2400 x 3600 fp64, 44 steps, ms per filter call; profiles/variants_r02.md).  What the
// round-2 A/B kept is unconditional code now: the landing tiles are re-armed by warp 0 (an edge warp with a third of
// the inner warps' work: -4.4 %), the periodic index arithmetic of the loaders wraps once instead of dividing (-2.8 %),
// FLUX hoists the 64-bit address arithmetic of load_bar / store, REGULAR5 resolves the mask flag per step (-2 %).
// What lost is gone: sanitized state in registers (SANSTATE: +60 %, spills), per-row NaN branches (ROWNAN: +13 %), a
// land-select publish with a running finiteness check instead of nan_to_num (FASTSAN: +7 % with tensor maps, +113 %
// without; 24-42 bytes of spills), skipping the last publish, the warp-vote nan_to_num shortcut, bar L2 prefetch,
// polling waits, a suspend-time hint on the parking waits (2 us / 20 us: +-0.1 %), contracting bar += p*T.
// The FLUX kernel sits at the 128-register cap of a 512-thread CTA: anything that costs registers spills.
#ifndef GCMF_OPT_TMAP
// Interior tiles (no periodic wrap, no fold row) are staged by ONE cp.async.bulk.tensor (SASS UTMALDG) per array from a
// cuTensorMapEncodeTiled descriptor instead of one cp.async.bulk per tile row issued by 32 lanes.
#define GCMF_OPT_TMAP 1
#endif

// XS: how the tile row is split over threads.  1: one 16-byte vector per thread and row (512 threads, used by
// the register-heavy FLUX kernel); 2: half a vector (1024 threads: the light REGULAR5 steps hide their latency
// better with twice the warps -- measured +8 % on cfg2, while FLUX loses 30 % with 64 registers per thread).
template <int KIND> struct FusedSplit { static constexpr int value = KIND == 1 /* FK_REG5 */ ? 2 : 1; };

template <typename T, int XS = 1> struct FusedGeom {
    static constexpr int AV = 16 / (int)sizeof(T);  // elements per 16 bytes: alignment unit of the bulk copies
    static constexpr int VX = AV / XS;              // consecutive columns per thread
    static constexpr int H = FUSED_H;
    static constexpr int NTX = 64 * XS;             // threads along x
    static constexpr int TW = NTX * VX;             // tile width  incl. halo: 128 (f64) / 256 (f32)
#ifndef GCMF_FUSED_R
#define GCMF_FUSED_R 4
#define GCMF_FUSED_NTY 8
#endif
    static constexpr int R = GCMF_FUSED_R;          // consecutive rows per thread
    static constexpr int NTY = GCMF_FUSED_NTY;      // threads along y
    static constexpr int TH = R * NTY;              // tile height incl. halo: 32
    static constexpr int CW = TW - 2 * H;           // core width  120 / 248
    static constexpr int CH = TH - 2 * H;           // core height 24
    static constexpr int NTHREADS = NTX * NTY;      // 512
    static constexpr int PLANE = TH * TW;           // elements per shared-memory tile (32 KiB)
    // tiles: X, Y (TMA landing of T1, T2), S0, S1 (sanitized ping-pong) [+ ce, cn, ra for FLUX]
    static GCMF_HD constexpr int ntiles(int kind) { return kind == FK_FLUX ? 7 : 4; }
    static GCMF_HD constexpr size_t smem_bytes(int kind) { return (size_t)ntiles(kind) * PLANE * sizeof(T) + 1024; }
};

template <typename T> struct FusedParams {
    Geo g;
    PlaneRef plane[3];          // FLUX: ce, cn, ra.  REG5: plane[0] = uint8 wet mask (if FL_MASK)
    FieldRef<const T> t1_in;    // T_{i-1}
    FieldRef<const T> t2_in;    // T_{i-2}
    FieldRef<T> t1_out;         // T_{i+k-1}
    FieldRef<T> t2_out;         // T_{i+k-2}
    FieldRef<T> bar;            // bar += sum_s p[s] T_{i+s}
    double c;
    double p[FUSED_H];          // Chebyshev coefficients of the k steps
    int32_t k;                  // fused steps, 1..H
    int32_t first;              // block starts at recurrence step 1: t1_in = prepared field x, no T_{i-2}, no bar yet
    int32_t last;               // block ends at step n_steps: bar is finalized (/area) and no T is stored
    double p0;                  // p[0] (first block)
    int32_t ncx, ncy;           // core tiles along x / y
    int64_t nb;                 // batch slices
    int32_t levels_per_cta;
};

// Tensor-map descriptors (CUtensorMap, 128 opaque bytes, encoded on the host with cuTensorMapEncodeTiled) of the
// arrays a launch stages through the TMA engine: boxes of one whole tile (TW x TH x 1 level).  A second
// __grid_constant__ kernel parameter: cp.async.bulk.tensor takes the descriptor's address in parameter space.
enum : int { TMAP_STATE = 1, TMAP_COEF = 2 };
struct FusedMaps {
    TmaDesc t1, t2;    // (nx, ny, nb) views of t1_in / t2_in
    TmaDesc coef[3];   // (nx, ny) views of ce, cn, ra (FLUX)
    int32_t use;       // TMAP_STATE | TMAP_COEF: which descriptors are valid (0: row-wise bulk copies everywhere)
};

template <typename T, int XS> struct FusedThread {  // per-thread registers that live across phases
    T t1[FusedGeom<T, XS>::R][FusedGeom<T, XS>::VX];   // raw T_{i-1} of the own points
    T t2[FusedGeom<T, XS>::R][FusedGeom<T, XS>::VX];   // raw T_{i-2}
    T acc[FusedGeom<T, XS>::R][FusedGeom<T, XS>::VX];  // running bar (owned points)
    uint32_t mbits;                             // REG5: wet bit of own point (q*VX+v)
    uint64_t wfbits;                            // REG5: wet_fac (0..4) of own point, 4 bits each
};

// Periodic index arithmetic of the tile loaders: a fused plan's grid is at least one tile wide and high, so every
// index handed to wrap_index lies in (-n, 2n) and one conditional add / subtract is the modulo (a `%` costs ~35
// instructions, four to six of them per lane and level on the path of the warp that re-arms the landing tiles).
GCMF_HD int wrap_index(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

// ---- bulk async copy global -> shared, completing on an mbarrier (device) / memcpy (host emulator) ----
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* mb, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mb)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mb, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mb)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mb) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mb)) : "memory");
}
// parking wait: the hardware suspends the warp until the phase completes or a time limit expires
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
// non-blocking probe of a phase
__device__ __forceinline__ unsigned mbar_test(uint64_t* mb, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(mb)), "r"(parity) : "memory");
    return ok;
}
// ---- drain counter of the landing tiles (acquire-release at CTA scope) ----
__device__ __forceinline__ uint32_t atom_add_acqrel_u32(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t ld_acquire_cta_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one whole tile (box TW x TH [x 1 level]) through a tensor map: SASS UTMALDG
__device__ __forceinline__ void tma_load_tile_3d(void* dst_smem, const TmaDesc* map, int x, int y, int z, uint64_t* mb) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(mb))
        : "memory");
}
__device__ __forceinline__ void tma_load_tile_2d(void* dst_smem, const TmaDesc* map, int x, int y, uint64_t* mb) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(mb))
        : "memory");
}
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

// EDGE: which ends of the recurrence the block touches, known at compile time so that the common mid block
// carries none of it: bit 0 = the block starts at step 1 (P.first), bit 1 = it ends at step n_steps (P.last).
template <typename T, int KIND, int EDGE> struct FusedTile {
    static constexpr int XS = FusedSplit<KIND>::value;
    using G = FusedGeom<T, XS>;
    using Thread = FusedThread<T, XS>;
    const FusedParams<T>& P;
    const FusedMaps* M;  // tensor maps of this launch (device: address in parameter space); nullptr in the host emulator
    int gy0, gx0;      // global coordinates of tile element (0,0) (may be negative: wraps)
    int cy0, cx0;      // global coordinates of the first core element
    T* smem;
    bool masked;       // REG5 with a wet mask (nan_to_num + mask), else raw values (NaNs spread)
    bool interior;     // the whole tile (halo included) lies inside a periodic / tripolar grid: no wrap, no fold row
    static GCMF_HD constexpr bool is_first() { return (EDGE & 1) != 0; }
    static GCMF_HD constexpr bool is_last() { return (EDGE & 2) != 0; }

    GCMF_HD FusedTile(const FusedParams<T>& P_, int tile, T* smem_, const FusedMaps* M_ = nullptr)
        : P(P_), M(M_), smem(smem_) {
        const int cx = tile % P.ncx, cy = tile / P.ncx;
        cy0 = cy * G::CH;
        cx0 = cx * G::CW;
        gy0 = cy0 - G::H;
        gx0 = cx0 - G::H;
        masked = (P.g.flags & FL_MASK) != 0;
        interior = (P.g.flags & FL_WRAP_Y) && gx0 >= 0 && gx0 + G::TW <= P.g.nx && gy0 >= 0 && gy0 + G::TH <= P.g.ny;
    }
    GCMF_HD T* tileX() const { return smem; }
    GCMF_HD T* tileY() const { return smem + (size_t)G::PLANE; }
    GCMF_HD T* tileS(int which) const { return smem + (size_t)(2 + which) * G::PLANE; }
    GCMF_HD T* tileC(int which) const { return smem + (size_t)(4 + which) * G::PLANE; }
    // does this tile take the tensor-map path for its state / coefficient tiles?
    GCMF_HD bool tmap(int what) const {
#if defined(__CUDA_ARCH__) && GCMF_OPT_TMAP
        return interior && (M->use & what) != 0;
#else
        (void)what;
        return false;
#endif
    }

    // value published to the neighbours: what the reference's Laplacian differences
    GCMF_HD T sanitize(T x, bool wet) const { return sanitize(x, wet, masked); }
    GCMF_HD T sanitize(T x, bool wet, bool msk) const {
        if (KIND == FK_FLUX) return nan2num(x);                    // kernels.py:300, 566
        if (msk) return wet ? nan2num(x) : T(0);                   // kernels.py:175-176
        return x;                                                  // kernels.py:113-121: NaNs spread
    }

    // ---- tripolar grids (FL_FOLD_N | FL_CUT_S; kernels.py:33-40) -------------------------------------------
    // Tile rows above the top of the grid are VIRTUAL: row ny-1+r is the mirror image of row ny-r with the
    // columns reversed, (ny-1+r, i) == (ny-r, nx-1-i).  A virtual cell is the real cell seen upside down: its
    // east face is the real cell's west face and its north face the real cell's south face, so the halo above
    // the fold evolves exactly like the real cells it mirrors.  Virtual rows cannot be bulk-copied (reversed):
    // the issuing thread gathers them element by element (only the top row of tiles pays for this).
    GCMF_HD bool fold() const { return (P.g.flags & FL_FOLD_N) != 0; }
    GCMF_HD bool virtual_row(int r) const { return fold() && gy0 + r >= P.g.ny; }
    GCMF_HD bool below_cut(int r) const { return (P.g.flags & FL_CUT_S) && gy0 + r < 0; }
    GCMF_HD int image_row(int r) const { return 2 * P.g.ny - 1 - (gy0 + r); }          // real row of a virtual tile row
    GCMF_HD int image_col(int c) const { return P.g.nx - 1 - wrap_index(gx0 + c, P.g.nx); }  // real column

    // Global row of tile row r.  Periodic grids wrap; a latitude band (FL_WRAP_Y clear) has FUSED_H ghost rows
    // physically present below row 0 and above row ny-1 (rows further out lie outside the dependency cone of
    // every owned row and are clamped onto the last ghost row).
    GCMF_HD int source_row(int r) const {
        const int g = gy0 + r;
        if (P.g.flags & FL_WRAP_Y) return wrap_index(g, P.g.ny);
        return g < -G::H ? -G::H : (g > P.g.ny + G::H - 1 ? P.g.ny + G::H - 1 : g);
    }

    // bytes of row r that arrive through bulk copies (the mbarrier's expected transaction count)
    GCMF_HD unsigned row_tx_bytes(int r) const { return virtual_row(r) ? 0u : (unsigned)(G::TW * sizeof(T)); }

    // Stage tile row r of a (2-D slice of a) global array: <= 2 bulk copies (split at the x wrap); the virtual rows
    // of a tile next to the fold are gathered by gather_virtual().
    GCMF_HD void copy_row(T* dst_tile, const T* src_slice, int64_t pitch, int r, uint64_t* mb) const {
        if (virtual_row(r)) return;
        const int gy = source_row(r);
        const int gx = wrap_index(gx0, P.g.nx);
        const T* row = src_slice + (int64_t)gy * pitch;
        const int n1 = (P.g.nx - gx) < G::TW ? (P.g.nx - gx) : G::TW;
        bulk_copy_g2s(dst_tile + r * G::TW, row + gx, (unsigned)(n1 * sizeof(T)), mb);
        if (n1 < G::TW) bulk_copy_g2s(dst_tile + r * G::TW + n1, row, (unsigned)((G::TW - n1) * sizeof(T)), mb);
    }
    // Virtual rows of one tile, gathered by the TH issuing lanes together (lane strides over the columns).
    // dj / di: index shift applied to the image cell (coefficient faces).
    GCMF_HD void gather_virtual(int lane, T* dst_tile, const T* src_slice, int64_t pitch, int dj = 0, int di = 0) const {
        if (!fold()) return;
        int r0 = P.g.ny - gy0;  // first virtual tile row
        if (r0 < 0) r0 = 0;
        for (int r = r0; r < G::TH; ++r) {
            const T* row = src_slice + (int64_t)(image_row(r) + dj) * pitch;
            for (int c = lane; c < G::TW; c += G::TH) dst_tile[r * G::TW + c] = row[wrap_index(image_col(c) + di, P.g.nx)];
        }
    }
    // phase: lane r < TH stages the coefficient rows (FLUX) and arrives on the barrier with the bytes it expects;
    // an interior tile is three tensor copies issued by lane 0 instead.
    GCMF_HD unsigned coef_tx_bytes(int r) const {
        return virtual_row(r) ? 0u : (below_cut(r) ? 2u : 3u) * (unsigned)(G::TW * sizeof(T));
    }
    GCMF_HD void issue_coef(int r, uint64_t* mb) const {
#if defined(__CUDA_ARCH__) && GCMF_OPT_TMAP
        if (tmap(TMAP_COEF)) {
            if (r == 0) {
                mbar_expect_tx(mb, 3u * (unsigned)(G::PLANE * sizeof(T)));
                for (int s = 0; s < 3; ++s) tma_load_tile_2d(tileC(s), &M->coef[s], gx0, gy0, mb);
            } else {
                mbar_arrive(mb);
            }
            return;
        }
#endif
        const T* ce = reinterpret_cast<const T*>(P.plane[0].p);
        const T* cn = reinterpret_cast<const T*>(P.plane[1].p);
        const T* ra = reinterpret_cast<const T*>(P.plane[2].p);
        // virtual cell: east face = image's west face (di = -1), north face = image's south face (dj = -1)
        gather_virtual(r, tileC(0), ce, P.plane[0].pitch, 0, -1);
        gather_virtual(r, tileC(1), cn, P.plane[1].pitch, -1, 0);
        gather_virtual(r, tileC(2), ra, P.plane[2].pitch);
        copy_row(tileC(0), ce, P.plane[0].pitch, r, mb);
        if (below_cut(r)) {  // no flux across the southern edge of row 0 (and nothing below it matters)
            for (int c = 0; c < G::TW; ++c) tileC(1)[r * G::TW + c] = T(0);
        } else {
            copy_row(tileC(1), cn, P.plane[1].pitch, r, mb);
        }
        copy_row(tileC(2), ra, P.plane[2].pitch, r, mb);
#ifdef __CUDA_ARCH__
        mbar_expect_tx(mb, coef_tx_bytes(r));  // generic writes (virtual / cut rows) precede this releasing arrive
#endif
    }
    // phase: lane r < TH issues row r of T1(level) -> X and T2(level) -> Y (interior tile: lane 0 issues two boxes)
    GCMF_HD unsigned state_tx_bytes(int r) const { return (is_first() ? 1u : 2u) * row_tx_bytes(r); }
    GCMF_HD void issue_state(int r, int64_t level, uint64_t* mb) const {
#if defined(__CUDA_ARCH__) && GCMF_OPT_TMAP
        if (tmap(TMAP_STATE)) {
            if (r == 0) {
                mbar_expect_tx(mb, (is_first() ? 1u : 2u) * (unsigned)(G::PLANE * sizeof(T)));
                tma_load_tile_3d(tileX(), &M->t1, gx0, gy0, (int)level, mb);
                if (!is_first()) tma_load_tile_3d(tileY(), &M->t2, gx0, gy0, (int)level, mb);
            } else {
                mbar_arrive(mb);
            }
            return;
        }
#endif
        gather_virtual(r, tileX(), P.t1_in.p + level * P.t1_in.bstride, P.t1_in.pitch);
        if (!is_first()) gather_virtual(r, tileY(), P.t2_in.p + level * P.t2_in.bstride, P.t2_in.pitch);
        copy_row(tileX(), P.t1_in.p + level * P.t1_in.bstride, P.t1_in.pitch, r, mb);
        if (!is_first()) copy_row(tileY(), P.t2_in.p + level * P.t2_in.bstride, P.t2_in.pitch, r, mb);
#ifdef __CUDA_ARCH__
        mbar_expect_tx(mb, state_tx_bytes(r));
#endif
    }

    GCMF_HD bool owns_cols(int tx) const {
        const int lc0 = tx * G::VX;
        return lc0 >= G::H && lc0 < G::TW - G::H && (cx0 + lc0 - G::H) < P.g.nx;
    }
    GCMF_HD bool owns_row(int lr) const { return lr >= G::H && lr < G::TH - G::H && (cy0 + lr - G::H) < P.g.ny; }

    // phase (once per CTA, REG5 masked): wet bit and wet_fac of the own points from the global uint8 mask
    GCMF_HD void load_mask(int tid, Thread& st) const {
        st.mbits = 0xffffffffu;
        st.wfbits = 0x4444444444444444ull;  // unmasked: every neighbour is wet
        if (KIND != FK_REG5 || !masked) return;
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const uint8_t* m = reinterpret_cast<const uint8_t*>(P.plane[0].p);
        const int64_t pitch = P.plane[0].pitch;
        uint32_t mb = 0;
        uint64_t wf = 0;
        const int ny = P.g.ny, nx = P.g.nx;
        for (int q = 0; q < G::R; ++q) {
            const int r = ty * G::R + q;
            const bool virt = virtual_row(r);
            const int gy = virt ? image_row(r) : source_row(r);
            for (int v = 0; v < G::VX; ++v) {
                const int c = tx * G::VX + v;
                const int gx = virt ? image_col(c) : wrap_index(gx0 + c, nx);
                const int gxe = gx + 1 == nx ? 0 : gx + 1, gxw = gx == 0 ? nx - 1 : gx - 1;
                const int idx = q * G::VX + v;
                if (m[(int64_t)gy * pitch + gx]) mb |= 1u << idx;
                // north neighbour of the top row is its mirror image across the fold (kernels.py:461-467)
                const bool top = fold() && gy == ny - 1;
                const bool wrap_y = (P.g.flags & FL_WRAP_Y) != 0;
                // A latitude band (no wrap) has H ghost rows in memory on either side; the outermost one has no
                // neighbour beyond it, and needs none: tile rows 0 and TH-1 are never computed (ASAN find).
                const int gyn = top ? ny - 1 : (wrap_y ? (gy + 1 == ny ? 0 : gy + 1)
                                                       : (gy + 1 > ny + G::H - 1 ? ny + G::H - 1 : gy + 1));
                const int gxn = top ? nx - 1 - gx : gx;
                // row 0 of a tripolar grid is land: the wrap below it is inert
                const int gys = wrap_y ? (gy == 0 ? ny - 1 : gy - 1) : (gy - 1 < -G::H ? -G::H : gy - 1);
                const uint64_t cnt = (m[(int64_t)gy * pitch + gxe] != 0) + (m[(int64_t)gy * pitch + gxw] != 0) +
                                     (m[(int64_t)gyn * pitch + gxn] != 0) + (m[(int64_t)gys * pitch + gx] != 0);
                wf |= cnt << (4 * idx);
            }
        }
        st.mbits = mb;
        st.wfbits = wf;
    }

    // FLUX hoists the 64-bit address arithmetic of load_bar / store: one full "level * bstride + first_row * pitch +
    // column" per array and level, then + q * pitch with the unrolled q (A/B: FLUX gains with the hoisting in place,
    // the REGULAR5 kernel loses 3 % with it and keeps the per-row form).
    static constexpr bool ROWPTR = KIND == FK_FLUX;

    // phase: bar of the owned points from HBM into registers
    GCMF_HD void load_bar(int tid, int64_t level, Thread& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const int lc0 = tx * G::VX;
        const bool oc = owns_cols(tx);
        const int64_t ob = level * P.bar.bstride + (int64_t)(cy0 + ty * G::R - G::H) * P.bar.pitch +
                           (cx0 + lc0 - G::H);  // element offset of the thread's first row
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int lr = ty * G::R + q;
            if (oc && owns_row(lr) && !is_first()) {
                if (ROWPTR)
                    Ld<T, G::VX>::go(P.bar.p + (ob + (int64_t)q * P.bar.pitch), st.acc[q]);
                else
                    Ld<T, G::VX>::go(P.bar.p + level * P.bar.bstride + (int64_t)(cy0 + lr - G::H) * P.bar.pitch +
                                         (cx0 + lc0 - G::H), st.acc[q]);
            } else {
#pragma unroll
                for (int v = 0; v < G::VX; ++v) st.acc[q][v] = T(0);
            }
        }
    }

    // phase: lift the raw own points out of the landing tiles, publish the sanitized T1 in S0
    GCMF_HD void extract(int tid, Thread& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const int lc0 = tx * G::VX;
        const T* X = tileX();
        const T* Y = tileY();
        T* S0 = tileS(0);
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int off = (ty * G::R + q) * G::TW + lc0;
            Ld<T, G::VX>::go(X + off, st.t1[q]);
            if (!is_first()) {
                Ld<T, G::VX>::go(Y + off, st.t2[q]);
            } else {
#pragma unroll
                for (int v = 0; v < G::VX; ++v) st.t2[q][v] = T(0);
            }
            T o[G::VX];
#pragma unroll
            for (int v = 0; v < G::VX; ++v) o[v] = sanitize(st.t1[q][v], (st.mbits >> (q * G::VX + v)) & 1u);
            St<T, G::VX>::go(S0 + off, o);
        }
    }

    // phase: recurrence step s (1-based) on the region [s, TH-s) x [s, TW-s); S = source work tile, D = destination.
    // X1 holds T_{i-1} (raw), X2 holds T_{i-2}; the new T_i is written over X2, so consecutive steps just swap
    // the roles of the two register arrays (no moves).  ALLROWS: the thread's R rows all lie in the region for
    // every s <= H (its rows are core rows), which removes every branch from the row loop.
    // MSK (REGULAR5): "has a wet mask", resolved per step instead of per point (straight-line row loop).
    template <bool ALLROWS, bool MSK>
    GCMF_HD void step_rows(int tid, int s, const T* __restrict__ S, T* __restrict__ D, T (&X1)[G::R][G::VX],
                           T (&X2)[G::R][G::VX], Thread& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const int lc0 = tx * G::VX;
        const int lr0 = ty * G::R;
        const int off0 = lr0 * G::TW + lc0;
        const T* Sc = S + off0;
        const T* Sw = Sc - (lc0 > 0 ? 1 : 0);                          // west neighbour of column lc0 (clamped)
        const T* Se = Sc + (lc0 + G::VX < G::TW ? G::VX : G::VX - 1);  // east neighbour of column lc0+VX-1 (clamped)
        const T c = (T)P.c;
        const double pk = P.p[s - 1];
        const bool start = is_first() && s == 1;  // recurrence step 1: T_1 = A(x), bar = p0 x + p1 T_1
        T o[G::R][G::VX], os[G::VX], on[G::VX];
#pragma unroll
        for (int q = 0; q < G::R; ++q) Ld<T, G::VX>::go(Sc + q * G::TW, o[q]);  // = sanitize(X1), published by this thread
        if (ALLROWS || lr0 > 0) Ld<T, G::VX>::go(Sc - G::TW, os);
        if (ALLROWS || lr0 + G::R < G::TH) Ld<T, G::VX>::go(Sc + G::R * G::TW, on);
        T cn_prev[G::VX];
        bool have_prev = false;
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int lr = lr0 + q;
            if (!ALLROWS && (lr < s || lr >= G::TH - s)) {
                have_prev = false;
                continue;
            }
            const T ow = Sw[q * G::TW];
            const T oe = Se[q * G::TW];
            T t0[G::VX], pub[G::VX];
            if (KIND == FK_FLUX) {
                const T* CE = tileC(0) + off0 + q * G::TW;
                const T* CN = tileC(1) + off0 + q * G::TW;
                const T* RA = tileC(2) + off0 + q * G::TW;
                T ce[G::VX], cn[G::VX], cs[G::VX], ra[G::VX];
                Ld<T, G::VX>::go(CE, ce);
                const T cew = *(CE - (lc0 > 0 ? 1 : 0));
                Ld<T, G::VX>::go(CN, cn);
                if ((ALLROWS && q > 0) || have_prev) {
#pragma unroll
                    for (int v = 0; v < G::VX; ++v) cs[v] = cn_prev[v];
                } else {
                    Ld<T, G::VX>::go(CN - G::TW, cs);
                }
                Ld<T, G::VX>::go(RA, ra);
#pragma unroll
                for (int v = 0; v < G::VX; ++v) {
                    const T o_e = v == G::VX - 1 ? oe : o[q][v + 1 < G::VX ? v + 1 : v];
                    const T o_w = v == 0 ? ow : o[q][v > 0 ? v - 1 : 0];
                    const T o_n = q == G::R - 1 ? on[v] : o[q + 1 < G::R ? q + 1 : q][v];
                    const T o_s = q == 0 ? os[v] : o[q > 0 ? q - 1 : 0][v];
                    const T lap = flux_lap<T>(o[q][v], o_w, o_e, o_n, o_s, ce[v], v == 0 ? cew : ce[v > 0 ? v - 1 : 0],
                                              cn[v], cs[v], ra[v]);
                    const T a = shifted_flux<T>(X1[q][v], c, lap);  // filter.py:171
                    t0[v] = start ? a : cheb_next<T>(a, X2[q][v]);  // filter.py:192-194 / 197-203
                    cn_prev[v] = cn[v];
                }
                have_prev = true;
            } else {
#pragma unroll
                for (int v = 0; v < G::VX; ++v) {
                    const T o_e = v == G::VX - 1 ? oe : o[q][v + 1 < G::VX ? v + 1 : v];
                    const T o_w = v == 0 ? ow : o[q][v > 0 ? v - 1 : 0];
                    const T o_n = q == G::R - 1 ? on[v] : o[q + 1 < G::R ? q + 1 : q][v];
                    const T o_s = q == 0 ? os[v] : o[q > 0 ? q - 1 : 0][v];
                    const int idx = q * G::VX + v;
                    T lap;
                    if (MSK) {  // kernels.py:178-186
                        const T wf = (T)(int)((st.wfbits >> (4 * idx)) & 0xfull);
                        const T r = (((-wf * o[q][v] + o_e) + o_w) + o_n) + o_s;
                        lap = ((st.mbits >> idx) & 1u) ? r : T(0);
                    } else {       // kernels.py:115-121
                        lap = (((T(-4) * o[q][v] + o_e) + o_w) + o_n) + o_s;
                    }
                    const T a = -X1[q][v] - c * lap;
                    t0[v] = start ? a : cheb_next<T>(a, X2[q][v]);
                }
            }
#pragma unroll
            for (int v = 0; v < G::VX; ++v) {
                const double b0 = start ? P.p0 * (double)X1[q][v] : (double)st.acc[q][v];
                st.acc[q][v] = (T)bar_update(b0, pk, (double)t0[v]);  // filter.py:195 / 204
                X2[q][v] = t0[v];                                     // T_i replaces T_{i-2}
            }
#pragma unroll
            for (int v = 0; v < G::VX; ++v) pub[v] = sanitize(t0[v], (st.mbits >> (q * G::VX + v)) & 1u, MSK);
            St<T, G::VX>::go(D + off0 + q * G::TW, pub);
        }
    }

    template <bool MSK> GCMF_HD void step_msk(int tid, int s, bool inner, Thread& st) const {
        // odd steps read T_{i-1} from st.t1 and overwrite st.t2; even steps the other way round
        if (s & 1) {
            if (inner) step_rows<true, MSK>(tid, s, tileS(0), tileS(1), st.t1, st.t2, st);
            else step_rows<false, MSK>(tid, s, tileS(0), tileS(1), st.t1, st.t2, st);
        } else {
            if (inner) step_rows<true, MSK>(tid, s, tileS(1), tileS(0), st.t2, st.t1, st);
            else step_rows<false, MSK>(tid, s, tileS(1), tileS(0), st.t2, st.t1, st);
        }
    }

    GCMF_HD void step(int tid, int s, Thread& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        const int lc0 = tx * G::VX;
        if (lc0 + G::VX <= s || lc0 >= G::TW - s) return;  // column group outside the region
        const bool inner = ty * G::R >= G::H && (ty + 1) * G::R <= G::TH - G::H;  // rows inside for every s <= H
        if (KIND == FK_REG5 && masked) step_msk<true>(tid, s, inner, st);
        else step_msk<false>(tid, s, inner, st);
    }

    // phase: write the owned core points of T_{i+k-1}, T_{i+k-2} and bar back to HBM (from registers)
    GCMF_HD void store(int tid, int64_t level, const Thread& st) const {
        const int tx = tid % G::NTX, ty = tid / G::NTX;
        if (!owns_cols(tx)) return;
        const int lc0 = tx * G::VX;
        const int gx = cx0 + lc0 - G::H;
        const int gyf = cy0 + ty * G::R - G::H;  // first row of the thread
        // element offsets of the thread's first row in the output arrays (ROWPTR form)
        const int64_t ob0 = level * P.bar.bstride + (int64_t)gyf * P.bar.pitch + gx;
        const int64_t o10 = is_last() ? 0 : level * P.t1_out.bstride + (int64_t)gyf * P.t1_out.pitch + gx;
        const int64_t o20 = is_last() ? 0 : level * P.t2_out.bstride + (int64_t)gyf * P.t2_out.pitch + gx;
#pragma unroll
        for (int q = 0; q < G::R; ++q) {
            const int lr = ty * G::R + q;
            if (!owns_row(lr)) continue;
            const int gy = cy0 + lr - G::H;
            T outv[G::VX];
#pragma unroll
            for (int v = 0; v < G::VX; ++v) outv[v] = st.acc[q][v];
            if (is_last()) {
                if (P.g.flags & FL_AREA) {  // finalize: divide by the cell area (kernels.py:103-104)
                    T ar[G::VX];
                    Ld<T, G::VX>::go(reinterpret_cast<const T*>(P.plane[1].p) + (int64_t)gy * P.plane[1].pitch + gx, ar);
#pragma unroll
                    for (int v = 0; v < G::VX; ++v) outv[v] = outv[v] / ar[v];
                }
            } else {
                // after an odd number of steps the newest T sits in st.t2 (see step_msk()).  Value selects: a select
                // between the two register ARRAYS can push the whole per-thread state into local memory (ptxas -v
                // once showed a 128-byte stack frame for fused_kernel<double, REG5>).
                T n1[G::VX], n2[G::VX];
#pragma unroll
                for (int v = 0; v < G::VX; ++v) {
                    n1[v] = (P.k & 1) ? st.t2[q][v] : st.t1[q][v];
                    n2[v] = (P.k & 1) ? st.t1[q][v] : st.t2[q][v];
                }
                if (ROWPTR) {
                    St<T, G::VX>::go(P.t1_out.p + (o10 + (int64_t)q * P.t1_out.pitch), n1);
                    St<T, G::VX>::go(P.t2_out.p + (o20 + (int64_t)q * P.t2_out.pitch), n2);
                } else {
                    St<T, G::VX>::go(P.t1_out.p + level * P.t1_out.bstride + (int64_t)gy * P.t1_out.pitch + gx, n1);
                    St<T, G::VX>::go(P.t2_out.p + level * P.t2_out.bstride + (int64_t)gy * P.t2_out.pitch + gx, n2);
                }
            }
            if (ROWPTR) St<T, G::VX>::go(P.bar.p + (ob0 + (int64_t)q * P.bar.pitch), outv);
            else St<T, G::VX>::go(P.bar.p + level * P.bar.bstride + (int64_t)gy * P.bar.pitch + gx, outv);
        }
    }
};

#ifdef __CUDACC__
template <typename T, int KIND, int EDGE>
__global__ void __launch_bounds__(FusedGeom<T, FusedSplit<KIND>::value>::NTHREADS, 1)
    fused_kernel(const __grid_constant__ FusedParams<T> P, const __grid_constant__ FusedMaps M) {
    using G = FusedGeom<T, FusedSplit<KIND>::value>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    uint64_t* mb = reinterpret_cast<uint64_t*>(smem_raw + (size_t)G::ntiles(KIND) * G::PLANE * sizeof(T));  // [coef, XY]
    const int tid = threadIdx.x;
    const int ntiles = P.ncx * P.ncy;
    const int tile = blockIdx.x % ntiles;
    const int grp = blockIdx.x / ntiles;
    const int64_t l0 = (int64_t)grp * P.levels_per_cta;
    const int64_t l1 = l0 + P.levels_per_cta < P.nb ? l0 + P.levels_per_cta : P.nb;
    if (l0 >= l1) return;
    FusedTile<T, KIND, EDGE> tl(P, tile, smem, &M);
    typename FusedTile<T, KIND, EDGE>::Thread st;
    if (tid == 0) {
        mbar_init(&mb[0], G::TH);
        mbar_init(&mb[1], G::TH);
        fence_mbar_init();
    }
    if (tid < G::NTHREADS / 32) {  // per-warp progress barriers: expected arrivals = number of neighbour warps
        constexpr int WPR_ = G::NTX / 32;
        const int wy_ = tid / WPR_, wx_ = tid % WPR_;
        const unsigned cnt = (wx_ > 0) + (wx_ < WPR_ - 1) + (wy_ > 0) + (wy_ < G::NTY - 1);
        mbar_init(&mb[2 + 2 * tid], cnt);
        mbar_init(&mb[2 + 2 * tid + 1], cnt);
    }
    if (tid == 32) *reinterpret_cast<uint32_t*>(mb + 2 + 2 * 32) = 0u;  // drain counter
    if (tid < 32) fence_mbar_init();
    static_assert(G::TH <= 32 && G::NTHREADS <= 1024, "one refill lane per tile row");
    __syncthreads();
    if (tid < G::TH) {
        if (KIND == FK_FLUX) tl.issue_coef(tid, &mb[0]);
        tl.issue_state(tid, l0, &mb[1]);
    }
    tl.load_mask(tid, st);
    // FLUX steps are long enough for neighbour-only synchronisation to pay; the light REGULAR5 steps keep
    // the cheaper CTA barrier.
    constexpr bool FINE_SYNC = KIND == FK_FLUX;
    if (!FINE_SYNC) {
    int it = 0;
    for (int64_t l = l0; l < l1; ++l, ++it) {
        tl.load_bar(tid, l, st);
        mbar_wait(&mb[1], (unsigned)(it & 1));
        tl.extract(tid, st);
        __syncthreads();  // S0 complete; landing tiles consumed
        if (l + 1 < l1 && tid < G::TH) {  // next level's tiles fly during the k steps
            fence_proxy_async();
            tl.issue_state(tid, l + 1, &mb[1]);
        }
#pragma unroll 1
        for (int s = 1; s <= P.k; ++s) {
            tl.step(tid, s, st);
            __syncthreads();
        }
        tl.store(tid, l, st);
    }
    } else {
    // Neighbour-only synchronisation instead of a CTA barrier per step.  A warp covers 32 threads x R rows of a
    // tile row group; in phase g (g = it*(k+1) + s, s = 0 for extract) it reads rows published
    // in phase g-1 by the warps above / below it and by the other half of its own rows, and overwrites
    // rows those same warps read in phase g-1.  Both hazards are covered by one rule: start phase g only
    // when these (up to four) neighbour warps have completed phase g-1.  Warps therefore drift apart by up to one phase per
    // hop, which spreads shared-memory and fp64 work in time instead of convoying at a barrier.
    // Progress is signalled through mbarriers rather than polled flags: a warp that has completed phase g
    // arrives (release) on barrier [g & 1] of each neighbour; a warp about to start phase g+1 waits (acquire, the
    // hardware parks it) on its own barrier [g & 1], whose expected arrival count is its
    // number of neighbours.  Two barriers per warp suffice because a neighbour can complete phase g+2 only after
    // this warp has completed g+1, i.e. after it has consumed phase g of the same barrier.
    uint64_t* nbar = mb + 2;                                          // [NWARPS][2]
    uint32_t* xcount = reinterpret_cast<uint32_t*>(mb + 2 + 2 * 32);  // warps that have drained the landing tiles
    constexpr int NWARPS = G::NTHREADS / 32;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int WPR = G::NTX / 32;  // warps per row group
    const int wy = warp / WPR, wx = warp % WPR;
    int nb_warp = -1;  // lanes 0..3 each serve one neighbour warp (west, east, south, north)
    if (lane == 0 && wx > 0) nb_warp = warp - 1;
    if (lane == 1 && wx < WPR - 1) nb_warp = warp + 1;
    if (lane == 2 && wy > 0) nb_warp = warp - WPR;
    if (lane == 3 && wy < G::NTY - 1) nb_warp = warp + WPR;
    auto wait_neighbours = [&](uint32_t need) {  // all neighbours have completed phase `need`
        if (need == 0) return;
        mbar_wait(&nbar[warp * 2 + (need & 1u)], ((need - 1u) >> 1) & 1u);
    };
    auto publish = [&](uint32_t done) {  // this warp has completed phase `done`
        __syncwarp();
        if (nb_warp >= 0) mbar_arrive(&nbar[nb_warp * 2 + (done & 1u)]);
    };
    int it = 0;
    for (int64_t l = l0; l < l1; ++l, ++it) {
        const uint32_t g0 = (uint32_t)it * (uint32_t)(P.k + 1);
        tl.load_bar(tid, l, st);
        if (it == 0) mbar_wait(&mb[0], 0);
        mbar_wait(&mb[1], (unsigned)(it & 1));
        // (Dropping this wait for even k -- extract only writes the thread's own points of S0, which no neighbour
        // reads after step k-1 -- is NOT safe: a neighbour could then arrive for phase g+2 on a barrier whose phase g
        // is still open, and the counting barrier cannot tell the two apart.  tests/tools/sync_model.py --latewait.)
        wait_neighbours(g0);
        tl.extract(tid, st);
        publish(g0 + 1);
        // The landing tiles are refilled for the next level once every warp has drained them (32 lanes = TH rows).
        // Warp 0 re-arms them -- an edge warp that owns halo rows only and does 6 of the inner warps' 16 row-steps per
        // level (measured -4.4 % against "whichever warp drains them last", whose ~300 instructions of address
        // arithmetic and copy issue landed on the critical path of a random inner warp).  It looks at the drain counter
        // before each of its steps and, at the latest, waits for it after its last one (every other warp reaches its
        // own extract without any further help from warp 0; model-checked in tests/tools/sync_model.py).
        if (lane == 0) atom_add_acqrel_u32(xcount, 1u);
        const uint32_t drained = (uint32_t)NWARPS * (uint32_t)(it + 1);  // counter value when all warps are through
        bool refill_due = warp == 0 && l + 1 < l1;
        auto try_refill = [&](bool block) {
            if (!refill_due) return;
            uint32_t seen = 0;
            do {
                if (lane == 0) seen = ld_acquire_cta_u32(xcount);
                seen = __shfl_sync(0xffffffffu, seen, 0);
            } while (block && seen < drained);
            if (seen >= drained) {
                fence_proxy_async();
                if (lane < G::TH) tl.issue_state(lane, l + 1, &mb[1]);
                refill_due = false;
            }
        };
#pragma unroll 1
        for (int s = 1; s <= P.k; ++s) {
            try_refill(false);
            wait_neighbours(g0 + (uint32_t)s);
            tl.step(tid, s, st);
            publish(g0 + (uint32_t)s + 1u);
        }
        try_refill(true);
        tl.store(tid, l, st);
    }
    }
}
#endif

}  // namespace gcmf

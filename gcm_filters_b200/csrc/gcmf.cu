// gcmf.cu -- libgcmf.so: C ABI (include/gcmf.h) + the one-step Chebyshev/Laplacian kernels.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared -Xcompiler -fPIC
// (see gcm_filters_b200/build.py).  No torch, no Python.h: plain CUDA runtime.
//
// The same file builds the TEST-ONLY host emulator (tests/hostemu/build.sh: g++ -DGCMF_HOSTEMU -x c++):
// there every "launch" is a plain loop over (b, j, i0) on host pointers, so the C ABI, the step
// sequencing and the index handling can be checked on a machine without a GPU.  The emulator is
// never built into, nor loaded by, the product (gcm_filters_b200 loads libgcmf.so only).
#ifdef GCMF_HOSTEMU
#define GCMF_HD inline
#else
#include <cuda.h>  // CUtensorMap types only: cuTensorMapEncodeTiled is resolved at run time, libcuda is not linked
#include <cuda_runtime.h>
#endif

#include <atomic>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gcmf.h"
#include "gcmf_internal.h"
#include "gcmf_stencils.cuh"
#include "gcmf_fused.cuh"
#ifndef GCMF_HOSTEMU
#include "gcmf_march.cuh"
#endif

using namespace gcmf;

// ------------------------------------------------------------------ errors / bookkeeping
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

int gcmf_set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
void gcmf_count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#ifdef GCMF_HOSTEMU
typedef void* cudaStream_t;
#define CUDA_TRY(expr) do { } while (0)
#else
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return gcmf_set_error(GCMF_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));     \
    } while (0)
#endif

extern "C" int gcmf_version(void) { return GCMF_VERSION; }
#ifdef GCMF_HOSTEMU
extern "C" int gcmf_sm_arch(void) { return 0; }  // 0 = host emulator (tests only)
#else
extern "C" int gcmf_sm_arch(void) { return 100; }
#endif
extern "C" const char* gcmf_last_error(void) { return g_err.c_str(); }
extern "C" int64_t gcmf_launch_count(void) { return g_launches.load(); }

static int n_planes_required(int op, int flags) {
    switch (op) {
        case GCMF_OP_REGULAR5: return (flags & GCMF_FLAG_AREA) ? 2 : ((flags & GCMF_FLAG_MASK) ? 1 : 0);
        case GCMF_OP_FLUX: return 3;
        case GCMF_OP_VECTOR_B: return 8;
        case GCMF_OP_VECTOR_C: return 14;
    }
    return -1;
}

// ------------------------------------------------------------------ plan
extern "C" int gcmf_plan_create(const gcmf_plan_desc* d, gcmf_plan** out) {
    if (!d || !out) return gcmf_set_error(GCMF_EINVAL, "null argument");
    if (d->op < 0 || d->op > GCMF_OP_VECTOR_C) return gcmf_set_error(GCMF_EINVAL, "unknown op %d", d->op);
    if (d->dtype != GCMF_F32 && d->dtype != GCMF_F64) return gcmf_set_error(GCMF_EINVAL, "unknown dtype %d", d->dtype);
    if (d->ny < 1 || d->nx < 2) return gcmf_set_error(GCMF_EINVAL, "grid %d x %d too small", d->ny, d->nx);
    if ((d->flags & GCMF_FLAG_MASK) && d->op != GCMF_OP_REGULAR5)
        return gcmf_set_error(GCMF_EINVAL, "GCMF_FLAG_MASK is only meaningful for GCMF_OP_REGULAR5");
    if ((d->flags & GCMF_FLAG_AREA) && d->op != GCMF_OP_REGULAR5)
        return gcmf_set_error(GCMF_EINVAL, "GCMF_FLAG_AREA is only meaningful for GCMF_OP_REGULAR5");
    if ((d->flags & (GCMF_FLAG_FOLD_N | GCMF_FLAG_CUT_S)) && d->op >= GCMF_OP_VECTOR_B)
        return gcmf_set_error(GCMF_EINVAL, "the tripolar fold is not defined for the vector operators");
    if ((d->flags & GCMF_FLAG_FOLD_N) && (d->flags & GCMF_FLAG_WRAP_Y) && !(d->flags & GCMF_FLAG_CUT_S))
        return gcmf_set_error(GCMF_EINVAL, "a folded grid that wraps in y needs GCMF_FLAG_CUT_S");
#ifndef GCMF_HOSTEMU
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (d->device < 0 || d->device >= ndev) return gcmf_set_error(GCMF_EINVAL, "device %d not present", d->device);
#endif
    gcmf_plan* p = new gcmf_plan();
    p->desc = *d;
    p->ncomp = d->op >= GCMF_OP_VECTOR_B ? 2 : 1;
    p->n_planes = n_planes_required(d->op, d->flags);
    memset(p->plane, 0, sizeof p->plane);
    p->n_steps = 0;
    p->c = 0.0;
    p->steps_per_block = 0;
    for (auto& m : p->coef_maps) m.p = nullptr;
#ifdef GCMF_HOSTEMU
    p->sm_count = 1;
#else
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, d->device));
    p->sm_count = prop.multiProcessorCount;
#endif
    *out = p;
    return GCMF_OK;
}

extern "C" int gcmf_plan_destroy(gcmf_plan* p) {
    delete p;
    return GCMF_OK;
}

extern "C" int gcmf_plan_set_plane(gcmf_plan* p, int slot, const void* dptr, int64_t pitch, int64_t bstride,
                                   int32_t plane_nb) {
    if (!p) return gcmf_set_error(GCMF_EINVAL, "null plan");
    if (slot < 0 || slot >= p->n_planes)
        return gcmf_set_error(GCMF_EINVAL, "plane slot %d out of range (op needs %d planes)", slot, p->n_planes);
    if (!dptr || pitch < p->desc.nx || plane_nb < 1)
        return gcmf_set_error(GCMF_EINVAL, "bad plane (ptr %p pitch %lld nb %d)", dptr, (long long)pitch, plane_nb);
    p->plane[slot] = PlaneRef{dptr, pitch, bstride, plane_nb};
    return GCMF_OK;
}

extern "C" int gcmf_plan_set_filter(gcmf_plan* p, int32_t n_steps, const double* coef, double c) {
    if (!p || !coef) return gcmf_set_error(GCMF_EINVAL, "null argument");
    if (n_steps < 2) return gcmf_set_error(GCMF_EINVAL, "n_steps must be >= 2 (got %d)", n_steps);
    p->n_steps = n_steps;
    p->p.assign(coef, coef + n_steps + 1);
    p->c = c;
    return GCMF_OK;
}

static int check_planes(const gcmf_plan* p) {
    for (int s = 0; s < p->n_planes; ++s) {
        if (p->desc.op == GCMF_OP_REGULAR5 && s == 0 && !(p->desc.flags & GCMF_FLAG_MASK)) continue;
        if (!p->plane[s].p) return gcmf_set_error(GCMF_ESTATE, "coefficient plane %d has not been set", s);
    }
    if (p->desc.op == GCMF_OP_VECTOR_C)
        for (int s = 1; s < 14; ++s)
            if (p->plane[s].pitch != p->plane[0].pitch)
                return gcmf_set_error(GCMF_EINVAL, "GCMF_OP_VECTOR_C planes must share one pitch");
    return GCMF_OK;
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t buffer_bytes(const gcmf_plan* p, int64_t nb) {
    const size_t w = p->desc.dtype == GCMF_F64 ? 8 : 4;
    return round_up((size_t)nb * p->desc.ny * p->desc.nx * w, 256);
}

static bool fused_eligible(const gcmf_plan* p);
static bool cg2_eligible(const gcmf_plan* p);
static bool plan_uses_fused(const gcmf_plan* p) { return p->steps_per_block != 1 && fused_eligible(p); }
// vector plans (VECTOR_C, VECTOR_B) that run two Chebyshev steps per launch (gcmf_vec2.cuh)
static bool plan_uses_cg2(const gcmf_plan* p) { return p->steps_per_block != 1 && cg2_eligible(p); }
static bool is_band_plan(const gcmf_plan* p) { return !(p->desc.flags & GCMF_FLAG_WRAP_Y); }

extern "C" int gcmf_plan_set_steps_per_block(gcmf_plan* p, int32_t k) {
    if (!p) return gcmf_set_error(GCMF_EINVAL, "null plan");
    if (k < 0 || k > FusedGeom<double>::H)
        return gcmf_set_error(GCMF_EINVAL, "steps_per_block must be 0 (auto) or 1..%d", FusedGeom<double>::H);
    p->steps_per_block = k;
    return GCMF_OK;
}

extern "C" int gcmf_workspace_bytes(const gcmf_plan* p, int64_t nb, size_t* bytes) {
    if (!p || !bytes || nb < 1) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    *bytes = (size_t)((plan_uses_fused(p) || plan_uses_cg2(p)) ? 4 : 2) * p->ncomp * buffer_bytes(p, nb);
    return GCMF_OK;
}

// ------------------------------------------------------------------ kernels
constexpr int BX = 32, BY = 8;

// ---- halo flags (system scope: the neighbour GPU writes / reads them over NVLink) ----
#ifndef GCMF_HOSTEMU
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// CTAs of the first / last row band: wait until the neighbour's rows have landed in my ghost rows
template <typename T> __device__ __forceinline__ void halo_wait(const HaloRef<T>& h, bool bottom, bool top) {
    if (!h.enabled || !(bottom || top)) return;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        unsigned spins = 0;
        if (bottom && h.wait_s)
            while ((int32_t)(ld_acquire_sys(h.wait_s) - h.wait_v) < 0)
                if (++spins > (1u << 26)) asm volatile("trap;");
        if (top && h.wait_n)
            while ((int32_t)(ld_acquire_sys(h.wait_n) - h.wait_v) < 0)
                if (++spins > (1u << 26)) asm volatile("trap;");
    }
    __syncthreads();
}
// ... and tell the neighbours once every border CTA of this launch has pushed its rows ("last block" pattern)
template <typename T>
__device__ __forceinline__ void halo_signal(const HaloRef<T>& h, bool bottom, bool top, unsigned n_border) {
    if (!h.enabled || !(bottom || top)) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        if (top && h.sig_n && atomicAdd(&h.counters[0], 1u) == n_border - 1) {
            h.counters[0] = 0;
            __threadfence_system();
            st_release_sys(h.sig_n, h.sig_v);
        }
        if (bottom && h.sig_s && atomicAdd(&h.counters[1], 1u) == n_border - 1) {
            h.counters[1] = 0;
            __threadfence_system();
            st_release_sys(h.sig_s, h.sig_v);
        }
    }
}
#endif

// 1-D grid; block id decodes as (x-block fastest, then batch, then row band): every batch slice
// of a row band is swept before the next band, so the band's coefficient-plane rows are read
// from HBM once and hit in L2 for the other nb-1 slices.
#ifndef GCMF_HOSTEMU
template <typename T, int VX, class OP, int MODE, bool HALO>
#ifndef GCMF_STEP_MINBLOCKS
#define GCMF_STEP_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(BX* BY, GCMF_STEP_MINBLOCKS) step_kernel(const __grid_constant__ StepParams<T> P, unsigned nxb) {
    unsigned bid = blockIdx.x;  // 32-bit decode: the grid has < 2^31 blocks
    const unsigned xb = bid % nxb;
    bid /= nxb;
    const unsigned nbu = (unsigned)P.nb;
    const int b = (int)(bid % nbu);
    const int yb = (int)(bid / nbu);
    const int i0 = (int)(xb * BX + threadIdx.x) * VX;
    const int j = yb * BY + (int)threadIdx.y;
    const bool bottom = yb == 0, top = (yb + 1) * BY >= P.g.ny;
    if (HALO) halo_wait<T>(P.halo, bottom, top);
    if (i0 < P.g.nx && j < P.g.ny) step_body<T, VX, OP, MODE, HALO>(P, b, j, i0);
    if (HALO) halo_signal<T>(P.halo, bottom, top, nxb * nbu);
}

// VECTOR_C: stresses once per point into shared memory, then the divergence (see CgridTile)
#ifndef GCMF_CGRID_MINBLOCKS
#define GCMF_CGRID_MINBLOCKS 6  // <= 40 registers: measured best on B200 (cfg5: 0.46 ms vs 0.72 ms per step at 92 registers)
#endif
template <typename T, int MODE, bool HALO>
__global__ void __launch_bounds__(CgridTile<T>::NTHREADS, GCMF_CGRID_MINBLOCKS) cgrid_kernel(const __grid_constant__ StepParams<T> P, unsigned nxb) {
    using C = CgridTile<T>;
    __shared__ T sm[C::SMEM_ELEMS];
    unsigned bid = blockIdx.x;
    const unsigned xb = bid % nxb;
    bid /= nxb;
    const unsigned nbu = (unsigned)P.nb;
    const int b = (int)(bid % nbu);
    const int yb = (int)(bid / nbu);
    const int j0 = yb * C::TY, i0 = (int)xb * C::TX;
    const int tid = threadIdx.x;
    const bool bottom = yb == 0, top = (yb + 1) * C::TY >= P.g.ny;
    if (HALO) halo_wait<T>(P.halo, bottom, top);
    for (int e = tid; e < C::SW * C::SH; e += C::NTHREADS) C::stress(P, b, j0, i0, e, sm);
    __syncthreads();
    const int ty = tid / C::TX, tx = tid % C::TX;
    const int j = j0 + ty, i = i0 + tx;
    if (j < P.g.ny && i < P.g.nx) {
        T lap[2][1], x[2][1];
        C::divergence(P, b, j, i, ty, tx, sm, lap, x);
        step_tail<T, 1, 2, false, MODE, HALO>(P, b, j, i, lap, x);
    }
    if (HALO) halo_signal<T>(P.halo, bottom, top, nxb * nbu);
}

// VECTOR_C, row-marching form (the default).  A warp owns CG_COLS = 30 output columns (lanes 1..30; lanes 0 and 31 carry
// the west / east neighbour column) and marches north through a band of rows, one row per iteration, keeping in
// registers everything of row j that row j+1 needs again: the point products b = v/dxCv, c = v/dyCv, e = u/dxCu, the raw
// (u, v), the four reciprocal spacings, the weighted stresses dyT^2*sxx, dxT^2*sxx of row j and dxBu^2*sxy of row j-1.
// West / east neighbours come from the adjacent lanes (four fp64 shuffles per row).  Every field and coefficient value
// is loaded from memory ONCE per output point (plus one priming row per band and two halo lanes per warp): 160 B read
// + 32 B written per point-step = B_alg, where the tiled kernel re-read ~24 values per stress entry through L1 with
// ~300 instructions of 64-bit address arithmetic and spilled at its 40-register cap.  Same expressions in the same
// order as OpVectorC / CgridTile (kernels.py:647-696): results are bit-identical.
// Band order of a launch that exchanges ghost rows (HALO): the bottom band first, the TOP band second -- both raise the
// neighbours' flags, and the neighbours' first CTAs of the next step wait for them -- then the interior bands.
__device__ __forceinline__ int halo_band_order(int raw, int nbands) {
    return raw == 0 ? 0 : (raw == 1 ? nbands - 1 : raw - 1);
}
constexpr int CG_WARPS = 4;   // warps per CTA: neighbouring column tiles of one row band (their halo columns hit in L1)
constexpr int CG_COLS = 30;   // output columns per warp
#ifndef GCMF_CGM_MINBLOCKS
#define GCMF_CGM_MINBLOCKS 4  // 127 registers, no spills, 16 warps per SM: A/B on cfg5 (ms per step) 1: 0.501, 4: 0.389, 6: 0.520, 8: 0.433
#endif
// read-only loads (ld.global.nc): the inputs of a step are never written by it, so the compiler may hoist the next
// row's loads above the current row's stores (GCMF_CGM_LDG=0: plain loads, ordered with the stores)
#ifndef GCMF_CGM_LDG
#define GCMF_CGM_LDG 1
#endif
#if GCMF_CGM_LDG
#define LDRO(p) __ldg(p)
#else
#define LDRO(p) (*(p))
#endif
template <typename T, int MODE, bool HALO>
__global__ void __launch_bounds__(32 * CG_WARPS, GCMF_CGM_MINBLOCKS) cgrid_march_kernel(const __grid_constant__ StepParams<T> P, unsigned ctas_x,
                                                                   int ry) {
    unsigned bid = blockIdx.x;
    const unsigned cx = bid % ctas_x;
    bid /= ctas_x;
    const unsigned nbu = (unsigned)P.nb;
    const int b = (int)(bid % nbu);
    const int ny = P.g.ny, nx = P.g.nx;
    const int band = HALO ? halo_band_order((int)(bid / nbu), (ny + ry - 1) / ry) : (int)(bid / nbu);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = band * ry, j1 = j0 + ry < ny ? j0 + ry : ny;
    const bool bottom = j0 == 0, top = j1 >= ny;
    if (HALO) halo_wait<T>(P.halo, bottom, top);
    const int i = (int)(cx * CG_WARPS + warp) * CG_COLS + lane - 1;  // may be -1 or >= nx: wrapped once (nx >= 32)
    const bool active = (int)(cx * CG_WARPS + warp) * CG_COLS < nx;
    if (active) {
        const int ic = i < 0 ? i + nx : (i >= nx ? i - nx : i);
        const bool wrap = (P.g.flags & FL_WRAP_Y) != 0;
        const T* U = P.t1[0].p + (int64_t)b * P.t1[0].bstride + ic;
        const T* V = P.t1[1].p + (int64_t)b * P.t1[1].bstride + ic;
        const int64_t fp = P.t1[0].pitch, pp = P.plane[0].pitch;
        const T* K[14];
#pragma unroll
        for (int s = 0; s < 14; ++s) K[s] = plane_base<T>(P.plane[s], b) + ic;
        const bool emit_lane = lane >= 1 && lane <= CG_COLS && i < nx;
        // row r of the arrays (periodic y, or ghost rows -1 / ny present in memory for a latitude band)
        auto rowidx = [&](int r) { return wrap ? (r < 0 ? r + ny : (r >= ny ? r - ny : r)) : r; };
        // state of "row j"
        T u_j, v_j, b_j, c_j, e_j, k0_j, k1_j, k2_j, k3_j, p1_j = T(0), p2_j = T(0), p3_jm = T(0);
        {
            const int64_t r = rowidx(j0 - 1);
            u_j = LDRO(U + r * fp);
            v_j = LDRO(V + r * fp);
            k0_j = LDRO(K[0] + r * pp); k1_j = LDRO(K[1] + r * pp); k2_j = LDRO(K[2] + r * pp); k3_j = LDRO(K[3] + r * pp);
            const T zu = nan2num(u_j), zv = nan2num(v_j);
            b_j = zv * k1_j; c_j = zv * k2_j; e_j = zu * k3_j;
        }
        for (int j = j0 - 1; j < j1; ++j) {
            const int64_t rn = rowidx(j + 1), rj = rowidx(j);
            // row j+1: raw values, reciprocal spacings, point products (kernels.py:653-656, 663-666)
            const T u_n = LDRO(U + rn * fp), v_n = LDRO(V + rn * fp);
            const T k0_n = LDRO(K[0] + rn * pp), k1_n = LDRO(K[1] + rn * pp), k2_n = LDRO(K[2] + rn * pp), k3_n = LDRO(K[3] + rn * pp);
            const T k4 = LDRO(K[4] + rn * pp), k5 = LDRO(K[5] + rn * pp), k8 = LDRO(K[8] + rn * pp), k9 = LDRO(K[9] + rn * pp);
            const T k6 = LDRO(K[6] + rj * pp), k7 = LDRO(K[7] + rj * pp), k10 = LDRO(K[10] + rj * pp), k11 = LDRO(K[11] + rj * pp);
            const T zu = nan2num(u_n), zv = nan2num(v_n);
            const T a_n = zu * k0_n, b_n = zv * k1_n, c_n = zv * k2_n, e_n = zu * k3_n;
            // str_xx at T point (j+1, i)  (kernels.py:653-661)
            const T a_w = __shfl_up_sync(0xffffffffu, a_n, 1);
            const T sxx = -(k4 * (a_n - a_w) - k5 * (b_n - b_j));
            const T p1_n = k8 * sxx, p2_n = k9 * sxx;   // dy2h * sxx, dx2h * sxx
            // str_xy at q point (j, i)  (kernels.py:663-670)
            const T c_e = __shfl_down_sync(0xffffffffu, c_j, 1);
            const T sxy = -(k6 * (c_e - c_j) + k7 * (e_n - e_j));
            const T p3_j = k10 * sxy, p4_j = k11 * sxy;  // dx2q * sxy, dy2q * sxy
            const T p1_e = __shfl_down_sync(0xffffffffu, p1_j, 1);
            const T p4_w = __shfl_up_sync(0xffffffffu, p4_j, 1);
            if (j >= j0 && emit_lane) {  // divergence at (j, i)  (kernels.py:672-694)
                T lap[2][1], x[2][1];
                T uc = k0_j * (p1_j - p1_e);
                uc = uc + k3_j * (p3_jm - p3_j);
                lap[0][0] = uc * LDRO(K[12] + rj * pp);
                T vc = k2_j * (p4_w - p4_j);
                vc = vc - k1_j * (p2_j - p2_n);
                lap[1][0] = vc * LDRO(K[13] + rj * pp);
                x[0][0] = u_j;
                x[1][0] = v_j;
                step_tail<T, 1, 2, false, MODE, HALO>(P, b, j, i, lap, x);
            }
            u_j = u_n; v_j = v_n; b_j = b_n; c_j = c_n; e_j = e_n;
            k0_j = k0_n; k1_j = k1_n; k2_j = k2_n; k3_j = k3_n;
            p1_j = p1_n; p2_j = p2_n; p3_jm = p3_j;
        }
    }
    if (HALO) halo_signal<T>(P.halo, bottom, top, ctas_x * nbu);
}

#undef LDRO
#include "gcmf_vec2.cuh"  // two-step vector kernel: uses halo_wait / halo_signal / halo_band_order from above
// VECTOR_C, TMA-pipelined row streaming (the default for 16-byte aligned arrays).  ncu of the marching kernel above:
// 89 % of the warp stalls are long_scoreboard, DRAM at 54 % -- its loads live in registers, so the bytes in flight are
// capped by the register file (16 warps x 24 loads).  Here the loads live in shared memory instead: a producer lane
// streams whole rows of all 20 arrays of a step (u, v, the 14 coefficient planes, T_{i-2} and bar of both components;
// 240 output columns + an aligned halo) through a 5-deep ring with the TMA engine (cp.async.bulk -> UBLKCP, one copy per
// array and row, mbarrier complete_tx), three rows ahead of the eight consumer warps, which run the marching
// recurrence of cgrid_march_kernel on shared-memory reads and release a row's slot through an "empty" mbarrier.
// ~117 KB in flight per SM, no register cost.  Same expressions and order: bit-identical results.
constexpr int CGT_WARPS = 8;                  // consumer warps per CTA (+ 1 producer warp)
constexpr int CGT_COLS = CGT_WARPS * CG_COLS; // 240 output columns per CTA
template <typename T> struct CgtGeom {
    static constexpr int AV = 16 / (int)sizeof(T);       // elements per 16 bytes
    static constexpr int HALO = AV;                      // aligned halo columns on either side (one is needed)
    static constexpr int LW = CGT_COLS + 2 * HALO;       // staged columns per row: 244 (f64) / 248 (f32)
    static constexpr int NARR = 20;                      // u, v, K0..K13, t2u, t2v, bar_u, bar_v
    static constexpr int STAGE_ELEMS = NARR * LW;
    static constexpr int STAGES = 5;
    static constexpr size_t smem_bytes() { return (size_t)STAGES * STAGE_ELEMS * sizeof(T) + 2 * STAGES * sizeof(uint64_t) + 128; }
};

template <typename T, int MODE, bool HALO>
__device__ __forceinline__ void cgt_tail(const StepParams<T>& P, int b, int j, int i, const T (&lap)[2], const T (&x)[2],
                                         const T (&t2)[2], const T (&bar)[2]) {
    const T c = (T)P.c;  // same arithmetic as step_tail (filter.py:225-283), operands already in registers
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        T* t0p = P.t0[k].p + (int64_t)b * P.t0[k].bstride + (int64_t)j * P.t0[k].pitch + i;
        if (MODE == MODE_LAP) {
            *t0p = lap[k];
            continue;
        }
        const T a = -x[k] - c * lap[k];  // shifted Laplacian, filter.py:232-236
        T* barp = P.bar[k].p + (int64_t)b * P.bar[k].bstride + (int64_t)j * P.bar[k].pitch + i;
        if (MODE == MODE_FIRST) {
            *t0p = a;
            if (HALO) { const T v1[1] = {a}; halo_push_row<T, 1>(P, k, b, j, i, v1); }
            *barp = (T)bar_update(P.p0 * (double)x[k], P.p1, (double)a);
        } else {
            const T t0 = cheb_next<T>(a, t2[k]);
            if (MODE == MODE_MID) {
                *t0p = t0;
                if (HALO) { const T v1[1] = {t0}; halo_push_row<T, 1>(P, k, b, j, i, v1); }
            }
            *barp = (T)bar_update((double)bar[k], P.p1, (double)t0);
        }
    }
}

template <typename T, int MODE, bool HALO>
__global__ void __launch_bounds__(32 * (CGT_WARPS + 1), 1) cgrid_tma_kernel(const __grid_constant__ StepParams<T> P,
                                                                            unsigned ctas_x, int ry) {
    using G = CgtGeom<T>;
    constexpr int NA = (MODE == MODE_MID || MODE == MODE_LAST) ? 20 : 16;  // arrays of a full row
    extern __shared__ __align__(128) unsigned char cgt_smem[];
    T* ring = reinterpret_cast<T*>(cgt_smem);
    uint64_t* full = reinterpret_cast<uint64_t*>(cgt_smem + (size_t)G::STAGES * G::STAGE_ELEMS * sizeof(T));
    uint64_t* empty = full + G::STAGES;
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
        for (int s = 0; s < G::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CGT_WARPS);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (HALO) halo_wait<T>(P.halo, bottom, top);
    const bool wrap = (P.g.flags & FL_WRAP_Y) != 0;
    auto rowidx = [&](int r) { return wrap ? (r < 0 ? r + ny : (r >= ny ? r - ny : r)) : r; };
    const int nstage = j1 - j0 + 2;  // rows j0-1 .. j1
    if (warp == CGT_WARPS) {
        // ---- producer warp: lane a owns array a (u, v, K0..K13, t2u, t2v, bar_u, bar_v) and issues that array's row
        // copy itself, so a row of the ring is armed by one warp-wide cp.async.bulk instead of a 20-iteration loop of a
        // single lane (measured: the single-lane producer took ~4000 cycles per row and starved the consumers)
        const T* base = nullptr;
        int64_t pitch = 0;
        if (lane < 2) {
            base = lane == 0 ? P.t1[0].p + (int64_t)b * P.t1[0].bstride : P.t1[1].p + (int64_t)b * P.t1[1].bstride;
            pitch = P.t1[0].pitch;
        }
#pragma unroll
        for (int k = 0; k < 14; ++k)
            if (lane == 2 + k) {
                base = plane_base<T>(P.plane[k], b);
                pitch = P.plane[k].pitch;
            }
        if (NA == 20) {
            if (lane == 16) { base = P.t2[0].p + (int64_t)b * P.t2[0].bstride; pitch = P.t2[0].pitch; }
            if (lane == 17) { base = P.t2[1].p + (int64_t)b * P.t2[1].bstride; pitch = P.t2[1].pitch; }
            if (lane == 18) { base = P.bar[0].p + (int64_t)b * P.bar[0].bstride; pitch = P.bar[0].pitch; }
            if (lane == 19) { base = P.bar[1].p + (int64_t)b * P.bar[1].bstride; pitch = P.bar[1].pitch; }
        }
        if (HALO) asm volatile("fence.proxy.async;" ::: "memory");  // ghost rows written by the peer, acquired above
        const int col0 = cx * CGT_COLS - G::HALO;
        const int gx = col0 < 0 ? col0 + nx : col0;
        const int n1 = (nx - gx) < G::LW ? (nx - gx) : G::LW;
        for (int s = 0; s < nstage; ++s) {
            const int slot = s % G::STAGES;
            if (s >= G::STAGES) {
                mbar_wait(&empty[slot], (unsigned)(((s / G::STAGES) - 1) & 1));
                fence_proxy_async();
            }
            const int r = j0 - 1 + s;
            const int na = (NA == 20 && r >= j0 && r < j1) ? 20 : 16;  // T_{i-2} / bar only exist for the owned rows
            if (lane == 0) mbar_expect_tx(&full[slot], (unsigned)(na * G::LW * sizeof(T)));
            __syncwarp();
            if (lane < na) {
                const T* row = base + (int64_t)rowidx(r) * pitch;
                T* dst = ring + (size_t)slot * G::STAGE_ELEMS + lane * G::LW;
                bulk_copy_g2s(dst, row + gx, (unsigned)(n1 * sizeof(T)), &full[slot]);
                if (n1 < G::LW) bulk_copy_g2s(dst + n1, row, (unsigned)((G::LW - n1) * sizeof(T)), &full[slot]);
            }
        }
    } else {  // ---- consumers: the marching recurrence of cgrid_march_kernel on shared-memory rows
        const int lc = G::HALO - 1 + warp * CG_COLS + lane;
        const int i = cx * CGT_COLS + warp * CG_COLS + lane - 1;
        const bool emit_lane = lane >= 1 && lane <= CG_COLS && i < nx;
        mbar_wait(&full[0], 0);
        const T* s0 = ring + lc;
        T u_j = s0[0], v_j = s0[G::LW];
        T k0_j = s0[2 * G::LW], k1_j = s0[3 * G::LW], k2_j = s0[4 * G::LW], k3_j = s0[5 * G::LW];
        T b_j, c_j, e_j, p1_j = T(0), p2_j = T(0), p3_jm = T(0);
        {
            const T zu = nan2num(u_j), zv = nan2num(v_j);
            b_j = zv * k1_j; c_j = zv * k2_j; e_j = zu * k3_j;
        }
        for (int s = 0; s + 1 < nstage; ++s) {
            const int j = j0 - 1 + s, sn = s + 1;
            mbar_wait(&full[sn % G::STAGES], (unsigned)((sn / G::STAGES) & 1));
            const T* cur = ring + (size_t)(s % G::STAGES) * G::STAGE_ELEMS + lc;
            const T* nxt = ring + (size_t)(sn % G::STAGES) * G::STAGE_ELEMS + lc;
            const T u_n = nxt[0], v_n = nxt[G::LW];
            const T k0_n = nxt[2 * G::LW], k1_n = nxt[3 * G::LW], k2_n = nxt[4 * G::LW], k3_n = nxt[5 * G::LW];
            const T k4 = nxt[6 * G::LW], k5 = nxt[7 * G::LW], k8 = nxt[10 * G::LW], k9 = nxt[11 * G::LW];
            const T k6 = cur[8 * G::LW], k7 = cur[9 * G::LW], k10 = cur[12 * G::LW], k11 = cur[13 * G::LW];
            const T zu = nan2num(u_n), zv = nan2num(v_n);
            const T a_n = zu * k0_n, b_n = zv * k1_n, c_n = zv * k2_n, e_n = zu * k3_n;
            const T a_w = __shfl_up_sync(0xffffffffu, a_n, 1);
            const T sxx = -(k4 * (a_n - a_w) - k5 * (b_n - b_j));      // kernels.py:653-661
            const T p1_n = k8 * sxx, p2_n = k9 * sxx;
            const T c_e = __shfl_down_sync(0xffffffffu, c_j, 1);
            const T sxy = -(k6 * (c_e - c_j) + k7 * (e_n - e_j));      // kernels.py:663-670
            const T p3_j = k10 * sxy, p4_j = k11 * sxy;
            const T p1_e = __shfl_down_sync(0xffffffffu, p1_j, 1);
            const T p4_w = __shfl_up_sync(0xffffffffu, p4_j, 1);
            if (j >= j0 && emit_lane) {  // kernels.py:672-694
                T lap[2], x[2] = {u_j, v_j}, t2[2] = {T(0), T(0)}, bar[2] = {T(0), T(0)};
                T uc = k0_j * (p1_j - p1_e);
                uc = uc + k3_j * (p3_jm - p3_j);
                lap[0] = uc * cur[14 * G::LW];
                T vc = k2_j * (p4_w - p4_j);
                vc = vc - k1_j * (p2_j - p2_n);
                lap[1] = vc * cur[15 * G::LW];
                if (NA == 20) {
                    t2[0] = cur[16 * G::LW]; t2[1] = cur[17 * G::LW];
                    bar[0] = cur[18 * G::LW]; bar[1] = cur[19 * G::LW];
                }
                cgt_tail<T, MODE, HALO>(P, b, j, i, lap, x, t2, bar);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s % G::STAGES]);  // this warp is done with row j's slot
            u_j = u_n; v_j = v_n; b_j = b_n; c_j = c_n; e_j = e_n;
            k0_j = k0_n; k1_j = k1_n; k2_j = k2_n; k3_j = k3_n;
            p1_j = p1_n; p2_j = p2_n; p3_jm = p3_j;
        }
    }
    if (HALO) halo_signal<T>(P.halo, bottom, top, ctas_x * nbu);
}

// Copy my first / last owned rows of a field into the neighbours' ghost rows and raise their flags (the
// exchange of the prepared input before step 1).  One CTA column per x-block; grid (nxb, nb).
template <typename T>
__global__ void halo_push_kernel(FieldRef<const T> f0, FieldRef<const T> f1, int ncomp, int ny, int nx, HaloRef<T> h) {
    halo_wait<T>(h, true, true);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i < nx) {
        for (int k = 0; k < ncomp; ++k) {
            const FieldRef<const T>& f = k ? f1 : f0;
            const T* base = f.p + (int64_t)b * f.bstride;
            if (h.north[k]) h.north[k][(int64_t)b * h.nbs + i] = base[(int64_t)(ny - 1) * f.pitch + i];
            if (h.south[k]) h.south[k][(int64_t)b * h.sbs + i] = base[i];
        }
    }
    halo_signal<T>(h, true, true, gridDim.x * gridDim.y);
}

template <typename T>
__global__ void prepare_kernel(const T* in, int64_t in_pitch, int64_t in_bs, T* out, int64_t out_pitch, int64_t out_bs,
                               PlaneRef area, int ny, int nx, int64_t nb, bool divide) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= nx) return;
    for (int64_t b = blockIdx.z; b < nb; b += gridDim.z) {
        const T* a = plane_base<T>(area, (int)b);
        prepare_body<T>(in, out, a, b * in_bs + (int64_t)j * in_pitch + i, b * out_bs + (int64_t)j * out_pitch + i,
                        (int64_t)j * area.pitch + i, divide);
    }
}

#endif  // !GCMF_HOSTEMU

template <typename T, int VX, class OP, int MODE>
static int launch_step(const StepParams<T>& P, cudaStream_t st) {
#ifdef GCMF_HOSTEMU
    (void)st;
    for (int64_t b = 0; b < P.nb; ++b)
        for (int j = 0; j < P.g.ny; ++j)
            for (int i0 = 0; i0 < P.g.nx; i0 += VX) {
                if (P.halo.enabled) step_body<T, VX, OP, MODE, true>(P, (int)b, j, i0);
                else step_body<T, VX, OP, MODE, false>(P, (int)b, j, i0);
            }
    gcmf_count_launch(1);
    return GCMF_OK;
#else
    const int nxb = (P.g.nx + BX * VX - 1) / (BX * VX);
    const int nyb = (P.g.ny + BY - 1) / BY;
    const int64_t nblk = (int64_t)nxb * nyb * P.nb;
    if (nblk > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "grid too large (%lld blocks)", (long long)nblk);
    if (P.halo.enabled) step_kernel<T, VX, OP, MODE, true><<<(unsigned)nblk, dim3(BX, BY), 0, st>>>(P, (unsigned)nxb);
    else step_kernel<T, VX, OP, MODE, false><<<(unsigned)nblk, dim3(BX, BY), 0, st>>>(P, (unsigned)nxb);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
#endif
}

template <typename T, int MODE> static int launch_cgrid(const StepParams<T>& P, cudaStream_t st) {
    using C = CgridTile<T>;
    const int nxb = (P.g.nx + C::TX - 1) / C::TX;
    const int nyb = (P.g.ny + C::TY - 1) / C::TY;
#ifdef GCMF_HOSTEMU
    (void)st;
    std::vector<T> sm(C::SMEM_ELEMS);
    for (int yb = 0; yb < nyb; ++yb)
        for (int64_t b = 0; b < P.nb; ++b)
            for (int xb = 0; xb < nxb; ++xb) {
                const int j0 = yb * C::TY, i0 = xb * C::TX;
                for (int e = 0; e < C::SW * C::SH; ++e) C::stress(P, (int)b, j0, i0, e, sm.data());
                for (int tid = 0; tid < C::NTHREADS; ++tid) {
                    const int ty = tid / C::TX, tx = tid % C::TX;
                    const int j = j0 + ty, i = i0 + tx;
                    if (j < P.g.ny && i < P.g.nx) {
                        T lap[2][1], x[2][1];
                        C::divergence(P, (int)b, j, i, ty, tx, sm.data(), lap, x);
                        if (P.halo.enabled) step_tail<T, 1, 2, false, MODE, true>(P, (int)b, j, i, lap, x);
                        else step_tail<T, 1, 2, false, MODE, false>(P, (int)b, j, i, lap, x);
                    }
                }
            }
    gcmf_count_launch(1);
    return GCMF_OK;
#else
    const int64_t nblk = (int64_t)nxb * nyb * P.nb;
    if (nblk > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "grid too large (%lld blocks)", (long long)nblk);
    if (P.halo.enabled) cgrid_kernel<T, MODE, true><<<(unsigned)nblk, C::NTHREADS, 0, st>>>(P, (unsigned)nxb);
    else cgrid_kernel<T, MODE, false><<<(unsigned)nblk, C::NTHREADS, 0, st>>>(P, (unsigned)nxb);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
#endif
}

#ifndef GCMF_HOSTEMU
template <typename T, int MODE> static int launch_cgrid_march(const gcmf_plan* pl, const StepParams<T>& P, cudaStream_t st) {
    const int nxt = (P.g.nx + CG_COLS - 1) / CG_COLS;
    const int ctas_x = (nxt + CG_WARPS - 1) / CG_WARPS;
    // rows per band: long enough to amortise the priming row of a band, short enough for ~16 CTAs per SM
    const int64_t target = (int64_t)pl->sm_count * 16;
    int64_t nbands = (target + (int64_t)ctas_x * P.nb - 1) / ((int64_t)ctas_x * P.nb);
    if (nbands < 1) nbands = 1;
    int ry = (int)((P.g.ny + nbands - 1) / nbands);
    if (const char* e = getenv("GCMF_CGRID_ROWS")) {
        const int v = atoi(e);
        if (v > 0) ry = v;
    } else {
        ry = ry < 8 ? 8 : (ry > 64 ? 64 : ry);
    }
    nbands = (P.g.ny + ry - 1) / ry;
    const int64_t nblk = (int64_t)ctas_x * P.nb * nbands;
    if (nblk > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "grid too large (%lld blocks)", (long long)nblk);
    if (P.halo.enabled)
        cgrid_march_kernel<T, MODE, true><<<(unsigned)nblk, 32 * CG_WARPS, 0, st>>>(P, (unsigned)ctas_x, ry);
    else
        cgrid_march_kernel<T, MODE, false><<<(unsigned)nblk, 32 * CG_WARPS, 0, st>>>(P, (unsigned)ctas_x, ry);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
}
#endif

#ifndef GCMF_HOSTEMU
template <typename T, int MODE, bool HALO>
static int launch_cgrid_tma_t(const gcmf_plan* pl, const StepParams<T>& P, cudaStream_t st) {
    using G = CgtGeom<T>;
    static bool attr_done[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(cgrid_tma_kernel<T, MODE, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)G::smem_bytes()));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int ctas_x = (P.g.nx + CGT_COLS - 1) / CGT_COLS;
    // One CTA per SM (the ring takes most of the shared memory).  Rows per band, measured on cfg5 (ms per step): 17 rows
    // 0.333, 34: 0.316, 68: 0.332, 135: 0.379, 270: 0.375 -- bands of ~34 rows (two priming rows = 6 %), shorter ones only
    // when the launch would otherwise not fill the device.
    int ry = 34;
    {
        const char* e = getenv("GCMF_CGRID_ROWS");
        const int v = e ? atoi(e) : 0;
        if (v > 0) {
            ry = v;
        } else {
            const int64_t per_band = (int64_t)ctas_x * P.nb;
            while (ry > 8 && per_band * ((P.g.ny + ry - 1) / ry) < pl->sm_count) ry = (ry + 1) / 2;
        }
    }
    const int64_t nbands = (P.g.ny + ry - 1) / ry;
    const int64_t nblk = (int64_t)ctas_x * P.nb * nbands;
    if (nblk > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "grid too large (%lld blocks)", (long long)nblk);
    cgrid_tma_kernel<T, MODE, HALO><<<(unsigned)nblk, 32 * (CGT_WARPS + 1), G::smem_bytes(), st>>>(P, (unsigned)ctas_x, ry);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
}
template <typename T, int MODE> static int launch_cgrid_tma(const gcmf_plan* pl, const StepParams<T>& P, cudaStream_t st) {
    if (P.halo.enabled) return launch_cgrid_tma_t<T, MODE, true>(pl, P, st);
    return launch_cgrid_tma_t<T, MODE, false>(pl, P, st);
}
#endif

// VECTOR_C dispatch.  aligned: every array of the launch is 16-byte aligned with vector-multiple strides (bulk copies).
template <typename T>
static int launch_cgrid_mode(const gcmf_plan* pl, const StepParams<T>& P, int mode, cudaStream_t st, bool aligned16) {
#ifndef GCMF_HOSTEMU
    // The tiled kernel covers grids narrower than a warp (and is the emulator's form of the operator: bit-identical,
    // checked on the GPU by tests/cabi/gpu_vs_emu.c); GCMF_CGRID_KERNEL=tiled|march|tma forces one form (A/B, tests).
    static const char* force = getenv("GCMF_CGRID_KERNEL");
    const bool want_tiled = force && !strcmp(force, "tiled");
    const bool want_march = force && !strcmp(force, "march");
    if (!want_tiled && !want_march && aligned16 && P.g.nx >= CgtGeom<T>::LW) {
        switch (mode) {
            case MODE_LAP: return launch_cgrid_tma<T, MODE_LAP>(pl, P, st);
            case MODE_FIRST: return launch_cgrid_tma<T, MODE_FIRST>(pl, P, st);
            case MODE_MID: return launch_cgrid_tma<T, MODE_MID>(pl, P, st);
            case MODE_LAST: return launch_cgrid_tma<T, MODE_LAST>(pl, P, st);
        }
    }
    if (!want_tiled && P.g.nx >= 32) {  // the marching kernel wraps column indices once: grids at least one warp wide
        switch (mode) {
            case MODE_LAP: return launch_cgrid_march<T, MODE_LAP>(pl, P, st);
            case MODE_FIRST: return launch_cgrid_march<T, MODE_FIRST>(pl, P, st);
            case MODE_MID: return launch_cgrid_march<T, MODE_MID>(pl, P, st);
            case MODE_LAST: return launch_cgrid_march<T, MODE_LAST>(pl, P, st);
        }
    }
#else
    (void)pl; (void)aligned16;
#endif
    switch (mode) {
        case MODE_LAP: return launch_cgrid<T, MODE_LAP>(P, st);
        case MODE_FIRST: return launch_cgrid<T, MODE_FIRST>(P, st);
        case MODE_MID: return launch_cgrid<T, MODE_MID>(P, st);
        case MODE_LAST: return launch_cgrid<T, MODE_LAST>(P, st);
    }
    return gcmf_set_error(GCMF_EINVAL, "bad mode %d", mode);
}

template <typename T, int VX, class OP>
static int launch_mode(const StepParams<T>& P, int mode, cudaStream_t st) {
    switch (mode) {
        case MODE_LAP: return launch_step<T, VX, OP, MODE_LAP>(P, st);
        case MODE_FIRST: return launch_step<T, VX, OP, MODE_FIRST>(P, st);
        case MODE_MID: return launch_step<T, VX, OP, MODE_MID>(P, st);
        case MODE_LAST: return launch_step<T, VX, OP, MODE_LAST>(P, st);
    }
    return gcmf_set_error(GCMF_EINVAL, "bad mode %d", mode);
}

template <typename T> struct VecWidth;
template <> struct VecWidth<double> { static constexpr int value = 2; };
template <> struct VecWidth<float> { static constexpr int value = 4; };

template <typename T, int VX>
static int launch_op(const gcmf_plan* pl, const StepParams<T>& P, int mode, cudaStream_t st) {
    switch (pl->desc.op) {
        case GCMF_OP_REGULAR5:
            if (pl->desc.flags & GCMF_FLAG_MASK) return launch_mode<T, VX, OpRegular5<T, VX, true>>(P, mode, st);
            return launch_mode<T, VX, OpRegular5<T, VX, false>>(P, mode, st);
        case GCMF_OP_FLUX: return launch_mode<T, VX, OpFlux<T, VX>>(P, mode, st);
        case GCMF_OP_VECTOR_B: return launch_mode<T, VX, OpVectorB<T, VX>>(P, mode, st);
        case GCMF_OP_VECTOR_C: {
            // tiled kernel needs every neighbour inside the array: at least 2 rows/columns; the point-wise
            // OpVectorC kernel remains for degenerate grids (and as the test oracle of the tiled one)
            static const bool pointwise = getenv("GCMF_CGRID_POINTWISE") != nullptr;
            if (!pointwise && P.g.ny >= 2 && P.g.nx >= 2) return launch_cgrid_mode<T>(pl, P, mode, st, VX > 1);
            return launch_mode<T, 1, OpVectorC<T, 1>>(P, mode, st);
        }
    }
    return gcmf_set_error(GCMF_EINVAL, "bad op");
}

static bool aligned(const void* p, int64_t pitch, int64_t bstride, int vx, size_t elem) {
    return ((uintptr_t)p % (vx * elem) == 0) && pitch % vx == 0 && bstride % vx == 0;
}

// can every array of this launch be accessed with VX-wide vectors?
template <typename T> static bool can_vectorize(const gcmf_plan* pl, const StepParams<T>& P, int mode) {
    const int vx = VecWidth<T>::value;
    if (P.g.nx % vx) return false;
    const int nc = pl->ncomp;
    for (int k = 0; k < nc; ++k) {
        if (!aligned(P.t1[k].p, P.t1[k].pitch, P.t1[k].bstride, vx, sizeof(T))) return false;
        if (mode != MODE_LAST && !aligned(P.t0[k].p, P.t0[k].pitch, P.t0[k].bstride, vx, sizeof(T))) return false;
        if (mode >= MODE_MID && !aligned(P.t2[k].p, P.t2[k].pitch, P.t2[k].bstride, vx, sizeof(T))) return false;
        if (mode != MODE_LAP && !aligned(P.bar[k].p, P.bar[k].pitch, P.bar[k].bstride, vx, sizeof(T))) return false;
    }
    for (int s = 0; s < pl->n_planes; ++s) {
        if (!P.plane[s].p) continue;
        const bool is_mask = pl->desc.op == GCMF_OP_REGULAR5 && s == 0;
        if (!aligned(P.plane[s].p, P.plane[s].pitch, P.plane[s].bstride, vx, is_mask ? 1 : sizeof(T))) return false;
    }
    return true;
}

template <typename T> static HaloRef<T> make_halo(const gcmf_halo* h) {
    HaloRef<T> r;
    memset(&r, 0, sizeof r);
    if (!h) return r;
    for (int k = 0; k < 2; ++k) {
        r.north[k] = (T*)h->north_ghost[k];
        r.south[k] = (T*)h->south_ghost[k];
    }
    r.nbs = h->north_bstride;
    r.sbs = h->south_bstride;
    r.wait_n = h->wait_north;
    r.wait_s = h->wait_south;
    r.sig_n = h->signal_north;
    r.sig_s = h->signal_south;
    r.wait_v = h->wait_value;
    r.sig_v = h->signal_value;
    r.counters = h->counters;
    r.enabled = 1;
    return r;
}

template <typename T>
static int run_step_t(const gcmf_plan* pl, int64_t nb, int mode, const gcmf_field* t1, const gcmf_field* t2,
                      const gcmf_field* t0, const gcmf_field* bar, double p0, double p1, cudaStream_t st,
                      const gcmf_halo* halo = nullptr) {
    // The C-grid kernels address u and v (read next to each other at every point) with one row pitch.
    if (pl->desc.op == GCMF_OP_VECTOR_C && t1 && t1[0].pitch != t1[1].pitch)
        return gcmf_set_error(GCMF_EINVAL, "GCMF_OP_VECTOR_C: the u and v input fields must share one row pitch");
    StepParams<T> P;
    memset(&P, 0, sizeof P);
    P.halo = make_halo<T>(halo);
    P.g.ny = pl->desc.ny;
    P.g.nx = pl->desc.nx;
    P.g.flags = pl->desc.flags;
    for (int s = 0; s < pl->n_planes; ++s) P.plane[s] = pl->plane[s];
    for (int k = 0; k < pl->ncomp; ++k) {
        if (t1) P.t1[k] = FieldRef<const T>{(const T*)t1[k].ptr, t1[k].pitch, t1[k].bstride};
        if (t2) P.t2[k] = FieldRef<const T>{(const T*)t2[k].ptr, t2[k].pitch, t2[k].bstride};
        if (t0) P.t0[k] = FieldRef<T>{(T*)t0[k].ptr, t0[k].pitch, t0[k].bstride};
        if (bar) P.bar[k] = FieldRef<T>{(T*)bar[k].ptr, bar[k].pitch, bar[k].bstride};
    }
    P.c = pl->c;
    P.p0 = p0;
    P.p1 = p1;
    P.nb = nb;
    bool vec = can_vectorize<T>(pl, P, mode);
    if (vec && halo) {  // the neighbours' ghost rows must take vector stores too
        const int vx = VecWidth<T>::value;
        for (int k = 0; k < pl->ncomp; ++k)
            vec = vec && (uintptr_t)halo->north_ghost[k] % (vx * sizeof(T)) == 0 &&
                  (uintptr_t)halo->south_ghost[k] % (vx * sizeof(T)) == 0;
        vec = vec && halo->north_bstride % vx == 0 && halo->south_bstride % vx == 0;
    }
#ifdef GCMF_HOSTEMU
    int rc = vec ? launch_op<T, VecWidth<T>::value>(pl, P, mode, st) : launch_op<T, 1>(pl, P, mode, st);
    if (rc == GCMF_OK && halo) {  // the emulator has no concurrency: raise the flags after the "launch"
        if (halo->signal_north) *halo->signal_north = halo->signal_value;
        if (halo->signal_south) *halo->signal_south = halo->signal_value;
    }
    return rc;
#else
    if (vec) return launch_op<T, VecWidth<T>::value>(pl, P, mode, st);
    return launch_op<T, 1>(pl, P, mode, st);
#endif
}

static int run_step(const gcmf_plan* pl, int64_t nb, int mode, const gcmf_field* t1, const gcmf_field* t2,
                    const gcmf_field* t0, const gcmf_field* bar, double p0, double p1, cudaStream_t st,
                    const gcmf_halo* halo = nullptr) {
    if (pl->desc.dtype == GCMF_F64) return run_step_t<double>(pl, nb, mode, t1, t2, t0, bar, p0, p1, st, halo);
    return run_step_t<float>(pl, nb, mode, t1, t2, t0, bar, p0, p1, st, halo);
}

static int check_fields(const gcmf_plan* p, const gcmf_field* f, const char* what) {
    if (!f) return gcmf_set_error(GCMF_EINVAL, "%s: null field array", what);
    for (int k = 0; k < p->ncomp; ++k)
        if (!f[k].ptr || f[k].pitch < p->desc.nx)
            return gcmf_set_error(GCMF_EINVAL, "%s[%d]: bad field (ptr %p pitch %lld)", what, k, f[k].ptr,
                                  (long long)f[k].pitch);
    return GCMF_OK;
}

#define TRY(expr)                    \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != GCMF_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------ public compute entry points
extern "C" int gcmf_laplacian(gcmf_plan* p, int64_t nb, const gcmf_field* in, const gcmf_field* out, void* stream) {
    if (!p || nb < 1) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    TRY(check_planes(p));
    TRY(check_fields(p, in, "in"));
    TRY(check_fields(p, out, "out"));
    CUDA_TRY(cudaSetDevice(p->desc.device));
    return run_step(p, nb, MODE_LAP, in, nullptr, out, nullptr, 0.0, 0.0, (cudaStream_t)stream);
}

static int run_area_op(gcmf_plan* p, int64_t nb, const gcmf_field* in, const gcmf_field* out, bool divide, void* stream);

extern "C" int gcmf_prepare(gcmf_plan* p, int64_t nb, const gcmf_field* in, const gcmf_field* out, void* stream) {
    return run_area_op(p, nb, in, out, false, stream);
}
extern "C" int gcmf_finalize(gcmf_plan* p, int64_t nb, const gcmf_field* in, const gcmf_field* out, void* stream) {
    return run_area_op(p, nb, in, out, true, stream);
}

static int run_area_op(gcmf_plan* p, int64_t nb, const gcmf_field* in, const gcmf_field* out, bool divide, void* stream) {
    if (!p || nb < 1) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    TRY(check_fields(p, in, "in"));
    TRY(check_fields(p, out, "out"));
    CUDA_TRY(cudaSetDevice(p->desc.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t w = p->desc.dtype == GCMF_F64 ? 8 : 4;
    const int ny = p->desc.ny, nx = p->desc.nx;
#ifdef GCMF_HOSTEMU
    (void)st;
    for (int k = 0; k < p->ncomp; ++k)
        for (int64_t b = 0; b < nb; ++b)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const int64_t ii = b * in[k].bstride + (int64_t)j * in[k].pitch + i;
                    const int64_t oi = b * out[k].bstride + (int64_t)j * out[k].pitch + i;
                    if (!(p->desc.flags & GCMF_FLAG_AREA)) {
                        memcpy((char*)out[k].ptr + oi * w, (const char*)in[k].ptr + ii * w, w);
                    } else if (p->desc.dtype == GCMF_F64) {
                        prepare_body<double>((const double*)in[k].ptr, (double*)out[k].ptr,
                                             plane_base<double>(p->plane[1], (int)b), ii, oi,
                                             (int64_t)j * p->plane[1].pitch + i, divide);
                    } else {
                        prepare_body<float>((const float*)in[k].ptr, (float*)out[k].ptr,
                                            plane_base<float>(p->plane[1], (int)b), ii, oi,
                                            (int64_t)j * p->plane[1].pitch + i, divide);
                    }
                }
    if (p->desc.flags & GCMF_FLAG_AREA) gcmf_count_launch(p->ncomp);
    return GCMF_OK;
#else
    for (int k = 0; k < p->ncomp; ++k) {
        if (!(p->desc.flags & GCMF_FLAG_AREA)) {
            for (int64_t b = 0; b < nb; ++b)  // plain strided copy
                CUDA_TRY(cudaMemcpy2DAsync((char*)out[k].ptr + b * out[k].bstride * w, out[k].pitch * w,
                                           (const char*)in[k].ptr + b * in[k].bstride * w, in[k].pitch * w,
                                           (size_t)nx * w, ny, cudaMemcpyDeviceToDevice, st));
            continue;
        }
        TRY(check_planes(p));
        dim3 blk(256), grd((nx + 255) / 256, ny, (unsigned)(nb < 64 ? nb : 64));
        if (p->desc.dtype == GCMF_F64)
            prepare_kernel<double><<<grd, blk, 0, st>>>((const double*)in[k].ptr, in[k].pitch, in[k].bstride,
                                                        (double*)out[k].ptr, out[k].pitch, out[k].bstride,
                                                        p->plane[1], ny, nx, nb, divide);
        else
            prepare_kernel<float><<<grd, blk, 0, st>>>((const float*)in[k].ptr, in[k].pitch, in[k].bstride,
                                                       (float*)out[k].ptr, out[k].pitch, out[k].bstride, p->plane[1],
                                                       ny, nx, nb, divide);
        gcmf_count_launch(1);
        CUDA_TRY(cudaGetLastError());
    }
    return GCMF_OK;
#endif
}

extern "C" int gcmf_cheb_step(gcmf_plan* p, int64_t nb, int32_t step, const gcmf_field* t1_in, const gcmf_field* t2,
                              const gcmf_field* t0_out, const gcmf_field* bar, void* stream) {
    return gcmf_cheb_step_halo(p, nb, step, t1_in, t2, t0_out, bar, nullptr, stream);
}

template <typename T>
static int run_halo_push_t(const gcmf_plan* p, int64_t nb, const gcmf_field* f, const gcmf_halo* halo, cudaStream_t st) {
    HaloRef<T> h = make_halo<T>(halo);
    const int ny = p->desc.ny, nx = p->desc.nx;
#ifdef GCMF_HOSTEMU
    (void)st;
    for (int k = 0; k < p->ncomp; ++k)
        for (int64_t b = 0; b < nb; ++b)
            for (int i = 0; i < nx; ++i) {
                const T* base = (const T*)f[k].ptr + b * f[k].bstride;
                if (h.north[k]) h.north[k][b * h.nbs + i] = base[(int64_t)(ny - 1) * f[k].pitch + i];
                if (h.south[k]) h.south[k][b * h.sbs + i] = base[i];
            }
    if (h.sig_n) *h.sig_n = h.sig_v;
    if (h.sig_s) *h.sig_s = h.sig_v;
    gcmf_count_launch(1);
    return GCMF_OK;
#else
    FieldRef<const T> f0{(const T*)f[0].ptr, f[0].pitch, f[0].bstride};
    FieldRef<const T> f1 = p->ncomp > 1 ? FieldRef<const T>{(const T*)f[1].ptr, f[1].pitch, f[1].bstride} : f0;
    dim3 blk(256), grd((nx + 255) / 256, (unsigned)nb);
    halo_push_kernel<T><<<grd, blk, 0, st>>>(f0, f1, p->ncomp, ny, nx, h);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
#endif
}

extern "C" int gcmf_halo_push(gcmf_plan* p, int64_t nb, const gcmf_field* field, const gcmf_halo* halo, void* stream) {
    if (!p || nb < 1 || !halo) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    if (nb > 65535) return gcmf_set_error(GCMF_EINVAL, "halo push: nb too large");
    if (p->desc.flags & GCMF_FLAG_WRAP_Y) return gcmf_set_error(GCMF_EINVAL, "halo exchange needs a band plan (no GCMF_FLAG_WRAP_Y)");
    TRY(check_fields(p, field, "field"));
    CUDA_TRY(cudaSetDevice(p->desc.device));
    if (p->desc.dtype == GCMF_F64) return run_halo_push_t<double>(p, nb, field, halo, (cudaStream_t)stream);
    return run_halo_push_t<float>(p, nb, field, halo, (cudaStream_t)stream);
}

extern "C" int gcmf_cheb_step_halo(gcmf_plan* p, int64_t nb, int32_t step, const gcmf_field* t1_in, const gcmf_field* t2,
                                   const gcmf_field* t0_out, const gcmf_field* bar, const gcmf_halo* halo, void* stream) {
    if (!p || nb < 1) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    if (halo && (p->desc.flags & GCMF_FLAG_WRAP_Y))
        return gcmf_set_error(GCMF_EINVAL, "halo exchange needs a band plan (no GCMF_FLAG_WRAP_Y)");
    if (p->n_steps < 2) return gcmf_set_error(GCMF_ESTATE, "gcmf_plan_set_filter has not been called");
    if (step < 1 || step > p->n_steps) return gcmf_set_error(GCMF_EINVAL, "step %d outside 1..%d", step, p->n_steps);
    TRY(check_planes(p));
    TRY(check_fields(p, t1_in, "t1_in"));
    TRY(check_fields(p, bar, "bar"));
    const int mode = step == 1 ? MODE_FIRST : (step == p->n_steps ? MODE_LAST : MODE_MID);
    if (mode != MODE_FIRST) TRY(check_fields(p, t2, "t2"));
    if (mode != MODE_LAST) {
        TRY(check_fields(p, t0_out, "t0_out"));
        for (int k = 0; k < p->ncomp; ++k)
            if (t0_out[k].ptr == t1_in[k].ptr) return gcmf_set_error(GCMF_EINVAL, "t0_out must not alias t1_in");
    }
    CUDA_TRY(cudaSetDevice(p->desc.device));
    const double p0 = p->p[0], p1 = mode == MODE_FIRST ? p->p[1] : p->p[step];
    return run_step(p, nb, mode, t1_in, t2, t0_out, bar, p0, p1, (cudaStream_t)stream, halo);
}

// ------------------------------------------------------------------ temporally blocked steps
// Which plans can take the fused path: FLUX or REGULAR5 operator, doubly periodic (no fold, no band
// ghost rows), shared 2-D planes, rows that split into 16-byte vectors, grid at least one tile.
// Returns FK_FLUX / FK_REG5 or -1.
template <typename T> static int fused_kind_t(const gcmf_plan* p) {
    using G = FusedGeom<T>;
    const int fl = p->desc.flags;
    if (p->desc.nx % G::AV || p->desc.nx < G::TW || p->desc.ny < G::TH) return -1;
    const int tripolar = GCMF_FLAG_FOLD_N | GCMF_FLAG_CUT_S;
    // A latitude band (no WRAP_Y) may use the fused path too: the caller keeps FUSED_H ghost rows of the
    // fields and of every plane on each side of the band and exchanges them after every block.  A whole
    // tripolar grid carries both FOLD_N and CUT_S; of its bands only the top one folds and only the bottom
    // one is cut.
    const bool band = !(fl & GCMF_FLAG_WRAP_Y);
    if (!band && (fl & tripolar) != 0 && (fl & tripolar) != tripolar) return -1;
    const int base = (fl & ~tripolar) | GCMF_FLAG_WRAP_Y;
    if (p->desc.op == GCMF_OP_FLUX) {
        if (base != (GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM)) return -1;
        for (int s = 0; s < 3; ++s)
            if (!p->plane[s].p || p->plane[s].nb != 1 ||
                !aligned(p->plane[s].p, p->plane[s].pitch, 0, G::AV, sizeof(T)))
                return -1;
        return FK_FLUX;
    }
    if (p->desc.op == GCMF_OP_REGULAR5) {
        const int core = base & ~GCMF_FLAG_AREA;  // prepare / finalize happen outside the recurrence arithmetic
        // the last block divides by the area with vector loads: a strided / batched area plane takes the one-step path
        if ((fl & GCMF_FLAG_AREA) &&
            (!p->plane[1].p || p->plane[1].nb != 1 || !aligned(p->plane[1].p, p->plane[1].pitch, 0, G::AV, sizeof(T))))
            return -1;
        if (core == GCMF_FLAG_WRAP_Y && !(fl & tripolar)) return FK_REG5;
        if (core == (GCMF_FLAG_WRAP_Y | GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM) && p->plane[0].p && p->plane[0].nb == 1)
            return FK_REG5;
    }
    return -1;
}
static int fused_kind(const gcmf_plan* p) {
    return p->desc.dtype == GCMF_F64 ? fused_kind_t<double>(p) : fused_kind_t<float>(p);
}
static bool fused_eligible(const gcmf_plan* p) { return fused_kind(p) >= 0; }

#ifdef GCMF_HOSTEMU
template <typename T, int KIND, int EDGE> static void fused_host_t(const FusedParams<T>& P, int ncta) {
    using G = FusedGeom<T, FusedSplit<KIND>::value>;
    std::vector<T> smem((size_t)G::ntiles(KIND) * G::PLANE);
    std::vector<typename FusedTile<T, KIND, EDGE>::Thread> st(G::NTHREADS);
    const int ntiles = P.ncx * P.ncy;
    for (int cta = 0; cta < ncta; ++cta) {
        const int tile = cta % ntiles, grp = cta / ntiles;
        const int64_t l0 = (int64_t)grp * P.levels_per_cta;
        const int64_t l1 = l0 + P.levels_per_cta < P.nb ? l0 + P.levels_per_cta : P.nb;
        if (l0 >= l1) continue;
        for (auto& v : smem) v = T(12345);
        FusedTile<T, KIND, EDGE> tl(P, tile, smem.data());
        for (int r = 0; r < G::TH; ++r) {
            if (KIND == FK_FLUX) tl.issue_coef(r, nullptr);
            tl.issue_state(r, l0, nullptr);
        }
        for (int t = 0; t < G::NTHREADS; ++t) tl.load_mask(t, st[t]);
        for (int64_t l = l0; l < l1; ++l) {
            for (int t = 0; t < G::NTHREADS; ++t) tl.load_bar(t, l, st[t]);
            for (int t = 0; t < G::NTHREADS; ++t) tl.extract(t, st[t]);
            if (l + 1 < l1) for (int r = 0; r < G::TH; ++r) tl.issue_state(r, l + 1, nullptr);
            for (int s = 1; s <= P.k; ++s)
                for (int t = 0; t < G::NTHREADS; ++t) tl.step(t, s, st[t]);
            for (int t = 0; t < G::NTHREADS; ++t) tl.store(t, l, st[t]);
        }
    }
}
template <typename T, int KIND> static void fused_host(const FusedParams<T>& P, int ncta) {
    switch ((P.first ? 1 : 0) | (P.last ? 2 : 0)) {
        case 0: fused_host_t<T, KIND, 0>(P, ncta); break;
        case 1: fused_host_t<T, KIND, 1>(P, ncta); break;
        case 2: fused_host_t<T, KIND, 2>(P, ncta); break;
        default: fused_host_t<T, KIND, 3>(P, ncta); break;
    }
}
#else
// ---- tensor maps (TMA descriptors) of whole-tile boxes -------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc) && alignof(CUtensorMap) <= alignof(TmaDesc), "TmaDesc mirrors CUtensorMap");

static EncodeTiledFn tensor_map_encoder() {
    static const EncodeTiledFn fn = [] {
        if (getenv("GCMF_NO_TMAP")) return (EncodeTiledFn) nullptr;  // A/B and test knob: row-wise bulk copies everywhere
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// (nx, ny[, nb]) view of an array with boxes of one tile (tw x th x 1).  False when the array cannot be described
// (the kernel then stages that tile row by row, as it does next to the periodic boundaries).
template <typename T>
static bool encode_tile_map(TmaDesc* out, const void* base, int nx, int ny, int64_t nb, int64_t pitch, int64_t bstride,
                            int tw, int th) {
    const EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    if (nb > 1 && bstride < (int64_t)ny * pitch) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)(nb > 0 ? nb : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(T),
                                   (cuuint64_t)(nb > 1 ? bstride : (int64_t)ny * pitch) * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)th, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const cuuint32_t rank = nb > 0 ? 3u : 2u;
    return enc(reinterpret_cast<CUtensorMap*>(out), dt, rank, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T, int KIND>
static const TmaDesc* state_map(gcmf_plan* pl, const FieldRef<const T>& f, int64_t nb) {
    using G = FusedGeom<T, FusedSplit<KIND>::value>;
    for (auto& m : pl->state_maps)
        if (m.p == f.p && m.pitch == f.pitch && m.bstride == f.bstride && m.nb == nb) return &m.d;
    gcmf_plan::MapEntry e{f.p, f.pitch, f.bstride, nb, {}};
    if (!encode_tile_map<T>(&e.d, f.p, pl->desc.nx, pl->desc.ny, nb, f.pitch, f.bstride, G::TW, G::TH)) return nullptr;
    if (pl->state_maps.size() >= 8) pl->state_maps.erase(pl->state_maps.begin());
    pl->state_maps.push_back(e);
    return &pl->state_maps.back().d;
}

// Fill the launch's descriptor block.  Only whole (periodic / tripolar) grids use tensor maps; inside them only the
// tiles that touch no boundary (91 % of the tiles of a 2400 x 3600 grid) -- the kernel decides per tile.
template <typename T, int KIND> static void make_maps(gcmf_plan* pl, const FusedParams<T>& P, FusedMaps& M) {
    using G = FusedGeom<T, FusedSplit<KIND>::value>;
    memset(&M, 0, sizeof M);
    if (!(P.g.flags & FL_WRAP_Y) || !tensor_map_encoder()) return;
    const TmaDesc* d1 = state_map<T, KIND>(pl, P.t1_in, P.nb);
    const TmaDesc* d2 = P.first ? d1 : state_map<T, KIND>(pl, P.t2_in, P.nb);
    d1 = state_map<T, KIND>(pl, P.t1_in, P.nb);  // (the second look-up may have evicted / moved the first entry)
    if (d1 && d2) {
        M.t1 = *d1;
        M.t2 = *d2;
        M.use |= TMAP_STATE;
    }
    if (KIND == FK_FLUX) {
        bool ok = true;
        for (int s = 0; s < 3 && ok; ++s) {
            gcmf_plan::MapEntry& e = pl->coef_maps[s];
            if (e.p != P.plane[s].p || e.pitch != P.plane[s].pitch) {
                e.p = nullptr;
                ok = encode_tile_map<T>(&e.d, P.plane[s].p, pl->desc.nx, pl->desc.ny, 0, P.plane[s].pitch, 0, G::TW, G::TH);
                if (ok) {
                    e.p = P.plane[s].p;
                    e.pitch = P.plane[s].pitch;
                }
            }
            if (ok) M.coef[s] = e.d;
        }
        if (ok) M.use |= TMAP_COEF;
    }
}

template <typename T, int KIND, int EDGE>
static int launch_fused_kernel_t(gcmf_plan* pl, const FusedParams<T>& P, int64_t ncta, cudaStream_t st) {
    using G = FusedGeom<T, FusedSplit<KIND>::value>;
    // the opt-in to > 48 KB of dynamic shared memory is a per-device property of the function: a process that
    // filters on several devices (tensors on cuda:0 and cuda:1) has to set it on each of them
    static bool attr_done[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(fused_kernel<T, KIND, EDGE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)G::smem_bytes(KIND)));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    FusedMaps M;
    make_maps<T, KIND>(pl, P, M);
    fused_kernel<T, KIND, EDGE><<<(unsigned)ncta, G::NTHREADS, G::smem_bytes(KIND), st>>>(P, M);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
}
template <typename T, int KIND>
static int launch_fused_kernel(gcmf_plan* pl, const FusedParams<T>& P, int64_t ncta, cudaStream_t st) {
    switch ((P.first ? 1 : 0) | (P.last ? 2 : 0)) {
        case 0: return launch_fused_kernel_t<T, KIND, 0>(pl, P, ncta, st);
        case 1: return launch_fused_kernel_t<T, KIND, 1>(pl, P, ncta, st);
        case 2: return launch_fused_kernel_t<T, KIND, 2>(pl, P, ncta, st);
    }
    return launch_fused_kernel_t<T, KIND, 3>(pl, P, ncta, st);
}
#endif

#ifndef GCMF_HOSTEMU
// ---- row-streaming form of the fused FLUX steps (gcmf_march.cuh) ----------------------------------------------
// The DEFAULT for fp64 FLUX plans on whole doubly periodic grids (cfg3) since its consumer keeps the row windows in
// row-keyed register slots (gcmf_march.cuh): measured on a B200 against the tile form (profiles/variants_r02c_march.log,
// variants_r02d_march_lv2.log, ncu_r02d_march_final_cfg3.*): cfg3 nb = 62: 72.7 vs 100.9 ms per filter call (324 vs 234 G
// units/s), nb = 8: 10.6 vs 13.9 ms.  The tile form keeps the tripolar grids (the fold), the latitude bands, grids
// narrower than one strip, single-level launches and fp32 (63.8 vs 77.8 ms on the cfg3 shape: twice the columns per tile
// row).  GCMF_FUSED_FORM=tile|march forces one form where both are eligible (A/B, tests).
template <typename T> static bool march_eligible(const gcmf_plan* pl, int64_t nb) {
    static const char* force = getenv("GCMF_FUSED_FORM");
    if (force && !strcmp(force, "tile")) return false;
    if (pl->desc.op != GCMF_OP_FLUX) return false;
    if (pl->desc.flags & (GCMF_FLAG_FOLD_N | GCMF_FLAG_CUT_S)) return false;
    if (pl->desc.nx < MARCH_W) return false;
    if (force && !strcmp(force, "march")) return true;  // incl. latitude bands and fp32
    // a CTA marches MARCH_LV levels at once: a single level leaves half of its level slots empty (with three levels per
    // CTA: 1, 2 and 4 levels)
    if (sizeof(T) != 8 || !(pl->desc.flags & GCMF_FLAG_WRAP_Y)) return false;
    return MARCH_LV == 2 ? nb >= 2 : (nb >= 3 && nb != 4);
}

template <typename T, int EDGE, int K> static int launch_march_t(const gcmf_plan* pl, const FusedParams<T>& P, cudaStream_t st) {
    using G = MarchGeom<T>;
    static bool attr_done[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(march_kernel<T, EDGE, K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)G::smem_bytes()));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int nstrips = (P.g.nx + G::SW - 1) / G::SW;
    const int nlg = (int)((P.nb + MARCH_LV - 1) / MARCH_LV);
    // Rows per band (the sweeps below: three levels per CTA): 3(K-1) priming iterations per band against enough CTAs for nine rounds of the device (the tail of
    // a launch is one CTA long): the tallest band of at most 400 rows that leaves that many.  Measured on cfg3, ms per
    // filter call: nb = 62 (630 CTAs per band): 200 rows 82.3, 400: 80.5, 800: 81.6 (before the steady-state
    // iterations: 60: 95.5, 120: 89.1, 240: 85.4, 400: 84.7, 600: 84.9, 1200: 86.9, 2400: 98.8); nb = 8 (90 CTAs per
    // band): 50 rows 13.5, 75: 12.7, 100: 12.3, 150: 12.1, 200: 12.5, 300: 12.5.
    int ry = 400;
    {
        const char* e = getenv("GCMF_MARCH_ROWS");
        const int v = e ? atoi(e) : 0;
        if (v > 0) {
            ry = v;
        } else {
            const int64_t per_band = (int64_t)nstrips * nlg;
            while (ry > 24 && per_band * ((P.g.ny + ry - 1) / ry) < 9 * (int64_t)pl->sm_count) --ry;
        }
    }
    const int64_t nbands = (P.g.ny + ry - 1) / ry;
    const int64_t ncta = (int64_t)nstrips * nlg * nbands;
    if (ncta > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "fused step: grid too large");
    march_kernel<T, EDGE, K><<<(unsigned)ncta, G::NTHREADS, G::smem_bytes(), st>>>(P, nstrips, nlg, ry);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
}
template <typename T, int EDGE> static int launch_march_e(const gcmf_plan* pl, const FusedParams<T>& P, cudaStream_t st) {
    switch (P.k) {
        case 1: return launch_march_t<T, EDGE, 1>(pl, P, st);
        case 2: return launch_march_t<T, EDGE, 2>(pl, P, st);
        case 3: return launch_march_t<T, EDGE, 3>(pl, P, st);
    }
    return launch_march_t<T, EDGE, 4>(pl, P, st);
}
template <typename T> static int launch_march(const gcmf_plan* pl, const FusedParams<T>& P, cudaStream_t st) {
    switch ((P.first ? 1 : 0) | (P.last ? 2 : 0)) {
        case 0: return launch_march_e<T, 0>(pl, P, st);
        case 1: return launch_march_e<T, 1>(pl, P, st);
        case 2: return launch_march_e<T, 2>(pl, P, st);
    }
    return launch_march_e<T, 3>(pl, P, st);
}
#endif

// ---- vector operators: two Chebyshev steps per launch (gcmf_vec2.cuh) -------------------------------------------
// Eligible: VECTOR_C / VECTOR_B on whole doubly periodic grids (gcmf_filter) and on latitude bands (gcmf_cheb_fused on a
// band plan: the caller keeps TWO ghost rows of the fields and of every plane on each side and refreshes those of
// T_{i+1} and T_i after every block), rows that split into 16-byte vectors, at least one strip wide, every coefficient
// plane 16-byte aligned.  Forcing a one-step C-grid kernel form with GCMF_CGRID_KERNEL (A/B, tests) switches it off.
template <typename T> static bool cg2_eligible_t(const gcmf_plan* p) {
#ifdef GCMF_HOSTEMU
    // mbarriers + TMA are device-only: the emulator runs a block as its two one-step launches (bit-identical by
    // construction), so that the CPU suite covers the blocked control flow of gcmf_filter and of the band scheduler
    return (p->desc.op == GCMF_OP_VECTOR_C || p->desc.op == GCMF_OP_VECTOR_B) && p->desc.ny >= 4;
#else
    using G = Cg2Geom<T>;
    static const bool forced = getenv("GCMF_CGRID_KERNEL") != nullptr;
    if (p->desc.op != GCMF_OP_VECTOR_C && p->desc.op != GCMF_OP_VECTOR_B) return false;
    if (forced) return false;
    if (p->desc.nx % G::AV || p->desc.nx < G::LW || p->desc.ny < 4) return false;
    for (int s = 0; s < p->n_planes; ++s)
        if (!p->plane[s].p || !aligned(p->plane[s].p, p->plane[s].pitch, p->plane[s].nb > 1 ? p->plane[s].bstride : 0,
                                       G::AV, sizeof(T)))
            return false;
    return true;
#endif
}
static bool cg2_eligible(const gcmf_plan* p) {
    return p->desc.dtype == GCMF_F64 ? cg2_eligible_t<double>(p) : cg2_eligible_t<float>(p);
}

#ifndef GCMF_HOSTEMU
template <typename T, template <typename, int> class OPT, int EDGE, bool HALO>
static int launch_cg2_t(const gcmf_plan* pl, const Cg2Params<T>& P, cudaStream_t st) {
    using G = Vec2Geom<T, OPT<T, 1>::NC>;
    static bool attr_done[64] = {false};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(vec2_kernel<T, OPT, EDGE, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)G::smem_bytes()));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int ny = P.g.ny;
    const int64_t ctas_x = (P.g.nx + CG2_WARPS * P.cpw - 1) / (CG2_WARPS * P.cpw);
    // Rows per band.  A band stages 4 priming rows (and pays ~2 rows of prologue) on top of its own, so bands should be
    // tall; but one CTA per SM runs at a time, so the launch takes ceil(CTAs / SMs) rounds of the longest CTA: pick the
    // height that minimises rounds x (rows + 6) among heights up to 64 (measured on cfg5, ms per launch: 24 rows 0.409,
    // 36: 0.404, 48: 0.410, 59 (exactly 5 rounds): 0.400, 72: 0.436, 99 (3 rounds, the model's optimum without the
    // cap): 0.433, 180: 0.478 -- few long CTAs end raggedly).  GCMF_CGRID_ROWS overrides (tuning, tests).
    int ry = 0;
    if (const char* e = getenv("GCMF_CGRID_ROWS")) ry = atoi(e);
    if (ry <= 0) {
        double best = 1e300;
        const int lo = ny < 12 ? ny : 12, hi = ny < 64 ? ny : 64;
        for (int cand = lo; cand <= hi; ++cand) {
            // with the exchange fused in, the first and the last row-band own the two border rows they push
            if (HALO && (cand < 2 || ny % cand == 1)) continue;
            const int64_t ncta = ctas_x * P.nb * ((ny + cand - 1) / cand);
            const int64_t rounds = (ncta + pl->sm_count - 1) / pl->sm_count;
            const double cost = (double)rounds * (cand + 6);
            if (cost < best) {
                best = cost;
                ry = cand;
            }
        }
    }
    if (ry <= 0 || ry > ny) ry = ny;
    if (HALO)
        while (ry < ny && (ry < 2 || ny % ry == 1)) ++ry;
    const int64_t nbands = (ny + ry - 1) / ry;
    const int64_t nblk = ctas_x * P.nb * nbands;
    if (nblk > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "grid too large (%lld blocks)", (long long)nblk);
    vec2_kernel<T, OPT, EDGE, HALO><<<(unsigned)nblk, G::NTHREADS, G::smem_bytes(), st>>>(P, (unsigned)ctas_x, ry);
    gcmf_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return GCMF_OK;
}
template <typename T, template <typename, int> class OPT, bool HALO>
static int launch_cg2_h(const gcmf_plan* pl, const Cg2Params<T>& P, bool first, bool last, cudaStream_t st) {
    switch ((first ? 1 : 0) | (last ? 2 : 0)) {
        case 0: return launch_cg2_t<T, OPT, 0, HALO>(pl, P, st);
        case 1: return launch_cg2_t<T, OPT, 1, HALO>(pl, P, st);
        case 2: return launch_cg2_t<T, OPT, 2, HALO>(pl, P, st);
    }
    return launch_cg2_t<T, OPT, 3, HALO>(pl, P, st);
}
template <typename T, template <typename, int> class OPT>
static int launch_cg2(const gcmf_plan* pl, const Cg2Params<T>& P, bool first, bool last, cudaStream_t st) {
    if (P.halo[0].enabled) return launch_cg2_h<T, OPT, true>(pl, P, first, last, st);
    return launch_cg2_h<T, OPT, false>(pl, P, first, last, st);
}
#endif

// steps step0 and step0+1 of a vector plan in one launch
template <typename T>
static int run_cg2_t(const gcmf_plan* pl, int64_t nb, int step0, const gcmf_field* t1, const gcmf_field* t2,
                     const gcmf_field* t1o, const gcmf_field* t2o, const gcmf_field* bar, cudaStream_t st,
                     const gcmf_halo* halo1 = nullptr, const gcmf_halo* halo2 = nullptr) {
#ifdef GCMF_HOSTEMU
    // T_i passes through t2_out, or through a scratch copy when the block ends the recurrence (no T is stored then)
    const bool first = step0 == 1, last = step0 + 1 == pl->n_steps;
    const int64_t plane = (int64_t)pl->desc.ny * pl->desc.nx;
    std::vector<T> scratch;
    gcmf_field tmp[2];
    const gcmf_field* ti = t2o;
    if (last) {
        scratch.resize((size_t)(2 * nb * plane));
        for (int k = 0; k < 2; ++k) tmp[k] = gcmf_field{scratch.data() + (size_t)k * nb * plane, pl->desc.nx, plane};
        ti = tmp;
    }
    if (!(pl->desc.flags & GCMF_FLAG_WRAP_Y)) {
        // Latitude band (two ghost rows per side present in memory): step i+1 needs T_i on rows -1 .. ny too, so step i
        // runs on the band extended by one row on either side -- a copy of the plan with ny + 2 rows whose planes and
        // fields start one row further south -- into scratch arrays with room for those rows.
        const int ny = pl->desc.ny, nx = pl->desc.nx;
        const int64_t eplane = (int64_t)(ny + 2) * nx;
        std::vector<T> Ti((size_t)(2 * nb * eplane)), Be((size_t)(2 * nb * eplane));
        gcmf_plan ext = *pl;
        ext.desc.ny = ny + 2;
        for (int s2 = 0; s2 < ext.n_planes; ++s2)
            if (ext.plane[s2].p) ext.plane[s2].p = (const T*)ext.plane[s2].p - ext.plane[s2].pitch;
        gcmf_field e1[2], e2[2], tif[2], bef[2], ti1[2];
        for (int k = 0; k < 2; ++k) {
            e1[k] = gcmf_field{(T*)t1[k].ptr - t1[k].pitch, t1[k].pitch, t1[k].bstride};
            if (!first) e2[k] = gcmf_field{(T*)t2[k].ptr - t2[k].pitch, t2[k].pitch, t2[k].bstride};
            tif[k] = gcmf_field{Ti.data() + (size_t)k * nb * eplane, nx, eplane};
            bef[k] = gcmf_field{Be.data() + (size_t)k * nb * eplane, nx, eplane};
            ti1[k] = gcmf_field{Ti.data() + (size_t)k * nb * eplane + nx, nx, eplane};  // row 0 of the band
            if (!first)
                for (int64_t b = 0; b < nb; ++b)
                    for (int j = 0; j < ny; ++j)
                        memcpy((T*)bef[k].ptr + b * eplane + (int64_t)(j + 1) * nx,
                               (const T*)bar[k].ptr + b * bar[k].bstride + (int64_t)j * bar[k].pitch, (size_t)nx * sizeof(T));
        }
        int rc = run_step_t<T>(&ext, nb, first ? MODE_FIRST : MODE_MID, e1, first ? nullptr : e2, tif, bef, pl->p[0], pl->p[step0], st);
        if (rc != GCMF_OK) return rc;
        for (int k = 0; k < 2; ++k)
            for (int64_t b = 0; b < nb; ++b)
                for (int j = 0; j < ny; ++j) {
                    memcpy((T*)bar[k].ptr + b * bar[k].bstride + (int64_t)j * bar[k].pitch,
                           (const T*)bef[k].ptr + b * eplane + (int64_t)(j + 1) * nx, (size_t)nx * sizeof(T));
                    if (!last)
                        memcpy((T*)t2o[k].ptr + b * t2o[k].bstride + (int64_t)j * t2o[k].pitch,
                               (const T*)ti1[k].ptr + b * eplane + (int64_t)j * nx, (size_t)nx * sizeof(T));
                }
        rc = run_step_t<T>(pl, nb, last ? MODE_LAST : MODE_MID, ti1, t1, last ? nullptr : t1o, bar, pl->p[0], pl->p[step0 + 1], st);
        gcmf_count_launch(-1);  // one launch on the device
        if (rc == GCMF_OK && halo1) {
            // the exchange fused into the device kernel: the emulator has no concurrency, so the border rows are copied
            // into the "neighbours'" ghost rows after the block and the flags raised (as for gcmf_cheb_step_halo)
            const gcmf_halo* hs[2] = {halo1, halo2 ? halo2 : halo1};
            const gcmf_field* outs[2] = {t1o, t2o};
            for (int a = 0; a < 2 && !last; ++a)
                for (int k = 0; k < 2; ++k)
                    for (int64_t b = 0; b < nb; ++b)
                        for (int r = 0; r < 2; ++r) {
                            const T* src_s = (const T*)outs[a][k].ptr + b * outs[a][k].bstride + (int64_t)r * outs[a][k].pitch;
                            const T* src_n = (const T*)outs[a][k].ptr + b * outs[a][k].bstride + (int64_t)(ny - 2 + r) * outs[a][k].pitch;
                            if (hs[a]->south_ghost[k])
                                memcpy((T*)hs[a]->south_ghost[k] + b * hs[a]->south_bstride + (int64_t)r * outs[a][k].pitch, src_s, (size_t)nx * sizeof(T));
                            if (hs[a]->north_ghost[k])
                                memcpy((T*)hs[a]->north_ghost[k] + b * hs[a]->north_bstride + (int64_t)r * outs[a][k].pitch, src_n, (size_t)nx * sizeof(T));
                        }
            if (halo1->signal_north) *halo1->signal_north = halo1->signal_value;
            if (halo1->signal_south) *halo1->signal_south = halo1->signal_value;
        }
        return rc;
    }
    if (halo1) return gcmf_set_error(GCMF_EINVAL, "halo exchange needs a band plan (no GCMF_FLAG_WRAP_Y)");
    int rc = run_step_t<T>(pl, nb, first ? MODE_FIRST : MODE_MID, t1, first ? nullptr : t2, ti, bar, pl->p[0], pl->p[step0], st);
    if (rc != GCMF_OK) return rc;
    rc = run_step_t<T>(pl, nb, last ? MODE_LAST : MODE_MID, ti, t1, last ? nullptr : t1o, bar, pl->p[0], pl->p[step0 + 1], st);
    gcmf_count_launch(-1);  // one launch on the device
    return rc;
#else
    using G = Cg2Geom<T>;
    const bool first = step0 == 1, last = step0 + 1 == pl->n_steps;
    const gcmf_field* all[5] = {t1, first ? nullptr : t2, last ? nullptr : t1o, last ? nullptr : t2o, bar};
    for (const gcmf_field* f : all)
        for (int k = 0; f && k < 2; ++k)
            if (!aligned(f[k].ptr, f[k].pitch, f[k].bstride, G::AV, sizeof(T)))
                return gcmf_set_error(GCMF_EINVAL, "fused step: fields must be 16-byte aligned with vector-multiple strides");
    for (int k = 0; k < 2; ++k) {
        if (bar[k].ptr == t1[k].ptr || (!first && bar[k].ptr == t2[k].ptr))
            return gcmf_set_error(GCMF_EINVAL, "fused step: bar must not alias the inputs");
        if (!last)
            for (int m = 0; m < 2; ++m)  // neighbouring strips / bands read the inputs while this one stores
                if (t1o[k].ptr == t1[m].ptr || t2o[k].ptr == t1[m].ptr ||
                    (!first && (t1o[k].ptr == t2[m].ptr || t2o[k].ptr == t2[m].ptr)))
                    return gcmf_set_error(GCMF_EINVAL, "fused step: outputs must not alias inputs");
    }
    Cg2Params<T> P;
    memset(&P, 0, sizeof P);
    P.g.ny = pl->desc.ny;
    P.g.nx = pl->desc.nx;
    P.g.flags = pl->desc.flags;
    for (int s = 0; s < pl->n_planes; ++s) P.plane[s] = pl->plane[s];
    for (int k = 0; k < 2; ++k) {
        P.t1[k] = FieldRef<const T>{(const T*)t1[k].ptr, t1[k].pitch, t1[k].bstride};
        if (!first) P.t2[k] = FieldRef<const T>{(const T*)t2[k].ptr, t2[k].pitch, t2[k].bstride};
        if (!last) {
            P.t1o[k] = FieldRef<T>{(T*)t1o[k].ptr, t1o[k].pitch, t1o[k].bstride};
            P.t2o[k] = FieldRef<T>{(T*)t2o[k].ptr, t2o[k].pitch, t2o[k].bstride};
        }
        P.bar[k] = FieldRef<T>{(T*)bar[k].ptr, bar[k].pitch, bar[k].bstride};
    }
    if (halo1) {
        P.halo[0] = make_halo<T>(halo1);
        P.halo[1] = make_halo<T>(halo2 ? halo2 : halo1);
    }
    P.c = pl->c;
    P.p0 = pl->p[0];
    P.pa = pl->p[step0];
    P.pb = pl->p[step0 + 1];
    P.nb = nb;
    // strips of equal width: the fewest CTAs per row that 28 columns per warp allow, then the columns spread evenly
    const int strip_max = CG2_WARPS * CG2_COLS;
    const int ctas_x = (pl->desc.nx + strip_max - 1) / strip_max;
    P.cpw = (pl->desc.nx + ctas_x * CG2_WARPS - 1) / (ctas_x * CG2_WARPS);
    P.lw = CG2_WARPS * P.cpw + 2 * G::HALO;
    if (pl->desc.op == GCMF_OP_VECTOR_B) return launch_cg2<T, BgOp>(pl, P, first, last, st);
    return launch_cg2<T, CgOp>(pl, P, first, last, st);
#endif
}

template <typename T>
static int run_fused_t(gcmf_plan* pl, int64_t nb, int step0, int k, const gcmf_field* t1, const gcmf_field* t2,
                       const gcmf_field* t1o, const gcmf_field* t2o, const gcmf_field* bar, cudaStream_t st) {
    using G = FusedGeom<T>;
    const int kind = fused_kind_t<T>(pl);
    const bool first = step0 == 1, last = step0 + k - 1 == pl->n_steps;
    const gcmf_field* all[5] = {t1, first ? nullptr : t2, last ? nullptr : t1o, last ? nullptr : t2o, bar};
    for (const gcmf_field* f : all)
        if (f && !aligned(f->ptr, f->pitch, f->bstride, G::AV, sizeof(T)))
            return gcmf_set_error(GCMF_EINVAL, "fused step: fields must be 16-byte aligned with vector-multiple strides");
    if (!last && (t1o->ptr == t1->ptr || t2o->ptr == t1->ptr || (!first && (t1o->ptr == t2->ptr || t2o->ptr == t2->ptr))))
        return gcmf_set_error(GCMF_EINVAL, "fused step: outputs must not alias inputs (neighbouring tiles read them)");
    if (bar->ptr == t1->ptr) return gcmf_set_error(GCMF_EINVAL, "fused step: bar must not alias the input");
    if (last && (pl->desc.flags & GCMF_FLAG_AREA) &&
        (pl->plane[1].nb != 1 || !aligned(pl->plane[1].p, pl->plane[1].pitch, 0, G::AV, sizeof(T))))
        return gcmf_set_error(GCMF_EINVAL, "fused step: the area plane must be a shared, vector-aligned 2-D plane");
    FusedParams<T> P;
    memset(&P, 0, sizeof P);
    P.g.ny = pl->desc.ny;
    P.g.nx = pl->desc.nx;
    P.g.flags = pl->desc.flags;
    for (int s = 0; s < 3; ++s) P.plane[s] = pl->plane[s];
    P.t1_in = FieldRef<const T>{(const T*)t1->ptr, t1->pitch, t1->bstride};
    if (!first) P.t2_in = FieldRef<const T>{(const T*)t2->ptr, t2->pitch, t2->bstride};
    if (!last) {
        P.t1_out = FieldRef<T>{(T*)t1o->ptr, t1o->pitch, t1o->bstride};
        P.t2_out = FieldRef<T>{(T*)t2o->ptr, t2o->pitch, t2o->bstride};
    }
    P.first = first;
    P.last = last;
    P.p0 = pl->p[0];
    P.bar = FieldRef<T>{(T*)bar->ptr, bar->pitch, bar->bstride};
    P.c = pl->c;
    for (int s = 0; s < k; ++s) P.p[s] = pl->p[step0 + s];
    P.k = k;
    P.ncx = (pl->desc.nx + G::CW - 1) / G::CW;
    P.ncy = (pl->desc.ny + G::CH - 1) / G::CH;
    P.nb = nb;
    // Level slabs.  One CTA keeps its coefficient tiles for a whole slab of levels and pays its prologue (barrier
    // set-up, coefficient staging, the first un-overlapped tile load: about the cost of one level, c = 1) once per
    // slab, so slabs should be long; but with dynamic CTA dispatch the tail of the launch is about one CTA
    // duration, so they should not be too long either.  time ~ ntiles*(nb + g*c)/SMs + (nb/g + c) is minimal at
    // g = sqrt(nb*SMs / (ntiles*c)) slabs per tile.  (Measured on cfg3: 1 slab 10.7 ms, 2 slabs 10.4 ms per launch at
    // nb = 62; chunks of 7 levels from the host pipeline want a single slab.)  GCMF_FUSED_GROUPS overrides.
    const int64_t ntiles = (int64_t)P.ncx * P.ncy;
    int64_t groups = (int64_t)(sqrt((double)nb * pl->sm_count / (double)ntiles) + 0.5);
    if (const char* e = getenv("GCMF_FUSED_GROUPS")) {
        const long long v = atoll(e);
        if (v > 0) groups = v;
    }
    if (groups < 1) groups = 1;
    if (groups > nb) groups = nb;
    P.levels_per_cta = (int32_t)((nb + groups - 1) / groups);
    if (const char* e = getenv("GCMF_FUSED_LEVELS_PER_CTA")) {  // test / tuning knob: force the slab length
        const long v = atol(e);
        if (v > 0) P.levels_per_cta = (int32_t)(v < nb ? v : nb);
    }
    groups = (nb + P.levels_per_cta - 1) / P.levels_per_cta;
    const int64_t ncta = ntiles * groups;
    if (ncta > 0x7fffffffLL) return gcmf_set_error(GCMF_EINVAL, "fused step: grid too large");
#ifdef GCMF_HOSTEMU
    (void)st;
    if (kind == FK_FLUX) fused_host<T, FK_FLUX>(P, (int)ncta);
    else fused_host<T, FK_REG5>(P, (int)ncta);
    gcmf_count_launch(1);
    return GCMF_OK;
#else
    if (kind == FK_FLUX && march_eligible<T>(pl, nb)) return launch_march<T>(pl, P, st);
    if (kind == FK_FLUX) return launch_fused_kernel<T, FK_FLUX>(pl, P, ncta, st);
    return launch_fused_kernel<T, FK_REG5>(pl, P, ncta, st);
#endif
}

extern "C" int gcmf_fused_max_steps(const gcmf_plan* p) {
    if (!p) return 0;
    if (cg2_eligible(p)) return 2;
    return fused_eligible(p) ? FusedGeom<double>::H : 0;
}

extern "C" int gcmf_cheb_fused(gcmf_plan* p, int64_t nb, int32_t step, int32_t k, const gcmf_field* t1_in,
                               const gcmf_field* t2_in, const gcmf_field* t1_out, const gcmf_field* t2_out,
                               const gcmf_field* bar, void* stream) {
    if (!p || nb < 1) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    if (p->n_steps < 2) return gcmf_set_error(GCMF_ESTATE, "gcmf_plan_set_filter has not been called");
    if (cg2_eligible(p)) {
        // vector operators: exactly two steps per launch; a trailing single step (odd n_steps) is the one-step LAST kernel
        if (k == 1 && step == p->n_steps) return gcmf_cheb_step(p, nb, step, t1_in, t2_in, t1_out, bar, stream);
        if (k != 2) return gcmf_set_error(GCMF_EINVAL, "vector plans fuse exactly 2 steps (k = %d)", k);
        if (step < 1 || step + 1 > p->n_steps)
            return gcmf_set_error(GCMF_EINVAL, "fused steps %d..%d must lie inside 1..%d", step, step + 1, p->n_steps);
        TRY(check_planes(p));
        TRY(check_fields(p, t1_in, "t1_in"));
        if (step > 1) TRY(check_fields(p, t2_in, "t2_in"));
        if (step + 1 < p->n_steps) {
            TRY(check_fields(p, t1_out, "t1_out"));
            TRY(check_fields(p, t2_out, "t2_out"));
        }
        TRY(check_fields(p, bar, "bar"));
        CUDA_TRY(cudaSetDevice(p->desc.device));
        if (p->desc.dtype == GCMF_F64)
            return run_cg2_t<double>(p, nb, step, t1_in, t2_in, t1_out, t2_out, bar, (cudaStream_t)stream);
        return run_cg2_t<float>(p, nb, step, t1_in, t2_in, t1_out, t2_out, bar, (cudaStream_t)stream);
    }
    if (!fused_eligible(p)) return gcmf_set_error(GCMF_EINVAL, "this plan has no fused path (see gcmf_fused_max_steps)");
    if (k < 1 || k > FusedGeom<double>::H) return gcmf_set_error(GCMF_EINVAL, "k = %d outside 1..%d", k, FusedGeom<double>::H);
    if (step < 1 || step + k - 1 > p->n_steps)
        return gcmf_set_error(GCMF_EINVAL, "fused steps %d..%d must lie inside 1..%d", step, step + k - 1, p->n_steps);
    TRY(check_planes(p));
    TRY(check_fields(p, t1_in, "t1_in"));
    if (step > 1) TRY(check_fields(p, t2_in, "t2_in"));
    if (step + k - 1 < p->n_steps) {
        TRY(check_fields(p, t1_out, "t1_out"));
        TRY(check_fields(p, t2_out, "t2_out"));
    }
    TRY(check_fields(p, bar, "bar"));
    CUDA_TRY(cudaSetDevice(p->desc.device));
    if (p->desc.dtype == GCMF_F64)
        return run_fused_t<double>(p, nb, step, k, t1_in, t2_in, t1_out, t2_out, bar, (cudaStream_t)stream);
    return run_fused_t<float>(p, nb, step, k, t1_in, t2_in, t1_out, t2_out, bar, (cudaStream_t)stream);
}

// gcmf_cheb_fused on a latitude band with the ghost-row exchange fused into the kernel (vector operators, k = 2)
extern "C" int gcmf_cheb_fused_halo(gcmf_plan* p, int64_t nb, int32_t step, int32_t k, const gcmf_field* t1_in,
                                    const gcmf_field* t2_in, const gcmf_field* t1_out, const gcmf_field* t2_out,
                                    const gcmf_field* bar, const gcmf_halo* halo_t1, const gcmf_halo* halo_t2, void* stream) {
    if (!p || nb < 1 || !halo_t1 || !halo_t2) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    if (p->desc.flags & GCMF_FLAG_WRAP_Y) return gcmf_set_error(GCMF_EINVAL, "halo exchange needs a band plan (no GCMF_FLAG_WRAP_Y)");
    if (p->n_steps < 2) return gcmf_set_error(GCMF_ESTATE, "gcmf_plan_set_filter has not been called");
    if (!cg2_eligible(p))
        return gcmf_set_error(GCMF_EINVAL, "gcmf_cheb_fused_halo: vector plans with a two-step kernel only (see gcmf_fused_max_steps)");
    if (k != 2 || step < 1 || step + 1 > p->n_steps)
        return gcmf_set_error(GCMF_EINVAL, "gcmf_cheb_fused_halo: k must be 2 and steps %d..%d inside 1..%d", step, step + 1, p->n_steps);
    if (p->desc.ny < 4) return gcmf_set_error(GCMF_EINVAL, "gcmf_cheb_fused_halo: a band needs at least 4 rows");
    if (halo_t1->wait_north != halo_t2->wait_north || halo_t1->wait_south != halo_t2->wait_south ||
        halo_t1->north_bstride != halo_t2->north_bstride || halo_t1->south_bstride != halo_t2->south_bstride)
        return gcmf_set_error(GCMF_EINVAL, "gcmf_cheb_fused_halo: the two halos must share flags and batch strides");
    TRY(check_planes(p));
    TRY(check_fields(p, t1_in, "t1_in"));
    if (step > 1) TRY(check_fields(p, t2_in, "t2_in"));
    if (step + 1 < p->n_steps) {
        TRY(check_fields(p, t1_out, "t1_out"));
        TRY(check_fields(p, t2_out, "t2_out"));
    }
    TRY(check_fields(p, bar, "bar"));
    CUDA_TRY(cudaSetDevice(p->desc.device));
    if (p->desc.dtype == GCMF_F64)
        return run_cg2_t<double>(p, nb, step, t1_in, t2_in, t1_out, t2_out, bar, (cudaStream_t)stream, halo_t1, halo_t2);
    return run_cg2_t<float>(p, nb, step, t1_in, t2_in, t1_out, t2_out, bar, (cudaStream_t)stream, halo_t1, halo_t2);
}

extern "C" int gcmf_filter(gcmf_plan* p, int64_t nb, const gcmf_field* in, const gcmf_field* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
    if (!p || nb < 1) return gcmf_set_error(GCMF_EINVAL, "bad argument");
    if (p->n_steps < 2) return gcmf_set_error(GCMF_ESTATE, "gcmf_plan_set_filter has not been called");
    TRY(check_planes(p));
    TRY(check_fields(p, in, "in"));
    TRY(check_fields(p, out, "out"));
    size_t need = 0;
    TRY(gcmf_workspace_bytes(p, nb, &need));
    if (!workspace || workspace_bytes < need)
        return gcmf_set_error(GCMF_EINVAL, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
    if ((uintptr_t)workspace % 256) return gcmf_set_error(GCMF_EINVAL, "workspace must be 256-byte aligned");
    for (int k = 0; k < p->ncomp; ++k)
        if (in[k].ptr == out[k].ptr) return gcmf_set_error(GCMF_EINVAL, "out must not alias in");
    const int nc = p->ncomp;
    const int64_t pitch = p->desc.nx, bs = (int64_t)p->desc.ny * p->desc.nx;
    const size_t bb = buffer_bytes(p, nb);
    gcmf_field A[2], B[2], X[2];
    for (int k = 0; k < nc; ++k) {
        A[k] = gcmf_field{(char*)workspace + (size_t)k * bb, pitch, bs};
        B[k] = gcmf_field{(char*)workspace + (size_t)(nc + k) * bb, pitch, bs};
        X[k] = in[k];
    }
    const bool area = (p->desc.flags & GCMF_FLAG_AREA) != 0;
    if (area) {  // x = f * area, materialised once in B (kernels.py:100-101)
        TRY(gcmf_prepare(p, nb, in, B, stream));
        for (int k = 0; k < nc; ++k) X[k] = B[k];
    }
    const int n = p->n_steps;
    if (is_band_plan(p))
        return gcmf_set_error(GCMF_EINVAL, "gcmf_filter runs whole (periodic / tripolar) grids; drive a latitude band with "
                                           "gcmf_cheb_step / gcmf_cheb_fused and exchange its ghost rows between calls");
    // The fused kernels move whole 16-byte vectors (bulk copies, vector stores): caller-provided fields that are not
    // 16-byte aligned with vector-multiple strides take the one-step kernels (which have a scalar form) instead.
    const size_t es = p->desc.dtype == GCMF_F64 ? 8 : 4;
    bool fused_ok = plan_uses_fused(p);
    for (int k = 0; fused_ok && k < nc; ++k)
        fused_ok = aligned(out[k].ptr, out[k].pitch, out[k].bstride, (int)(16 / es), es) &&
                   (area || aligned(in[k].ptr, in[k].pitch, in[k].bstride, (int)(16 / es), es));
    if (fused_ok) {
        // Temporally blocked path: the whole recurrence (filter.py:191-206) as ceil(n/kmax) launches; the first
        // block performs step 1, the last one finalizes bar.  State ping-pongs between two workspace pairs.
        const int kmax = p->steps_per_block ? p->steps_per_block : FusedGeom<double>::H;
        gcmf_field pair[2][2] = {{A[0], B[0]},
                                 {gcmf_field{(char*)workspace + (size_t)2 * bb, pitch, bs},
                                  gcmf_field{(char*)workspace + (size_t)3 * bb, pitch, bs}}};
        int cur = area ? 1 : 0;  // the prepared field lives in B (pair 0) when area-weighted
        gcmf_field T1 = X[0], T2 = X[0];
        for (int i = 1; i <= n;) {
            const int kk = (n - i + 1) < kmax ? (n - i + 1) : kmax;
            TRY(gcmf_cheb_fused(p, nb, i, kk, &T1, &T2, &pair[cur][0], &pair[cur][1], out, stream));
            T1 = pair[cur][0];
            T2 = pair[cur][1];
            cur ^= 1;
            i += kk;
        }
        return GCMF_OK;
    }
    if (plan_uses_cg2(p)) {
        // vector operators, temporally blocked: the recurrence (filter.py:225-283) as n/2 two-step launches (+ the one-step LAST
        // kernel when n is odd).  State ping-pongs between two workspace pairs; the user's input is only read.
        bool ok = true;
        for (int k = 0; ok && k < nc; ++k)
            ok = aligned(out[k].ptr, out[k].pitch, out[k].bstride, (int)(16 / es), es) &&
                 aligned(in[k].ptr, in[k].pitch, in[k].bstride, (int)(16 / es), es);
        if (ok) {
            gcmf_field pair[2][2][2];
            for (int j = 0; j < 4; ++j)
                for (int k = 0; k < nc; ++k)
                    pair[j >> 1][j & 1][k] = gcmf_field{(char*)workspace + (size_t)(j * nc + k) * bb, pitch, bs};
            const gcmf_field* T1 = X;
            const gcmf_field* T2 = X;
            int cur = 0;
            for (int i = 1; i <= n;) {
                const int kk = (n - i + 1) < 2 ? 1 : 2;
                TRY(gcmf_cheb_fused(p, nb, i, kk, T1, T2, pair[cur][0], pair[cur][1], out, stream));
                T1 = pair[cur][0];
                T2 = pair[cur][1];
                cur ^= 1;
                i += kk;
            }
            return GCMF_OK;
        }
    }
    // step 1: T1 = A(x) -> A ; bar = p0 x + p1 T1 -> out        (filter.py:191-195)
    TRY(gcmf_cheb_step(p, nb, 1, X, nullptr, A, out, stream));
    gcmf_field T1[2], T2[2];
    for (int k = 0; k < nc; ++k) { T1[k] = A[k]; T2[k] = X[k]; }
    for (int i = 2; i <= n; ++i) {  // filter.py:196-206; pointer rotation replaces the two .copy() per step
        gcmf_field Dst[2];
        for (int k = 0; k < nc; ++k) Dst[k] = (i == 2 && !area) ? B[k] : T2[k];  // never write the user's input
        TRY(gcmf_cheb_step(p, nb, i, T1, T2, Dst, out, stream));
        for (int k = 0; k < nc; ++k) { T2[k] = T1[k]; T1[k] = Dst[k]; }
    }
    return GCMF_OK;
}

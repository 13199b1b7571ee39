// gcmf_stencils.cuh -- point-wise arithmetic of the gcm-filters Laplacians and of one Chebyshev step.
//
// Everything here is a pure __host__ __device__ function of (parameters, batch b, row j, first
// column i0): the CUDA kernels in gcmf.cu call it per thread, and the test-only host emulator
// (tests/hostemu) compiles the very same code with g++ to check the index handling on a machine
// without a GPU.  Compiled with -fmad=false: multiplications and additions round separately, in
// the reference's evaluation order, so REGULAR5 and VECTOR_B are bit-identical to numpy.
//
// Index conventions: f[j,i], j = row (axis -2), i = column (axis -1);
// E = [j,i+1], W = [j,i-1], N = [j+1,i], S = [j-1,i]; x always wraps; y wraps when WRAP_Y is set,
// otherwise rows -1 and ny are ghost rows that exist in memory (latitude-band decomposition).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef GCMF_HD
#define GCMF_HD __host__ __device__ __forceinline__
#endif

namespace gcmf {

enum : int { OP_REGULAR5 = 0, OP_FLUX = 1, OP_VECTOR_B = 2, OP_VECTOR_C = 3 };
enum : int { FL_MASK = 1, FL_NAN2NUM = 2, FL_FOLD_N = 4, FL_CUT_S = 8, FL_WRAP_Y = 16, FL_AREA = 32 };
enum : int { MODE_LAP = 0, MODE_FIRST = 1, MODE_MID = 2, MODE_LAST = 3 };
constexpr int MAX_PLANES = 16;

template <typename T> struct Lim;
template <> struct Lim<float> { static GCMF_HD float big() { return 3.402823466e+38f; } };
template <> struct Lim<double> { static GCMF_HD double big() { return 1.7976931348623157e+308; } };

// numpy.nan_to_num: NaN -> 0, +-inf -> +-largest finite value.  Branch-free on the device (integer
// selects on the bit pattern) so that it never splits a warp or blocks instruction scheduling.
GCMF_HD double nan2num(double x) {
#ifdef __CUDA_ARCH__
    const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
    const bool nonfin = (hi & 0x7ff00000u) == 0x7ff00000u;
    const bool isnan = ((hi & 0x000fffffu) | lo) != 0u;  // only meaningful when nonfin
    const unsigned fhi = isnan ? 0u : ((hi & 0x80000000u) | 0x7fefffffu);
    const unsigned flo = isnan ? 0u : 0xffffffffu;
    return __hiloint2double((int)(nonfin ? fhi : hi), (int)(nonfin ? flo : lo));
#else
    if (!(x - x == 0.0)) {
        if (x != x) return 0.0;
        return x > 0.0 ? Lim<double>::big() : -Lim<double>::big();
    }
    return x;
#endif
}
GCMF_HD float nan2num(float x) {
#ifdef __CUDA_ARCH__
    const unsigned u = __float_as_uint(x);
    const bool nonfin = (u & 0x7f800000u) == 0x7f800000u;
    const bool isnan = (u & 0x007fffffu) != 0u;
    const unsigned f = isnan ? 0u : ((u & 0x80000000u) | 0x7f7fffffu);
    return __uint_as_float(nonfin ? f : u);
#else
    if (!(x - x == 0.0f)) {
        if (x != x) return 0.0f;
        return x > 0.0f ? Lim<float>::big() : -Lim<float>::big();
    }
    return x;
#endif
}

struct PlaneRef {
    const void* p;    // element (b, j, i) at p[(b % nb)*bstride + j*pitch + i]
    int64_t pitch;
    int64_t bstride;
    int32_t nb;
};

// CUtensorMap (128 opaque bytes, 64-byte aligned): a TMA descriptor encoded on the host, see gcmf_fused.cuh
struct alignas(64) TmaDesc { uint64_t opaque[16]; };

template <typename T> struct FieldRef {
    T* p;
    int64_t pitch;
    int64_t bstride;
};

struct Geo {
    int32_t ny, nx;
    int32_t flags;
};

// Ghost-row exchange of a latitude band through peer memory (NVLink): while a step kernel writes T_i it also
// stores its first / last owned row straight into the neighbouring GPU's ghost row, then raises a flag in
// the neighbour's memory; the next step on that GPU waits for the flag before it touches its ghost rows.
template <typename T> struct HaloRef {
    T* north[2];             // neighbour's ghost row that mirrors my row ny-1 (element b=0, i=0), per component
    T* south[2];             // neighbour's ghost row that mirrors my row 0
    int64_t nbs, sbs;        // batch strides of the neighbours' arrays
    const uint32_t* wait_n;  // LOCAL flags raised by the neighbours: ghost rows of t1 are valid once >= wait_v
    const uint32_t* wait_s;
    uint32_t* sig_n;         // flags in the neighbours' memory, set to sig_v when my rows have landed there
    uint32_t* sig_s;
    uint32_t wait_v, sig_v;
    uint32_t* counters;      // LOCAL [2]: border CTAs done (top, bottom); self-resetting
    int32_t enabled;
};

// Everything one launch needs.  NC = number of field components (1 scalar, 2 vector).
template <typename T> struct StepParams {
    Geo g;
    PlaneRef plane[MAX_PLANES];
    FieldRef<const T> t1[2];  // field the Laplacian acts on (LAP: input)
    FieldRef<const T> t2[2];  // T_{i-2}                       (MID, LAST)
    FieldRef<T> t0[2];        // LAP: output; FIRST: T_1 out; MID: T_i out (may alias t2)
    FieldRef<T> bar[2];       // running filtered field; LAST writes the finalized result here
    double c;                 // 2/s_max  or 2/(s_max dx_min^2)        (filter.py:168-173)
    double p0, p1;            // FIRST: p[0], p[1];  MID/LAST: p1 = p[i]
    int64_t nb;
    HaloRef<T> halo;          // enabled only for band-decomposed plans driven through gcmf_cheb_step_halo
};

// Per-thread point context: rows/columns of the neighbours after wrap / fold handling.
struct Pt {
    int32_t b, j, i0;
    int32_t jm, jp;   // row indices of the S and N neighbours
    int32_t im, ip;   // column indices of the W neighbour of column i0 and the E neighbour of i0+VX-1
    bool fold_top;    // N neighbour of (j, i) is (j, nx-1-i)
    bool cut_south;   // no S neighbour (flux-form: zero south flux)
};

template <int VX> GCMF_HD Pt make_pt(const Geo& g, int b, int j, int i0) {
    Pt q;
    q.b = b; q.j = j; q.i0 = i0;
    const bool wrap = (g.flags & FL_WRAP_Y) != 0;
    q.jm = (j == 0 && wrap) ? g.ny - 1 : j - 1;
    q.jp = (j == g.ny - 1 && wrap) ? 0 : j + 1;
    q.im = (i0 == 0) ? g.nx - 1 : i0 - 1;
    q.ip = (i0 + VX >= g.nx) ? 0 : i0 + VX;
    q.fold_top = (g.flags & FL_FOLD_N) && j == g.ny - 1;
    q.cut_south = (g.flags & FL_CUT_S) && j == 0;
    return q;
}

// ---- (vector) loads of VX consecutive elements; callers guarantee alignment when VX > 1 ----
template <typename S, int VX> struct Ld {
    static GCMF_HD void go(const S* p, S (&o)[VX]) {
#pragma unroll
        for (int v = 0; v < VX; ++v) o[v] = p[v];
    }
};
template <typename S, int VX> struct St {
    static GCMF_HD void go(S* p, const S (&o)[VX]) {
#pragma unroll
        for (int v = 0; v < VX; ++v) p[v] = o[v];
    }
};
#ifdef __CUDACC__
template <> struct Ld<double, 2> {
    static GCMF_HD void go(const double* p, double (&o)[2]) {
        const double2 t = *reinterpret_cast<const double2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
};
template <> struct Ld<float, 4> {
    static GCMF_HD void go(const float* p, float (&o)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
};
template <> struct Ld<float, 2> {
    static GCMF_HD void go(const float* p, float (&o)[2]) {
        const float2 t = *reinterpret_cast<const float2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
};
template <> struct St<float, 2> {
    static GCMF_HD void go(float* p, const float (&o)[2]) { *reinterpret_cast<float2*>(p) = make_float2(o[0], o[1]); }
};
template <> struct Ld<uint8_t, 4> {
    static GCMF_HD void go(const uint8_t* p, uint8_t (&o)[4]) {
        const uchar4 t = *reinterpret_cast<const uchar4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    }
};
template <> struct Ld<uint8_t, 2> {
    static GCMF_HD void go(const uint8_t* p, uint8_t (&o)[2]) {
        const uchar2 t = *reinterpret_cast<const uchar2*>(p);
        o[0] = t.x; o[1] = t.y;
    }
};
template <> struct St<double, 2> {
    static GCMF_HD void go(double* p, const double (&o)[2]) {
        *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
    }
};
template <> struct St<float, 4> {
    static GCMF_HD void go(float* p, const float (&o)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    }
};
#endif

// store the freshly computed row segment into the neighbours' ghost rows (peer memory)
template <typename T, int VX>
GCMF_HD void halo_push_row(const StepParams<T>& P, int k, int b, int j, int i0, const T (&v)[VX]) {
    if (!P.halo.enabled) return;
    if (j == P.g.ny - 1 && P.halo.north[k]) St<T, VX>::go(P.halo.north[k] + (int64_t)b * P.halo.nbs + i0, v);
    if (j == 0 && P.halo.south[k]) St<T, VX>::go(P.halo.south[k] + (int64_t)b * P.halo.sbs + i0, v);
}

// 5-point neighbourhood of VX consecutive points of row j.
template <typename S, int VX> struct Nb {
    S w, e;        // [j, i0-1], [j, i0+VX]
    S c[VX];       // [j, i0+v]
    S n[VX], s[VX];
    GCMF_HD S east(int v) const { return v == VX - 1 ? e : c[v + 1]; }
    GCMF_HD S west(int v) const { return v == 0 ? w : c[v - 1]; }
};

template <typename S, int VX>
GCMF_HD void load_nb(Nb<S, VX>& o, const S* base, int64_t pitch, const Pt& q, int nx) {
    const S* rc = base + (int64_t)q.j * pitch;
    Ld<S, VX>::go(rc + q.i0, o.c);
    o.w = rc[q.im];
    o.e = rc[q.ip];
    Ld<S, VX>::go(base + (int64_t)q.jm * pitch + q.i0, o.s);
    if (q.fold_top) {
#pragma unroll
        for (int v = 0; v < VX; ++v) o.n[v] = rc[nx - 1 - (q.i0 + v)];
    } else {
        Ld<S, VX>::go(base + (int64_t)q.jp * pitch + q.i0, o.n);
    }
}

template <typename S> GCMF_HD const S* plane_base(const PlaneRef& pl, int b) {
    return reinterpret_cast<const S*>(pl.p) + (int64_t)(pl.nb > 1 ? (int)((unsigned)b % (unsigned)pl.nb) : 0) * pl.bstride;
}

// One flux-form Laplacian value, shared by the one-step and the fused kernels:
//   Lap = (Fe - Fw + Fn - Fs) * ra with F = difference * face coefficient (kernels.py:297-315, 564-585),
// evaluated as a chain of fused multiply-adds (explicit fma: the library is built with -fmad=false so that
// nothing else contracts).  The flux family is not bit-comparable with the reference anyway (the face
// coefficients are precombined, SURVEY note N3: ~1e-16 relative); the fma chain is the more accurate form.
GCMF_HD double fma_(double a, double b, double c) { return ::fma(a, b, c); }
GCMF_HD float fma_(float a, float b, float c) { return ::fmaf(a, b, c); }
template <typename T>
GCMF_HD T flux_lap(T oc, T ow, T oe, T on, T os, T ce, T cew, T cn, T cs, T ra) {
    T sum = (oe - oc) * ce;
    sum = fma_(ow - oc, cew, sum);   // - (oc - ow) * cew
    sum = fma_(on - oc, cn, sum);
    sum = fma_(os - oc, cs, sum);    // - (oc - os) * cs
    return sum * ra;
}
// shifted Laplacian A(x) = -x - c*Lap (filter.py:171) for the flux family: one fma
template <typename T> GCMF_HD T shifted_flux(T x, T c, T lap) { return fma_(-c, lap, -x); }

// T_i = 2 A(T_{i-1}) - T_{i-2} (filter.py:197-203) as one fma: 2*a is exact, so the bits are those of (2*a) - t2.
template <typename T> GCMF_HD T cheb_next(T a, T t2) { return fma_(T(2), a, -t2); }
// bar + p_i T_i in fp64 (filter.py:204; 195 with bar = p0 x): numpy's two roundings, in every kernel (contracting it
// into one fma was measured within noise and would only move the results away from the reference's).
GCMF_HD double bar_update(double bar, double p, double t0) { return bar + p * t0; }

// exponent all ones: NaN or +-inf (what nan2num changes)
GCMF_HD bool nonfinite(double x) {
#ifdef __CUDA_ARCH__
    return ((unsigned)__double2hiint(x) & 0x7ff00000u) == 0x7ff00000u;
#else
    return !(x - x == 0.0);
#endif
}
GCMF_HD bool nonfinite(float x) {
#ifdef __CUDA_ARCH__
    return (__float_as_uint(x) & 0x7f800000u) == 0x7f800000u;
#else
    return !(x - x == 0.0f);
#endif
}

// =====================================================================================
// Operators.  apply(P, q, lap, x): lap[c][v] = Laplacian at the VX points, x[c][v] = raw centre
// values of P.t1 (NaNs kept: the reference's `-field` term uses the raw field, filter.py:171).
// =====================================================================================

// ---- REGULAR5 (kernels.py:107-124 unmasked; :150-190 masked; :435-492 masked + fold) ----
template <typename T, int VX, bool MASKED> struct OpRegular5 {
    static constexpr int NC = 1;
    static constexpr bool FMA_SHIFT = false;
    static GCMF_HD void apply(const StepParams<T>& P, const Pt& q, T (&lap)[1][VX], T (&x)[1][VX]) {
        Nb<T, VX> f;
        load_nb<T, VX>(f, P.t1[0].p + (int64_t)q.b * P.t1[0].bstride, P.t1[0].pitch, q, P.g.nx);
#pragma unroll
        for (int v = 0; v < VX; ++v) x[0][v] = f.c[v];
        if (!MASKED) {
#pragma unroll
            for (int v = 0; v < VX; ++v)  // kernels.py:115-121: -4 f + E + W + N + S
                lap[0][v] = (((T(-4) * f.c[v] + f.east(v)) + f.west(v)) + f.n[v]) + f.s[v];
        } else {
            Nb<uint8_t, VX> m;
            load_nb<uint8_t, VX>(m, plane_base<uint8_t>(P.plane[0], q.b), P.plane[0].pitch, q, P.g.nx);
            // o = wet_mask * nan_to_num(f)   (kernels.py:175-176, 472-473)
            Nb<T, VX> o;
            o.w = m.w ? nan2num(f.w) : T(0);
            o.e = m.e ? nan2num(f.e) : T(0);
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                o.c[v] = m.c[v] ? nan2num(f.c[v]) : T(0);
                o.n[v] = m.n[v] ? nan2num(f.n[v]) : T(0);
                o.s[v] = m.s[v] ? nan2num(f.s[v]) : T(0);
            }
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                // wet_fac = mE + mW + mN + mS (kernels.py:165-170, 462-467)
                const int wf = (m.east(v) != 0) + (m.west(v) != 0) + (m.n[v] != 0) + (m.s[v] != 0);
                const T r = (((-T(wf) * o.c[v] + o.east(v)) + o.west(v)) + o.n[v]) + o.s[v];  // :178-184
                lap[0][v] = m.c[v] ? r : T(0);  // :186
            }
        }
    }
};

// ---- FLUX form: Lap = (((Fe - Fw) + Fn) - Fs) * ra,  Fe[j,i] = (o[j,i+1]-o[j,i])*ce[j,i],
//      Fn[j,i] = (o[j+1,i]-o[j,i])*cn[j,i], Fw = Fe[j,i-1], Fs = Fn[j-1,i]
//      (kernels.py:297-315, 351-372, 408-429, 564-585 with precombined face coefficients) ----
template <typename T, int VX> struct OpFlux {
    static constexpr int NC = 1;
    static constexpr bool FMA_SHIFT = true;
    static GCMF_HD void apply(const StepParams<T>& P, const Pt& q, T (&lap)[1][VX], T (&x)[1][VX]) {
        Nb<T, VX> f;
        load_nb<T, VX>(f, P.t1[0].p + (int64_t)q.b * P.t1[0].bstride, P.t1[0].pitch, q, P.g.nx);
        T oc[VX], on[VX], os[VX];
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            x[0][v] = f.c[v];
            oc[v] = nan2num(f.c[v]);
            on[v] = nan2num(f.n[v]);
            os[v] = nan2num(f.s[v]);
        }
        const T ow = nan2num(f.w), oe = nan2num(f.e);
        const T* ce = plane_base<T>(P.plane[0], q.b);
        const T* cn = plane_base<T>(P.plane[1], q.b);
        const T* ra = plane_base<T>(P.plane[2], q.b);
        T cev[VX], cnv[VX], csv[VX], rav[VX];
        const T* cer = ce + (int64_t)q.j * P.plane[0].pitch;
        Ld<T, VX>::go(cer + q.i0, cev);
        const T cew = cer[q.im];
        Ld<T, VX>::go(cn + (int64_t)q.j * P.plane[1].pitch + q.i0, cnv);
        if (q.cut_south) {
#pragma unroll
            for (int v = 0; v < VX; ++v) csv[v] = T(0);
        } else {
            Ld<T, VX>::go(cn + (int64_t)q.jm * P.plane[1].pitch + q.i0, csv);
        }
        Ld<T, VX>::go(ra + (int64_t)q.j * P.plane[2].pitch + q.i0, rav);
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            const T o_e = v == VX - 1 ? oe : oc[v + 1 < VX ? v + 1 : v];
            const T o_w = v == 0 ? ow : oc[v > 0 ? v - 1 : 0];
            lap[0][v] = flux_lap<T>(oc[v], o_w, o_e, on[v], os[v], cev[v], v == 0 ? cew : cev[v > 0 ? v - 1 : 0], cnv[v],
                                    csv[v], rav[v]);
        }
    }
};

// ---- VECTOR_B (kernels.py:740-837): 10-term sum, left to right ----
template <typename T, int VX> struct OpVectorB {
    static constexpr int NC = 2;
    static constexpr bool FMA_SHIFT = false;
    static GCMF_HD void apply(const StepParams<T>& P, const Pt& q, T (&lap)[2][VX], T (&x)[2][VX]) {
        Nb<T, VX> u, w;
        load_nb<T, VX>(u, P.t1[0].p + (int64_t)q.b * P.t1[0].bstride, P.t1[0].pitch, q, P.g.nx);
        load_nb<T, VX>(w, P.t1[1].p + (int64_t)q.b * P.t1[1].bstride, P.t1[1].pitch, q, P.g.nx);
#pragma unroll
        for (int v = 0; v < VX; ++v) { x[0][v] = u.c[v]; x[1][v] = w.c[v]; }
        // nan_to_num both components (kernels.py:743-744)
        u.w = nan2num(u.w); u.e = nan2num(u.e); w.w = nan2num(w.w); w.e = nan2num(w.e);
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            u.c[v] = nan2num(u.c[v]); u.n[v] = nan2num(u.n[v]); u.s[v] = nan2num(u.s[v]);
            w.c[v] = nan2num(w.c[v]); w.n[v] = nan2num(w.n[v]); w.s[v] = nan2num(w.s[v]);
        }
        T k[8][VX];
#pragma unroll
        for (int s = 0; s < 8; ++s)
            Ld<T, VX>::go(plane_base<T>(P.plane[s], q.b) + (int64_t)q.j * P.plane[s].pitch + q.i0, k[s]);
#pragma unroll
        for (int v = 0; v < VX; ++v) {
            const T cc = k[0][v], dun = k[1][v], dus = k[2][v], due = k[3][v], duw = k[4][v];
            const T dmc = k[5][v], dmn = k[6][v], dme = k[7][v];
            const T dms = -dmn, dmw = -dme;  // kernels.py:804-805
            lap[0][v] = ((((((((cc * u.c[v] + dun * u.n[v]) + dus * u.s[v]) + due * u.east(v)) + duw * u.west(v))
                            + dmc * w.c[v]) + dmn * w.n[v]) + dms * w.s[v]) + dme * w.east(v)) + dmw * w.west(v);
            lap[1][v] = ((((((((cc * w.c[v] + dun * w.n[v]) + dus * w.s[v]) + due * w.east(v)) + duw * w.west(v))
                            + dmc * u.c[v]) + dmn * u.n[v]) + dms * u.s[v]) + dme * u.east(v)) + dmw * u.west(v);
        }
    }
};

// ---- VECTOR_C (kernels.py:647-696).  Compact 3x3 stencil: two chained half-cell stages.  VX = 1. ----
template <typename T, int VX> struct OpVectorC {
    static_assert(VX == 1, "VECTOR_C is instantiated with one point per thread");
    static constexpr int NC = 2;
    static constexpr bool FMA_SHIFT = false;
    // value of a field or plane at (j+dj, i+di), dj, di in {-1,0,1}: base[off.o[dj+1][di+1]]
    struct Off9 { int64_t o[3][3]; };
    static GCMF_HD Off9 offsets(int64_t pitch, const Pt& q) {
        Off9 f;
        const int64_t r[3] = {(int64_t)q.jm * pitch, (int64_t)q.j * pitch, (int64_t)q.jp * pitch};
        const int32_t col[3] = {q.im, q.i0, q.ip};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) f.o[a][b] = r[a] + col[b];
        return f;
    }
    struct At {
        const T* base; const Off9* off;
        GCMF_HD T operator()(int dj, int di) const { return base[off->o[dj + 1][di + 1]]; }
    };
    // all 14 coefficient planes share one pitch (checked on the host); u and v share one pitch too
    static GCMF_HD void apply(const StepParams<T>& P, const Pt& q, T (&lap)[2][1], T (&x)[2][1]) {
        const Off9 fo = offsets(P.t1[0].pitch, q);
        const Off9 po = offsets(P.plane[0].pitch, q);
        const At U{P.t1[0].p + (int64_t)q.b * P.t1[0].bstride, &fo};
        const At V{P.t1[1].p + (int64_t)q.b * P.t1[1].bstride, &fo};
        At K[14];
#pragma unroll
        for (int s = 0; s < 14; ++s) K[s] = At{plane_base<T>(P.plane[s], q.b), &po};
        x[0][0] = U(0, 0);
        x[1][0] = V(0, 0);
        // a = u/dyCu, b = v/dxCv, c = v/dyCv, e = u/dxCu  (as products with the reciprocal planes)
        auto a = [&](int dj, int di) { return nan2num(U(dj, di)) * K[0](dj, di); };
        auto b = [&](int dj, int di) { return nan2num(V(dj, di)) * K[1](dj, di); };
        auto c = [&](int dj, int di) { return nan2num(V(dj, di)) * K[2](dj, di); };
        auto e = [&](int dj, int di) { return nan2num(U(dj, di)) * K[3](dj, di); };
        // str_xx at T point (j+dj, i+di)  (kernels.py:653-661)
        auto sxx = [&](int dj, int di) {
            return -(K[4](dj, di) * (a(dj, di) - a(dj, di - 1)) - K[5](dj, di) * (b(dj, di) - b(dj - 1, di)));
        };
        // str_xy at q point (j+dj, i+di)  (kernels.py:663-670)
        auto sxy = [&](int dj, int di) {
            return -(K[6](dj, di) * (c(dj, di + 1) - c(dj, di)) + K[7](dj, di) * (e(dj + 1, di) - e(dj, di)));
        };
        const T sxx_c = sxx(0, 0), sxx_e = sxx(0, 1), sxx_n = sxx(1, 0);
        const T sxy_c = sxy(0, 0), sxy_s = sxy(-1, 0), sxy_w = sxy(0, -1);
        // kernels.py:672-682
        T uc = K[0](0, 0) * (K[8](0, 0) * sxx_c - K[8](0, 1) * sxx_e);
        uc = uc + K[3](0, 0) * (K[10](-1, 0) * sxy_s - K[10](0, 0) * sxy_c);
        lap[0][0] = uc * K[12](0, 0);
        // kernels.py:684-694
        T vc = K[2](0, 0) * (K[11](0, -1) * sxy_w - K[11](0, 0) * sxy_c);
        vc = vc - K[1](0, 0) * (K[9](0, 0) * sxx_c - K[9](1, 0) * sxx_n);
        lap[1][0] = vc * K[13](0, 0);
    }
};

// =====================================================================================
// One Chebyshev step at VX points (filter.py:162-175, 185-206 scalar; :225-283 vector).
// =====================================================================================
template <typename T, int VX, int NC, bool FMA_SHIFT, int MODE, bool HALO>
GCMF_HD void step_tail(const StepParams<T>& P, int b, int j, int i0, const T (&lap)[NC][VX], const T (&x)[NC][VX]);

template <typename T, int VX, class OP, int MODE, bool HALO = false>
GCMF_HD void step_body(const StepParams<T>& P, int b, int j, int i0) {
    constexpr int NC = OP::NC;
    const Pt q = make_pt<VX>(P.g, b, j, i0);
    T lap[NC][VX], x[NC][VX];
    OP::apply(P, q, lap, x);
    step_tail<T, VX, NC, OP::FMA_SHIFT, MODE, HALO>(P, b, j, i0, lap, x);
}

// The recurrence arithmetic that follows the Laplacian, shared by every one-step kernel.
// HALO: also store the new row into the neighbouring GPUs' ghost rows (band decomposition).
template <typename T, int VX, int NC, bool FMA_SHIFT, int MODE, bool HALO>
GCMF_HD void step_tail(const StepParams<T>& P, int b, int j, int i0, const T (&lap)[NC][VX], const T (&x)[NC][VX]) {
    const T c = (T)P.c;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        T outv[VX];
        if (MODE == MODE_LAP) {
#pragma unroll
            for (int v = 0; v < VX; ++v) outv[v] = lap[k][v];
            St<T, VX>::go(P.t0[k].p + (int64_t)b * P.t0[k].bstride + (int64_t)j * P.t0[k].pitch + i0, outv);
            continue;
        }
        T a[VX];
#pragma unroll
        for (int v = 0; v < VX; ++v)  // shifted Laplacian, filter.py:171/173
            a[v] = FMA_SHIFT ? shifted_flux<T>(x[k][v], c, lap[k][v]) : (-x[k][v] - c * lap[k][v]);
        T* barp = P.bar[k].p + (int64_t)b * P.bar[k].bstride + (int64_t)j * P.bar[k].pitch + i0;
        if (MODE == MODE_FIRST) {
            St<T, VX>::go(P.t0[k].p + (int64_t)b * P.t0[k].bstride + (int64_t)j * P.t0[k].pitch + i0, a);
            if (HALO) halo_push_row<T, VX>(P, k, b, j, i0, a);
#pragma unroll
            for (int v = 0; v < VX; ++v)  // filter.py:195
                outv[v] = (T)bar_update(P.p0 * (double)x[k][v], P.p1, (double)a[v]);
            St<T, VX>::go(barp, outv);
        } else {
            T t2[VX], t0[VX], bar[VX];
            Ld<T, VX>::go(P.t2[k].p + (int64_t)b * P.t2[k].bstride + (int64_t)j * P.t2[k].pitch + i0, t2);
            Ld<T, VX>::go(barp, bar);
#pragma unroll
            for (int v = 0; v < VX; ++v) {
                t0[v] = cheb_next<T>(a[v], t2[v]);                              // filter.py:197-203
                outv[v] = (T)bar_update((double)bar[v], P.p1, (double)t0[v]);  // filter.py:204
            }
            if (MODE == MODE_MID) {
                St<T, VX>::go(P.t0[k].p + (int64_t)b * P.t0[k].bstride + (int64_t)j * P.t0[k].pitch + i0, t0);
                if (HALO) halo_push_row<T, VX>(P, k, b, j, i0, t0);
            } else if (P.g.flags & FL_AREA) {  // finalize: divide by the cell area (kernels.py:103-104)
                T ar[VX];
                Ld<T, VX>::go(plane_base<T>(P.plane[1], b) + (int64_t)j * P.plane[1].pitch + i0, ar);
#pragma unroll
                for (int v = 0; v < VX; ++v) outv[v] = outv[v] / ar[v];
            }
            St<T, VX>::go(barp, outv);
        }
    }
}

// =====================================================================================
// VECTOR_C, tiled: the C-grid operator is two chained half-cell stages (stress tensor at T and q points, then
// its divergence at u and v points; kernels.py:647-696).  A CTA first evaluates the four weighted stresses
//   P1 = dyT^2*sxx, P2 = dxT^2*sxx (T points),  P3 = dxBu^2*sxy, P4 = dyBu^2*sxy (q points)
// once per point of its tile (+1 row / column) into shared memory, then every thread differences them.
// Same expressions, same order as OpVectorC::apply (bit-identical results), one third of the stress work.
// =====================================================================================
template <typename T> struct CgridTile {
#ifndef GCMF_CGRID_TX
#define GCMF_CGRID_TX 32  // tile geometry: A/B knobs (tests/tools/build_variant.py -DGCMF_CGRID_TX=.. -DGCMF_CGRID_TY=..);
#define GCMF_CGRID_TY 8   // 33 x 9 = 297 stress entries for 256 threads: the second pass of stage A runs 41 lanes
#endif
    static constexpr int TX = GCMF_CGRID_TX, TY = GCMF_CGRID_TY;  // output points per CTA
    static constexpr int SW = TX + 1, SH = TY + 1;      // stress tiles: one extra column / row
    static constexpr int NTHREADS = TX * TY;
    static constexpr int SMEM_ELEMS = 4 * SW * SH;

    // row / column index with the plan's boundary rule (periodic x; periodic y or ghost rows).  A tile may stick out
    // of the grid by up to TY / TX points; the stress entries out there feed discarded outputs only, so they are
    // clamped onto the last addressable row / column instead of being wrapped a second time (found by the ASAN
    // build of the host emulator: the single wrap read past the end of the planes for nx < 32 or ny < 9, and past
    // the ghost row of a latitude band whose height is not a multiple of TY).
    static GCMF_HD int wrap_row(const Geo& g, int j) {
        if (!(g.flags & FL_WRAP_Y)) return j > g.ny ? g.ny : j;  // band: rows -1 and ny are ghost rows in memory
        j = j < 0 ? j + g.ny : (j >= g.ny ? j - g.ny : j);
        return j >= g.ny ? g.ny - 1 : j;
    }
    static GCMF_HD int wrap_col(const Geo& g, int i) {
        i = i < 0 ? i + g.nx : (i >= g.nx ? i - g.nx : i);
        return i >= g.nx ? g.nx - 1 : i;
    }

    // stage A: entry e of the stress tiles.  sxx-type entries sit at (j0 + r, i0 + cc), sxy-type at (j0-1+r, i0-1+cc).
    static GCMF_HD void stress(const StepParams<T>& P, int b, int j0, int i0, int e, T* sm) {
        const int r = e / SW, cc = e % SW;
        const T* U = P.t1[0].p + (int64_t)b * P.t1[0].bstride;
        const T* V = P.t1[1].p + (int64_t)b * P.t1[1].bstride;
        const int64_t fp = P.t1[0].pitch, pp = P.plane[0].pitch;
        const T* K[12];
#pragma unroll
        for (int s = 0; s < 12; ++s) K[s] = plane_base<T>(P.plane[s], b);
        {   // T point (J, I): str_xx  (kernels.py:653-661)
            const int J = wrap_row(P.g, j0 + r), I = wrap_col(P.g, i0 + cc);
            const int Jm = wrap_row(P.g, j0 + r - 1), Im = wrap_col(P.g, i0 + cc - 1);
            const T a_c = nan2num(U[(int64_t)J * fp + I]) * K[0][(int64_t)J * pp + I];
            const T a_w = nan2num(U[(int64_t)J * fp + Im]) * K[0][(int64_t)J * pp + Im];
            const T b_c = nan2num(V[(int64_t)J * fp + I]) * K[1][(int64_t)J * pp + I];
            const T b_s = nan2num(V[(int64_t)Jm * fp + I]) * K[1][(int64_t)Jm * pp + I];
            const T sxx = -(K[4][(int64_t)J * pp + I] * (a_c - a_w) - K[5][(int64_t)J * pp + I] * (b_c - b_s));
            sm[0 * SW * SH + e] = K[8][(int64_t)J * pp + I] * sxx;   // dy2h * sxx
            sm[1 * SW * SH + e] = K[9][(int64_t)J * pp + I] * sxx;   // dx2h * sxx
        }
        {   // q point (J, I): str_xy  (kernels.py:663-670)
            const int J = wrap_row(P.g, j0 - 1 + r), I = wrap_col(P.g, i0 - 1 + cc);
            const int Jp = wrap_row(P.g, j0 + r), Ip = wrap_col(P.g, i0 + cc);
            const T c_c = nan2num(V[(int64_t)J * fp + I]) * K[2][(int64_t)J * pp + I];
            const T c_e = nan2num(V[(int64_t)J * fp + Ip]) * K[2][(int64_t)J * pp + Ip];
            const T e_c = nan2num(U[(int64_t)J * fp + I]) * K[3][(int64_t)J * pp + I];
            const T e_n = nan2num(U[(int64_t)Jp * fp + I]) * K[3][(int64_t)Jp * pp + I];
            const T sxy = -(K[6][(int64_t)J * pp + I] * (c_e - c_c) + K[7][(int64_t)J * pp + I] * (e_n - e_c));
            sm[2 * SW * SH + e] = K[10][(int64_t)J * pp + I] * sxy;  // dx2q * sxy
            sm[3 * SW * SH + e] = K[11][(int64_t)J * pp + I] * sxy;  // dy2q * sxy
        }
    }

    // stage B: the divergence at output point (j0+ty, i0+tx)  (kernels.py:672-694)
    static GCMF_HD void divergence(const StepParams<T>& P, int b, int j, int i, int ty, int tx, const T* sm,
                                   T (&lap)[2][1], T (&x)[2][1]) {
        const int64_t pp = P.plane[0].pitch;
        const int64_t o = (int64_t)j * pp + i;
        const T* P1 = sm;
        const T* P2 = sm + SW * SH;
        const T* P3 = sm + 2 * SW * SH;
        const T* P4 = sm + 3 * SW * SH;
        x[0][0] = P.t1[0].p[(int64_t)b * P.t1[0].bstride + (int64_t)j * P.t1[0].pitch + i];
        x[1][0] = P.t1[1].p[(int64_t)b * P.t1[1].bstride + (int64_t)j * P.t1[1].pitch + i];
        const T k0 = plane_base<T>(P.plane[0], b)[o], k1 = plane_base<T>(P.plane[1], b)[o];
        const T k2 = plane_base<T>(P.plane[2], b)[o], k3 = plane_base<T>(P.plane[3], b)[o];
        // T-point entries (r, cc) = (ty, tx); q-point entries are shifted by one: (j, i) -> (ty+1, tx+1)
        const int t = ty * SW + tx, qd = (ty + 1) * SW + (tx + 1);
        T uc = k0 * (P1[t] - P1[t + 1]);                 // 1/dyCu * (dy2h sxx - (dy2h sxx)[j,i+1])
        uc = uc + k3 * (P3[qd - SW] - P3[qd]);           // + 1/dxCu * ((dx2q sxy)[j-1,i] - dx2q sxy)
        lap[0][0] = uc * plane_base<T>(P.plane[12], b)[o];
        T vc = k2 * (P4[qd - 1] - P4[qd]);               // 1/dyCv * ((dy2q sxy)[j,i-1] - dy2q sxy)
        vc = vc - k1 * (P2[t] - P2[t + SW]);             // - 1/dxCv * (dx2h sxx - (dx2h sxx)[j+1,i])
        lap[1][0] = vc * plane_base<T>(P.plane[13], b)[o];
    }
};

// prepare: x = f * area (kernels.py:100-101); finalize: f = x / area (kernels.py:103-104)
template <typename T>
GCMF_HD void prepare_body(const T* in, T* out, const T* area, int64_t idx_in, int64_t idx_out, int64_t idx_area,
                          bool divide = false) {
    out[idx_out] = divide ? in[idx_in] / area[idx_area] : in[idx_in] * area[idx_area];
}

}  // namespace gcmf

/* TEST-ONLY: drive the C ABI of libgcmf.so on the GPU from plain C (no Python, no torch: starts in about a second)
 * and compare with the host emulator of the same sources (tests/hostemu/libgcmf_hostemu.so) on the same inputs.
 * Everything on the path is IEEE +, -, *, / and explicit fma, so the expectation is bit-identical results.
 *
 *   gcc -std=c99 -O1 tests/cabi/gpu_vs_emu.c -I include -I /usr/local/cuda/include -L /usr/local/cuda/lib64 \
 *       -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -ldl -lm -o tests/cabi/gpu_vs_emu
 *   ./tests/cabi/gpu_vs_emu            (on a GPU box, from the repository root)
 *   ./tests/cabi/gpu_vs_emu --emu-only (anywhere: emulator against itself, checks this program)
 *   ./tests/cabi/gpu_vs_emu build/variants/libgcmf_ss.so [emulator.so]   (an A/B variant: 3 s instead of a pytest run)
 *
 * Cases: every operator family, one-step and fused kernels, whole grids, tripolar folds and latitude bands with ghost
 * rows, fp32 and fp64 (the first nine are the ones run on a B200 at the end of round 1, profiles/gpu_vs_emu_r01.log:
 * the kernels whose device code had changed after the last full GPU test run).
 */
#define _POSIX_C_SOURCE 200112L
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "gcmf.h"

typedef struct {
    void* h;
    int (*plan_create)(const gcmf_plan_desc*, gcmf_plan**);
    int (*plan_destroy)(gcmf_plan*);
    int (*set_plane)(gcmf_plan*, int, const void*, int64_t, int64_t, int32_t);
    int (*set_filter)(gcmf_plan*, int32_t, const double*, double);
    int (*ws_bytes)(const gcmf_plan*, int64_t, size_t*);
    int (*filter)(gcmf_plan*, int64_t, const gcmf_field*, const gcmf_field*, void*, size_t, void*);
    int (*cheb_fused)(gcmf_plan*, int64_t, int32_t, int32_t, const gcmf_field*, const gcmf_field*, const gcmf_field*,
                      const gcmf_field*, const gcmf_field*, void*);
    const char* (*last_error)(void);
    int (*sm_arch)(void);
    int device; /* 1: pointers handed to this library are device pointers */
} Api;

static int load(Api* a, const char* path, int device) {
    a->h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!a->h) {
        fprintf(stderr, "dlopen %s: %s\n", path, dlerror());
        return 1;
    }
    *(void**)&a->plan_create = dlsym(a->h, "gcmf_plan_create");
    *(void**)&a->plan_destroy = dlsym(a->h, "gcmf_plan_destroy");
    *(void**)&a->set_plane = dlsym(a->h, "gcmf_plan_set_plane");
    *(void**)&a->set_filter = dlsym(a->h, "gcmf_plan_set_filter");
    *(void**)&a->ws_bytes = dlsym(a->h, "gcmf_workspace_bytes");
    *(void**)&a->filter = dlsym(a->h, "gcmf_filter");
    *(void**)&a->cheb_fused = dlsym(a->h, "gcmf_cheb_fused");
    *(void**)&a->last_error = dlsym(a->h, "gcmf_last_error");
    *(void**)&a->sm_arch = dlsym(a->h, "gcmf_sm_arch");
    a->device = device;
    return !(a->plan_create && a->filter && a->cheb_fused && a->last_error);
}

static uint64_t rng_state = 88172645463325252ull;
static double urand(void) { /* xorshift64 */
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}

static void* to_side(const Api* a, const void* host, size_t bytes) {
    void* p = NULL;
    if (a->device) {
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return NULL;
        if (bytes) cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice);
    } else {
        if (posix_memalign(&p, 256, bytes ? bytes : 16)) return NULL;
        memcpy(p, host, bytes);
    }
    return p;
}
static void from_side(const Api* a, void* host, const void* p, size_t bytes) {
    if (a->device) {
        cudaDeviceSynchronize();
        cudaMemcpy(host, p, bytes, cudaMemcpyDeviceToHost);
    } else {
        memcpy(host, p, bytes);
    }
}
static void free_side(const Api* a, void* p) {
    if (a->device) cudaFree(p);
    else free(p);
}

#define CHECK(api, call)                                                              \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_) {                                                                    \
            fprintf(stderr, "%s failed: %d %s\n", #call, rc_, (api)->last_error()); \
            return -1.0;                                                              \
        }                                                                             \
    } while (0)

/* One problem on one library.  planes: nplanes host arrays of `ny_alloc * nx` elements (uint8 for a mask plane, else
 * the field type), fields: ncomp host arrays of nb * ny_alloc * nx elements.  ghost > 0: a latitude band -- the arrays
 * carry `ghost` rows on either side of the ny owned rows and one fused block (steps 1..4) is run through
 * gcmf_cheb_fused, whose three outputs are returned back to back; otherwise gcmf_filter.  Returns 0 or -1. */
static double run(const Api* a, int op, int dtype, int ny, int nx, int flags, int ghost, int nplanes, void* const* planes,
                  const int* plane_is_mask, int nb, int ncomp, void* const* fields, int n_steps, void* out_host,
                  size_t* out_bytes) {
    const size_t es = dtype == GCMF_F64 ? 8 : 4;
    const int nya = ny + 2 * ghost;
    const size_t fbytes = (size_t)nb * nya * nx * es;
    gcmf_plan_desc d;
    gcmf_plan* plan = NULL;
    double p[64];
    void* dpl[16] = {0};
    void* dfi[2] = {0};
    void* dou[6] = {0};
    gcmf_field fin[2], fout[2];
    int k, s;
    memset(&d, 0, sizeof d);
    d.op = op;
    d.dtype = dtype;
    d.ny = ny;
    d.nx = nx;
    d.flags = flags;
    d.device = 0;
    CHECK(a, a->plan_create(&d, &plan));
    for (s = 0; s < nplanes; ++s) {
        const size_t pes = plane_is_mask[s] ? 1 : es;
        if (!planes[s]) continue;
        dpl[s] = to_side(a, planes[s], (size_t)nya * nx * pes);
        CHECK(a, a->set_plane(plan, s, (char*)dpl[s] + (size_t)ghost * nx * pes, nx, (int64_t)nya * nx, 1));
    }
    for (k = 0; k <= n_steps; ++k) p[k] = (k % 2 ? -1.0 : 1.0) / (k + 2.0);
    CHECK(a, a->set_filter(plan, n_steps, p, 0.1));
    for (k = 0; k < ncomp; ++k) {
        dfi[k] = to_side(a, fields[k], fbytes);
        fin[k].ptr = (char*)dfi[k] + (size_t)ghost * nx * es;
        fin[k].pitch = nx;
        fin[k].bstride = (int64_t)nya * nx;
    }
    if (!ghost) {
        size_t wsb = 0;
        void* ws;
        void* zero = calloc(1, fbytes);
        CHECK(a, a->ws_bytes(plan, nb, &wsb));
        if (a->device) {
            if (cudaMalloc(&ws, wsb ? wsb : 256) != cudaSuccess) return -1.0;
        } else if (posix_memalign(&ws, 256, wsb ? wsb : 256)) {
            return -1.0;
        }
        for (k = 0; k < ncomp; ++k) {
            dou[k] = to_side(a, zero, fbytes);
            fout[k].ptr = dou[k];
            fout[k].pitch = nx;
            fout[k].bstride = (int64_t)ny * nx;
        }
        CHECK(a, a->filter(plan, nb, fin, fout, ws, wsb, NULL));
        for (k = 0; k < ncomp; ++k) from_side(a, (char*)out_host + k * fbytes, dou[k], fbytes);
        *out_bytes = ncomp * fbytes;
        free_side(a, ws);
        free(zero);
    } else { /* one fused block on a band: T_{4}, T_{3} (ghosted arrays) and bar (owned rows) */
        const size_t obytes = (size_t)nb * ny * nx * es;
        void* zero = calloc(1, fbytes);
        gcmf_field t1o, t2o, bar;
        dou[0] = to_side(a, zero, fbytes);
        dou[1] = to_side(a, zero, fbytes);
        dou[2] = to_side(a, zero, obytes);
        t1o.ptr = (char*)dou[0] + (size_t)ghost * nx * es;
        t2o.ptr = (char*)dou[1] + (size_t)ghost * nx * es;
        t1o.pitch = t2o.pitch = bar.pitch = nx;
        t1o.bstride = t2o.bstride = (int64_t)nya * nx;
        bar.ptr = dou[2];
        bar.bstride = (int64_t)ny * nx;
        CHECK(a, a->cheb_fused(plan, nb, 1, 4, fin, NULL, &t1o, &t2o, &bar, NULL));
        from_side(a, out_host, dou[0], fbytes);
        from_side(a, (char*)out_host + fbytes, dou[1], fbytes);
        from_side(a, (char*)out_host + 2 * fbytes, dou[2], obytes);
        *out_bytes = 2 * fbytes + obytes;
        free(zero);
    }
    for (k = 0; k < 6; ++k)
        if (dou[k]) free_side(a, dou[k]);
    for (k = 0; k < ncomp; ++k) free_side(a, dfi[k]);
    for (s = 0; s < nplanes; ++s)
        if (dpl[s]) free_side(a, dpl[s]);
    a->plan_destroy(plan);
    return 0.0;
}

static int compare(const char* name, int dtype, const void* x, const void* y, size_t bytes) {
    size_t n = bytes / (dtype == GCMF_F64 ? 8 : 4), i, nbad = 0, nnan = 0;
    double worst = 0.0;
    for (i = 0; i < n; ++i) {
        const double u = dtype == GCMF_F64 ? ((const double*)x)[i] : ((const float*)x)[i];
        const double v = dtype == GCMF_F64 ? ((const double*)y)[i] : ((const float*)y)[i];
        if (isnan(u) || isnan(v)) {
            nnan += isnan(u) && isnan(v);
            nbad += isnan(u) != isnan(v);
            continue;
        }
        if (u != v) {
            const double dd = fabs(u - v) / (fabs(v) > 1e-300 ? fabs(v) : 1.0);
            if (dd > worst) worst = dd;
            ++nbad;
        }
    }
    printf("%-34s %9zu values, %zu NaN on both sides, %zu differ, worst relative difference %.3e  %s\n", name, n, nnan,
           nbad, worst, nbad ? "DIFFERENT" : "bit-identical");
    return nbad != 0;
}

int main(int argc, char** argv) {
    const int emu_only = argc > 1 && !strcmp(argv[1], "--emu-only");
    /* optional: the CUDA library under test (e.g. an A/B variant from build/variants) and the emulator to compare with */
    const char* gpu_path = (argc > 1 && !emu_only) ? argv[1] : "gcm_filters_b200/libgcmf.so";
    const char* emu_path = (argc > 2 && !emu_only) ? argv[2] : "tests/hostemu/libgcmf_hostemu.so";
    Api gpu, emu;
    int bad = 0, c;
    if (load(&emu, emu_path, 0)) return 2;
    if (load(&gpu, emu_only ? emu_path : gpu_path, !emu_only)) return 2;
    printf("library under test: sm_arch %d%s\n", gpu.sm_arch(), emu_only ? " (emulator against itself)" : "");
    if (!emu_only && cudaSetDevice(0) != cudaSuccess) {
        fprintf(stderr, "no CUDA device\n");
        return 3;
    }
    /* case table: op, dtype, ny, nx, flags, ghost rows, nb, n_steps */
    struct {
        const char* name;
        int op, dtype, ny, nx, flags, ghost, nb, n_steps;
    } cases[] = {
        {"cgrid f64 37x54 (odd, 2 tiles wide)", GCMF_OP_VECTOR_C, GCMF_F64, 37, 54, GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM, 0, 2, 6},
        {"cgrid f32 20x24 (narrower than a tile)", GCMF_OP_VECTOR_C, GCMF_F32, 20, 24, GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM, 0, 1, 5},
        {"cgrid f64 7x70 (lower than a tile)", GCMF_OP_VECTOR_C, GCMF_F64, 7, 70, GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM, 0, 1, 4},
        {"reg5 masked f64 40x264 fused", GCMF_OP_REGULAR5, GCMF_F64, 40, 264, GCMF_FLAG_WRAP_Y | GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM, 0, 3, 9},
        {"reg5 masked f32 40x264 fused", GCMF_OP_REGULAR5, GCMF_F32, 40, 264, GCMF_FLAG_WRAP_Y | GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM, 0, 3, 9},
        {"reg5 unmasked f64 36x128 fused", GCMF_OP_REGULAR5, GCMF_F64, 36, 128, GCMF_FLAG_WRAP_Y, 0, 1, 5},
        {"reg5 masked f64 band 40(+8)x264", GCMF_OP_REGULAR5, GCMF_F64, 40, 264, GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM, 4, 2, 9},
        {"reg5 masked f32 band 36(+8)x264", GCMF_OP_REGULAR5, GCMF_F32, 36, 264, GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM, 4, 2, 9},
        {"flux f64 48x256 fused (control)", GCMF_OP_FLUX, GCMF_F64, 48, 256, GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM, 0, 2, 9},
        {"flux f32 40x264 fused", GCMF_OP_FLUX, GCMF_F32, 40, 264, GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM, 0, 3, 10},
        {"flux f64 48x256 fused, tripolar fold", GCMF_OP_FLUX, GCMF_F64, 48, 256,
         GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM | GCMF_FLAG_FOLD_N | GCMF_FLAG_CUT_S, 0, 2, 9},
        {"flux f64 37x54 one-step (odd)", GCMF_OP_FLUX, GCMF_F64, 37, 54, GCMF_FLAG_WRAP_Y | GCMF_FLAG_NAN2NUM, 0, 2, 7},
        {"flux f64 band 40(+8)x256 fused", GCMF_OP_FLUX, GCMF_F64, 40, 256, GCMF_FLAG_NAN2NUM, 4, 2, 9},
        {"vector B f64 37x54", GCMF_OP_VECTOR_B, GCMF_F64, 37, 54, GCMF_FLAG_WRAP_Y, 0, 2, 6},
        {"vector B f32 33x64", GCMF_OP_VECTOR_B, GCMF_F32, 33, 64, GCMF_FLAG_WRAP_Y, 0, 1, 5},
        {"reg5 masked+area f64 40x264 fused", GCMF_OP_REGULAR5, GCMF_F64, 40, 264,
         GCMF_FLAG_WRAP_Y | GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM | GCMF_FLAG_AREA, 0, 2, 9},
        {"reg5 masked f64 40x264 tripolar fused", GCMF_OP_REGULAR5, GCMF_F64, 40, 264,
         GCMF_FLAG_WRAP_Y | GCMF_FLAG_MASK | GCMF_FLAG_NAN2NUM | GCMF_FLAG_FOLD_N | GCMF_FLAG_CUT_S, 0, 2, 9},
    };
    for (c = 0; c < (int)(sizeof cases / sizeof cases[0]); ++c) {
        const int op = cases[c].op, dt = cases[c].dtype, ny = cases[c].ny, nx = cases[c].nx, gh = cases[c].ghost;
        const int nb = cases[c].nb, nya = ny + 2 * gh;
        const size_t es = dt == GCMF_F64 ? 8 : 4, npl = (size_t)nya * nx, nf = (size_t)nb * npl;
        const int ncomp = (op == GCMF_OP_VECTOR_C || op == GCMF_OP_VECTOR_B) ? 2 : 1;
        const int nplanes = op == GCMF_OP_VECTOR_C ? 14 : op == GCMF_OP_VECTOR_B ? 8 : op == GCMF_OP_FLUX ? 3
                            : (cases[c].flags & GCMF_FLAG_AREA) ? 2 : ((cases[c].flags & GCMF_FLAG_MASK) ? 1 : 0);
        void* planes[16] = {0};
        int is_mask[16] = {0};
        void* fields[2] = {0};
        unsigned char* mask = NULL;
        void *o1, *o2;
        size_t b1 = 0, b2 = 0, i;
        int s, k;
        for (s = 0; s < nplanes; ++s) {
            if (op == GCMF_OP_REGULAR5 && s == 0) {
                if (!(cases[c].flags & GCMF_FLAG_MASK)) continue; /* slot 0 unused without a mask */
                mask = (unsigned char*)malloc(npl);
                for (i = 0; i < npl; ++i) mask[i] = urand() > 0.25;
                if (cases[c].flags & GCMF_FLAG_CUT_S)
                    for (i = 0; i < (size_t)nx; ++i) mask[i] = 0; /* tripolar grids: row 0 is land */
                planes[s] = mask;
                is_mask[s] = 1;
            } else {
                planes[s] = malloc(npl * es);
                for (i = 0; i < npl; ++i) {
                    double v = 0.5 + urand();
                    if (op == GCMF_OP_FLUX && s < 2 && urand() < 0.2) v = 0.0; /* closed faces */
                    if (dt == GCMF_F64) ((double*)planes[s])[i] = v;
                    else ((float*)planes[s])[i] = (float)v;
                }
            }
        }
        for (k = 0; k < ncomp; ++k) {
            fields[k] = malloc(nf * es);
            for (i = 0; i < nf; ++i) {
                double v = urand();
                if (mask && !mask[i % npl]) v = NAN; /* NaN on land */
                if (op == GCMF_OP_FLUX && urand() < 0.05) v = urand() < 0.8 ? NAN : INFINITY; /* nan_to_num at work */
                if (dt == GCMF_F64) ((double*)fields[k])[i] = v;
                else ((float*)fields[k])[i] = (float)v;
            }
        }
        o1 = calloc(3 * nf + 16, es * 2);
        o2 = calloc(3 * nf + 16, es * 2);
        if (run(&gpu, op, dt, ny, nx, cases[c].flags, gh, nplanes, planes, is_mask, nb, ncomp, fields, cases[c].n_steps, o1, &b1) < 0 ||
            run(&emu, op, dt, ny, nx, cases[c].flags, gh, nplanes, planes, is_mask, nb, ncomp, fields, cases[c].n_steps, o2, &b2) < 0 ||
            b1 != b2) {
            printf("%-34s FAILED TO RUN\n", cases[c].name);
            bad = 1;
        } else {
            bad |= compare(cases[c].name, dt, o1, o2, b1);
        }
        free(o1);
        free(o2);
        for (s = 0; s < nplanes; ++s) free(planes[s]);
        for (k = 0; k < ncomp; ++k) free(fields[k]);
    }
    if (!emu_only) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("CUDA error at the end: %s\n", cudaGetErrorString(e));
            bad = 1;
        }
    }
    printf(bad ? "RESULT: differences\n" : "RESULT: all cases bit-identical\n");
    return bad;
}

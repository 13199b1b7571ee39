/*
 * gcmf.h -- C ABI of libgcmf.so: the B200 (sm_100a) implementation of the gcm-filters
 * iterative Laplacian filter hot path.
 *
 * The reference (ocean-eddy-cpt/gcm-filters) is pure Python: it has no FFI.  This header is
 * the boundary a maintainer would bind (ctypes / cffi) to replace
 *   - the numpy/cupy Laplacian kernels          gcm_filters/kernels.py:107-840
 *   - the Chebyshev step loop                   gcm_filters/filter.py:154-291
 *   - the numpy-or-cupy backend switch          gcm_filters/gpu_compat.py:5-10
 * Each entry point cites the reference interface it stands in for.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, a GCMF_E* code otherwise, and
 *     gcmf_last_error() returns a thread-local message.  No exceptions cross the boundary.
 *   - all data pointers are caller-owned DEVICE pointers (the Python host keeps them alive in
 *     torch tensors); a plan owns nothing but its own small descriptor tables.
 *   - calls are asynchronous on the cudaStream_t passed as `void* stream`.
 *   - a plan is bound to one device and allows one in-flight call at a time; distinct plans
 *     are independent.
 *   - fields are (nb, ny, nx) arrays addressed as  ptr[b*bstride + j*pitch + i]  (units:
 *     elements).  Axis -2 is y (rows), axis -1 is x, as in the reference ("dimension order
 *     matters", filter.py:444-446).  Leading batch axes are flattened into nb.
 */
#ifndef GCMF_H
#define GCMF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCMF_VERSION 1

/* ---- status codes ---- */
enum {
    GCMF_OK = 0,
    GCMF_EINVAL = 1,   /* bad argument / unsupported combination */
    GCMF_ECUDA = 2,    /* a CUDA runtime / driver call failed     */
    GCMF_ESTATE = 3    /* plan not fully configured               */
};

/* ---- element types ---- */
enum { GCMF_F32 = 0, GCMF_F64 = 1 };

/* ---- device operator families.  The 11 reference GridTypes (kernels.py:13-28) map onto
 *      four stencil families; the host precombines the reference's grid variables into the
 *      coefficient planes each family reads (gcm_filters_b200/kernels.py).
 *
 *  GCMF_OP_REGULAR5  5-point Laplacian on a unit grid, optional uint8 wet mask:
 *        REGULAR, REGULAR_AREA_WEIGHTED (kernels.py:107-147)                 no mask
 *        REGULAR_WITH_LAND[_AREA_WEIGHTED] (kernels.py:150-219)              mask
 *        TRIPOLAR_REGULAR_WITH_LAND_AREA_WEIGHTED (kernels.py:435-492)       mask + fold
 *      planes: [0] wet mask (uint8, value != 0 is ocean; only with GCMF_FLAG_MASK)
 *              [1] cell area (only with GCMF_FLAG_AREA; AreaWeightedMixin kernels.py:89-104)
 *  GCMF_OP_FLUX      flux-form Laplacian with east-face, north-face and 1/area planes:
 *        IRREGULAR_WITH_LAND (kernels.py:222-318), MOM5U/MOM5T (kernels.py:321-432),
 *        TRIPOLAR_POP_WITH_LAND (kernels.py:495-588, with fold)
 *      planes: [0] ce  [1] cn  [2] ra
 *  GCMF_OP_VECTOR_B  B-grid vector Laplacian (kernels.py:702-840)
 *      planes: [0] cc [1] DUN [2] DUS [3] DUE [4] DUW [5] DMC [6] DMN [7] DME
 *  GCMF_OP_VECTOR_C  C-grid vector Laplacian (kernels.py:591-699)
 *      planes: [0] 1/dyCu [1] 1/dxCv [2] 1/dyCv [3] 1/dxCu
 *              [4] (kappa_iso+kappa_aniso/2)*dyT/dxT*mask_t [5] (..)*dxT/dyT*mask_t
 *              [6] kappa_iso*dyBu/dxBu*mask_q               [7] kappa_iso*dxBu/dyBu*mask_q
 *              [8] dyT^2 [9] dxT^2 [10] dxBu^2 [11] dyBu^2 [12] 1/area_u (0 if area_u<=0) [13] 1/area_v
 */
enum { GCMF_OP_REGULAR5 = 0, GCMF_OP_FLUX = 1, GCMF_OP_VECTOR_B = 2, GCMF_OP_VECTOR_C = 3 };
#define GCMF_MAX_PLANES 16

/* ---- plan flags ---- */
enum {
    GCMF_FLAG_MASK = 1,      /* REGULAR5: plane 0 is a uint8 wet mask                               */
    GCMF_FLAG_NAN2NUM = 2,   /* nan_to_num() the field before differencing (kernels.py:175,300,..)  */
    GCMF_FLAG_FOLD_N = 4,    /* tripolar fold: north neighbour of (ny-1,i) is (ny-1,nx-1-i)
                                (_prepare_tripolar_exchanges, kernels.py:33-40)                     */
    GCMF_FLAG_CUT_S = 8,     /* no south neighbour below row 0 (tripolar grids: row 0 is land)       */
    GCMF_FLAG_WRAP_Y = 16,   /* y is periodic (np.roll on axis -2).  Clear for a latitude band whose
                                ghost rows j=-1 and j=ny are physically present in memory           */
    GCMF_FLAG_AREA = 32      /* area-weighted prepare/finalize (kernels.py:100-104)                 */
};

typedef struct gcmf_plan gcmf_plan; /* opaque */

typedef struct gcmf_plan_desc {
    int32_t op;      /* GCMF_OP_*   */
    int32_t dtype;   /* GCMF_F32 / GCMF_F64: arithmetic and storage type of fields and planes */
    int32_t ny, nx;  /* rows and columns of the (band of the) grid this plan computes         */
    int32_t flags;   /* GCMF_FLAG_* */
    int32_t device;  /* CUDA device ordinal the plan (and all pointers given to it) lives on  */
} gcmf_plan_desc;

/* One (nb, ny, nx) field: element (b, j, i) is at ptr[b*bstride + j*pitch + i].  Any pitch >= nx and
 * bstride >= ny*pitch is accepted: 16-byte aligned pointers with pitch and bstride multiples of 16 bytes take the
 * vector / temporally blocked kernels, anything else their scalar forms (same results).  GCMF_OP_VECTOR_C reads u and
 * v next to each other at every point and wants ONE row pitch for the two input components (EINVAL otherwise), as
 * it wants one pitch for its 14 planes. */
typedef struct gcmf_field {
    void *ptr;
    int64_t pitch;
    int64_t bstride;
} gcmf_field;

/* Library/ABI version (GCMF_VERSION) and the SM architecture the kernels were built for (100). */
int gcmf_version(void);
int gcmf_sm_arch(void);
const char *gcmf_last_error(void);

/* Replaces `Laplacian(**grid_vars)` (filter.py:181-183): create the per-grid operator state. */
int gcmf_plan_create(const gcmf_plan_desc *desc, gcmf_plan **out);
int gcmf_plan_destroy(gcmf_plan *plan);

/* Attach coefficient plane `slot` (see the per-op plane lists above).  The plane has `plane_nb`
 * batch entries; field batch index b reads entry b % plane_nb (plane_nb = 1: shared 2-D plane,
 * the reference's broadcasting of (y,x) grid variables against (...,y,x) fields).
 * Stands in for the dataclass fields + __post_init__ derived arrays of each Laplacian
 * (e.g. kernels.py:250-295, 510-543, 630-645, 730-738). */
int gcmf_plan_set_plane(gcmf_plan *plan, int slot, const void *dptr, int64_t pitch, int64_t bstride,
                        int32_t plane_nb);

/* Chebyshev polynomial of the filter: FilterSpec(n_steps, s_max, p, dx_min_sq) (filter.py:92-151)
 * reduced to what the loop needs: p[0..n_steps] and c = 2/s_max (dimensional Laplacians) or
 * 2/(s_max*dx_min_sq) (filter.py:168-173).  `p` is a HOST pointer, copied. */
int gcmf_plan_set_filter(gcmf_plan *plan, int32_t n_steps, const double *p, double c);

/* Bytes of caller-provided device scratch that gcmf_filter needs for nb batch slices. */
int gcmf_workspace_bytes(const gcmf_plan *plan, int64_t nb, size_t *bytes);

/* out = Laplacian(in): one call of the reference operator's __call__ (kernels.py:113,172,297,351,
 * 408,469,564,647,740).  in/out hold ncomp fields (1 scalar, 2 for the vector ops: u then v). */
int gcmf_laplacian(gcmf_plan *plan, int64_t nb, const gcmf_field *in, const gcmf_field *out, void *stream);

/* out = filter(in): the whole of filter_func / filter_func_vec (filter.py:177-212, 242-289):
 * prepare, the n_steps Chebyshev recurrence over the shifted Laplacian, finalize.
 * `in` is not modified; `out` may not alias `in`. */
int gcmf_filter(gcmf_plan *plan, int64_t nb, const gcmf_field *in, const gcmf_field *out, void *workspace,
                size_t workspace_bytes, void *stream);

/* One step of the recurrence, for callers that interleave halo exchanges between steps (the
 * latitude-band decomposition).  step = 1 .. n_steps (filter.py:192-206):
 *   step 1            : t1 = A(x);              bar = p0*x + p1*t1     (x = `t1_in`, the prepared field)
 *   1 < step < n_steps: t0 = 2*A(t1_in) - t2;   bar += p[step]*t0
 *   step == n_steps   : as above, then bar = finalize(bar) and t0 is not stored
 * with A(x) = -x - c*Laplacian(x).  `t0_out` may alias `t2` (in-place rotation), never `t1_in`.
 * Arrays are ncomp-long.  For step 1 `t2` is ignored. */
int gcmf_cheb_step(gcmf_plan *plan, int64_t nb, int32_t step, const gcmf_field *t1_in, const gcmf_field *t2,
                   const gcmf_field *t0_out, const gcmf_field *bar, void *stream);

/* Temporal blocking (north_star item 2; no counterpart in the reference, which makes ~26 array passes
 * per step).  gcmf_fused_max_steps: how many recurrence steps this plan can fuse into one HBM round
 * trip (0 = no fused path for this operator / grid; the one-step kernels are used).
 * gcmf_plan_set_steps_per_block: 0 = auto (default), 1 = never fuse, 2..4 = cap the block length.
 * gcmf_cheb_fused: recurrence steps step .. step+k-1 (1 <= step, step+k-1 <= n_steps) in one launch:
 *   reads T_{step-1} (`t1_in`; the prepared field when step == 1) and T_{step-2} (`t2_in`, ignored when
 *   step == 1), writes T_{step+k-1} (`t1_out`) and T_{step+k-2} (`t2_out`) unless the block reaches
 *   n_steps, and updates `bar` (initialised by step 1, finalized by step n_steps).  Outputs must not alias
 *   inputs (neighbouring tiles read the inputs' halos).  Scalar operators fuse 1..4 steps (tiles + 4-cell
 *   halos); the vector operators (each field argument is an array of two gcmf_field: u, v) fuse exactly
 *   k = 2 steps per launch (rows streamed once for both steps; k = 1 is accepted for the trailing step of
 *   an odd n_steps and runs the one-step kernel).  On a band plan (GCMF_FLAG_WRAP_Y clear) the caller
 *   keeps gcmf_fused_max_steps() ghost rows of the fields and of every plane on either side of the band
 *   and refreshes those of t1_out / t2_out after every block. */
int gcmf_fused_max_steps(const gcmf_plan *plan);
int gcmf_plan_set_steps_per_block(gcmf_plan *plan, int32_t k);
int gcmf_cheb_fused(gcmf_plan *plan, int64_t nb, int32_t step, int32_t k, const gcmf_field *t1_in,
                    const gcmf_field *t2_in, const gcmf_field *t1_out, const gcmf_field *t2_out,
                    const gcmf_field *bar, void *stream);

/* ---- ghost-row exchange through peer memory (latitude-band decomposition over NVLink) ----------------------
 * No counterpart in the reference (it cannot split the filtered dimensions at all).  A band-decomposed plan
 * (GCMF_FLAG_WRAP_Y clear) owns rows 0..ny-1 and reads ghost rows -1 and ny.  With a gcmf_halo the step kernel
 * itself stores its first / last row of T_i into the neighbouring GPUs' ghost rows (plain stores to
 * peer-mapped addresses) and raises a flag in the neighbour's memory; the next step over there waits for the
 * flag before touching its ghost rows.  No separate pack / send / receive / unpack launches, no host round trip.
 * All pointers are device pointers; `north_ghost` / `south_ghost` / `signal_*` point into the NEIGHBOURS'
 * memory (peer mapped, e.g. torch symmetric memory or CUDA IPC), `wait_*` and `counters` are local.
 * Flag protocol: flags only grow; a launch waits until *wait_x - wait_value >= 0 (as int32) and, when all of
 * its border CTAs have pushed, sets *signal_x = signal_value.  NULL pointers disable a direction. */
typedef struct gcmf_halo {
    void *north_ghost[2];          /* per component: neighbour's ghost row mirroring my row ny-1 (b = 0, i = 0) */
    void *south_ghost[2];          /* neighbour's ghost row mirroring my row 0 */
    int64_t north_bstride, south_bstride; /* batch strides (elements) of the neighbours' arrays */
    const uint32_t *wait_north, *wait_south;
    uint32_t *signal_north, *signal_south;
    uint32_t wait_value, signal_value;
    uint32_t *counters;            /* local scratch, 2 x uint32, zero before first use */
} gcmf_halo;

/* gcmf_cheb_step with the exchange fused in: waits for the ghost rows of `t1_in`, computes step `step`, pushes
 * the border rows of T_step into the neighbours' `t0_out` ghost rows, signals.  The last step pushes nothing
 * but still signals (so that the neighbours know this rank has finished reading). */
int gcmf_cheb_step_halo(gcmf_plan *plan, int64_t nb, int32_t step, const gcmf_field *t1_in, const gcmf_field *t2,
                        const gcmf_field *t0_out, const gcmf_field *bar, const gcmf_halo *halo, void *stream);
/* gcmf_cheb_fused (vector operators, k = 2) on a band plan with the exchange fused in: the launch waits for the two
 * ghost rows per side of `t1_in` / `t2_in` (flags of `halo_t1`), runs steps step and step+1, stores its first / last
 * two rows of T_{step+1} into the neighbours' ghost rows described by `halo_t1` and those of T_step into the ones
 * described by `halo_t2` (north_ghost[k]: the neighbour's ghost row mirroring my row ny-2, south_ghost[k]: the one
 * mirroring my row 0; consecutive ghost rows are `t1_out[k].pitch` elements apart), and raises the neighbours' flags.
 * The two halos share their flags, counters and batch strides.  A block that reaches n_steps pushes nothing but
 * still signals.  One launch per two Chebyshev steps and rank, no NCCL call, no host synchronisation. */
int gcmf_cheb_fused_halo(gcmf_plan *plan, int64_t nb, int32_t step, int32_t k, const gcmf_field *t1_in,
                         const gcmf_field *t2_in, const gcmf_field *t1_out, const gcmf_field *t2_out,
                         const gcmf_field *bar, const gcmf_halo *halo_t1, const gcmf_halo *halo_t2, void *stream);
/* Push the border rows of `field` (the prepared input, before step 1) into the neighbours' ghost rows. */
int gcmf_halo_push(gcmf_plan *plan, int64_t nb, const gcmf_field *field, const gcmf_halo *halo, void *stream);

/* x = field * area (AreaWeightedMixin.prepare, kernels.py:100-101); a copy when the plan has no
 * GCMF_FLAG_AREA.  gcmf_filter calls this itself. */
int gcmf_prepare(gcmf_plan *plan, int64_t nb, const gcmf_field *in, const gcmf_field *out, void *stream);
/* field = x / area (AreaWeightedMixin.finalize, kernels.py:103-104); a copy without GCMF_FLAG_AREA.  The filter
 * entry points finalize by themselves; this serves direct callers of the operator protocol. */
int gcmf_finalize(gcmf_plan *plan, int64_t nb, const gcmf_field *in, const gcmf_field *out, void *stream);

/* Number of kernel launches issued through this library by the calling process so far
 * (bench.py reports it as gpu_launches). */
int64_t gcmf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GCMF_H */

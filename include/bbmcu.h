/* bbmcu.h - C ABI of libbbmcu.so, the B200 (sm_100a) CUDA backbone for bbm's data-parallel hot path:
 * batched BSDF eval / sample / pdf / reflectance for the 34 analytic models, the MERL and spherical
 * linearizers, the six fitting metrics (loss + analytic parameter gradient) and .fit / BSDF-string I/O.
 *
 * The reference (bsdfbenchmark/bbm 0.5.1) has no C ABI on this path: its boundary is a set of C++20
 * concepts.  Each entry point below names the reference interface it stands in for; INTEGRATION.md
 * shows the C++ adapter a bbm maintainer would add on top (backbone/cuda).
 *
 * Conventions
 *  - every function returns 0 on success; on failure a non-zero bbmcu_status and a message readable
 *    through bbmcu_last_error(ctx) (the reference throws std::invalid_argument from its parsers,
 *    core/stringconvert.h:68,517,540,565, and std::runtime_error elsewhere, core/error.h:42-46).
 *  - direction / spectrum buffers are struct-of-arrays: an `xyz` (or `rgb`) pointer addresses three
 *    consecutive planes of n floats (x[n], y[n], z[n]); `uv` two planes.  The reference is
 *    array-of-structs (std::array<float,3>, backbone/native/include/backbone/array.h:27); adapters
 *    transpose at the host boundary only.
 *  - data pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes).  Device
 *    pointers run in place on the context's stream.  Host pointers are processed in chunks of 2^21
 *    elements through device staging buffers, the copies of a chunk overlapping the kernels and
 *    copies of its neighbours: pinned or registered host memory (cudaMallocHost, cudaHostRegister,
 *    bbmcu_host_register) is DMA'd in place; pageable memory goes through an internal pinned ring
 *    filled and drained by a few host threads.  The caller owns every buffer.
 *  - OUTPUT pointers of the batched calls may be NULL where documented: that plane is neither
 *    stored by the kernel nor copied back.  Host outputs of the fused pass that are cheap functions of
 *    others do not cross PCIe as they are: the flag plane travels as one byte per element, and
 *    sample_pdf - where the model defines it as pdf(sample.direction, out) - is written on the host
 *    from the pdf plane.  The caller's buffers receive the same bits either way
 *    (BBMCU_HOST_TRANSFER_PLAIN=1 in the environment sends every plane as it is).
 *  - planes of one argument are n floats apart unless bbmcu_set_plane_stride says otherwise.
 *  - there is no CPU fallback: without a usable CUDA device bbmcu_init fails.
 *  - one context per host thread per device; calls on one context are serialised on its stream.
 */
#ifndef BBMCU_H_
#define BBMCU_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define BBMCU_API __attribute__((visibility("default")))
#else
#define BBMCU_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bbmcu_ctx bbmcu_ctx;      /* device, stream, staging buffers, precomputed tables           */
typedef struct bbmcu_bsdf bbmcu_bsdf;    /* = bbm::bsdf_ptr<floatRGB>  (include/bbm/bsdf_ptr.h:20-165)     */
typedef struct bbmcu_loss bbmcu_loss;    /* = a bbm::sampledlossfunction (include/bbm/sampledlossfunction.h:26-95) */

typedef enum { BBMCU_OK = 0, BBMCU_INVALID_ARGUMENT = 1, BBMCU_RUNTIME_ERROR = 2, BBMCU_CUDA_ERROR = 3, BBMCU_OUT_OF_RANGE = 4 } bbmcu_status;

/* bbm::bsdf_flag (include/bbm/bsdf_flag.h:21-27) and bbm::unit_t (include/bbm/unit.h:20-24) */
enum { BBMCU_NONE = 0, BBMCU_DIFFUSE = 1, BBMCU_SPECULAR = 2, BBMCU_ALL = 3 };
enum { BBMCU_RADIANCE = 0, BBMCU_IMPORTANCE = 1 };
/* bbm::bsdf_attr (include/bbm/bsdf_attr_flag.h:17-31) */
enum { BBMCU_ATTR_DIFFUSE_SCALE = 1, BBMCU_ATTR_DIFFUSE_PARAMETER = 2, BBMCU_ATTR_SPECULAR_SCALE = 4,
       BBMCU_ATTR_SPECULAR_PARAMETER = 8, BBMCU_ATTR_DEPENDENT = 16, BBMCU_ATTR_ALL = 15 };
/* the six metrics of include/loss/cosine_weighted_l2.h and cosine_weighted_log.h */
enum { BBMCU_NGAN_L2 = 0, BBMCU_LOW_L2 = 1, BBMCU_BIERON_L2 = 2, BBMCU_LOW_LOG = 3, BBMCU_BIERON_LOG = 4, BBMCU_STANDARD_LOG = 5 };
/* which vector bbmcu_bsdf_get_params returns: parameter_values / _default_values / _lower_bound / _upper_bound
 * (include/bbm/bsdf_enumerate.h:102-237) */
enum { BBMCU_PARAM_VALUE = 0, BBMCU_PARAM_DEFAULT = 1, BBMCU_PARAM_LOWER = 2, BBMCU_PARAM_UPPER = 3 };

/* ---- context -------------------------------------------------------------------------------------- */
BBMCU_API int  bbmcu_init(int device, bbmcu_ctx** out);
BBMCU_API void bbmcu_destroy(bbmcu_ctx* ctx);
BBMCU_API const char* bbmcu_last_error(bbmcu_ctx* ctx);          /* ctx may be NULL: error of the last failed call on this thread */
BBMCU_API int  bbmcu_synchronize(bbmcu_ctx* ctx);
BBMCU_API void* bbmcu_stream(bbmcu_ctx* ctx);                    /* the cudaStream_t all work of this context is issued on */
BBMCU_API uint64_t bbmcu_launch_count(bbmcu_ctx* ctx);           /* kernels launched by this context so far */
/* planes of every SoA argument of the following batched calls are `ld` floats apart (ld >= n; 0 = n, the default).  With
 * ld a multiple of 4 (and 16-byte aligned base pointers) the kernels keep their 16-byte vector accesses for ANY n; with
 * the default stride a batch whose n is not a multiple of 4 has misaligned planes and runs on scalar accesses. */
BBMCU_API int  bbmcu_set_plane_stride(bbmcu_ctx* ctx, size_t ld);
/* page-lock caller memory once so the host-pointer path DMAs it in place (cudaHostRegister / cudaHostUnregister) */
BBMCU_API int  bbmcu_host_register(bbmcu_ctx* ctx, void* ptr, size_t bytes);
BBMCU_API int  bbmcu_host_unregister(bbmcu_ctx* ctx, void* ptr);

/* ---- model registry (replaces the BBM_EXPORT_BSDFMODEL tables, include/export/bbm_fromstring.h:48-49) -- */
typedef struct {
  const char* name;      /* attribute name as printed by toString                                    */
  int width;             /* scalars in this attribute                                                  */
  int rows;              /* 1, or 2 for [[..],[..]] attributes (complex RGB ior, Bagher F0/F1)        */
  int flag;              /* one BBMCU_ATTR_* bit                                                       */
  int offset;            /* first float of the attribute in the model's attribute block               */
} bbmcu_attr;
BBMCU_API int  bbmcu_model_count(void);                          /* 35: the 34 analytic models + Merl (staticmodel/merl.h) */
BBMCU_API const char* bbmcu_model_name(int model_id);
BBMCU_API int  bbmcu_model_lookup(const char* name, int* model_id);
BBMCU_API int  bbmcu_model_layout(int model_id, bbmcu_attr* attrs, int* n_attrs);   /* reflection order (util/reflection.h:141-148) */

/* ---- BSDF objects: bsdf_import / toString / parameter enumeration ------------------------------------ */
/* bbm::bsdf_import<floatRGB>(str) (include/bbm/bsdf_import.h:22-26): "Model(args)" or "Aggregate(m1, m2, ...)" */
BBMCU_API int  bbmcu_bsdf_from_string(bbmcu_ctx* ctx, const char* str, bbmcu_bsdf** out);
/* which of the reference's native configurations (backbone/native/include/backbone.h:41-42) the HOST side of the object
 * mirrors: floatRGB parses with std::stof and prints floats (a value beyond FLT_MAX is an error, as in the reference:
 * fits/bagher_sgd.fit); doubleRGB parses with std::stod and keeps doubles in the parameter vectors and strings.  Kernels
 * compute in FP32 either way (attribute values are rounded when they travel to the device). */
enum { BBMCU_FLOAT_RGB = 0, BBMCU_DOUBLE_RGB = 1 };
BBMCU_API int  bbmcu_bsdf_from_string_ex(bbmcu_ctx* ctx, const char* str, int config, bbmcu_bsdf** out);
BBMCU_API void bbmcu_bsdf_free(bbmcu_bsdf* bsdf);
BBMCU_API int  bbmcu_bsdf_to_string(const bbmcu_bsdf* bsdf, char* buf, size_t cap);           /* bsdf_ptr::toString */
BBMCU_API int  bbmcu_bsdf_param_count(const bbmcu_bsdf* bsdf, int attr_flags);
BBMCU_API int  bbmcu_bsdf_get_params(const bbmcu_bsdf* bsdf, int which, int attr_flags, double* values, int* count);
BBMCU_API int  bbmcu_bsdf_set_params(bbmcu_bsdf* bsdf, int attr_flags, const double* values, int count);

/* ---- batched BSDF concept (include/concepts/bsdfmodel.h:32-146; include/bbm/bsdf_base.h:76-129) --------- */
BBMCU_API int  bbmcu_eval(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
                const float* in_xyz, const float* out_xyz, size_t n, float* rgb);
BBMCU_API int  bbmcu_sample(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
                  const float* out_xyz, const float* xi_uv, size_t n, float* dir_xyz, float* pdf, int32_t* flag);
BBMCU_API int  bbmcu_pdf(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
               const float* in_xyz, const float* out_xyz, size_t n, float* pdf);
BBMCU_API int  bbmcu_reflectance(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
                       const float* out_xyz, size_t n, float* rgb);
/* one fused pass of the checkBsdf-style inner loop (bin/checkBsdf.cpp:206-267):
 *   s = sample(out, xi); rgb = eval(s.direction, out); pdf = pdf(s.direction, out)          */
BBMCU_API int  bbmcu_sample_eval_pdf(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
                           const float* out_xyz, const float* xi_uv, size_t n,
                           float* dir_xyz, float* sample_pdf, int32_t* flag, float* rgb, float* pdf);

/* (any of dir_xyz, sample_pdf, flag, rgb, pdf may be NULL: not stored, not copied)
 * The same pass over inputs drawn ON THE DEVICE (0 bytes in): element i uses a counter-based generator (Philox4x32-10 of
 * (seed, first + i)): `out` uniform on the upper hemisphere as bin/checkBsdf.cpp:38-45 draws it, xi uniform in [0,1)^2.
 * gen_out_xyz / gen_xi_uv (may be NULL) receive the generated inputs, so a checker can evaluate the reference at exactly those. */
BBMCU_API int  bbmcu_sample_eval_pdf_generated(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
                                     uint64_t seed, uint64_t first, size_t n, float* gen_out_xyz, float* gen_xi_uv,
                                     float* dir_xyz, float* sample_pdf, int32_t* flag, float* rgb, float* pdf);
/* eval(in, out) for (in, out) = merl_linearizer(idx), idx = first .. first+n-1, the linearizer fused into the kernel
 * (include/linearizer/merl_linearizer.h:50-83): 0 bytes in, 12 bytes out per eval.  in_xyz / out_xyz (may be NULL) receive
 * the generated directions; they are bit-identical to bbmcu_merl_dirs. */
BBMCU_API int  bbmcu_eval_merl_grid(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, int component, int unit,
                          uint32_t first, size_t n, float* rgb, float* in_xyz, float* out_xyz);

/* ---- linearizers (include/linearizer/merl_linearizer.h:50-123, spherical_linearizer.h:37-111) --------- */
#define BBMCU_MERL_BINS 1458000u
/* merl_linearizer(in, out): bin index, BBMCU_MERL_BINS for pairs below the horizon, 0xFFFFFFFF for NaN input */
BBMCU_API int  bbmcu_merl_index(bbmcu_ctx* ctx, const float* in_xyz, const float* out_xyz, size_t n, uint32_t* index);
/* merl_linearizer(idx) for idx = first .. first+n-1 */
BBMCU_API int  bbmcu_merl_dirs(bbmcu_ctx* ctx, uint32_t first, size_t n, float* in_xyz, float* out_xyz);
typedef struct {
  uint32_t samples_in[2], samples_out[2];   /* (phi, theta) counts                                     */
  float start_in[2], end_in[2];             /* (phi, theta) ranges; Hemisphere = (2 pi, pi/2)          */
  float start_out[2], end_out[2];
} bbmcu_spherical_grid;
BBMCU_API void bbmcu_spherical_grid_default(bbmcu_spherical_grid* g, uint32_t in_phi, uint32_t in_theta, uint32_t out_phi, uint32_t out_theta);
BBMCU_API int  bbmcu_spherical_dirs(bbmcu_ctx* ctx, const bbmcu_spherical_grid* grid, uint64_t first, size_t n, float* in_xyz, float* out_xyz);

/* ---- measured data (include/staticmodel/merl.h:78-96,173-206) --------------------------------------------- */
/* read a MERL binary (3 x u32 dims = 90,90,180 then 3 planar double[N]; negative -> 0; scaled by
 * 1/1500, 1.15/1500, 1.66/1500) into `rgb` (3 planes of BBMCU_MERL_BINS floats, HOST memory) */
BBMCU_API int  bbmcu_merl_read(bbmcu_ctx* ctx, const char* filename, float* rgb);
BBMCU_API int  bbmcu_merl_write(bbmcu_ctx* ctx, const char* filename, const float* rgb);     /* inverse of the above */

/* ---- Holzschuch-Pacanowski precompute (precompute/HolzschuchPacanowski/G1.cpp) --------------------------------------- */
/* regenerates the 100 x 1000 G1 table of include/precomputed/holzschuchpacanowski/G1.h (row = 5/p - 1, column = the
 * table's tan(theta) map) into `table` (HOST memory, 100000 floats): ~1e9 quadrature terms on the GPU */
BBMCU_API int  bbmcu_hp_precompute_g1(bbmcu_ctx* ctx, float* table);
/* regenerates the 100 x 100 x 100 renormalisation table sigma_rel^2 / sigma_s^2 of precompute/HolzschuchPacanowski/
 * normalization.cpp (index = (b, c, sin theta_i) as its main() maps them, :214-236; integralSH :133-155: 5.7e9 terms of a
 * double pow and acos) into `table` (HOST memory, 1 000 000 floats, sin theta_i fastest).  The reference ships no copy of
 * this table (it is one of its missing large blobs); entries follow the generator's own operation order. */
BBMCU_API int  bbmcu_hp_precompute_normalization(bbmcu_ctx* ctx, float* table);

/* ---- losses (include/loss/{cosine_weighted_l2,cosine_weighted_log}.h, include/bbm/sampledlossfunction.h:62-87) ---------------------------------- */
/* The reference evaluates  loss = (1/N) sum_i e(in_i, out_i, fitted.eval(in_i,out_i), reference.eval(in_i,out_i))
 * over a linearizer.  A bbmcu_loss fixes metric, linearizer, component and the reference operand; the
 * reference values are tabulated once per sample index (they never change during a fit).
 *   grid == NULL: merl_linearizer (N = BBMCU_MERL_BINS), else the spherical grid.
 *   reference is ONE of: an analytic bsdf (reference_bsdf), or a measured MERL table
 *   (reference_merl_rgb: 3 planes of BBMCU_MERL_BINS floats, host or device), looked up per sample with
 *   merl_linearizer(in, out) exactly as merl_data::eval does.
 *   first/count select the shard [first, first+count) of the sample axis this context owns (multi-GPU);
 *   count == 0 means all N.  Losses and gradients are returned as SUMS over the shard divided by the
 *   FULL N, so summing over shards (ncclAllReduce) gives the reference's mean. */
BBMCU_API int  bbmcu_loss_create(bbmcu_ctx* ctx, int metric, const bbmcu_spherical_grid* grid, int component, int unit,
                       const bbmcu_bsdf* reference_bsdf, const float* reference_merl_rgb,
                       uint64_t first, uint64_t count, bbmcu_loss** out);
/* The same with (a) a BATCH of n_materials measured tables sharing one linearizer and one launch (configs[4] of the
 * benchmark: many materials x many parameter sets per launch; reference_merl_rgb = array of n_materials table pointers;
 * an analytic reference_bsdf requires n_materials == 1), and (b) flags.  By default the kernels GENERATE each sample's
 * direction pair from its linearizer index in registers (merl_linearizer(idx), include/linearizer/merl_linearizer.h:50-83,
 * through a 900-float separable table staged in shared memory; spherical_linearizer(idx) directly): a pass then reads only
 * the 12 B per sample of tabulated reference data and a loss object costs 17.5 MB per material instead of 52.5 MB.  The
 * generated directions are bit-identical to bbmcu_merl_dirs / bbmcu_spherical_dirs.  BBMCU_LOSS_MATERIALISE_DIRECTIONS
 * keeps the direction planes in device memory and reads them instead (36 B per sample; same results bit for bit).
 * BBMCU_LOSS_SHARD_INTERLEAVED changes what (first, count) mean: first = rank, count = world, and the shard is every
 * world-th BLOCK of 1024 consecutive samples (blocks rank, rank + world, ...).  Samples of the MERL grid differ in cost by
 * region (pairs below the horizon leave the model early; the first eighth of the index range has none), so contiguous
 * eighths are unequal work and the slowest shard sets the pace of a multi-GPU step; interleaved shards are equal.
 * bbmcu_loss_terms then returns the shard's samples in its own (block-interleaved) order. */
enum { BBMCU_LOSS_MATERIALISE_DIRECTIONS = 1, BBMCU_LOSS_SHARD_INTERLEAVED = 2 };
BBMCU_API int  bbmcu_loss_create_ex(bbmcu_ctx* ctx, int metric, const bbmcu_spherical_grid* grid, int component, int unit,
                          const bbmcu_bsdf* reference_bsdf, const float* const* reference_merl_rgb, int n_materials,
                          uint64_t first, uint64_t count, unsigned flags, bbmcu_loss** out);
BBMCU_API void bbmcu_loss_free(bbmcu_loss* loss);
BBMCU_API uint64_t bbmcu_loss_samples(const bbmcu_loss* loss);      /* N of the linearizer (sampledlossfunction::samples) */
BBMCU_API uint64_t bbmcu_loss_shard_count(const bbmcu_loss* loss);  /* samples of this shard = floats bbmcu_loss_terms writes */
BBMCU_API int  bbmcu_loss_materials(const bbmcu_loss* loss);
/* the six metrics differ only in the per-sample functor: switching is free (the tabulated reference is metric independent) */
BBMCU_API int  bbmcu_loss_set_metric(bbmcu_loss* loss, int metric);
/* loss (and, if grad != NULL, d loss / d parameter) of `bsdf` at K parameter vectors.
 * params: K x P row-major, P = bbmcu_bsdf_param_count(bsdf, BBMCU_ATTR_ALL), forward enumeration order;
 * params == NULL with K == 1 evaluates the bsdf's current parameters.  loss: K doubles; grad: K x P doubles.
 * Host pointers.  device_out (may be NULL): if given, a DEVICE buffer of K*(1+P) doubles that receives
 * [loss_k, grad_k...] rows and no host copy/synchronisation is done (for an NCCL all-reduce by the caller, or - after
 * bbmcu_loss_peer_connect - already combined over the shards). */
BBMCU_API int  bbmcu_loss_eval(bbmcu_loss* loss, const bbmcu_bsdf* bsdf, const double* params, size_t K,
                     double* loss_out, double* grad_out, double* device_out);
/* K parameter vectors for EACH of the loss's M materials in one launch (the loop of include/bbm/sampledlossfunction.h:62-87
 * over M x K (reference, parameter) pairs): params M x K x P material-major (row m*K + k belongs to material m),
 * loss M x K, grad M x K x P, device_out M x K x (1+P). */
BBMCU_API int  bbmcu_loss_eval_multi(bbmcu_loss* loss, const bbmcu_bsdf* bsdf, const double* params, size_t K,
                           double* loss_out, double* grad_out, double* device_out);
/* gradient: 1 = compute it, 0 = value-only kernel (device_out rows then carry zeros in the gradient columns), -1 = decide from
 * the output pointers as bbmcu_loss_eval does (gradient iff grad_out or device_out is given) */
BBMCU_API int  bbmcu_loss_eval_multi_ex(bbmcu_loss* loss, const bbmcu_bsdf* bsdf, const double* params, size_t K,
                              double* loss_out, double* grad_out, double* device_out, int gradient);
/* per-sample terms l(idx) of the shard (sampledlossfunction::operator()(idx)); `terms` = bbmcu_loss_shard_count floats
 * (host or device); _at: against material `material` of a batched loss */
BBMCU_API int  bbmcu_loss_terms(bbmcu_loss* loss, const bbmcu_bsdf* bsdf, float* terms);
BBMCU_API int  bbmcu_loss_terms_at(bbmcu_loss* loss, const bbmcu_bsdf* bsdf, int material, float* terms);

/* ---- multi-GPU combine of the sample-axis shards over NVLink peer memory -------------------------------------------------
 * (no reference counterpart: bbm is single-threaded; this is the exchange step of SURVEY.md section 8e.)
 * One process (or context) per GPU of ONE node owns one shard of the same loss.  After the calls below every
 * bbmcu_loss_eval of that loss is a COLLECTIVE over the `world` shards - all of them must call it with the same bsdf shape
 * and K, in the same order - and returns the sum over the shards (= the reference's mean): the kernel that finishes a
 * shard's K x (1+P) totals stores each of them into every peer's exchange window as two 8-byte words that carry the batch
 * number next to the payload (plain remote stores over NVLink, nothing waits), and a gather kernel polls the `world` rows
 * of its own window until every word shows this batch and adds them in rank order.  No NCCL call, no fence, flag or host
 * round trip; every rank gets bit-identical results.  A peer that does not arrive within ~10 s makes the call fail
 * (BBMCU_RUNTIME_ERROR at the next synchronising call) instead of hanging the device.
 *   peer_init:     allocates this shard's window for batches of up to max_values = K*(1+P) doubles and returns its
 *                  cudaIpcMemHandle (64 bytes) for the other PROCESSES, and its device address for contexts of the same
 *                  process;
 *   peer_connect:  handles = world x 64 bytes in rank order (own entry ignored), as exchanged by the host's launcher
 *                  (torch.distributed all_gather, MPI, a pipe ...);
 *   peer_connect_ptrs: the same for shards living in one process (windows[r] = peer r's device address). */
BBMCU_API int  bbmcu_loss_peer_init(bbmcu_loss* loss, int rank, int world, size_t max_values, unsigned char handle_out[64], void** window_out);
BBMCU_API int  bbmcu_loss_peer_connect(bbmcu_loss* loss, const unsigned char* handles);
BBMCU_API int  bbmcu_loss_peer_connect_ptrs(bbmcu_loss* loss, void* const* windows);

/* ---- .fit files (include/io/fit.h:34-77) -------------------------------------------------------------- */
typedef struct bbmcu_fit bbmcu_fit;
BBMCU_API int  bbmcu_fit_import(bbmcu_ctx* ctx, const char* filename, bbmcu_fit** out);
BBMCU_API int  bbmcu_fit_import_ex(bbmcu_ctx* ctx, const char* filename, int config, bbmcu_fit** out);   /* config: BBMCU_FLOAT_RGB / BBMCU_DOUBLE_RGB */
BBMCU_API int  bbmcu_fit_count(const bbmcu_fit* fit);
BBMCU_API const char* bbmcu_fit_key(const bbmcu_fit* fit, int i);
BBMCU_API int  bbmcu_fit_bsdf(const bbmcu_fit* fit, int i, bbmcu_bsdf** out);     /* a copy the caller frees */
BBMCU_API int  bbmcu_fit_create(bbmcu_fit** out);
BBMCU_API int  bbmcu_fit_add(bbmcu_fit* fit, const char* key, const bbmcu_bsdf* bsdf);
BBMCU_API int  bbmcu_fit_export(bbmcu_ctx* ctx, const bbmcu_fit* fit, const char* filename, const char* comment);
BBMCU_API void bbmcu_fit_free(bbmcu_fit* fit);

/* ---- checkBsdf on the device (bin/checkBsdf.cpp:51-430) -----------------------------------------------------------------
 * The reference's six consistency tests as fused generate -> evaluate -> reduce kernels: every sample of the reference's
 * scalar loops is one item, the per-sample terms are summed on the device (float terms, double sums).  `rng` selects
 * where the random numbers come from:
 *   BBMCU_RNG_PHILOX   drawn in the kernel from a counter-based generator of `seed` (0 bytes in; any sample count);
 *   BBMCU_RNG_MT19937  the reference's own stream - a default-seeded std::mt19937 read through
 *                      std::uniform_real_distribution<float>(0,1) in the reference's call order (bin/checkBsdf.cpp:18-25;
 *                      rndVec2d draws its second component first) - drawn on the host and uploaded, so the printed numbers
 *                      can be compared with the reference's for the same command line (`seed` ignored).
 * Sums are double where the reference accumulates in float, so agreement is to the float accumulation error of the
 * reference (~1e-4 relative at 10^5 samples), not bit for bit.  All output pointers are HOST memory. */
enum { BBMCU_RNG_PHILOX = 0, BBMCU_RNG_MT19937 = 1 };
/* testReflectance (:51-97): for out = (theta_idx * pi/2 / n_theta, phi = 0): estimate[3 i ..] = mean of eval * cos / pdf over
 * `samples` directions (uniform over the sphere, or bsdf.sample(out, xi) if importance), reflectance[3 i ..] = bsdf.reflectance(out),
 * out_dirs[3 i ..] = out */
BBMCU_API int  bbmcu_check_reflectance(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int n_theta, int importance,
                                       int rng, uint64_t seed, double* estimate, float* reflectance, float* out_dirs);
/* testReciprocity (:102-152) and testAdjoint (:157-201): mean[3] and max_diff[3] of |eval(a, b) - eval(b, a)| over uniform
 * direction pairs, max_pair[6] = (a, b) of the first pair with the largest channel sum (zeros if every difference is 0).
 * Radiance, importance and adjoint give the same numbers: no model of the reference reads unit_t. */
BBMCU_API int  bbmcu_check_reciprocity(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int rng, uint64_t seed,
                                       double* mean, float* max_diff, float* max_pair);
/* testPdf (:206-267): counts[4] = negative pdfs (radiance, importance), sampled directions below the horizon (radiance,
 * importance; 0 unless check_below_horizon); mismatch[2] = mean |sample.pdf - pdf(sample.direction, view)|.  The whole
 * sample count is evaluated (the reference stops once a counter reaches maxError).  offenders (may be NULL): up to
 * max_offenders records of 8 floats {kind (0/1 below horizon rad/imp, 2/3 negative pdf rad/imp), pdf, direction[3], view[3]};
 * n_offenders receives how many were written. */
BBMCU_API int  bbmcu_check_pdf(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int sample_sphere, int check_below_horizon,
                               int rng, uint64_t seed, uint64_t* counts, double* mismatch, float* offenders, int max_offenders, int* n_offenders);
/* testPdfInt (:272-316): integral[t] = MC integral over the sphere of pdf(., dirs[3 t ..]) for `trials` random directions */
BBMCU_API int  bbmcu_check_pdf_integral(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t samples, int trials, int sample_sphere,
                                        int rng, uint64_t seed, double* integral, float* dirs);
/* testSample (:321-430): per trial, the pdf integrated over n_theta x n_phi bins (pdf_samples per bin), the histogram of
 * `samples` sampled directions, chi2 / df as the reference forms them and P = gamma_q((df - 1)/2, chi2/2) (NaN when df <= 1).
 * bin_pdf / bin_count (may be NULL): trials x n_theta x n_phi. */
BBMCU_API int  bbmcu_check_sample(bbmcu_ctx* ctx, const bbmcu_bsdf* bsdf, uint64_t pdf_samples, uint64_t samples, int n_theta, int n_phi,
                                  int trials, int sample_sphere, int include_zero_pdf_samples, int rng, uint64_t seed,
                                  double* chi2, double* df, double* P, float* dirs, double* bin_pdf, uint64_t* bin_count);

#ifdef __cplusplus
}
#endif
#endif /* BBMCU_H_ */

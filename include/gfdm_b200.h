/*
 * gfdm_b200.h -- C ABI of the B200-native GFDM baseband engine.
 *
 * This is the drop-in boundary for gr-gfdm's GNU-Radio-free kernel layer.  Each
 * group of entry points replaces one reference class; the reference interface
 * it replaces is cited as <file>:<line> (paths relative to the reference tree).
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Conventions
 *  - gfdm_complex is layout-identical to std::complex<float> / fftwf_complex.
 *  - Every function returning `int` returns a gfdm_status; on failure
 *    gfdm_last_error() holds the text the reference would have thrown
 *    (GFDM_ERR_INVALID_ARGUMENT <-> std::invalid_argument,
 *     GFDM_ERR_RUNTIME <-> std::runtime_error).
 *  - Per-frame entry points take HOST pointers and return when the result is
 *    in `out` (same contract as the reference's generic_work()).
 *  - `*_batch` entry points process `n_frames` contiguous frames
 *    (frame f uses in + f*in_size, out + f*out_size -- the loop the reference's
 *    GNU Radio blocks run, e.g. lib/simple_modulator_cc_impl.cc:72-76).
 *    `mem` selects where the pointers live: GFDM_MEM_HOST (copied through
 *    pinned staging, synchronous) or GFDM_MEM_DEVICE (zero-copy, asynchronous
 *    on the handle's stream; call gfdm_sync()).
 *  - One handle <-> one CUDA stream; a handle is not thread-safe (same rule as
 *    the reference, whose kernels own member scratch buffers).
 *  - The same ABI is exported by the test oracles under oracle/ (CPU only,
 *    GFDM_MEM_DEVICE rejected) so that parity tests drive all three through
 *    identical calls.  The product library never falls back to a CPU path.
 */
#ifndef INCLUDED_GFDM_B200_H
#define INCLUDED_GFDM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GFDM_B200_API
#else
#define GFDM_B200_API __attribute__((visibility("default")))
#endif

typedef struct gfdm_complex {
    float re;
    float im;
} gfdm_complex;

typedef enum gfdm_status {
    GFDM_OK = 0,
    GFDM_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument in the reference */
    GFDM_ERR_RUNTIME = 2,          /* std::runtime_error in the reference bindings */
    GFDM_ERR_CUDA = 3,             /* CUDA runtime/driver failure, no device, ... */
    GFDM_ERR_UNSUPPORTED = 4       /* e.g. GFDM_MEM_DEVICE on a CPU oracle */
} gfdm_status;

typedef enum gfdm_mem {
    GFDM_MEM_HOST = 0,
    GFDM_MEM_DEVICE = 1,
    /* HOST pointers, asynchronous: the call returns once the chunked copies and kernels of the batch are enqueued
     * (host->device on one copy stream, kernels on the handle's stream, device->host on a second copy stream); the
     * output is valid -- and the input may be reused -- after gfdm_sync(handle).  Buffers MUST be page-locked
     * (cudaHostAlloc / cudaHostRegister), otherwise the copies degrade to synchronous ones.  Lets a caller overlap the
     * device->host leg of one handle (a modulator) with the host->device leg of another (a receiver), i.e. use both
     * PCIe directions at once.  Accepted by the pipelined batch entries only: gfdm_modulator_work[_chunks]_batch[_sc16],
     * gfdm_receiver_*_batch[_sc16], gfdm_receiver_work_decide_batch[_sc16], gfdm_advanced_receiver_work_batch,
     * gfdm_transmitter_work_chunks_batch, gfdm_remove_prefix_work_batch, gfdm_resource_mapper_demap_chunks_batch. */
    GFDM_MEM_HOST_ASYNC = 2
} gfdm_mem;

/* ---- library-wide ------------------------------------------------------- */
GFDM_B200_API const char* gfdm_last_error(void);  /* thread-local, never NULL */
GFDM_B200_API const char* gfdm_backend(void);     /* "cuda-sm_100a" | "oracle-port" | "oracle-ref" */
GFDM_B200_API int gfdm_device_count(void);        /* 0 on the CPU oracles */
GFDM_B200_API int gfdm_set_device(int device);    /* device used by handles created afterwards */
/* Any handle type below may be passed.  `cuda_stream` is a cudaStream_t. */
GFDM_B200_API int gfdm_set_stream(void* handle, void* cuda_stream);
GFDM_B200_API int gfdm_sync(void* handle);
/* number of CUDA kernels this handle has launched so far (0 on the oracles) */
GFDM_B200_API long long gfdm_launch_count(void* handle);
/* name of the kernel variant the last call on this handle dispatched to */
GFDM_B200_API const char* gfdm_last_kernel(void* handle);

/* gfdm_kernel_utils::calculate_signal_energy -- lib/gfdm_kernel_utils.cc:59-65 */
GFDM_B200_API int gfdm_calculate_signal_energy(float* energy, const gfdm_complex* in, int n);

/* Unnormalised c2c DFT, the engine that replaces gfdm_kernel_utils::initialize_fft
 * + fftwf_execute (lib/gfdm_kernel_utils.cc:32-57).  forward!=0: exp(-j..). */
typedef struct gfdm_fft gfdm_fft;
GFDM_B200_API int gfdm_fft_create(gfdm_fft** out, int fft_size, int forward);
GFDM_B200_API void gfdm_fft_destroy(gfdm_fft* h);
GFDM_B200_API int gfdm_fft_execute_batch(gfdm_fft* h, gfdm_complex* out, const gfdm_complex* in,
                                         int n_transforms, int mem);

/* ---- modulator_kernel_cc -- include/gfdm/modulator_kernel_cc.h:41-51,
 *      lib/modulator_kernel_cc.cc:30-141 --------------------------------- */
typedef struct gfdm_modulator gfdm_modulator;
GFDM_B200_API int gfdm_modulator_create(gfdm_modulator** out, int n_timeslots, int n_subcarriers,
                                        int overlap, const gfdm_complex* frequency_taps, int n_taps);
GFDM_B200_API void gfdm_modulator_destroy(gfdm_modulator* h);
GFDM_B200_API int gfdm_modulator_block_size(const gfdm_modulator* h);
GFDM_B200_API int gfdm_modulator_filter_taps(const gfdm_modulator* h, gfdm_complex* taps_out); /* M*L, normalised */
GFDM_B200_API int gfdm_modulator_work(gfdm_modulator* h, gfdm_complex* out, const gfdm_complex* in); /* generic_work */
GFDM_B200_API int gfdm_modulator_work_batch(gfdm_modulator* h, gfdm_complex* out, const gfdm_complex* in,
                                            int n_frames, int mem);

/* ---- receiver_kernel_cc -- include/gfdm/receiver_kernel_cc.h:53-89,
 *      lib/receiver_kernel_cc.cc:31-334 ----------------------------------- */
typedef struct gfdm_receiver gfdm_receiver;
GFDM_B200_API int gfdm_receiver_create(gfdm_receiver** out, int n_timeslots, int n_subcarriers,
                                       int overlap, const gfdm_complex* frequency_taps, int n_taps);
GFDM_B200_API void gfdm_receiver_destroy(gfdm_receiver* h);
GFDM_B200_API int gfdm_receiver_block_size(const gfdm_receiver* h);
GFDM_B200_API int gfdm_receiver_timeslots(const gfdm_receiver* h);
GFDM_B200_API int gfdm_receiver_subcarriers(const gfdm_receiver* h);
GFDM_B200_API int gfdm_receiver_overlap(const gfdm_receiver* h);
GFDM_B200_API int gfdm_receiver_filter_taps(const gfdm_receiver* h, gfdm_complex* taps_out);    /* M*L */
GFDM_B200_API int gfdm_receiver_ic_filter_taps(const gfdm_receiver* h, gfdm_complex* taps_out); /* M */
/* generic_work / generic_work_equalize (:322-334) */
GFDM_B200_API int gfdm_receiver_work(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in);
GFDM_B200_API int gfdm_receiver_work_equalize(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                              const gfdm_complex* f_eq_in);
/* f_eq_in == NULL selects the non-equalising path; f_eq advances per frame
 * (lib/advanced_receiver_sb_cc_impl.cc:98-104). */
GFDM_B200_API int gfdm_receiver_work_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                           const gfdm_complex* f_eq_in, int n_frames, int mem);
/* the separately callable stages (:301-320, :211-225, :274-299) */
GFDM_B200_API int gfdm_receiver_fft_filter_downsample(gfdm_receiver* h, gfdm_complex* out,
                                                      const gfdm_complex* in);
GFDM_B200_API int gfdm_receiver_fft_equalize_filter_downsample(gfdm_receiver* h, gfdm_complex* out,
                                                               const gfdm_complex* in,
                                                               const gfdm_complex* f_eq_in);
GFDM_B200_API int gfdm_receiver_fft_filter_downsample_batch(gfdm_receiver* h, gfdm_complex* out,
                                                            const gfdm_complex* in,
                                                            const gfdm_complex* f_eq_in, int n_frames,
                                                            int mem);
GFDM_B200_API int gfdm_receiver_transform_subcarriers_to_td(gfdm_receiver* h, gfdm_complex* out,
                                                            const gfdm_complex* in);
GFDM_B200_API int gfdm_receiver_transform_subcarriers_to_td_batch(gfdm_receiver* h, gfdm_complex* out,
                                                                  const gfdm_complex* in, int n_frames,
                                                                  int mem);
GFDM_B200_API int gfdm_receiver_cancel_sc_interference(gfdm_receiver* h, gfdm_complex* out,
                                                       const gfdm_complex* td_in,
                                                       const gfdm_complex* fd_in);
GFDM_B200_API int gfdm_receiver_cancel_sc_interference_batch(gfdm_receiver* h, gfdm_complex* out,
                                                             const gfdm_complex* td_in,
                                                             const gfdm_complex* fd_in, int n_frames,
                                                             int mem);

/* ---- advanced_receiver_kernel_cc -- include/gfdm/advanced_receiver_kernel_cc.h:37-61,
 *      lib/advanced_receiver_kernel_cc.cc:32-123 -------------------------- */
/* GNU-Radio-free stand-in for gr::digital::constellation_sptr (points() +
 * decision_maker(), call sites lib/advanced_receiver_kernel_cc.cc:114-120). */
typedef enum gfdm_decision_rule {
    GFDM_DECISION_NEAREST = 0,   /* argmin |s - p_i|^2, first minimum wins */
    GFDM_DECISION_QPSK_SIGN = 1  /* idx = 2*(im > 0) + (re > 0)  (gr::digital::constellation_qpsk) */
} gfdm_decision_rule;
typedef struct gfdm_constellation {
    const gfdm_complex* points;
    int n_points;
    int decision_rule; /* gfdm_decision_rule */
} gfdm_constellation;

typedef struct gfdm_advanced_receiver gfdm_advanced_receiver;
GFDM_B200_API int gfdm_advanced_receiver_create(gfdm_advanced_receiver** out, int timeslots,
                                                int subcarriers, int overlap,
                                                const gfdm_complex* frequency_taps, int n_taps,
                                                const int* subcarrier_map, int n_map, int ic_iter,
                                                const gfdm_constellation* constellation,
                                                int do_phase_compensation);
GFDM_B200_API void gfdm_advanced_receiver_destroy(gfdm_advanced_receiver* h);
GFDM_B200_API int gfdm_advanced_receiver_block_size(const gfdm_advanced_receiver* h);
GFDM_B200_API int gfdm_advanced_receiver_set_ic(gfdm_advanced_receiver* h, int ic_iter);
GFDM_B200_API int gfdm_advanced_receiver_get_ic(const gfdm_advanced_receiver* h);
GFDM_B200_API int gfdm_advanced_receiver_set_phase_compensation(gfdm_advanced_receiver* h, int v);
GFDM_B200_API int gfdm_advanced_receiver_get_phase_compensation(const gfdm_advanced_receiver* h);
GFDM_B200_API int gfdm_advanced_receiver_work(gfdm_advanced_receiver* h, gfdm_complex* out,
                                              const gfdm_complex* in);
GFDM_B200_API int gfdm_advanced_receiver_work_equalize(gfdm_advanced_receiver* h, gfdm_complex* out,
                                                       const gfdm_complex* in,
                                                       const gfdm_complex* f_eq_in);
GFDM_B200_API int gfdm_advanced_receiver_work_batch(gfdm_advanced_receiver* h, gfdm_complex* out,
                                                    const gfdm_complex* in,
                                                    const gfdm_complex* f_eq_in, int n_frames,
                                                    int mem);

/* ---- resource_mapper_kernel_cc -- include/gfdm/resource_mapper_kernel_cc.h:38-58,
 *      lib/resource_mapper_kernel_cc.cc:30-162 ---------------------------- */
typedef struct gfdm_resource_mapper gfdm_resource_mapper;
GFDM_B200_API int gfdm_resource_mapper_create(gfdm_resource_mapper** out, int timeslots,
                                              int subcarriers, int active_subcarriers,
                                              const int* subcarrier_map, int n_map, int per_timeslot,
                                              int is_mapper);
GFDM_B200_API void gfdm_resource_mapper_destroy(gfdm_resource_mapper* h);
GFDM_B200_API size_t gfdm_resource_mapper_frame_size(const gfdm_resource_mapper* h);
GFDM_B200_API size_t gfdm_resource_mapper_block_size(const gfdm_resource_mapper* h);
GFDM_B200_API size_t gfdm_resource_mapper_input_vector_size(const gfdm_resource_mapper* h);
GFDM_B200_API size_t gfdm_resource_mapper_output_vector_size(const gfdm_resource_mapper* h);
GFDM_B200_API int gfdm_resource_mapper_map_to_resources(gfdm_resource_mapper* h, gfdm_complex* out,
                                                        const gfdm_complex* in, size_t ninput_size);
GFDM_B200_API int gfdm_resource_mapper_demap_from_resources(gfdm_resource_mapper* h, gfdm_complex* out,
                                                            const gfdm_complex* in,
                                                            size_t noutput_size);
/* batch: every frame carries `size_per_frame` symbols (in stride for map, out stride for demap) */
GFDM_B200_API int gfdm_resource_mapper_map_to_resources_batch(gfdm_resource_mapper* h,
                                                              gfdm_complex* out,
                                                              const gfdm_complex* in,
                                                              size_t size_per_frame, int n_frames,
                                                              int mem);
GFDM_B200_API int gfdm_resource_mapper_demap_from_resources_batch(gfdm_resource_mapper* h,
                                                                  gfdm_complex* out,
                                                                  const gfdm_complex* in,
                                                                  size_t size_per_frame,
                                                                  int n_frames, int mem);

/* ---- add_cyclic_prefix_cc -- include/gfdm/add_cyclic_prefix_cc.h:38-57,
 *      lib/add_cyclic_prefix_cc.cc:30-104 --------------------------------- */
typedef struct gfdm_cyclic_prefixer gfdm_cyclic_prefixer;
GFDM_B200_API int gfdm_cyclic_prefixer_create(gfdm_cyclic_prefixer** out, int block_len, int cp_len,
                                              int cs_len, int ramp_len,
                                              const gfdm_complex* window_taps, int n_window_taps,
                                              int cyclic_shift);
GFDM_B200_API void gfdm_cyclic_prefixer_destroy(gfdm_cyclic_prefixer* h);
GFDM_B200_API int gfdm_cyclic_prefixer_block_size(const gfdm_cyclic_prefixer* h);
GFDM_B200_API int gfdm_cyclic_prefixer_frame_size(const gfdm_cyclic_prefixer* h);
GFDM_B200_API int gfdm_cyclic_prefixer_cyclic_shift(const gfdm_cyclic_prefixer* h);
GFDM_B200_API int gfdm_cyclic_prefixer_work(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                            const gfdm_complex* in); /* generic_work */
GFDM_B200_API int gfdm_cyclic_prefixer_add_cyclic_prefix(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                                         const gfdm_complex* in, int cyclic_shift);
GFDM_B200_API int gfdm_cyclic_prefixer_remove_cyclic_prefix(gfdm_cyclic_prefixer* h, gfdm_complex* out,
                                                            const gfdm_complex* in);
GFDM_B200_API int gfdm_cyclic_prefixer_add_cyclic_prefix_batch(gfdm_cyclic_prefixer* h,
                                                               gfdm_complex* out,
                                                               const gfdm_complex* in,
                                                               int cyclic_shift, int n_frames,
                                                               int mem);
GFDM_B200_API int gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(gfdm_cyclic_prefixer* h,
                                                                  gfdm_complex* out,
                                                                  const gfdm_complex* in,
                                                                  int n_frames, int mem);

/* ---- preamble_channel_estimator_cc -- include/gfdm/preamble_channel_estimator_cc.h:44-77,
 *      lib/preamble_channel_estimator_cc.cc:34-294 ------------------------ */
typedef struct gfdm_channel_estimator gfdm_channel_estimator;
GFDM_B200_API int gfdm_channel_estimator_create(gfdm_channel_estimator** out, int timeslots,
                                                int fft_len, int active_subcarriers, int is_dc_free,
                                                int which_estimator, const gfdm_complex* preamble,
                                                int n_preamble /* >= 2*fft_len */);
GFDM_B200_API void gfdm_channel_estimator_destroy(gfdm_channel_estimator* h);
GFDM_B200_API int gfdm_channel_estimator_fft_len(const gfdm_channel_estimator* h);
GFDM_B200_API int gfdm_channel_estimator_timeslots(const gfdm_channel_estimator* h);
GFDM_B200_API int gfdm_channel_estimator_frame_len(const gfdm_channel_estimator* h);
GFDM_B200_API int gfdm_channel_estimator_active_subcarriers(const gfdm_channel_estimator* h);
GFDM_B200_API int gfdm_channel_estimator_is_dc_free(const gfdm_channel_estimator* h);
GFDM_B200_API int gfdm_channel_estimator_preamble_filter_taps(const gfdm_channel_estimator* h,
                                                              float* taps_out /* 9 */);
/* sub-steps (:121-132, :145-185, :238-274, :276-282) */
GFDM_B200_API int gfdm_channel_estimator_estimate_preamble_channel(gfdm_channel_estimator* h,
                                                                   gfdm_complex* fd_preamble_channel,
                                                                   const gfdm_complex* rx_preamble);
GFDM_B200_API int gfdm_channel_estimator_filter_preamble_estimate(gfdm_channel_estimator* h,
                                                                  gfdm_complex* filtered,
                                                                  const gfdm_complex* estimate);
GFDM_B200_API int gfdm_channel_estimator_interpolate_frame(gfdm_channel_estimator* h,
                                                           gfdm_complex* frame_estimate,
                                                           const gfdm_complex* estimate);
GFDM_B200_API int gfdm_channel_estimator_prepare_for_zf(gfdm_channel_estimator* h,
                                                        gfdm_complex* transformed_frame,
                                                        const gfdm_complex* frame_estimate);
/* estimate_frame (:285-294): in 2*fft_len samples -> out frame_len bins.
 * Bins the reference never writes (is_dc_free == 0) are left untouched. */
GFDM_B200_API int gfdm_channel_estimator_estimate_frame(gfdm_channel_estimator* h,
                                                        gfdm_complex* frame_estimate,
                                                        const gfdm_complex* rx_preamble);
GFDM_B200_API int gfdm_channel_estimator_estimate_frame_batch(gfdm_channel_estimator* h,
                                                              gfdm_complex* frame_estimate,
                                                              const gfdm_complex* rx_preamble,
                                                              int n_frames, int mem);
/* estimate_snr (:187-235): snr_lin[1], cnrs[active_subcarriers] (cnrs may be NULL) */
GFDM_B200_API int gfdm_channel_estimator_estimate_snr(gfdm_channel_estimator* h, float* snr_lin,
                                                      float* cnrs, const gfdm_complex* rx_preamble);
GFDM_B200_API int gfdm_channel_estimator_estimate_snr_batch(gfdm_channel_estimator* h, float* snr_lin,
                                                            float* cnrs,
                                                            const gfdm_complex* rx_preamble,
                                                            int n_frames, int mem);

/* ---- transmitter_kernel -- include/gfdm/transmitter_kernel.h:43-69,
 *      lib/transmitter_kernel.cc:34-107 ----------------------------------- */
typedef struct gfdm_transmitter gfdm_transmitter;
GFDM_B200_API int gfdm_transmitter_create(gfdm_transmitter** out, int timeslots, int subcarriers,
                                          int active_subcarriers, int cp_len, int cs_len, int ramp_len,
                                          const int* subcarrier_map, int n_map, int per_timeslot,
                                          int overlap, const gfdm_complex* frequency_taps, int n_taps,
                                          const gfdm_complex* window_taps, int n_window_taps,
                                          const int* cyclic_shifts, int n_cyclic_shifts,
                                          const gfdm_complex* const* preambles,
                                          const int* preamble_sizes, int n_preambles);
GFDM_B200_API void gfdm_transmitter_destroy(gfdm_transmitter* h);
GFDM_B200_API int gfdm_transmitter_input_vector_size(const gfdm_transmitter* h);
GFDM_B200_API int gfdm_transmitter_output_vector_size(const gfdm_transmitter* h);
GFDM_B200_API int gfdm_transmitter_n_cyclic_shifts(const gfdm_transmitter* h);
GFDM_B200_API int gfdm_transmitter_cyclic_shifts(const gfdm_transmitter* h, int* shifts_out);
GFDM_B200_API int gfdm_transmitter_work(gfdm_transmitter* h, gfdm_complex* out, const gfdm_complex* in,
                                        int ninput_size); /* generic_work :101-107 */
GFDM_B200_API int gfdm_transmitter_modulate(gfdm_transmitter* h, gfdm_complex* out,
                                            const gfdm_complex* in, int ninput_size); /* :78-84 */
GFDM_B200_API int gfdm_transmitter_add_frame(gfdm_transmitter* h, gfdm_complex* out,
                                             const gfdm_complex* in, int cyclic_shift); /* :92-98 */
/* frames of cyclic_shifts[0] only: out[n_frames][output_vector_size] */
GFDM_B200_API int gfdm_transmitter_work_batch(gfdm_transmitter* h, gfdm_complex* out,
                                              const gfdm_complex* in, int ninput_size, int n_frames,
                                              int mem);
/* all antennas, the loop of lib/transmitter_cc_impl.cc:165-177:
 * out[n_cyclic_shifts][n_frames][output_vector_size] */
GFDM_B200_API int gfdm_transmitter_work_all_batch(gfdm_transmitter* h, gfdm_complex* out,
                                                  const gfdm_complex* in, int ninput_size,
                                                  int n_frames, int mem);
/* The CUDA library runs mapper + modulator + preamble + prefixer as ONE kernel where the shape allows
 * (8*(n_in + P+cp+N+cs) bytes of HBM traffic per frame and antenna); on = 0 forces the four separate
 * kernels (used by the parity tests: both forms must agree bit for bit).  No-op on the CPU oracles. */
GFDM_B200_API int gfdm_transmitter_set_chain_fusion(gfdm_transmitter* h, int on);

/* ======================================================================== */
/* The rows either side of the hot path.  The reference implements
 * them as GNU Radio blocks / stock gr-digital blocks / pygfdm helpers; here they are batched kernels fed by
 * plain arrays instead of stream tags.                                                                  */

/* ---- remove_prefix_cc -- lib/remove_prefix_cc_impl.cc:84-115 ------------
 * out[f][0:block_len] = in[f][offset : offset+block_len], frames of frame_len samples
 * (constructor: include/gfdm/remove_prefix_cc.h, make(frame_len, block_len, offset, ...)). */
typedef struct gfdm_remove_prefix gfdm_remove_prefix;
GFDM_B200_API int gfdm_remove_prefix_create(gfdm_remove_prefix** out, int frame_len, int block_len,
                                            int offset);
GFDM_B200_API void gfdm_remove_prefix_destroy(gfdm_remove_prefix* h);
GFDM_B200_API int gfdm_remove_prefix_work_batch(gfdm_remove_prefix* h, gfdm_complex* out,
                                                const gfdm_complex* in, int n_frames, int mem);

/* ---- extract_burst_cc -- lib/extract_burst_cc_impl.cc:43-242 ------------
 * One general_work() call (:117-242) over a window of `n_in` stream samples.  The stream tags become
 * parallel HOST arrays sorted by offset (the reference sorts them, :131): burst_starts[i] = tag offset
 * relative to in[0]; scale_factors[i] = the tag's "scale_factor" (NULL: 1.0, :78-86);
 * phase_rotations[i] = the tag's "phase_rotation" (NULL: 1+0j, :88-96).  For every tag, in order:
 *   actual_start = burst_start - tag_backoff;
 *   if (n_in - burst_start >= burst_len && produced + burst_len <= max_bursts*burst_len)
 *        out burst = scale * in[actual_start : actual_start+burst_len]   (negative start: zeros in front
 *        and the burst is cut short, :196-205); with CFO correction on, sample i of the burst is rotated
 *        by inc^i, inc = conj(phase_rotation)/|phase_rotation| (:88-96,107-115);
 *        consumed = burst_start + burst_len;
 *   else consumed = max(0, burst_start) and the loop stops (:224-238).
 * n_produced = bursts written to out[n_produced][burst_len]; n_consumed = items the block would consume
 * (n_in when every tag was served).  `mem` applies to in/out only.                                     */
typedef struct gfdm_extract_burst gfdm_extract_burst;
GFDM_B200_API int gfdm_extract_burst_create(gfdm_extract_burst** out, int burst_len, int tag_backoff,
                                            int activate_cfo_correction);
GFDM_B200_API void gfdm_extract_burst_destroy(gfdm_extract_burst* h);
GFDM_B200_API int gfdm_extract_burst_activate_cfo_compensation(gfdm_extract_burst* h, int on); /* :98-105 */
GFDM_B200_API int gfdm_extract_burst_work(gfdm_extract_burst* h, gfdm_complex* out, int max_bursts,
                                          const gfdm_complex* in, long long n_in,
                                          const long long* burst_starts, const float* scale_factors,
                                          const gfdm_complex* phase_rotations, int n_tags,
                                          int* n_produced, long long* n_consumed, int mem);

/* ---- symbol mapping -- python/pygfdm/symbolmapping.py:27-47, python/pygfdm/utils.py:47-51,
 *      gr::digital chunks_to_symbols_bc / constellation decoder as used by
 *      python/qa_advanced_receiver_sb_cc.py:97-99 ---------------------------------------------------
 * chunk = index of a constellation point (one unsigned char per symbol); bits = one unsigned char (0/1)
 * per bit, MSB first, log2(n_points) bits per symbol (pack_bits :27-31, unpackbits :44).
 * map: out = points[chunk] (chunk >= n_points gives 0+0j); decide: the constellation's decision rule. */
typedef struct gfdm_symbol_mapper gfdm_symbol_mapper;
GFDM_B200_API int gfdm_symbol_mapper_create(gfdm_symbol_mapper** out, const gfdm_constellation* constellation);
GFDM_B200_API void gfdm_symbol_mapper_destroy(gfdm_symbol_mapper* h);
GFDM_B200_API int gfdm_symbol_mapper_n_points(const gfdm_symbol_mapper* h);
GFDM_B200_API int gfdm_symbol_mapper_bits_per_symbol(const gfdm_symbol_mapper* h); /* 0 unless n_points = 2^b */
GFDM_B200_API int gfdm_symbol_mapper_points(const gfdm_symbol_mapper* h, gfdm_complex* points_out);
GFDM_B200_API int gfdm_symbol_mapper_decision_rule(const gfdm_symbol_mapper* h);
GFDM_B200_API int gfdm_symbol_mapper_map_chunks_batch(gfdm_symbol_mapper* h, gfdm_complex* out,
                                                      const unsigned char* chunks, size_t n_symbols, int mem);
GFDM_B200_API int gfdm_symbol_mapper_decide_batch(gfdm_symbol_mapper* h, unsigned char* chunks_out,
                                                  const gfdm_complex* in, size_t n_symbols, int mem);
GFDM_B200_API int gfdm_symbol_mapper_bits2symbols_batch(gfdm_symbol_mapper* h, gfdm_complex* out,
                                                        const unsigned char* bits, size_t n_symbols, int mem);
GFDM_B200_API int gfdm_symbol_mapper_symbols2bits_batch(gfdm_symbol_mapper* h, unsigned char* bits_out,
                                                        const gfdm_complex* in, size_t n_symbols, int mem);

/* The same mapping fused into the kernels of the path, so that the symbol side of a frame crosses HBM
 * (and PCIe) as one byte per symbol instead of eight:
 *   modulator   <- chunks[n_frames][block_size]                       (8N+N instead of 16N bytes per frame)
 *   transmitter <- chunks[n_frames][ninput_size] (compact, as map_to_resources takes them)
 *   receiver    -> chunks[n_frames][block_size]  = decide(generic_work[_equalize] output), full grid
 *   demapper    on chunk grids (demap_from_resources on bytes) completes the receive chain.        */
GFDM_B200_API int gfdm_modulator_work_chunks_batch(gfdm_modulator* h, const gfdm_symbol_mapper* sm,
                                                   gfdm_complex* out, const unsigned char* chunks,
                                                   int n_frames, int mem);
GFDM_B200_API int gfdm_transmitter_work_chunks_batch(gfdm_transmitter* h, const gfdm_symbol_mapper* sm,
                                                     gfdm_complex* out, const unsigned char* chunks,
                                                     int ninput_size, int n_frames, int mem);
GFDM_B200_API int gfdm_receiver_work_decide_batch(gfdm_receiver* h, const gfdm_symbol_mapper* sm,
                                                  unsigned char* chunks_out, const gfdm_complex* in,
                                                  const gfdm_complex* f_eq_in, int n_frames, int mem);
GFDM_B200_API int gfdm_resource_mapper_demap_chunks_batch(gfdm_resource_mapper* h, unsigned char* out,
                                                          const unsigned char* in, size_t size_per_frame,
                                                          int n_frames, int mem);

/* remove_prefix_cc (lib/remove_prefix_cc_impl.cc:84-115) fused into the receiver's loads: frame f of the batch is the
 * block_size samples at in + f*in_stride + in_offset (in_stride >= in_offset + block_size), i.e. the frames still carry
 * preamble / cyclic prefix / suffix and are never copied to a packed array (single-pass shapes read them in place;
 * elsewhere the gather kernel runs first).  f_eq_in (may be NULL) and out are packed [n_frames][block_size]. */
GFDM_B200_API int gfdm_receiver_work_strided_batch(gfdm_receiver* h, gfdm_complex* out, const gfdm_complex* in,
                                                   const gfdm_complex* f_eq_in, size_t in_stride, size_t in_offset,
                                                   int n_frames, int mem);

/* ---- short_burst_shaper: lib/short_burst_shaper_impl.cc:57-84 (ctor checks), :161-182 (work) ---------------------
 * The sample path of the block that follows the transmitter in the reference's flowgraphs: every burst becomes
 *   [pre_padding zeros | in * scale | post_padding zeros]      (volk_32fc_s32fc_multiply_32fc: plain complex product).
 * The block's timed-command / message handling (:185-230) is radio control, not sample processing, and is out of scope.
 * gfdm_transmitter_work_shaped_batch runs the transmitter chain with this step as the EPILOGUE of the same kernel
 * (single-pass shapes; composition of the two elsewhere):
 *   out [n_ant][n_frames][pre + output_vector_size + post], n_ant = all_antennas ? n_cyclic_shifts : 1. */
typedef struct gfdm_burst_shaper gfdm_burst_shaper;
GFDM_B200_API int gfdm_burst_shaper_create(gfdm_burst_shaper** out, int pre_padding, int post_padding, float scale_re,
                                           float scale_im);
GFDM_B200_API void gfdm_burst_shaper_destroy(gfdm_burst_shaper* h);
GFDM_B200_API int gfdm_burst_shaper_pre_padding(const gfdm_burst_shaper* h);
GFDM_B200_API int gfdm_burst_shaper_post_padding(const gfdm_burst_shaper* h);
GFDM_B200_API int gfdm_burst_shaper_work_batch(gfdm_burst_shaper* h, gfdm_complex* out, const gfdm_complex* in,
                                               int burst_len, int n_bursts, int mem);
GFDM_B200_API int gfdm_transmitter_work_shaped_batch(gfdm_transmitter* h, const gfdm_burst_shaper* shaper,
                                                     gfdm_complex* out, const gfdm_complex* in, int ninput_size,
                                                     int n_frames, int all_antennas, int mem);

/* ---- sc16 sample format on the host side of a batch ------------------------------------------------------
 * Time-domain samples cross the host<->device link as interleaved int16 I/Q ("sc16", the wire format of the SDR front
 * ends the reference's flowgraphs feed, e.g. UHD's sc16) instead of complex64: 4 instead of 8 bytes per sample on the
 * PCIe leg that bounds every HOST batch.  Arithmetic stays fp32 on the device:
 *   modulator: out_sc16[i] = saturate_int16(round_to_nearest_even(generic_work(in)[i] * scale))   (re, im separately)
 *   receiver : generic_work[_equalize] runs on in[i] = (float)in_sc16[i] * (1.0f / scale)
 * `mem` as for the complex64 entries (HOST, DEVICE, HOST_ASYNC).  The quantisation is part of the FORMAT, not of the
 * kernels: the complex64 entries remain the parity path (north_star tolerance); tests hold these entries to
 * +-1 LSB of the quantised oracle output (modulator) and to the usual tolerance on identical int16 input (receiver). */
GFDM_B200_API int gfdm_modulator_work_batch_sc16(gfdm_modulator* h, short* out_iq, const gfdm_complex* in, float scale,
                                                 int n_frames, int mem);
GFDM_B200_API int gfdm_modulator_work_chunks_batch_sc16(gfdm_modulator* h, const gfdm_symbol_mapper* sm, short* out_iq,
                                                        const unsigned char* chunks, float scale, int n_frames, int mem);
GFDM_B200_API int gfdm_receiver_work_batch_sc16(gfdm_receiver* h, gfdm_complex* out, const short* in_iq,
                                                const gfdm_complex* f_eq_in, float scale, int n_frames, int mem);
GFDM_B200_API int gfdm_receiver_work_decide_batch_sc16(gfdm_receiver* h, const gfdm_symbol_mapper* sm,
                                                       unsigned char* chunks_out, const short* in_iq,
                                                       const gfdm_complex* f_eq_in, float scale, int n_frames, int mem);

#ifdef __cplusplus
}
#endif

#endif /* INCLUDED_GFDM_B200_H */

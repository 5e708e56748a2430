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
    GFDM_MEM_DEVICE = 1
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

#ifdef __cplusplus
}
#endif

#endif /* INCLUDED_GFDM_B200_H */

/*
 * nis_sar.h -- C ABI of libnis_sar.so: the B200 (sm_100a) SAR hot path.
 *
 * The reference (noiseinspacechannel/NIS-SAR-AMTIGMTI-Video) has no FFI: its boundary is a set of
 * Python functions and inline numpy expressions.  Each entry point below names the reference code
 * it replaces (file:line, relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 (NIS_OK) or a negative error class; the message is read with
 *     nis_last_error() (thread-local).  Nothing throws across this boundary.
 *   - pointers marked "dev" are CUDA device pointers owned by the CALLER (a torch tensor's
 *     data_ptr() is fine); the library owns only ctx / plan handles and their private workspace.
 *   - all work is enqueued on the given cudaStream_t (passed as void*) and is asynchronous;
 *     the only hidden synchronisations are the explicitly named *_readback helpers.
 *   - there is NO CPU fallback: without a CUDA device nis_ctx_create() fails.
 *   - complex samples are interleaved (re, im) float32 pairs ("c32") on the device.
 */
#ifndef NIS_SAR_H
#define NIS_SAR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIS_SAR_ABI_VERSION 2

#if defined(__GNUC__)
#define NIS_API __attribute__((visibility("default")))
#else
#define NIS_API
#endif

enum {
    NIS_OK = 0,
    NIS_ERR_INVALID = -1,      /* bad argument */
    NIS_ERR_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
    NIS_ERR_UNSUPPORTED = -3,  /* size / mode not supported by this build */
    NIS_ERR_NOMEM = -4
};

typedef struct nis_ctx nis_ctx;
typedef struct nis_csa_plan nis_csa_plan;
typedef struct nis_rda_plan nis_rda_plan;
typedef struct nis_tdbp_plan nis_tdbp_plan;
typedef struct { float re, im; } nis_c32;
typedef void* nis_stream; /* cudaStream_t */

/* ------------------------------------------------------------------ library / context */
NIS_API int nis_version(void);
/* copies the calling thread's last error message (NUL terminated) and returns its length */
NIS_API size_t nis_last_error(char* buf, size_t cap);
/* one context per device.  Calls on one ctx may be issued from several streams (and host threads) at once: the only
 * library-owned scratch (pulse tables of nis_tdbp_backproject, histograms of nis_region_select) is kept per stream, is
 * never freed while the ctx lives (a captured CUDA graph may replay with its address) and refuses to grow while its
 * stream is capturing -- run the call once outside the capture first.  Every entry point expects the ctx's device to be
 * the calling thread's current CUDA device (plan / ctx creation switch to it themselves and restore the caller's). */
NIS_API int nis_ctx_create(int device, nis_ctx** out);
NIS_API int nis_ctx_destroy(nis_ctx* ctx);
/* number of kernels this library has launched on ctx since creation (bench.py "gpu_launches") */
NIS_API uint64_t nis_ctx_launch_count(const nis_ctx* ctx);

/* ------------------------------------------------------------------ K1: raw-echo synthesis
 * Replaces the per-pulse loops of
 *   run_bistatic_physics_gpu   sar_ati_dcpa_sim_csa.py:106-181   (bistatic = 1)
 *   run_physics_engine         sar_satellite_sim.py:211-305      (bistatic = 0, tgt_vel = 0)
 *   run_moving_physics         sar_satellite_moving_sim.py:111-159
 *   run_custom_physics         sar_vehicle_sim.py:83-126
 * raw[i][n] (+)= sum_b amp_b * exp(j 2 pi (-fc tau_bi + (k_rate/2) (t_n - tau_bi - t_p/2)^2))
 *               over scatterers with |t_n - tau_bi - t_p/2| <= t_p/2 (closed interval),
 * tau_bi = (|p_b(t_i) - pos_tx_i| + |p_b(t_i) - pos_rx_i|) / c, p_b(t) = pos0_b + vel_b t.
 * For a monostatic engine pass pos_rx = NULL (tau = 2 |p - pos_tx| / c).
 * The receiver positions are host-side geometry (p_tx + v/|v| * offset, :145-148) and arrive
 * precomputed.  Geometry is evaluated in fp64 per (scatterer, pulse); samples accumulate in fp32.
 */
typedef struct {
    double c;        /* propagation speed (m/s) */
    double fc;       /* carrier (Hz) */
    double k_rate;   /* chirp rate BW / T_p (Hz/s) */
    double t_p;      /* pulse width (s) */
    double t_start;  /* t_fast[0] (s) */
    double dt_fast;  /* nominal fast-time step: t_fast[n] ~= t_start + n * dt_fast */
    int32_t per_target_velocity; /* 0: tgt_vel is one xyz triple; 1: tgt_vel is [T*3] */
    int32_t samples_per_thread;  /* 0: library chooses; 8 or 16: samples per thread.  A CTA (<= 256 threads) covers one of
                                  * ceil(S / (256 * this)) equal chunks of the window; 8 suits scenes whose chirps
                                  * cover well under the whole window */
} nis_echo_params;

/* Spotlight engine, run_physics_spotlight (sar_batch_sim.py:83-169): start-stop corrected delay -- the receive position is
 * pos_sat + vel_sat * 2 d_tx / c (:133-137) --, one-way sinc^2 pattern gain((pi l_ant / lambda) sin(off-boresight angle))
 * of an aperture steered at the scene centre (:139-149), amplitude rcs * gain (not its square root), chirp CENTRED on the
 * delay: exp(j 2 pi (-fc tau + (k_rate/2)(t_n - tau)^2)) for |t_n - tau| <= t_p/2 (:151-155).  The heading rotation of the
 * target block (:92-100) is host geometry and arrives applied to tgt_pos0 / tgt_vel. */
NIS_API int nis_echo_spotlight(nis_ctx* ctx, const nis_echo_params* prm, const double* tgt_pos0, const double* tgt_vel,
                       const double* tgt_rcs /* dev [T] */, const double* pos_sat /* dev [P*3] */,
                       const double* vel_sat /* dev [P*3] */, const double* t_slow, const double* t_fast,
                       int32_t T, int32_t P0, int32_t P1, int32_t S, double ant_pi_l_over_lambda,
                       nis_c32* raw, int32_t accumulate, nis_stream stream);

NIS_API int nis_echo_accumulate(nis_ctx* ctx, const nis_echo_params* prm,
                        const double* tgt_pos0 /* dev [T*3] */,
                        const double* tgt_vel  /* dev [3] or [T*3] */,
                        const double* tgt_amp  /* dev [T]  sqrt(rcs) */,
                        const double* pos_tx   /* dev [P*3] */,
                        const double* pos_rx   /* dev [P*3] or NULL */,
                        const double* t_slow   /* dev [P] */,
                        const double* t_fast   /* dev [S] exact sample times (linspace grid) */,
                        int32_t T, int32_t P0, int32_t P1, int32_t S,
                        nis_c32* raw /* dev [P][S], rows P0..P1-1 are written */,
                        int32_t accumulate /* 0: overwrite rows, 1: add to them, 2: add atomically (several devices
                                              * reducing their scatterer shards into one buffer, local or peer-mapped) */,
                        nis_stream stream);

/* ------------------------------------------------------------------ K2: Chirp Scaling focusing
 * Replaces sar_focus_csa (sar_ati_dcpa_sim_csa.py:202-396): azimuth FFT, x Phi1, range FFT, x Phi2,
 * range IFFT, x Phi3, azimuth IFFT, all fftshift/ifftshift folded into index arithmetic.
 * A plan owns the twiddle tables, the fp64-derived per-Doppler phase coefficients and a workspace
 * of n_az*n_rg complex samples.  The focused image is written CORNER-TURNED, i.e. as the
 * [n_rg][n_az] array that the reference returns as `img.T` (:396).
 */
typedef struct {
    double c;        /* 299792458.0 in the reference (:211) */
    double lambda;   /* center_wavelength_m */
    double kr;       /* chirp_rate_hzpsec */
    double fs;       /* sample_rate_hz */
    double prf;      /* prf_hz */
    double vr;       /* platform_speed_mps */
    double r_ref;    /* range_ref_m */
    double t_start;  /* t_start_fast */
} nis_csa_params;

NIS_API int nis_csa_plan_create(nis_ctx* ctx, int32_t n_az, int32_t n_rg, const nis_csa_params* prm,
                        nis_csa_plan** out);
NIS_API int nis_csa_plan_destroy(nis_csa_plan* plan);
/* 1 if (n_az, n_rg) is handled by the fused power-of-two path, 2 for the general-size path, 0 = unsupported */
NIS_API int nis_csa_size_class(int32_t n_az, int32_t n_rg);
/* range_axis[n_rg] = c tau / 2 (:346,:388), cross_range[n_az] = (n/prf - mean) vr (:392-394); host buffers */
NIS_API int nis_csa_axes(const nis_csa_plan* plan, double* range_axis, double* cross_range);
/* phist: dev [n_az][pitch] complex64, pitch in elements (>= n_rg).  The DPCA co-registration of
 * :402-403 (rx1[1:], rx2[:-1]) is a pointer offset of one row by the caller.
 * slc: dev [n_rg][n_az].  max_sq (optional, dev, 1 double, reset by the call on the same stream): max |slc|^2 evaluated
 * in fp64 on the stored fp32 samples -- hand it to nis_gmti_fused to skip its pass over slc1. */
NIS_API int nis_csa_focus(nis_csa_plan* plan, const nis_c32* phist, int64_t pitch, nis_c32* slc,
                  double* max_sq, nis_stream stream);

/* Per-kernel timing for bench.py's roofline: when enabled every nis_csa_focus call records CUDA
 * events between its five kernels (az-outer-fwd, az-inner-fwd, range, az-inner-inv, az-outer-inv)
 * on the launching stream; nis_csa_stage_times waits for the call `calls_back` calls ago (0 = most
 * recent, ring of 64) and returns the five durations in milliseconds (host buffer). */
NIS_API int nis_csa_plan_set_profiling(nis_csa_plan* plan, int32_t enable);
NIS_API int nis_csa_stage_times(nis_csa_plan* plan, int32_t calls_back, float* ms5);

/* ------------------------------------------------------------------ K2b: Range-Doppler focusing
 * Replaces sar_focus_rda (sar_satellite_sim.py:356-448, sar_vehicle_sim.py:182-274,
 * sar_satellite_moving_sim.py:208-285): Hamming-weighted matched-filter range compression (the 'same' window of the
 * linear convolution), Hamming-weighted azimuth DFT, per-Doppler linear-interpolation RCMC with zero fill, range-dependent
 * azimuth compression, inverse azimuth DFT, magnitude.  All arrays are PULSE-MAJOR [n_pulses][n_ranges] -- the layout of
 * the echo engines and of the `.T` the reference returns; its [num_ranges, num_pulses] arrays are the transposed views.
 */
typedef struct {
    double c;          /* 299792458 in the reference (:359) */
    double lambda;     /* center_wavelength_m */
    double t_p;        /* pulse_width_sec */
    double kr;         /* chirp_rate_hzpsec */
    double fs;         /* sample_rate_hz */
    double prf;        /* prf_hz */
    double vr;         /* platform_speed_mps */
    double range_grp;  /* range_grp_m */
} nis_rda_params;

/* 1 if an (n_pulses, n_ranges) frame with this matched-filter length (taps = int(t_p * fs) + 1) can be focused:
 * n_ranges <= 24576; taps <= 14337 (one or several 16384-point blocks), or taps <= 32768 with n_ranges <= 16384 and
 * n_ranges + taps / 2 <= 32768 (one 32768-point block); n_pulses a power of two 64..32768 with n_ranges % 32 == 0, or any length
 * the row-DFT engines take (prime factors <= 13 up to 14000, anything else up to 8192). */
NIS_API int nis_rda_supported(int32_t n_pulses, int32_t n_ranges, const nis_rda_params* prm);
NIS_API int nis_rda_plan_create(nis_ctx* ctx, int32_t n_pulses, int32_t n_ranges, const nis_rda_params* prm,
                        nis_rda_plan** out);
NIS_API int nis_rda_plan_destroy(nis_rda_plan* plan);
/* host buffers: range_axis_centered[n_ranges] (:443-444), cross_range[n_pulses] (:441), doppler_freq[n_pulses] (:400-404) */
NIS_API int nis_rda_axes(const nis_rda_plan* plan, double* range_axis_centered, double* cross_range, double* doppler_freq);
/* phist: dev [n_pulses][pitch] complex64.  image_mag: dev [n_pulses][n_ranges] float (= the reference's
 * sar_image_mag.T).  Optional exports, each dev [n_pulses][n_ranges] complex64 or NULL: rc = phist_compressed.T,
 * rd = range_doppler.T, rcmc = range_doppler_rcmc.T, filt = range_doppler_filtered.T (Doppler rows in fftshift order). */
NIS_API int nis_rda_focus(nis_rda_plan* plan, const nis_c32* phist, int64_t pitch, float* image_mag, nis_c32* rc,
                  nis_c32* rd, nis_c32* rcmc, nis_c32* filt, nis_stream stream);

/* ------------------------------------------------------------------ time-domain backprojection (VideoSAR frames)
 * Replaces tdbp_gpu (sar_batch_sim.py:171-238).  Range compression: circular correlation of each pulse with the fftshifted
 * int(t_p fs)-tap reference chirp over the n_samples window (:177-185).  Backprojection: per pixel of the nx x ny grid
 * linspace(-scene_size/2, scene_size/2) (x along-track, y ground range, z = 0) and per pulse, fp64 two-way delay with the
 * start-stop correction at both ends and the pixel moving at vel_focus about the CPI centre (:207-223), range-Doppler
 * coupling shift (:214-217), float32 two-tap interpolation of the compressed pulse exactly as torch's grid_sample
 * (align_corners=False) evaluates it (:225-230), carrier phase exp(j 2 pi fc tau), complex128 sum over pulses (:232-235).
 */
typedef struct {
    double c, fc, k_rate, t_p, fs;
    double t_start;      /* first sample time of the receive window */
    double scene_size;   /* m */
    int32_t n_samples, nx, ny, reserved;
} nis_tdbp_params;

NIS_API int nis_tdbp_plan_create(nis_ctx* ctx, const nis_tdbp_params* prm, nis_tdbp_plan** out);
NIS_API int nis_tdbp_plan_destroy(nis_tdbp_plan* plan);
/* raw: dev [n_pulses][pitch] complex64 -> rc: dev [n_pulses][n_samples] complex64 */
NIS_API int nis_tdbp_range_compress(nis_tdbp_plan* plan, const nis_c32* raw, int64_t pitch, int32_t n_pulses, nis_c32* rc,
                            nis_stream stream);
/* pos_plat / vel_plat: dev [n_pulses*3], t_pulses: dev [n_pulses] doubles; t_centre = mean(t_pulses) of the whole CPI;
 * vel_focus: HOST [3].  Pulses [p_begin, p_end) are summed into image (dev [ny][nx] complex128 as double pairs):
 * accumulate = 0 overwrites, 1 adds -- a CPI can be split into pulse blocks, across calls or across GPUs. */
NIS_API int nis_tdbp_backproject(nis_tdbp_plan* plan, const nis_c32* rc, const double* pos_plat, const double* vel_plat,
                         const double* t_pulses, int32_t n_pulses, int32_t p_begin, int32_t p_end, double t_centre,
                         const double* vel_focus, double* image, int32_t accumulate, nis_stream stream);

/* ------------------------------------------------------------------ K3: DPCA + ATI + detection
 * Replaces the inline numpy passes sar_ati_dcpa_sim_csa.py:414-419, :447-449 and
 * SARData.compute_all (sar_ati_dcpa_viewer_csa.py:42-52):
 *   s2c = slc2 * exp(j cal_phase); interf = slc1 conj(s2c); phase = atan2; diff = slc1 - s2c;
 *   mask = |slc1| > thresh_frac * max|slc1| (strict, evaluated in fp64 on the fp32 samples);
 *   phase_masked = mask ? phase : 0; det_idx = ascending flat indices with mask set;
 *   peak = first index attaining max|slc1|.
 * Any output pointer may be NULL (that product is not materialised).
 * One pass over the pair (products, detection bitmap, per-tile counts) + a tail kernel that reads only the workspace and
 * writes the index list and the result record; preceded by a pass over slc1 for max|slc1| only when the caller does not
 * pass max_mag_sq_in.  No atomics and no inter-block waiting: the list is deterministic.
 * Buffers: every pointer aligned to its element size (8 B complex, 4 B float / index).  When additionally the complex
 * buffers are 16-byte, the float maps 8-byte and the mask 2-byte aligned -- true for any whole allocation -- the kernel
 * moves two pixels per access; views at odd element offsets take an element-wise variant of the same kernel.
 * n_pix < 2^32 - 1 (flat indices and det_count are 32-bit: every BASELINE size fits).
 * workspace: dev, caller-owned, nis_gmti_workspace_bytes(n_pix) bytes (264 bytes per 2048 pixels + 16), 8-byte aligned,
 * contents irrelevant on entry; it must not be shared by two calls that may run concurrently.
 */
NIS_API uint64_t nis_gmti_workspace_bytes(uint64_t n_pix);
typedef struct {
    uint32_t det_count;   /* number of detected pixels (may exceed det_cap: list is truncated) */
    uint32_t peak_idx;    /* flat index of the first maximum of |slc1| */
    double   max_mag_sq;  /* max |slc1|^2 (fp64 on the fp32 samples) */
} nis_gmti_result;

NIS_API int nis_gmti_fused(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix,
                   double thresh_frac, double cal_phase,
                   nis_c32* ati_interf, float* ati_phase, nis_c32* dpca_diff, float* dpca_mag,
                   float* slc1_mag, uint8_t* mag_mask, float* ati_phase_masked,
                   uint32_t* det_idx, uint32_t det_cap,
                   const double* max_mag_sq_in /* dev, optional: max |slc1|^2 from nis_csa_focus */,
                   void* workspace /* dev */, uint64_t workspace_bytes,
                   nis_gmti_result* result /* dev, 16 bytes, written by the call */, nis_stream stream);
/* viewer auto-balance (sar_ati_dcpa_viewer_csa.py:249-250): sum slc1 conj(slc2) in fp64;
 * out: dev [2] doubles (re, im) of the SUM (angle of the mean == angle of the sum) */
NIS_API int nis_gmti_balance_sum(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix,
                         double* out_sum, nis_stream stream);

/* ------------------------------------------------------------------ viewer data layer
 * SARData.compute_all (sar_ati_dcpa_viewer_csa.py:42-52): with s2c = slc2 * exp(j cal_phase):
 * |slc1|, arg slc1, |s2c|, arg s2c, |slc1 - s2c|, arg(slc1 - s2c), arg(slc1 conj(s2c)); any output may be NULL. */
NIS_API int nis_viewer_products(nis_ctx* ctx, const nis_c32* slc1, const nis_c32* slc2, uint64_t n_pix, double cal_phase,
                        float* ch1_mag, float* ch1_phase, float* ch2_mag, float* ch2_phase, float* dpca_mag,
                        float* dpca_phase, float* ati_phase, nis_stream stream);
/* Statistics of the visible region (:117-139): a rows x cols rectangle of a float map whose rows are `pitch` elements
 * apart; db_scale != 0 evaluates 20 log10(x + 1e-12) (:128).  out_dev: dev [4] doubles = sum, min, max, sum of squared
 * deviations about the mean (mean = sum / n, std = sqrt(ssd / n) as numpy's). */
NIS_API int nis_region_stats(nis_ctx* ctx, const float* map, int64_t pitch, int32_t rows, int32_t cols, int32_t db_scale,
                     double* out_dev, nis_stream stream);
/* Exact order statistics of the region by radix select (median :118,:136; 99.9th percentile :147,:178,:184):
 * out_dev[i] = the value of 0-based rank ranks[i] (host array, 1..4 ranks) in ascending order. */
NIS_API int nis_region_select(nis_ctx* ctx, const float* map, int64_t pitch, int32_t rows, int32_t cols,
                      const uint64_t* ranks, int32_t n_ranks, float* out_dev, nis_stream stream);

/* ------------------------------------------------------------------ noise and sea clutter on the device
 * Replaces add_ocean_noise (sar_satellite_sim.py:331-344 and its copies sar_vehicle_sim.py:152-165,
 * sar_satellite_moving_sim.py:188-206) and generate_noise_tensor (sar_batch_sim.py:65-81): complex Gaussian thermal
 * noise at P / 10^(snr_db/10) plus K-distributed clutter (Gamma(k_nu, 1/k_nu) texture x Exp(1) speckle, uniform phase)
 * at P / 10^(scr_db/10).  P = *power_dev * power_value when power_dev != NULL -- the output of nis_power_sum with
 * power_value = 1/n (mean power, add_ocean_noise) or of nis_power_max with power_value = 1 (peak power, sar_batch_sim.py:317):
 * no host round trip -- else P = power_value.  accumulate != 0: x += noise; 0: x = noise (generate_noise_tensor).
 * Counter-based generator (Philox4x32-10 keyed by seed, counter = sample index): reproducible, launch-shape independent.
 */
NIS_API int nis_power_sum(nis_ctx* ctx, const nis_c32* x, uint64_t n, double* sum_dev /* dev, 1 double */, nis_stream stream);
NIS_API int nis_power_max(nis_ctx* ctx, const nis_c32* x, uint64_t n, double* max_dev /* dev, 1 double */, nis_stream stream);
NIS_API int nis_noise_add(nis_ctx* ctx, nis_c32* x, uint64_t n, const double* power_dev, double power_value,
                  double snr_db, double scr_db, double k_nu, uint64_t seed, int32_t accumulate, nis_stream stream);

/* ------------------------------------------------------------------ peer memory (multi-GPU exchange steps)
 * Where the path exchanges data -- partial echoes of scatterer shards, the neighbour channel of a DPCA/ATI pair -- the
 * kernels above read / reduce into another GPU's HBM directly over NVLink: a peer-mapped pointer is passed like any other
 * buffer (nis_echo_accumulate(..., accumulate = 2) adds a rank's partial echo into the owner's rows; nis_gmti_fused takes
 * slc2 = the neighbour's image).  One process per GPU: the owner allocates and exports (64-byte CUDA IPC handle, sent
 * over the caller's control plane), the others map it into their current device's context. */
NIS_API int nis_peer_alloc(uint64_t bytes, void** ptr, uint8_t* handle64);
NIS_API int nis_peer_free(void* ptr);
NIS_API int nis_peer_open(const uint8_t* handle64, void** ptr);
NIS_API int nis_peer_close(void* ptr);

/* 1 when `device` performs atomics on `peer`'s memory natively over their link (cudaDevP2PAttrNativeAtomicSupported):
 * the precondition of accumulate = 2 into a peer-mapped buffer (system-scope RED.ADD); 0 -> use nis_echo_reduce. */
NIS_API int nis_peer_native_atomics(int32_t device, int32_t peer);

/* ------------------------------------------------------------------ collectives (the path's two exchange steps, NCCL)
 * One process per GPU; the 128-byte id is created by rank 0 (nis_comm_unique_id) and distributed over the caller's own
 * control plane (torch.distributed, MPI, a file ...).  nis_comm_init binds to the calling thread's current device.
 * NCCL is resolved at run time from the host process (libnccl.so.2); without it these return NIS_ERR_UNSUPPORTED.
 *   nis_echo_reduce   raw[n_samples] complex64 summed in place across ranks -- the reduction of scatterer-sharded echo
 *                     synthesis (config 3): root = -1 all-reduce, else reduce to `root`.  Sum order differs from the
 *                     one-GPU run: results agree to fp32 rounding, not bit-wise.
 *   nis_slc_exchange  ring shift for DPCA/ATI channel pairing (one receive channel per rank): rank k sends its focused
 *                     image to rank k-1 and receives channel k+1 into slc_next (NULL on the last rank). */
typedef struct nis_comm nis_comm;
NIS_API int nis_comm_unique_id(uint8_t* id128);
NIS_API int nis_comm_init(int32_t rank, int32_t nranks, const uint8_t* id128, nis_comm** out);
NIS_API int nis_comm_destroy(nis_comm* comm);
NIS_API int nis_echo_reduce(nis_comm* comm, nis_c32* raw, uint64_t n_samples, int32_t root, nis_stream stream);
NIS_API int nis_slc_exchange(nis_comm* comm, const nis_c32* slc_local, nis_c32* slc_next, uint64_t n_pix, nis_stream stream);

/* ------------------------------------------------------------------ buffer format helpers
 * The reference's arrays are complex128; these convert on the device so that host<->device copies
 * are the only host-side cost. */
NIS_API int nis_narrow_c128_to_c32(nis_ctx* ctx, const double* src /* dev [n*2] */, nis_c32* dst, uint64_t n, nis_stream stream);
NIS_API int nis_widen_c32_to_c128(nis_ctx* ctx, const nis_c32* src, double* dst /* dev [n*2] */, uint64_t n, nis_stream stream);
/* Host <-> device transfer of the reference's complex128 arrays in their NARROW form (half the PCIe bytes), converted
 * on the host cores while the DMA engine runs (csrc/hostcopy.cpp): chunks of 4 MiB through a ring of page-locked slots
 * owned by the library, `threads` worker threads (1..32), one whole chunk per thread.  Both calls BLOCK until the
 * destination is complete; host pointers may be pageable or page-locked.  The values are those of the device-side
 * helpers above (float -> double exact, double -> float round to nearest even).
 *   nis_d2h_widen   host_dst[2n doubles] = (double) dev_src[n complex64]
 *   nis_h2d_narrow  dev_dst[n complex64] = (float) host_src[2n doubles] */
NIS_API int nis_d2h_widen(nis_ctx* ctx, const nis_c32* dev_src, double* host_dst, uint64_t n, int32_t threads, nis_stream stream);
NIS_API int nis_h2d_narrow(nis_ctx* ctx, const double* host_src, nis_c32* dev_dst, uint64_t n, int32_t threads, nis_stream stream);
/* out[c][r] = in[r][c] for a rows x cols complex64 matrix (tiled, coalesced both sides) */
NIS_API int nis_transpose_c32(nis_ctx* ctx, const nis_c32* in, nis_c32* out, int32_t rows, int32_t cols, nis_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* NIS_SAR_H */

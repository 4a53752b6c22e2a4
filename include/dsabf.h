/*
 * dsabf.h -- C ABI of the B200-native DSA beamforming hot path.
 *
 * The reference (devincody/DSAbeamformer) has no plugin / FFI interface: its hot
 * path is the body of the per-stream loop in main()
 * (/root/reference/src/beamformer.cu:461-521) plus the set-up around it
 * (:163-326) and the event plumbing of observation_loop_state
 * (src/observation_loop.hh:71-119).  Each entry point below names the piece of
 * that code it replaces.  Plain pointers and sizes only; every function returns
 * 0 on success or a non-zero code (a cudaError_t value, or DSABF_E_*), with the
 * message available from dsabf_last_error_string().  A handle is owned by one
 * host thread (one handle per GPU), like the reference's single-threaded loop.
 *
 * Layouts (identical to the reference's):
 *   packed voltages : uint8 [chunk][channel][column][antenna]; one byte per complex
 *                     sample, re = high nibble, im = low nibble, two's complement;
 *                     column = out_time*(n_pol*n_averaging) + in_time*n_pol + pol
 *                     (src/test_data_generator.hh:83, src/beamformer.cuh:122-123)
 *   weights         : int8  [channel][antenna][beam][2] (re,im), beam fastest
 *                     == d_fourier_coefficients (src/beamformer.cu:230-241)
 *   power           : float [out_time][channel][beam]  == d_out / beam_out
 *                     (src/beamformer.cuh:125-126,147), includes alpha^2=(1/127)^2
 *   voltages (debug): int32 [channel][column][beam][2] == 127 * d_C
 */
#ifndef DSABF_H
#define DSABF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSABF_VERSION 1

/* error codes beyond cudaError_t values */
#define DSABF_E_INVALID   10001  /* bad argument / unsupported shape           */
#define DSABF_E_STATE     10002  /* call out of order (e.g. weights not set)   */
#define DSABF_E_WEIGHT    10003  /* weight component == -128 (cannot negate)   */

typedef struct dsabf_handle dsabf_handle;

/* Shape constants of /root/reference/src/beamformer.hh:47-60,108-147, made
 * run-time.  dsabf_default_config() fills in the reference's values. */
typedef struct dsabf_config {
    int n_antennas;           /* N_ANTENNAS 64 (64, 128, 192 or 256)                        */
    int n_beams;              /* N_BEAMS 256   (multiple of 64, <= 4096)                     */
    int n_channels;           /* N_FREQUENCIES 256 (any >= 1)                                */
    int n_pol;                /* N_POL 2                                                     */
    int n_averaging;          /* N_AVERAGING: 16 observation / 1 DEBUG (power of two <= 32) */
    int n_outputs_per_chunk;  /* N_OUTPUTS_PER_GEMM 8                                        */
    int n_chunks_per_block;   /* N_GEMMS_PER_BLOCK 32                                        */
    int n_ring_slots;         /* N_BLOCKS_ON_GPU 8: device ring of input blocks              */
    int device;               /* CUDA ordinal; -1 = current device                           */
    int exact_power;          /* 1: epilogue accumulates |V|^2 in int64 (exact), 0: fp32     */
    float power_scale;        /* multiplies the integer power; 0 => (float)(1.0/127)^2       */
} dsabf_config;

/* beamformer.hh:47-60: debug != 0 selects the DEBUG-build constants. */
int dsabf_default_config(dsabf_config *cfg, int debug);

/* ---- derived sizes (beamformer.hh:108-147) ---- */
size_t dsabf_chunk_bytes(const dsabf_config *cfg);        /* N_BYTES_PRE_EXPANSION_PER_GEMM  */
size_t dsabf_block_bytes(const dsabf_config *cfg);        /* N_BYTES_PRE_EXPANSION_PER_BLOCK */
size_t dsabf_chunk_out_floats(const dsabf_config *cfg);   /* N_F_PER_DETECT                  */
int    dsabf_columns_per_chunk(const dsabf_config *cfg);  /* N_TIMESTEPS_PER_GEMM            */

/* ---- life cycle: replaces the allocation / stream / handle set-up and the
 * teardown in main(), beamformer.cu:249-320 and :560-618 ---- */
int dsabf_create(const dsabf_config *cfg, dsabf_handle **out);
int dsabf_destroy(dsabf_handle *h);
const char *dsabf_last_error_string(void);

/* ---- steering weights ---- */
/* Host-side generator: the loop of beamformer.cu:230-241 (+ :173, hh:63-77).
 * pos: [n_ant][3] (x,y,z metres), dir: [n_beams][2] (theta,phi radians),
 * gpu: sub-band index (-g), w: [n_chan][n_ant][n_beams][2] int8. */
int dsabf_compute_weights(const float *pos, const float *dir, int n_ant, int n_beams,
                          int n_chan, int gpu, int8_t *w);
/* Upload (beamformer.cu:272) and re-lay the weights for the tensor cores. */
int dsabf_set_weights(dsabf_handle *h, const int8_t *w);

/* ---- ingest: replaces cudaMemcpyAsync on HtoDstream + generate_transfer_event
 * (beamformer.cu:389-396 / 425-431; observation_loop.hh:71-75) ---- */
void *dsabf_ring_slot_ptr(dsabf_handle *h, int slot);           /* &d_data[block_bytes*slot] */
/* copy `bytes` (<= block bytes) of caller-owned (ideally pinned) host memory into
 * ring slot `slot`; returns a ticket in *ticket for dsabf_query_transfer. */
int dsabf_enqueue_h2d(dsabf_handle *h, int slot, const void *host_src, size_t bytes, uint64_t *ticket);
/* cudaEventQuery semantics (observation_loop.hh:83-119): 1 done, 0 pending, <0 error */
int dsabf_query_transfer(dsabf_handle *h, uint64_t ticket);
int dsabf_query_analysis(dsabf_handle *h, uint64_t ticket);

/* ---- the hot path: replaces, for every chunk of the block,
 *   expand_input -> cublasGemmStridedBatchedEx -> detect_sum -> D2H
 *   [-> cublasSgemv -> D2H]                 (beamformer.cu:461-521)
 * with ONE fused kernel over the whole block.  The call waits (on the device)
 * for the slot's H2D ticket, so the caller need not poll before enqueueing.
 *   n_chunks           : chunks present in the slot (<= n_chunks_per_block)
 *   beam_out_host      : caller-owned pinned buffer, n_chunks*chunk_out_floats
 *                        floats, [chunk][out_time][chan][beam]; NULL = keep on device
 *   dedispersed_host   : NULL, or [n_chunks][n_beams] floats = DM-0 sum over the
 *                        channels of output time 0 of each chunk (the DEBUG gemv)
 *   n_valid_dedisp     : only the first n_valid_dedisp chunks' sums are copied
 *                        (the reference's `current_gemm < n_pt_sources` guard)
 *   ticket             : completion ticket for dsabf_query_analysis            */
int dsabf_enqueue_block(dsabf_handle *h, int slot, int n_chunks, float *beam_out_host,
                        float *dedispersed_host, int n_valid_dedisp, uint64_t *ticket);
/* device pointer of the power produced by the last dsabf_enqueue_block on `slot` */
float *dsabf_block_out_ptr(dsabf_handle *h, int slot);
/* Wait for everything queued on the handle (all streams fenced -- the reference
 * fences only stream[N_STREAMS-1], beamformer.cu:525). */
int dsabf_sync(dsabf_handle *h);

/* Device-resident entry (what bench.py's `value` times): packed voltages already
 * in HBM, power left in HBM.  d_packed: n_chunks chunks; d_out: [n_chunks*
 * n_outputs_per_chunk][chan][beam] floats; stream: cudaStream_t or NULL for the
 * handle's compute stream. */
int dsabf_beamform_device(dsabf_handle *h, const void *d_packed, int n_chunks, float *d_out, void *stream);
/* Host-buffer convenience (synchronous): H2D -> fused kernel -> D2H. */
int dsabf_beamform_host(dsabf_handle *h, const void *packed, int n_chunks, float *out);

/* ---- debug / parity taps (the reference never exposes d_C) ---- */
/* int32 beam voltages V = 127*d_C for n_chunks host chunks:
 * V [chunk][chan][column][beam][2]; optional exact int64 power
 * P [chunk*n_outputs_per_chunk][chan][beam] (NULL to skip). */
int dsabf_debug_voltages_host(dsabf_handle *h, const void *packed, int n_chunks, int32_t *V, int64_t *P);

/* ---- dedispersion (generalises the DEBUG cublasSgemv, beamformer.cu:498-504;
 * DM > 0 per sandbox/Dispersion Theory.ipynb cells 2,5) ----
 * d_power: [n_t][n_chan][n_beams] float in HBM; h_delay: [n_dm][n_chan] int32 >= 0
 * (host); d_out: [n_dm][n_t_out][n_beams] float in HBM:
 *   out[d][t][b] = sum_f power[t + delay[d][f]][f][b]   (terms with t+delay >= n_t dropped)
 * dsabf_dedisperse_device is stream-ordered (returns after enqueueing; h_delay is read by a stream-ordered copy: pageable memory may be reused at once, pinned memory only after the stream reaches this call);
 * stream NULL = the handle's compute stream, which dsabf_sync() drains. */
int dsabf_dedisperse_device(dsabf_handle *h, const float *d_power, int n_t, int n_chan, int n_beams,
                            const int32_t *h_delay, int n_dm, int n_t_out, float *d_out, void *stream);
int dsabf_dedisperse_host(dsabf_handle *h, const float *power, int n_t, int n_chan, int n_beams,
                          const int32_t *delay, int n_dm, int n_t_out, float *out);

/* ---- DEBUG-mode synthetic input, generated on the GPU (test_data_generator::generate_test_data,
 * src/test_data_generator.hh:63-95): batch `batch` of 1024 chunks, one point source per chunk (sources
 * [1024*batch, 1024*batch + 1024) of the n_sources given as host pairs (theta, phi) in radians; chunks past the last
 * source are zero bytes), pos = host [n_antennas][3] metres, gpu = sub-band index (-g).  Same bytes as the CPU loop.
 * _device writes 1024 * dsabf_chunk_bytes() bytes of device memory (16-byte aligned; stream NULL = compute stream);
 * _host additionally copies them into dst (host, ideally pinned) and returns when they are there. */
int dsabf_generate_test_data_device(dsabf_handle *h, const float *pos, const float *sources, int n_sources, int batch,
                                    int gpu, void *d_dst, void *stream);
int dsabf_generate_test_data_host(dsabf_handle *h, const float *pos, const float *sources, int n_sources, int batch,
                                  int gpu, void *dst);

/* Host side of the dedispersion stage: the DM trial list of sandbox/Dispersion Theory.ipynb cells 1-2 (2048 channels
 * to DM 2000: 1627 trials) and the integer delay table of its cell 5 in output samples (16 x 131 us).
 * dsabf_dm_trials returns the number of trials (also when it exceeds capacity; only the first `capacity` are
 * written), -1 on a bad argument.  delay: [n_dm][n_chan], freq_ghz: channel centre frequencies. */
int dsabf_dm_trials(double dm_max, int n_chan_total, double *dms, int capacity);
int dsabf_dm_delays(const double *dms, int n_dm, const double *freq_ghz, int n_chan, int32_t *delay);

/* ---- multi-GPU exchange in front of full-band dedispersion (the reference has none: each
 * `beam -g <k>` process sums only its own 256 channels, beamformer.cu:498-504) ----
 * peer_power[g]: device pointer, valid on THIS GPU, to GPU g's power [n_t][c_local][n_beams] (peer-mapped
 * memory, e.g. the buffer_ptrs of a torch symmetric-memory allocation; peer_power[rank] is the local one).
 * d_full: [n_t][world*c_local][n_beams/world] float in local HBM:
 *   full[t][g*c_local + c][b] = power_g[t][c][rank*(n_beams/world) + b]
 * One kernel, loads over NVLink, no staging.  The caller fences: every peer has finished writing its power
 * before the call, and nobody overwrites it until every rank's call has completed. */
int dsabf_transpose_pull_device(dsabf_handle *h, const void *const *peer_power, int world, int rank, int n_t,
                                int c_local, int n_beams, float *d_full, void *stream);

/* Full-band single-trial dedispersion (DM 0 = the DEBUG cublasSgemv over ALL sub-bands, or any one delay curve)
 * fused with the exchange: this GPU's n_beams/world beams summed over the world*c_local channels that live in the
 * peers' buffers, every remote element read once over NVLink, nothing materialised.
 * h_delay: [world*c_local] int32 (host), global channel order; d_out: [n_t_out][n_beams/world] float. */
int dsabf_dedisperse_single_peers_device(dsabf_handle *h, const void *const *peer_power, int world, int rank, int n_t,
                                         int c_local, int n_beams, const int32_t *h_delay, int n_t_out, float *d_out,
                                         void *stream);

/* ---- introspection ---- */
/* number of this library's kernels launched on the handle since creation */
uint64_t dsabf_kernel_launches(dsabf_handle *h);
/* average device time (ms, CUDA events on the launching stream) of the fused
 * kernel launches recorded since the last call with reset != 0; *n_launches gets
 * the number of launches averaged.  Used for bench.py's roofline.achieved. */
int dsabf_fused_kernel_ms(dsabf_handle *h, double *avg_ms, int *n_launches, int reset);
int dsabf_enable_kernel_timing(dsabf_handle *h, int on);

#ifdef __cplusplus
}
#endif
#endif /* DSABF_H */

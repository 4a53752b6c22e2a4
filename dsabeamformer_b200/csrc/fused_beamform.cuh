// fused_beamform.cuh -- launch interface of the fused unpack+beamform+detect kernel.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace dsabf {

struct FusedParams {
    const uint8_t *packed;   // [n_chunks][n_chan][cols_per_chunk][64] packed 4+4-bit voltages
    const uint8_t *wprep;    // [n_chan][n_beams/64][n_ant/64][16384] pre-swizzled A' tiles (launch_prep_weights)
    float *out;              // [n_chunks*cols_per_chunk/group][n_chan][n_beams] power
    int32_t *vdump;          // optional: [n_chan][n_chunks*cols_per_chunk][n_beams][2] int32 voltages
    long long *pdump;        // optional (exact mode): integer power, same shape as out
    long long *trace;        // optional: clock64 trace of CTA 0 (debug), 128 stages x 16 slots
    int n_chan;
    int n_chunks;
    int cols_per_chunk;      // N_TIMESTEPS_PER_GEMM
    int group;               // columns summed per output = n_pol * n_averaging
    int n_ant;               // antennas: 64 (resident-weights kernel) or 64..256 in steps of 64 (streamed-weights kernel)
    int n_beams;             // multiple of 64; <= 256 for the resident-weights kernel
    int n_tiles;             // ceil(n_chunks*cols_per_chunk / 256)
    float scale;             // power_scale: multiplies the integer power
};

cudaError_t launch_fused_beamform(const FusedParams &p, bool exact, int num_sms, cudaStream_t stream);
cudaError_t launch_fused_beamform_big(const FusedParams &p, bool exact, int num_sms, cudaStream_t stream);
cudaError_t launch_prep_weights(const int8_t *d_w, uint8_t *d_wprep, int n_chan, int n_ant, int n_beams, int *d_bad,
                                cudaStream_t stream);
cudaError_t launch_dm0(const float *d_out, float *d_ded, int n_chunks, int n_out_per_chunk, int n_chan, int n_beams,
                       cudaStream_t stream);
// dedisperse.cu: one trial over a channel axis spread across peers' buffers (exchange fused into the sum)
cudaError_t launch_dedisperse_single_peers(const float *const *peer_power, int world, int rank, int n_t, int c_local, int n_beams,
                                           const int32_t *d_delay, int n_t_out, float *d_out, cudaStream_t stream);
// exchange.cu: frequency -> beam transpose pulled from peer GPUs' memory (NVLink), see the file header
cudaError_t launch_transpose_pull(const float *const *peer_power, int world, int rank, int n_t, int c_local, int n_beams, float *d_full,
                                  int num_sms, cudaStream_t stream);
cudaError_t launch_dedisperse(const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_dm,
                              int n_t_out, float *d_out, cudaStream_t stream);

// test_data_gen.cu: the DEBUG-mode synthetic input on the GPU (test_data_generator.hh:63-95)
cudaError_t launch_generate_test_data(uint8_t *d_dst, const float *d_pos_xy, const float *d_src_sin, const float *d_wavelength, int n_chunks,
                                      int n_valid, int n_ant, int n_chan, int cols, cudaStream_t stream);

}  // namespace dsabf

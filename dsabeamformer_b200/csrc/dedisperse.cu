// dedisperse.cu -- incoherent dedispersion of beamformed power over DM trials.
//
// Generalises the reference's DEBUG-mode "dedispersion at DM = 0" -- a cublasSgemv
// with a ones vector over the channel axis, /root/reference/src/beamformer.cu:498-504
// -- to the delay table of sandbox/Dispersion Theory.ipynb cell 5:
//     out[d][t][b] = sum_f power[t + delay[d][f]][f][b]
// power is the fused kernel's output layout [time][channel][beam] (beam fastest).
//
// Blocking.  One CTA owns (a tile of TT output times) x (a tile of 64 beams) and
// loops over DM trials in groups of DMS.  For each channel the CTA needs the
// power rows t0 + delay .. t0 + delay + TT - 1; consecutive DM trials have
// nearly equal delays, so for a group of trials the union of needed rows is a
// short window [dmin, dmax + TT).  That window (<= WIN rows x 64 beams) is staged
// once in shared memory per channel (coalesced 256 B row segments) and every
// trial of the group reads its rows from there: HBM/L2 traffic per channel is
// one window instead of DMS x TT rows.  Channels are split over the 8 warps of
// the CTA (each warp accumulates a private partial sum for all (trial, t) pairs
// in registers) and the partial sums are combined at the end through shared
// memory.  Lanes map to beams (2 beams per lane), so every global and shared
// access is a contiguous row segment.
#include "fused_beamform.cuh"

namespace dsabf {
namespace {

constexpr int DD_BEAMS = 64;    // beams per CTA (2 per lane)
constexpr int DD_TT = 8;        // output times per CTA
constexpr int DD_DMS = 4;       // DM trials per pass
constexpr int DD_WARPS = 8;
constexpr int DD_WIN = 40;      // staged rows per channel: dmax - dmin + TT <= WIN, else fallback

// Fast path: valid when, for every channel, the spread of delays inside the DM
// group fits the staging window.  `ok` is precomputed per DM group on the host.
__global__ void __launch_bounds__(DD_WARPS * 32)
dedisperse_kernel(const float *__restrict__ power, int n_t, int n_chan, int n_beams, const int32_t *__restrict__ delay, int n_dm,
                  int n_t_out, float *__restrict__ out)
{
    // dynamic shared memory: per-warp staging windows (rows x 64 beams), later reused for the
    // cross-warp reduction
    extern __shared__ float2 dd_smem[];
    float2 (*win)[DD_WIN][32] = reinterpret_cast<float2 (*)[DD_WIN][32]>(dd_smem);
    float2 (*red)[DD_DMS * DD_TT][32] = reinterpret_cast<float2 (*)[DD_DMS * DD_TT][32]>(dd_smem);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * DD_BEAMS;
    const int t0 = blockIdx.y * DD_TT;
    const int d0 = blockIdx.z * DD_DMS;
    const int nd = min(DD_DMS, n_dm - d0);
    const bool beam_ok = (b0 + 2 * lane + 1) < n_beams;      // n_beams is even (multiple of 64 in practice)

    float2 acc[DD_DMS][DD_TT];
#pragma unroll
    for (int i = 0; i < DD_DMS; i++)
#pragma unroll
        for (int j = 0; j < DD_TT; j++) acc[i][j] = make_float2(0.f, 0.f);

    for (int f = warp; f < n_chan; f += DD_WARPS) {
        int dl[DD_DMS];
        int dmin = 0x7fffffff, dmax = 0;
#pragma unroll
        for (int i = 0; i < DD_DMS; i++) {
            dl[i] = (i < nd) ? delay[(size_t)(d0 + i) * n_chan + f] : -1;
            if (i < nd) { dmin = min(dmin, dl[i]); dmax = max(dmax, dl[i]); }
        }
        const int rows = dmax - dmin + DD_TT;
        if (rows <= DD_WIN) {
            // stage the window [t0 + dmin, t0 + dmin + rows) of channel f
            for (int r = 0; r < rows; r++) {
                const int tt = t0 + dmin + r;
                float2 v = make_float2(0.f, 0.f);
                if (tt < n_t && beam_ok) v = *reinterpret_cast<const float2 *>(power + ((size_t)tt * n_chan + f) * n_beams + b0 + 2 * lane);
                win[warp][r][lane] = v;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < DD_DMS; i++) {
                if (i < nd) {
                    const int off = dl[i] - dmin;
#pragma unroll
                    for (int j = 0; j < DD_TT; j++) {
                        const float2 v = win[warp][off + j][lane];
                        acc[i][j].x += v.x;
                        acc[i][j].y += v.y;
                    }
                }
            }
            __syncwarp();
        } else {
            // delays of this group are too spread out at this channel: read rows directly
#pragma unroll
            for (int i = 0; i < DD_DMS; i++) {
                if (i < nd) {
#pragma unroll
                    for (int j = 0; j < DD_TT; j++) {
                        const int tt = t0 + j + dl[i];
                        if (tt < n_t && beam_ok) {
                            const float2 v = *reinterpret_cast<const float2 *>(power + ((size_t)tt * n_chan + f) * n_beams + b0 + 2 * lane);
                            acc[i][j].x += v.x;
                            acc[i][j].y += v.y;
                        }
                    }
                }
            }
        }
    }

    // combine the 8 per-warp partial sums
    __syncthreads();
#pragma unroll
    for (int i = 0; i < DD_DMS; i++)
#pragma unroll
        for (int j = 0; j < DD_TT; j++) red[warp][i * DD_TT + j][lane] = acc[i][j];
    __syncthreads();
    for (int e = warp; e < DD_DMS * DD_TT; e += DD_WARPS) {
        const int i = e / DD_TT, j = e % DD_TT;
        float2 s = red[0][e][lane];
#pragma unroll
        for (int w = 1; w < DD_WARPS; w++) { s.x += red[w][e][lane].x; s.y += red[w][e][lane].y; }
        const int t = t0 + j;
        if (i < nd && t < n_t_out && beam_ok)
            *reinterpret_cast<float2 *>(out + ((size_t)(d0 + i) * n_t_out + t) * n_beams + b0 + 2 * lane) = s;
    }
}

}  // namespace

cudaError_t launch_dedisperse(const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_dm, int n_t_out,
                              float *d_out, cudaStream_t stream)
{
    if (n_dm <= 0 || n_t_out <= 0) return cudaSuccess;
    if (n_beams % 2) return cudaErrorInvalidValue;
    dim3 grid((n_beams + DD_BEAMS - 1) / DD_BEAMS, (n_t_out + DD_TT - 1) / DD_TT, (n_dm + DD_DMS - 1) / DD_DMS);
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
    constexpr size_t win_bytes = sizeof(float2) * DD_WARPS * DD_WIN * 32, red_bytes = sizeof(float2) * DD_WARPS * DD_DMS * DD_TT * 32;
    constexpr size_t smem = win_bytes > red_bytes ? win_bytes : red_bytes;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dedisperse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dedisperse_kernel<<<grid, DD_WARPS * 32, smem, stream>>>(d_power, n_t, n_chan, n_beams, d_delay, n_dm, n_t_out, d_out);
    return cudaGetLastError();
}

}  // namespace dsabf

// dedisperse.cu -- incoherent dedispersion of beamformed power over DM trials.
//
// Generalises the reference's DEBUG-mode "dedispersion at DM = 0" -- a cublasSgemv
// with a ones vector over the channel axis, /root/reference/src/beamformer.cu:498-504
// -- to the delay table of sandbox/Dispersion Theory.ipynb cell 5:
//     out[d][t][b] = sum_f power[t + delay[d][f]][f][b]
// power is the fused kernel's output layout [time][channel][beam] (beam fastest).
//
// Blocking.  One CTA = 8 warps owns (32 output times) x (32 beams) x (8 DM trials): warp w
// accumulates trial d0+w, lane l beam b0+l, 32 time accumulators per thread in registers.
// Channels are processed in chunks of 4.  Neighbouring DM trials have nearly equal delays,
// so for one channel the union of the rows the 8 trials need is a short window
// [t0+dmin, t0+dmax+32) of 128-byte rows; the window of each channel of the chunk is staged in
// shared memory with cp.async (double-buffered against the arithmetic on the previous chunk)
// and every warp then reads its 32 rows from there (consecutive lanes = consecutive beams:
// conflict-free).  Each staged value is used by up to 8 trials, so global traffic per add is
// ~1/16 of a float; every add costs one 4-byte shared-memory read per lane, which is the
// structural bound of brute-force dedispersion (32 adds/clk/SM).  With few trials the kernel is
// HBM-bound (power read once), with many it is shared-memory-bound.
#include "fused_beamform.cuh"

namespace dsabf {
namespace {

constexpr int DD_TT = 32;        // output times per CTA
constexpr int DD_BB = 32;        // beams per CTA (one per lane)
constexpr int DD_DMS = 8;        // DM trials per CTA (one per warp)
constexpr int DD_FC = 4;         // channels per staged chunk
constexpr int DD_WIN = 64;       // staged rows per channel: (dmax - dmin) + 32 <= 64, else direct global reads
constexpr int DD_THREADS = DD_DMS * 32;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ChunkMeta {
    int dmin[DD_FC];             // smallest delay of the CTA's trials, per channel of the chunk
    int rows[DD_FC];             // rows staged (0 = window too wide: read global memory directly)
    int off[DD_FC][DD_DMS];      // delay - dmin per trial
};

__global__ void __launch_bounds__(DD_THREADS)
dedisperse_kernel(const float *__restrict__ power, int n_t, int n_chan, int n_beams, const int32_t *__restrict__ delay, int n_dm,
                  int n_t_out, float *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t dd_smem[];
    float (*win)[DD_FC][DD_WIN][DD_BB] = reinterpret_cast<float (*)[DD_FC][DD_WIN][DD_BB]>(dd_smem);          // [2]
    ChunkMeta *meta = reinterpret_cast<ChunkMeta *>(dd_smem + sizeof(float) * 2 * DD_FC * DD_WIN * DD_BB);  // [2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * DD_BB, t0 = blockIdx.y * DD_TT, d0 = blockIdx.z * DD_DMS;
    const int nd = min(DD_DMS, n_dm - d0);
    const bool my_dm = warp < nd;
    const bool my_beam = (b0 + lane) < n_beams;
    const int n_chunks = (n_chan + DD_FC - 1) / DD_FC;

    float acc[DD_TT];
#pragma unroll
    for (int t = 0; t < DD_TT; t++) acc[t] = 0.f;

    // stage chunk c into buffer buf: metadata by the first 32 threads, rows by everybody (cp.async, 16 B)
    auto stage = [&](int c, int buf) {
        if (threadIdx.x < DD_FC * DD_DMS) {
            const int fi = threadIdx.x / DD_DMS, w = threadIdx.x % DD_DMS;
            const int f = c * DD_FC + fi;
            int dl = (w < nd && f < n_chan) ? delay[(size_t)(d0 + w) * n_chan + f] : 0x3fffffff;
            int mn = dl, mx = (w < nd && f < n_chan) ? dl : 0;
#pragma unroll
            for (int o = DD_DMS / 2; o > 0; o >>= 1) {     // the 8 trials of one channel sit in 8 consecutive lanes
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            meta[buf].off[fi][w] = dl - mn;
            if (w == 0) {
                const int rows = (f < n_chan) ? (mx - mn + DD_TT) : 0;
                meta[buf].dmin[fi] = mn;
                meta[buf].rows[fi] = (rows <= DD_WIN) ? rows : 0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int fi = 0; fi < DD_FC; fi++) {
            const int f = c * DD_FC + fi;
            const int rows = meta[buf].rows[fi];
            const int r_first = t0 + meta[buf].dmin[fi];
            for (int i = threadIdx.x; i < rows * (DD_BB / 4); i += DD_THREADS) {     // 8 x 16 B per 128-byte row
                const int r = i >> 3, q = i & 7;
                float *dst = &win[buf][fi][r][q * 4];
                const int tt = r_first + r;
                if (tt < n_t && b0 + q * 4 + 3 < n_beams && (n_beams & 3) == 0) {
                    cp_async16(dst, power + ((size_t)tt * n_chan + f) * n_beams + b0 + q * 4);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        dst[e] = (tt < n_t && b0 + q * 4 + e < n_beams) ? power[((size_t)tt * n_chan + f) * n_beams + b0 + q * 4 + e] : 0.f;
                }
            }
        }
        cp_async_commit();
    };

    stage(0, 0);
    for (int c = 0; c < n_chunks; c++) {
        const int buf = c & 1;
        if (c + 1 < n_chunks) { stage(c + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();                                   // chunk c (and its metadata) visible to all warps
        if (my_dm) {
#pragma unroll
            for (int fi = 0; fi < DD_FC; fi++) {
                const int f = c * DD_FC + fi;
                if (f >= n_chan) break;
                const int rows = meta[buf].rows[fi];
                const int off = meta[buf].off[fi][warp];
                if (rows > 0) {
                    const float *w0 = &win[buf][fi][off][lane];
#pragma unroll
                    for (int t = 0; t < DD_TT; t++) acc[t] += w0[t * DD_BB];
                } else if (my_beam) {
                    // this group's delays are too spread out at this channel: read the rows directly
                    const int dl = meta[buf].dmin[fi] + off;
#pragma unroll
                    for (int t = 0; t < DD_TT; t++) {
                        const int tt = t0 + t + dl;
                        if (tt < n_t) acc[t] += power[((size_t)tt * n_chan + f) * n_beams + b0 + lane];
                    }
                }
            }
        }
        __syncthreads();                                   // everybody done with buf before it is re-staged
    }

    if (my_dm && my_beam) {
        float *o = out + ((size_t)(d0 + warp) * n_t_out + t0) * n_beams + b0 + lane;
#pragma unroll
        for (int t = 0; t < DD_TT; t++)
            if (t0 + t < n_t_out) o[(size_t)t * n_beams] = acc[t];
    }
}

}  // namespace

cudaError_t launch_dedisperse(const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_dm, int n_t_out,
                              float *d_out, cudaStream_t stream)
{
    if (n_dm <= 0 || n_t_out <= 0) return cudaSuccess;
    dim3 grid((n_beams + DD_BB - 1) / DD_BB, (n_t_out + DD_TT - 1) / DD_TT, (n_dm + DD_DMS - 1) / DD_DMS);
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
    constexpr size_t smem = sizeof(float) * 2 * DD_FC * DD_WIN * DD_BB + 2 * sizeof(ChunkMeta);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dedisperse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dedisperse_kernel<<<grid, DD_THREADS, smem, stream>>>(d_power, n_t, n_chan, n_beams, d_delay, n_dm, n_t_out, d_out);
    return cudaGetLastError();
}

}  // namespace dsabf

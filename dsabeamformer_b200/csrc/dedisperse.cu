// dedisperse.cu -- incoherent dedispersion of beamformed power over DM trials.
//
// Generalises the reference's DEBUG-mode "dedispersion at DM = 0" -- a cublasSgemv
// with a ones vector over the channel axis, /root/reference/src/beamformer.cu:498-504
// -- to the delay table of sandbox/Dispersion Theory.ipynb cell 5:
//     out[d][t][b] = sum_f power[t + delay[d][f]][f][b]
// power is the fused kernel's output layout [time][channel][beam] (beam fastest).
//
// Blocking.  One CTA = 8 warps owns (DD_TT = 64 output times) x (32 beams) x (8 DM trials): a warp
// accumulates a pair of adjacent trials over half of the times, lane l beam b0+l, 64 accumulators per thread.
// Channels are processed in chunks of 4.  Neighbouring DM trials have nearly equal delays,
// so for one channel the union of the rows the 8 trials need is a short window
// [t0+dmin, t0+dmax+DD_TT) of 128-byte rows; the window of each channel of the chunk is staged in
// shared memory with cp.async (double-buffered against the arithmetic on the previous chunk)
// and every warp then reads its 32 rows from there (consecutive lanes = consecutive beams:
// conflict-free).  Each staged value is used by up to 8 trials, so global traffic per add is
// ~1/16 of a float; every add costs one 4-byte shared-memory read per lane, which is the
// structural bound of brute-force dedispersion (32 adds/clk/SM).  With few trials the kernel is
// HBM-bound (power read once), with many it is shared-memory-bound.  Fewer than 8 trials: the warps
// split the output times (TSPLIT); a grid under two waves runs a 4-deep cp.async pipeline (STAGES).
#include <cstdint>
#include <cstdlib>
#include "fused_beamform.cuh"

namespace dsabf {
namespace {

#ifndef DSABF_DD_TT
#define DSABF_DD_TT 64           // 64 x 96 rows against 32 x 64: staging, metadata and barriers per add halve (+10..14 % at 64 / 256 / 1627 trials)
#endif
constexpr int DD_TT = DSABF_DD_TT;   // output times per CTA
constexpr int DD_BB = 32;        // beams per CTA (one per lane)
constexpr int DD_DMS = 8;        // DM trials per CTA (one per warp)
#ifndef DSABF_DD_FC
#define DSABF_DD_FC 4
#endif
#ifndef DSABF_DD_WIN
#define DSABF_DD_WIN 96
#endif
constexpr int DD_FC = DSABF_DD_FC;   // channels per staged chunk
constexpr int DD_WIN = DSABF_DD_WIN; // staged rows per channel: (dmax - dmin) + DD_TT <= DD_WIN, else direct global reads
constexpr int DD_THREADS = DD_DMS * 32;
static_assert(DD_FC * DD_DMS == 32, "the chunk metadata is computed by exactly one full warp (full-mask shuffles): DD_FC * DD_DMS must be 32");

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

template <int STAGES, int TSPLIT>
__global__ void __launch_bounds__(DD_THREADS)
dedisperse_kernel(const float *__restrict__ power, int n_t, int n_chan, int n_beams, const int32_t *__restrict__ delay, int n_dm,
                  int n_t_out, float *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t dd_smem[];
    float (*win)[DD_FC][DD_WIN][DD_BB] = reinterpret_cast<float (*)[DD_FC][DD_WIN][DD_BB]>(dd_smem);               // [STAGES]
    ChunkMeta *meta = reinterpret_cast<ChunkMeta *>(dd_smem + sizeof(float) * STAGES * DD_FC * DD_WIN * DD_BB);  // [STAGES]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // DM slab fastest: the CTAs that read the same (time, beam) window of the power run together and share it in L2
    const int d0 = blockIdx.x * (DD_DMS / TSPLIT), t0 = blockIdx.y * DD_TT, b0 = blockIdx.z * DD_BB;
    const int nd = min(DD_DMS / TSPLIT, n_dm - d0);
    // With fewer than 8 trials per CTA the warps split the output times instead (TSPLIT slices of TW times)
    constexpr int TW = DD_TT / TSPLIT;
    const int dw = warp % (DD_DMS / TSPLIT), tofs = (warp / (DD_DMS / TSPLIT)) * TW;
    const bool my_dm = dw < nd;
    const bool my_beam = (b0 + lane) < n_beams;
    const int n_chunks = (n_chan + DD_FC - 1) / DD_FC;

    float acc[TW];
#pragma unroll
    for (int t = 0; t < TW; t++) acc[t] = 0.f;
    // TSPLIT == 1 (8 trials per CTA): a warp owns a PAIR of adjacent trials and half of the output times.  Adjacent
    // trials' delays at one channel differ by 0 or 1 sample almost always (725 samples spread over 1627 trials at the
    // bottom of the band), so the pair shares its shared-memory reads: 16 or 17 loads feed 32 adds instead of 32.
    // acc[] then holds trial A in [0, PT) and trial B in [PT, 2 PT).
    constexpr int PT = DD_TT / 2;
    const int pr = warp & 3, th = warp >> 2;
    const bool has_a = 2 * pr < nd, has_b = 2 * pr + 1 < nd;

    // delay of (channel fi of chunk c, trial w) for the 32 metadata threads; fetched one iteration ahead of its
    // use so that the global-load latency is off the per-chunk critical path
    auto load_delay = [&](int c) -> int {
        if (threadIdx.x >= DD_FC * DD_DMS) return 0;
        const int fi = threadIdx.x / DD_DMS, w = threadIdx.x % DD_DMS;
        const int f = c * DD_FC + fi;
        return (w < nd && f < n_chan) ? delay[(size_t)(d0 + w) * n_chan + f] : 0x3fffffff;
    };
    // stage chunk c into buffer buf: metadata by the first 32 threads, rows by everybody (cp.async, 16 B)
    auto stage = [&](int c, int buf, int dl) {
        if (threadIdx.x < DD_FC * DD_DMS) {
            const int fi = threadIdx.x / DD_DMS, w = threadIdx.x % DD_DMS;
            const int f = c * DD_FC + fi;
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

    // STAGES-deep cp.async pipeline: chunks c+1 .. c+STAGES-1 are in flight while chunk c is summed.  One commit
    // group per iteration (empty past the end) keeps the group arithmetic uniform.
    for (int c = 0; c < STAGES - 1; c++) {
        if (c < n_chunks) stage(c, c, load_delay(c)); else cp_async_commit();
    }
    int dl_next = load_delay(STAGES - 1);
    for (int c = 0; c < n_chunks; c++) {
        const int buf = c % STAGES;
        const int dl_now = dl_next;
        dl_next = load_delay(c + STAGES);
        if (c + STAGES - 1 < n_chunks) stage(c + STAGES - 1, (c + STAGES - 1) % STAGES, dl_now); else cp_async_commit();
        cp_async_wait<STAGES - 1>();
        __syncthreads();                                   // chunk c (and its metadata) visible to all warps
        if (TSPLIT == 1) {
            if (has_a) {
#pragma unroll
                for (int fi = 0; fi < DD_FC; fi++) {
                    const int f = c * DD_FC + fi;
                    if (f >= n_chan) break;
                    const int rows = meta[buf].rows[fi];
                    const int offa = meta[buf].off[fi][2 * pr];
                    const int offb = has_b ? meta[buf].off[fi][2 * pr + 1] : offa;
                    if (rows > 0) {
                        const float *wa = &win[buf][fi][offa + th * PT][lane];
                        const int delta = offb - offa;
                        if (delta == 0) {
#pragma unroll
                            for (int t = 0; t < PT; t++) { const float v = wa[t * DD_BB]; acc[t] += v; acc[PT + t] += v; }
                        } else if (delta == 1) {
                            float v[PT + 1];
#pragma unroll
                            for (int t = 0; t <= PT; t++) v[t] = wa[t * DD_BB];
#pragma unroll
                            for (int t = 0; t < PT; t++) { acc[t] += v[t]; acc[PT + t] += v[t + 1]; }
                        } else {
                            const float *wb = &win[buf][fi][offb + th * PT][lane];
#pragma unroll
                            for (int t = 0; t < PT; t++) { acc[t] += wa[t * DD_BB]; acc[PT + t] += wb[t * DD_BB]; }
                        }
                    } else if (my_beam) {
                        // this group's delays are too spread out at this channel: read the rows directly
                        const int dla = meta[buf].dmin[fi] + offa, dlb = meta[buf].dmin[fi] + offb;
#pragma unroll
                        for (int t = 0; t < PT; t++) {
                            const int ta = t0 + th * PT + t + dla, tb = t0 + th * PT + t + dlb;
                            if (ta < n_t) acc[t] += power[((size_t)ta * n_chan + f) * n_beams + b0 + lane];
                            if (tb < n_t) acc[PT + t] += power[((size_t)tb * n_chan + f) * n_beams + b0 + lane];
                        }
                    }
                }
            }
        } else if (my_dm) {
#pragma unroll
            for (int fi = 0; fi < DD_FC; fi++) {
                const int f = c * DD_FC + fi;
                if (f >= n_chan) break;
                const int rows = meta[buf].rows[fi];
                const int off = meta[buf].off[fi][dw];
                if (rows > 0) {
                    const float *w0 = &win[buf][fi][off + tofs][lane];
#pragma unroll
                    for (int t = 0; t < TW; t++) acc[t] += w0[t * DD_BB];
                } else if (my_beam) {
                    // this group's delays are too spread out at this channel: read the rows directly
                    const int dl = meta[buf].dmin[fi] + off;
#pragma unroll
                    for (int t = 0; t < TW; t++) {
                        const int tt = t0 + tofs + t + dl;
                        if (tt < n_t) acc[t] += power[((size_t)tt * n_chan + f) * n_beams + b0 + lane];
                    }
                }
            }
        }
        __syncthreads();                                   // everybody done with buf before it is re-staged
    }

    if (TSPLIT == 1) {
        if (my_beam) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (h == 0 ? has_a : has_b) {
                    float *o = out + ((size_t)(d0 + 2 * pr + h) * n_t_out + t0 + th * PT) * n_beams + b0 + lane;
#pragma unroll
                    for (int t = 0; t < PT; t++)
                        if (t0 + th * PT + t < n_t_out) o[(size_t)t * n_beams] = acc[h * PT + t];
                }
            }
        }
    } else if (my_dm && my_beam) {
        float *o = out + ((size_t)(d0 + dw) * n_t_out + t0 + tofs) * n_beams + b0 + lane;
#pragma unroll
        for (int t = 0; t < TW; t++)
            if (t0 + tofs + t < n_t_out) o[(size_t)t * n_beams] = acc[t];
    }
}

// One trial (DM 0 = the reference's DEBUG cublasSgemv, beamformer.cu:498-504, or any single delay curve): nothing is
// shared between trials, so skip shared memory.  Thread = (output time, 4 beams): 16-byte loads, coalesced across the
// warp (512 contiguous bytes), 8 channels in flight per thread, channels added in order.  Pure streaming: HBM-bound.
constexpr int DS_UNROLL = 8;
constexpr int DS_MAX_SEG = 16;
// The channel axis may be split into `n_seg` segments of c_seg channels living in different buffers -- the peers'
// power buffers mapped over NVLink: full-band DM 0 across GPUs is then ONE kernel that reads every remote element
// exactly once (exchange fused into the sum, nothing materialised).  Local case: one segment.
struct Segments {
    const float *p[DS_MAX_SEG];
    int n_seg, c_seg;          // segments, channels per segment (c_seg % DS_UNROLL == 0 when n_seg > 1)
    int row_floats, col_off;   // floats per (time, channel) row of a segment; first beam of this GPU's slice in it
};
// SPLIT > 1: small problems do not fill the machine with one thread per (time, 4 beams); SPLIT threads then share the
// channel range of an item and slice 0 adds the partial sums in slice order (deterministic).
template <int SPLIT>
__global__ void __launch_bounds__(256)
dedisperse_single_kernel(Segments seg, int n_t, int n_chan, int n_beams, const int32_t *__restrict__ delay, int n_t_out, float *__restrict__ out)
{
    constexpr int ITEMS = 256 / SPLIT;
    __shared__ float4 red[SPLIT > 1 ? SPLIT - 1 : 1][ITEMS];
    const int b4 = n_beams >> 2;                                      // float4 per output row
    const int li = threadIdx.x % ITEMS, slice = threadIdx.x / ITEMS;
    const long long idx = (long long)blockIdx.x * ITEMS + li;
    const bool live = idx < (long long)n_t_out * b4;
    const int t = live ? (int)(idx / b4) : 0, q = live ? (int)(idx - (long long)t * b4) : 0;
    const size_t rowf = (size_t)seg.row_floats;
    const size_t qoff = (size_t)seg.col_off + 4 * (size_t)q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int per = (n_chan + SPLIT - 1) / SPLIT;
    per = (per + DS_UNROLL - 1) / DS_UNROLL * DS_UNROLL;              // slices start on an unroll boundary (never straddle a segment)
    int f = slice * per;
    const int f_end = min(n_chan, f + per);
    auto ld = [](const float *a) {
        float4 v;
        asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a));
        return v;
    };
    if (live) {
        for (; f + DS_UNROLL <= f_end; f += DS_UNROLL) {
            const int g = (seg.n_seg == 1) ? 0 : f / seg.c_seg;
            const int c0 = f - g * seg.c_seg;
            const float *base = seg.p[g];
            float4 v[DS_UNROLL];
#pragma unroll
            for (int u = 0; u < DS_UNROLL; u++) {
                const int tt = t + __ldg(delay + f + u);
                v[u] = (tt < n_t) ? ld(base + ((size_t)tt * seg.c_seg + c0 + u) * rowf + qoff) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < DS_UNROLL; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; f < f_end; f++) {                                      // tail: only in the one-segment case
            const int tt = t + __ldg(delay + f);
            if (tt < n_t) {
                const float4 v = ld(seg.p[0] + ((size_t)tt * seg.c_seg + f) * rowf + qoff);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    if (SPLIT > 1) {
        if (slice > 0) red[slice - 1][li] = acc;
        __syncthreads();
        if (slice == 0) {
#pragma unroll
            for (int s2 = 0; s2 < SPLIT - 1; s2++) { const float4 r = red[s2][li]; acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w; }
        }
    }
    if (live && slice == 0) reinterpret_cast<float4 *>(out)[(size_t)t * b4 + q] = acc;
}

cudaError_t launch_single(const Segments &seg, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_t_out, float *d_out, cudaStream_t stream)
{
    const long long n = (long long)n_t_out * (n_beams >> 2);
    // fewer than ~4 resident warps' worth of items per SM: split the channel range over 4 or 8 threads
    if (n >= 65536) dedisperse_single_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seg, n_t, n_chan, n_beams, d_delay, n_t_out, d_out);
    else if (n >= 16384) dedisperse_single_kernel<4><<<(unsigned)((n + 63) / 64), 256, 0, stream>>>(seg, n_t, n_chan, n_beams, d_delay, n_t_out, d_out);
    else dedisperse_single_kernel<8><<<(unsigned)((n + 31) / 32), 256, 0, stream>>>(seg, n_t, n_chan, n_beams, d_delay, n_t_out, d_out);
    return cudaGetLastError();
}

// A few trials (2..8): still one pass over the power.  Same streaming structure as the single-trial kernel -- thread =
// (output time, 4 beams), 16-byte loads, channels in order -- with ND accumulators per thread.  The ND loads of a
// channel hit rows a few samples apart; they are allowed to allocate in L1 (the threads of a block that cover
// neighbouring times read the same rows for neighbouring trials) and otherwise come from L2, so HBM still sees every
// row about once.  The staged kernel below needs ~8 trials per CTA to pay for its shared-memory staging and
// barriers; under that it was latency-bound (1.3-1.6 TB/s of algorithmic bytes at 2-8 trials, r01).
template <int ND, int UNROLL, int BLK>
__global__ void __launch_bounds__(BLK)
dedisperse_few_kernel(const float *__restrict__ power, int n_t, int n_chan, int n_beams, const int32_t *__restrict__ delay, int n_t_out,
                      float *__restrict__ out)
{
    // Row-major items: a block of BLK threads covers BLK / (n_beams/4) consecutive output times of every beam quad, i.e. whole
    // 1 KB rows at 256 beams.  BLK = 1024 (16 times x 256 beams): the ND trials of a channel read rows a few samples apart,
    // which the other threads of the block have just pulled into L1, so L2 sees each row ~1.3 times instead of ND times.
    // (A block shape of 32 times x 8 quads -- 128-byte row pieces -- was measured slower than this one: 3.2 vs 5.4 TB/s at 2 trials.)
    const int b4 = n_beams >> 2;
    const long long idx = (long long)blockIdx.x * BLK + threadIdx.x;
    if (idx >= (long long)n_t_out * b4) return;
    const int t = (int)(idx / b4), q = (int)(idx - (long long)t * b4);
    const float4 *p4 = reinterpret_cast<const float4 *>(power) + q;
    const size_t row4 = (size_t)b4;                                  // float4 per (time, channel) row
    float4 acc[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) acc[d] = make_float4(0.f, 0.f, 0.f, 0.f);
    int f = 0;
    // UNROLL channels per step.  Trial 0's rows are loaded for all of them at once (UNROLL independent 16-byte loads in
    // flight); every further trial reuses the row held for the previous trial whenever its delay at that channel is the same
    // -- neighbouring trials of a fine DM grid differ at few channels, at none at the top of the band -- and reloads it
    // otherwise (from L1 / L2: a neighbouring thread has just read that row).  The delays are the same for every thread, so
    // the branches are uniform; each trial still adds its channels in ascending order (results unchanged bit for bit).
    const bool vec_delay = UNROLL == 4 && (n_chan & 3) == 0 && ((uintptr_t)delay & 15) == 0;
    for (; f + UNROLL <= n_chan; f += UNROLL) {
        float4 cur[UNROLL];
        int dlp[UNROLL];
#pragma unroll
        for (int d = 0; d < ND; d++) {
            int dl[UNROLL];
            if (vec_delay) {
                const int4 q4 = __ldg(reinterpret_cast<const int4 *>(delay + (size_t)d * n_chan + f));
                dl[0] = q4.x; dl[1 % UNROLL] = q4.y; dl[2 % UNROLL] = q4.z; dl[3 % UNROLL] = q4.w;
            } else {
#pragma unroll
                for (int u = 0; u < UNROLL; u++) dl[u] = __ldg(delay + (size_t)d * n_chan + f + u);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                if (d == 0 || dl[u] != dlp[u]) {
                    const int tt = t + dl[u];
                    cur[u] = (tt < n_t) ? __ldg(p4 + ((size_t)tt * n_chan + f + u) * row4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dlp[u] = dl[u];
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) { acc[d].x += cur[u].x; acc[d].y += cur[u].y; acc[d].z += cur[u].z; acc[d].w += cur[u].w; }
        }
    }
    for (; f < n_chan; f++)
#pragma unroll
        for (int d = 0; d < ND; d++) {
            const int tt = t + __ldg(delay + (size_t)d * n_chan + f);
            if (tt < n_t) {
                const float4 v = __ldg(p4 + ((size_t)tt * n_chan + f) * row4);
                acc[d].x += v.x; acc[d].y += v.y; acc[d].z += v.z; acc[d].w += v.w;
            }
        }
#pragma unroll
    for (int d = 0; d < ND; d++) reinterpret_cast<float4 *>(out)[((size_t)d * n_t_out + t) * b4 + q] = acc[d];
}

template <int ND, int UNROLL>
cudaError_t launch_few_u(int blk, const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_t_out, float *d_out, cudaStream_t stream)
{
    const long long n = (long long)n_t_out * (n_beams >> 2);
    if (blk == 1024)     dedisperse_few_kernel<ND, UNROLL, 1024><<<(unsigned)((n + 1023) / 1024), 1024, 0, stream>>>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out);
    else if (blk == 512) dedisperse_few_kernel<ND, UNROLL, 512><<<(unsigned)((n + 511) / 512), 512, 0, stream>>>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out);
    else                 dedisperse_few_kernel<ND, UNROLL, 256><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out);
    return cudaGetLastError();
}

template <int ND>
cudaError_t launch_few(const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_t_out, float *d_out, cudaStream_t stream)
{
    const char *e = getenv("DSABF_DEDISP_BLOCK");                      // experiment hooks: block size 256 / 512 / 1024, channels per step 2 / 4
    const char *eu = getenv("DSABF_DEDISP_UNROLL");
    // measured, 2048 times x 2048 channels x 256 beams, neighbouring trials (profiles/r02_dedisperse_timing.jsonl, group 4), TB/s of
    // algorithmic bytes at the setting chosen here: 2: 7.0, 3: 6.5, 4: 6.2, 5: 5.3, 6: 4.0, 7: 3.6, 8: 3.3
    const int blk = e ? atoi(e) : (ND <= 4 || ND == 8 ? 1024 : (ND == 7 ? 512 : 256));
    const int unroll = eu ? atoi(eu) : (ND == 6 ? 2 : 4);
    if (unroll == 2) return launch_few_u<ND, 2>(blk, d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
    return launch_few_u<ND, 4>(blk, d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
}

template <int STAGES, int TSPLIT>
cudaError_t launch_stages(dim3 grid, const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_dm, int n_t_out,
                          float *d_out, cudaStream_t stream)
{
    constexpr size_t smem = sizeof(float) * STAGES * DD_FC * DD_WIN * DD_BB + STAGES * sizeof(ChunkMeta);
    // The attribute belongs to the (function, device) pair and a process may hold handles on several GPUs from
    // several threads: set it on every launch for the current device (a cheap driver call) instead of caching it.
    cudaError_t e = cudaFuncSetAttribute(dedisperse_kernel<STAGES, TSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dedisperse_kernel<STAGES, TSPLIT><<<grid, DD_THREADS, smem, stream>>>(d_power, n_t, n_chan, n_beams, d_delay, n_dm, n_t_out, d_out);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_dedisperse(const float *d_power, int n_t, int n_chan, int n_beams, const int32_t *d_delay, int n_dm, int n_t_out,
                              float *d_out, cudaStream_t stream)
{
    if (n_dm <= 0 || n_t_out <= 0) return cudaSuccess;
    if (n_dm == 1 && (n_beams & 3) == 0 && ((uintptr_t)d_power & 15) == 0 && ((uintptr_t)d_out & 15) == 0) {
        Segments seg{};
        seg.p[0] = d_power; seg.n_seg = 1; seg.c_seg = n_chan; seg.row_floats = n_beams; seg.col_off = 0;
        return launch_single(seg, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
    }
    // 2..8 trials on a problem large enough to fill the machine with one thread per (time, 4 beams): one streaming pass
    // (measured, 2048 times x 2048 channels x 256 beams, TB/s of algorithmic bytes, neighbouring trials: 2 trials 7.0, 3: 6.5, 4: 6.2,
    // 5: 5.3, 6: 4.0, 7: 3.6, 8: 3.3; every 100th trial of the list, i.e. no shared rows: 5.2, 3.5, 3.3, 1.3 (6), 1.7 (8) -- the
    // staged kernel: 1.3-1.6 throughout; at 512 times, too few threads, the staged kernel is as fast or faster)
    const bool force_few = getenv("DSABF_DEDISP_FORCE_FEW") != nullptr;              // testing hook (read per call)
    const long long few_items = (long long)n_t_out * (n_beams >> 2);
    if (n_dm >= 2 && n_dm <= 8 && (n_beams & 3) == 0 && ((uintptr_t)d_power & 15) == 0 && ((uintptr_t)d_out & 15) == 0 &&
        (force_few ? few_items >= 32768 : few_items >= 98304)) {
        switch (n_dm) {
        case 2: return launch_few<2>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        case 3: return launch_few<3>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        case 4: return launch_few<4>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        case 5: return launch_few<5>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        case 6: return launch_few<6>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        case 7: return launch_few<7>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        default: return launch_few<8>(d_power, n_t, n_chan, n_beams, d_delay, n_t_out, d_out, stream);
        }
    }
    // fewer than 8 trials: the 8 warps of a CTA split the output times instead of idling (tsplit = 8, 4, 2 for 1, 2, <=4 trials)
    const int tsplit = n_dm >= 5 ? 1 : (n_dm >= 3 ? 2 : (n_dm == 2 ? 4 : 8));
    const int dms_per_cta = DD_DMS / tsplit;
    dim3 grid((n_dm + dms_per_cta - 1) / dms_per_cta, (n_t_out + DD_TT - 1) / DD_TT, (n_beams + DD_BB - 1) / DD_BB);
    if (grid.y > 65535 || grid.z > 65535) return cudaErrorInvalidValue;
    // Few CTAs (few DM trials): each walks all channels serially and the kernel is bound by load latency, so
    // keep 3 chunks (48 KB) in flight per CTA; many CTAs: 2 stages and 3 CTAs per SM hide it better.
    int num_sms = 0, dev = 0;       // of the current device (per call: handles on different GPUs may share this process)
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const long long n_cta = (long long)grid.x * grid.y * grid.z;
    const bool deep = n_cta <= 2LL * num_sms;
#define DD_LAUNCH(S, T) launch_stages<S, T>(grid, d_power, n_t, n_chan, n_beams, d_delay, n_dm, n_t_out, d_out, stream)
    switch (tsplit) {
    case 1: return deep ? DD_LAUNCH(4, 1) : DD_LAUNCH(2, 1);
    case 2: return deep ? DD_LAUNCH(4, 2) : DD_LAUNCH(2, 2);
    case 4: return deep ? DD_LAUNCH(4, 4) : DD_LAUNCH(2, 4);
    default: return deep ? DD_LAUNCH(4, 8) : DD_LAUNCH(2, 8);
    }
#undef DD_LAUNCH
}

// Full-band single-trial dedispersion with the channel axis spread over `world` peers' buffers
// ([n_t][c_local][n_beams] each): this GPU sums its n_beams/world beams over all world*c_local channels.
cudaError_t launch_dedisperse_single_peers(const float *const *peer_power, int world, int rank, int n_t, int c_local, int n_beams,
                                           const int32_t *d_delay, int n_t_out, float *d_out, cudaStream_t stream)
{
    if (world < 1 || world > DS_MAX_SEG || rank < 0 || rank >= world || n_beams % world || (n_beams / world) % 4 || c_local % DS_UNROLL)
        return cudaErrorInvalidValue;
    if (n_t_out <= 0) return cudaSuccess;
    Segments seg{};
    for (int g = 0; g < world; g++) {
        if (!peer_power[g] || ((uintptr_t)peer_power[g] & 15)) return cudaErrorInvalidValue;
        seg.p[g] = peer_power[g];
    }
    const int bs = n_beams / world;
    seg.n_seg = world; seg.c_seg = c_local; seg.row_floats = n_beams; seg.col_off = rank * bs;
    return launch_single(seg, n_t, world * c_local, bs, d_delay, n_t_out, d_out, stream);
}

}  // namespace dsabf

// fused_epilogue.cuh -- pieces shared by the two fused beamforming kernels (fused_beamform.cu: resident
// weights, 64 antennas x <= 256 beams; fused_beamform_big.cu: streamed weights, up to 256 antennas x 4096 beams):
// tile constants, warp-role layout, and the TMEM -> power epilogue arithmetic.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

#include "fused_beamform.cuh"
#include "sm100_ptx.cuh"

namespace dsabf {
namespace fused {

using namespace ptx;

constexpr int kTileN = 256;              // columns per work item (UMMA N)
constexpr int kRowBytes = 128;           // K' bytes per row = 2 * 64 antennas
constexpr int kATileBytes = 128 * kRowBytes;   // 16 KB: one M=128 weights tile
constexpr int kBTileBytes = kTileN * kRowBytes; // 32 KB: unpacked voltages tile
constexpr int kStageBytes = kTileN * 64;        // 16 KB: packed voltages tile
constexpr int kStages = 4;
constexpr int kMaxMT = 4;
constexpr int kAccCols = 256;            // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;

// warp roles: warpgroup 0 = {TMA, MMA, weights+TMEM alloc, idle}, warpgroup 1 = unpack,
// warpgroups 2-5 = epilogue (16 warps: TMEM lane quarter = warp % 4, 64 columns each)
constexpr int kUnpackWarp0 = 4, kNumUnpackWarps = 4;
constexpr int kEpiWarp0 = 8, kNumEpiWarps = 16;
constexpr int kThreads = (kEpiWarp0 + kNumEpiWarps) * 32;   // 768
constexpr int kEpiCols = kTileN / (kNumEpiWarps / 4);       // 64 columns per epilogue warp
constexpr int kRegsLight = 40, kRegsEpi = 96;               // setmaxnreg budgets (launch: 80)

// Square-accumulate of one row's voltages in the two supported arithmetic modes.
// Both keep several independent partial sums so that consecutive columns do not
// form one long dependent chain.
//  EXACT : int64 via mad.wide.s32 (IMAD.WIDE) -- integer power exact, slower.
//  fast  : fp32; two columns at a time are converted (I2FP) and fed to one packed
//          fma.rn.f32x2 (FFMA2).  Exact whenever the integer power is < 2^24
//          (every partial sum is then an exactly representable integer multiple
//          of 256); otherwise <= G * 2^-24 relative, ~1e-7 typical.
template <bool EXACT> struct Acc;
template <> struct Acc<true> {
    using sum_t = long long;
    long long s0, s1, s2, s3;
    __device__ __forceinline__ void zero() { s0 = s1 = s2 = s3 = 0; }
    template <int SLOT> __device__ __forceinline__ void add2(int v0, int v1)
    {
        if (SLOT == 0) {
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s0) : "r"(v0));
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s1) : "r"(v1));
        } else {
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s2) : "r"(v0));
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s3) : "r"(v1));
        }
    }
    // sum of the first NSLOT packed slots that were used since zero()
    template <int NSLOT> __device__ __forceinline__ long long sum() const { return NSLOT == 1 ? (s0 + s1) : (s0 + s1) + (s2 + s3); }
};
template <> struct Acc<false> {
    using sum_t = float;
    unsigned long long a0, a1;     // two packed float2 accumulators
    __device__ __forceinline__ void zero() { a0 = 0ull; a1 = 0ull; }
    template <int SLOT> __device__ __forceinline__ void add2(int v0, int v1)
    {
        float f0, f1;
        asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f0) : "r"(v0));
        asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f1) : "r"(v1));
        if (SLOT == 0) asm("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a0) : "f"(f0), "f"(f1));
        else           asm("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a1) : "f"(f0), "f"(f1));
    }
    // Biased columns (kBiasCols): a fifth MMA (all-255 u8 x all-255 u8, K = 32) has added kBiasValue = 2 080 800 >=
    // max |V'| to these accumulators, so the int32 is a positive number < 2^23 whose bit pattern, read as a float, is the
    // denormal acc * 2^-149.  One FMA turns it into the voltage: fma(bits, 2^127, -kBiasValue * 2^-22) = V' * 2^-22,
    // exactly -- on the FMA pipe instead of the half-rate ALU pipe (I2FP).  The squares then carry an exact factor
    // 2^-44, folded into the output scale: results stay bit-identical to the I2FP path.
    template <int SLOT> __device__ __forceinline__ void add2b(int v0, int v1)
    {
        const unsigned long long k2 = 0x7F0000007F000000ull;      // {2^127, 2^127}
        const unsigned long long c2 = 0xBEFE0100BEFE0100ull;      // {-2080800 * 2^-22, same}
        if (SLOT == 0) asm("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 t, t, %3, %4; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a0) : "r"(v0), "r"(v1), "l"(k2), "l"(c2));
        else           asm("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 t, t, %3, %4; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a1) : "r"(v0), "r"(v1), "l"(k2), "l"(c2));
    }
    template <int NSLOT> __device__ __forceinline__ float sum() const
    {
        float a, b, c, d;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(a0));
        if (NSLOT == 1) return a + b;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(c), "=f"(d) : "l"(a1));
        return (a + b) + (c + d);
    }
};

// Columns [0, kBiasCols) of every 256-column accumulator tile of the resident-weights kernel get the bias MMA and the
// FMA-pipe conversion (Acc<false>::add2b); a multiple of 64 = whole epilogue-warp slices.  0 = off.
#ifndef DSABF_EPI_BIAS_COLS
#define DSABF_EPI_BIAS_COLS 256      // all four slices: measured 1-2 % (observation shape) to 5 % (DEBUG shape) faster than I2FP, bit-identical
#endif
#if DSABF_EPI_BIAS_COLS > 0 && defined(__CUDA_FTZ)
#error "the bias-MMA epilogue reads int32 accumulators as denormal floats: do not compile with -ftz=true / -use_fast_math"
#endif
constexpr int kBiasCols = DSABF_EPI_BIAS_COLS;
static_assert(kBiasCols % 64 == 0 && kBiasCols >= 0 && kBiasCols <= 256, "bias columns: whole 64-column slices");
constexpr float kBiasUnscale = 17592186044416.0f;          // 2^44: the biased slices accumulate (V' * 2^-22)^2

// Work-item geometry shared by all roles.
struct Geom {
    int n_chan, n_chunks, cpc, n_tiles, n_cols_total;
    __device__ __forceinline__ Geom(const FusedParams &p)
        : n_chan(p.n_chan), n_chunks(p.n_chunks), cpc(p.cols_per_chunk), n_tiles(p.n_tiles),
          n_cols_total(p.n_chunks * p.cols_per_chunk) {}
};

// Per-(row, M-tile) epilogue state of one thread.
template <bool EXACT> struct EpiCtx {
    using sum_t = typename Acc<EXACT>::sum_t;
    Acc<EXACT> acc;       // running sum of squares of the current group of G columns
    sum_t pend;           // finished even group waiting for the odd group of its pair
    int lane_odd;         // odd lanes hold im rows
    int vg_lane;          // (valid groups of this warp's 64 columns) - lane_odd   (ragged last tile)
    float *op;            // running pointer: &out[next group pair + lane_odd][f][beam of this row]
    long long *pp;        // same for the exact-power dump (nullptr = off)
    size_t pair_stride;   // elements between consecutive group pairs = 2 output times
    float scale;
};

// Store the finished output of group pair PAIR (compile-time) and advance the pointers.
template <bool EXACT>
__device__ __forceinline__ void epi_store(EpiCtx<EXACT> &c, typename EpiCtx<EXACT>::sum_t tot, int first_grp_of_pair)
{
    if (first_grp_of_pair < c.vg_lane) {
        *c.op = (float)tot * c.scale;
        if (EXACT) {
            if (c.pp) *c.pp = (long long)tot >> 8;
        }
    }
    c.op += c.pair_stride;
    if (EXACT) {
        if (c.pp) c.pp += c.pair_stride;
    }
}

// 32 accumulator columns [BLK*32, BLK*32+32) of this warp's 64: square, sum per
// group of G columns, and on every second finished group exchange with the lane
// holding the other component (re rows = even lanes, im rows = odd lanes): even
// lanes finish the even group of the pair, odd lanes the odd group, so one
// shuffle yields two outputs and a warp stores 2 x 64 contiguous bytes.
// BIASED: the accumulators carry the bias MMA's offset (Acc<false>::add2b).
template <int G, bool EXACT, int BLK, bool BIASED = false>
__device__ __forceinline__ void epi_block(EpiCtx<EXACT> &c, const uint32_t (&v)[32])
{
    constexpr int kGroups = kEpiCols / G;                 // groups per warp: 32,16,8,4,2,1
    constexpr int kSlots = (G >= 4) ? 2 : 1;              // packed accumulator slots in use
#pragma unroll
    for (int jj = 0; jj < 32; jj += 2) {
        const int pr = jj >> 1;                           // pair index inside this block of 32 columns
        const int slot = (kSlots == 1 || (pr & 1) == 0) ? 0 : 1;
        if constexpr (!EXACT && BIASED) {
            if (slot == 0) c.acc.template add2b<0>((int)v[jj], (int)v[jj + 1]);
            else           c.acc.template add2b<1>((int)v[jj], (int)v[jj + 1]);
        } else {
            if (slot == 0) c.acc.template add2<0>((int)v[jj], (int)v[jj + 1]);
            else           c.acc.template add2<1>((int)v[jj], (int)v[jj + 1]);
        }
        const int cl = BLK * 32 + jj + 1;          // last column handled so far (compile-time after unrolling)
        if (((cl + 1) % G) == 0) {
            const int grp = cl / G;
            const typename EpiCtx<EXACT>::sum_t sgrp = c.acc.template sum<kSlots>();
            c.acc.zero();
            if (kGroups == 1) {
                const auto tot = sgrp + __shfl_xor_sync(0xffffffffu, sgrp, 1);
                // single group: only the even lanes (vg_lane = valid - 0) store
                if (!c.lane_odd) epi_store<EXACT>(c, tot, 0);
            } else if ((grp & 1) == 0) {
                c.pend = sgrp;
            } else {
                const auto send = c.lane_odd ? c.pend : sgrp;
                const auto mine = c.lane_odd ? sgrp : c.pend;
                const auto tot = mine + __shfl_xor_sync(0xffffffffu, send, 1);
                epi_store<EXACT>(c, tot, grp - 1);
            }
        }
    }
}

}  // namespace fused
}  // namespace dsabf

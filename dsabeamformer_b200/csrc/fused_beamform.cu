// fused_beamform.cu -- ONE kernel for unpack -> beamform -> detect/average.
//
// Replaces the reference's three-step sequence (and its two HBM round trips)
//   expand_input                      /root/reference/src/beamformer.cuh:66-109
//   cublasGemmStridedBatchedEx        /root/reference/src/beamformer.cu:470-477
//   detect_sum                        /root/reference/src/beamformer.cuh:115-155
// for a whole block of chunks per launch.
//
// Math.  Per channel f the reference computes C = (1/127) W_f X_f (complex) and
// sums |C|^2 over G = n_pol*n_averaging consecutive columns.  Here the complex
// product is two real int8 GEMMs on the 5th-gen tensor cores (tcgen05.mma
// kind::i8, int32 accumulators in TMEM):
//     [Vre; Vim] = [[Wre, -Wim], [Wim, Wre]] [Xre; Xim]
// A' (weights) rows are interleaved (row 2i = re of beam i, row 2i+1 = im), K' =
// 2*64 antennas = 128 bytes = exactly one SWIZZLE_128B row, so one M=128 tile is
// 64 beams and a 256-beam channel is 4 resident 16 KB tiles.
//
// The epilogue, not the tensor core, bounds this kernel: every int32 accumulator
// receives only K' = 128 MACs, so at tensor peak 64 accumulators per clock per SM
// have to be converted and squared.  Measured B200 pipe rates (profiles/
// r01_ubench_pipe_rates.txt): I2FP issues at half rate on the ALU pipe, FFMA/FFMA2
// at full rate on the FMA pipe, also with denormal operands.  So the work is split
// over both pipes (G <= 32):
//  * columns 128..255 of a tile: I2FP (ALU) + packed fma.rn.f32x2 square-accumulate;
//  * columns 0..127: one extra MMA (K = 32 bytes of constants: 64 x 64 for these columns, 64 x 0
//    for the others) pre-loads these accumulators with 2^17 > max|V|, so I = 2^17 + V is in (0, 2^23) and its bit
//    pattern read as fp32 is the DENORMAL I*2^-149 -- already a float, linear in V.
//    One FFMA2, f = I*2^-149 * 2^126 - 2^-6 = V*2^-23 (exact), replaces the two I2FP;
//    the factor 2^46 of |f|^2 is folded into the output scale of those columns.
// Each epilogue warp takes 32 columns of either kind, so both pipes stay busy.
//
// Data flow per CTA (persistent, one CTA per SM, contiguous range of work items;
// item = (channel, 256-column tile), 4 M-tiles per item):
//   warp 0      TMA producer: cp.async.bulk packed voltages -> staging ring (4 x 16 KB)
//   warps 1, 3  MMA issuers (one thread each, one per TMEM accumulator stage): 1 constant + 4 data
//               tcgen05.mma M128 N256 K32 per M-tile, accumulators double-buffered (2 x 256 TMEM columns)
//   warp 2      weights loader (cp.async.bulk of the pre-swizzled A' tiles) + TMEM alloc
//   warps 4-11  epilogue: tcgen05.ld -> square/accumulate over G columns -> one
//               shuffle to add re^2+im^2 rows -> coalesced float stores
//   warps 12-15 unpack: staging -> sign-preserving nibble split -> SW128 K-major B' tile
// Voltages are read from HBM once, weights once per (CTA, channel); nothing but
// the final float power is written.
#include "fused_beamform.cuh"
#include "sm100_ptx.cuh"
#include <cstddef>
#include <type_traits>

namespace dsabf {

using namespace ptx;

namespace {

constexpr int kTileN = 256;              // columns per work item (UMMA N)
constexpr int kRowBytes = 128;           // K' bytes per row = 2 * 64 antennas
constexpr int kATileBytes = 128 * kRowBytes;   // 16 KB: one M=128 weights tile
constexpr int kBTileBytes = kTileN * kRowBytes; // 32 KB: unpacked voltages tile
constexpr int kStageBytes = kTileN * 64;        // 16 KB: packed voltages tile
constexpr int kStages = 4;
constexpr int kMaxMT = 4;
constexpr int kBiasABytes = 128 * 32;    // 4 KB of 0x40: A operand of the bias MMA (M=128, K=32)
constexpr int kBiasBBytes = kTileN * 32; // 8 KB: B operand (N=256, K=32): rows 0..127 = 0x40, rows 128..255 = 0
constexpr int kBias = 1 << 17;           // 32 * 64 * 64: what the bias MMA puts into columns 0..127
constexpr int kAccCols = 256;            // TMEM columns per accumulator stage
constexpr int kTmemCols = 512;

// warp roles: warpgroup 0 = {TMA, MMA, weights+TMEM alloc, idle}, warpgroup 1 = unpack,
// warpgroups 2-5 = epilogue (16 warps: TMEM lane quarter = warp % 4, 64 columns each)
constexpr int kUnpackWarp0 = 4, kNumUnpackWarps = 4;
constexpr int kEpiWarp0 = 8, kNumEpiWarps = 16;
constexpr int kThreads = (kEpiWarp0 + kNumEpiWarps) * 32;   // 768
constexpr int kRegsLight = 40, kRegsEpi = 96;               // setmaxnreg budgets (launch: 80)

struct __align__(1024) Smem {
    uint8_t a[kMaxMT][kATileBytes];      // 64 KB  resident weights of the current channel
    uint8_t b[2][kBTileBytes];           // 64 KB  unpacked voltages, double buffered
    uint8_t stg[kStages][kStageBytes];   // 64 KB  packed voltages ring
    uint8_t bias_a[kBiasABytes];         //  4 KB  constant 64 (s8)
    uint8_t bias_b[kBiasBBytes];         //  8 KB  64 for columns 0..127, 0 for columns 128..255
    uint64_t stg_full[kStages], stg_empty[kStages];
    uint64_t b_full[2], b_empty[2];
    uint64_t a_full[kMaxMT], a_empty[kMaxMT];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// Square-accumulate of one row's voltages in the two supported arithmetic modes.  BIASED
// columns hold I = 2^17 + V, the others V.  Several independent partial sums keep the
// dependent chains short.
//  EXACT : int64 via mad.wide.s32 (IMAD.WIDE) -- integer power exact, slower.
//  fast  : fp32, two columns per packed fma.rn.f32x2 (FFMA2).  Exact whenever the integer power
//          is < 2^24 (every partial sum is then exactly representable); otherwise <= G * 2^-24
//          relative (~1e-7 typical).  Biased groups carry V*2^-23, i.e. power * 2^-46.
template <bool EXACT> struct Acc;
template <> struct Acc<true> {
    using sum_t = long long;
    long long s0, s1, s2, s3;
    __device__ __forceinline__ void zero() { s0 = s1 = s2 = s3 = 0; }
    template <int SLOT, bool BIASED> __device__ __forceinline__ void add2(int i0, int i1)
    {
        const int v0 = BIASED ? i0 - kBias : i0, v1 = BIASED ? i1 - kBias : i1;
        if (SLOT == 0) {
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s0) : "r"(v0));
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s1) : "r"(v1));
        } else {
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s2) : "r"(v0));
            asm("mad.wide.s32 %0, %1, %1, %0;" : "+l"(s3) : "r"(v1));
        }
    }
    template <int NSLOT> __device__ __forceinline__ long long sum() const { return NSLOT == 1 ? (s0 + s1) : (s0 + s1) + (s2 + s3); }
};
template <> struct Acc<false> {
    using sum_t = float;
    unsigned long long a0, a1;     // two packed float2 accumulators
    __device__ __forceinline__ void zero() { a0 = 0ull; a1 = 0ull; }
    template <int SLOT, bool BIASED> __device__ __forceinline__ void add2(int i0, int i1)
    {
        if (BIASED) {
            // t = (I * 2^-149) * 2^126 - 2^17 * 2^-23 = V * 2^-23; 0x7E800000 = 2^126, 0xBC800000 = -2^-6
            if (SLOT == 0)
                asm("{.reg .b64 t, s, c; mov.b64 t, {%1, %2}; mov.b64 s, 0x7E8000007E800000; mov.b64 c, 0xBC800000BC800000;"
                    " fma.rn.f32x2 t, t, s, c; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a0) : "r"(i0), "r"(i1));
            else
                asm("{.reg .b64 t, s, c; mov.b64 t, {%1, %2}; mov.b64 s, 0x7E8000007E800000; mov.b64 c, 0xBC800000BC800000;"
                    " fma.rn.f32x2 t, t, s, c; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a1) : "r"(i0), "r"(i1));
        } else {
            float f0, f1;
            asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f0) : "r"(i0));
            asm("cvt.rn.f32.s32 %0, %1;" : "=f"(f1) : "r"(i1));
            if (SLOT == 0) asm("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a0) : "f"(f0), "f"(f1));
            else           asm("{.reg .b64 t; mov.b64 t, {%1, %2}; fma.rn.f32x2 %0, t, t, %0;}" : "+l"(a1) : "f"(f0), "f"(f1));
        }
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

// four packed 4-bit two's complement values (low nibble of every byte of w) -> four int8
__device__ __forceinline__ uint32_t sext4(uint32_t w)
{
    return ((((w & 0x0F0F0F0Fu) ^ 0x88888888u) - 0x08080808u) ^ 0x80808080u);
}

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
    sum_t pend;           // finished group waiting for its partner group
    int lane_odd;         // odd lanes hold im rows
    int vg[2];            // valid groups in this warp's first / second 32-column block, minus lane_odd
    float *op[2];         // &out[first group of the block + lane_odd][f][beam of this row]
    long long *pp[2];     // same for the exact-power dump (nullptr = off)
    size_t tstride;       // elements between consecutive output times
    float scale[2];       // output scale of the two blocks (the biased block carries 2^-46)
};

template <bool EXACT>
__device__ __forceinline__ void epi_store(const EpiCtx<EXACT> &c, typename EpiCtx<EXACT>::sum_t tot, int blk, int first_grp_of_pair)
{
    if (first_grp_of_pair < c.vg[blk]) {
        const size_t o = (size_t)first_grp_of_pair * c.tstride;
        c.op[blk][o] = (float)tot * c.scale[blk];
        if (EXACT) {
            if (c.pp[blk]) c.pp[blk][o] = (long long)tot;
        }
    }
}

// One 32-column block (BLK 0/1 of this warp's two): square, sum per group of G columns, combine the
// re row (even lane) with the im row (odd lane) by one shuffle per two groups: even lanes finish the
// first group of a pair, odd lanes the second, so a warp stores 2 x 64 contiguous bytes per pair.
//  G < 32 : pairs are consecutive groups inside the block.
//  G = 32 : the pair is (block 0, block 1): block 0 parks its sum, block 1 exchanges and stores.
//  G = 64 : one group over both blocks; only even lanes store.
template <int G, bool EXACT, int BLK, bool BIASED>
__device__ __forceinline__ void epi_block(EpiCtx<EXACT> &c, const uint32_t (&v)[32])
{
    constexpr int kSlots = (G >= 4) ? 2 : 1;              // packed accumulator slots in use
#pragma unroll
    for (int jj = 0; jj < 32; jj += 2) {
        if (kSlots == 1 || ((jj >> 1) & 1) == 0) c.acc.template add2<0, BIASED>((int)v[jj], (int)v[jj + 1]);
        else                                     c.acc.template add2<1, BIASED>((int)v[jj], (int)v[jj + 1]);
        const int cl = jj + 1;                     // last column of the block handled so far (compile-time)
        if (G <= 32 && ((cl + 1) % G) == 0) {
            const int grp = cl / G;                // group index inside the block
            const auto sgrp = c.acc.template sum<kSlots>();
            c.acc.zero();
            if (G == 32) {
                if (BLK == 0) {
                    c.pend = sgrp;
                } else {
                    const auto send = c.lane_odd ? c.pend : sgrp;        // even lanes finish block 0's group, odd lanes block 1's
                    const auto mine = c.lane_odd ? sgrp : c.pend;
                    const auto tot = mine + __shfl_xor_sync(0xffffffffu, send, 1);
                    if (c.lane_odd) epi_store<EXACT>(c, tot, 1, 0);
                    else            epi_store<EXACT>(c, tot, 0, 0);
                }
            } else if ((grp & 1) == 0) {
                c.pend = sgrp;
            } else {
                const auto send = c.lane_odd ? c.pend : sgrp;
                const auto mine = c.lane_odd ? sgrp : c.pend;
                const auto tot = mine + __shfl_xor_sync(0xffffffffu, send, 1);
                epi_store<EXACT>(c, tot, BLK, grp - 1);
            }
        }
    }
    if (G == 64 && BLK == 1) {
        const auto sgrp = c.acc.template sum<kSlots>();
        const auto tot = sgrp + __shfl_xor_sync(0xffffffffu, sgrp, 1);
        if (!c.lane_odd) epi_store<EXACT>(c, tot, 0, 0);
    }
}

// MODE: 0 = fast fp32 power, 1 = exact int64 power (+ optional integer-power dump),
//       2 = exact + int32 voltage dump (debug / parity tap)
template <int G, int MODE>
__global__ void __launch_bounds__(kThreads, 1) fused_beamform_kernel(const FusedParams p)
{
    constexpr bool EXACT = MODE != 0;
    constexpr bool DUMP = MODE == 2;
    constexpr bool kMixed = G <= 32;      // columns 0..127 of every tile carry the 2^17 bias (see header)
    extern __shared__ uint8_t smem_raw[];
    // align in the shared address space (keeps every later address a cheap 32-bit add)
    uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + (sbase - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const Geom g(p);
    const int mt = p.n_beams >> 6;

    const long long items_total = (long long)g.n_chan * g.n_tiles;
    const int item_begin = (int)((items_total * blockIdx.x) / gridDim.x);
    const int item_end = (int)((items_total * (blockIdx.x + 1)) / gridDim.x);

    // ---- one-time set-up ----
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; i++) { mbar_init(&sm.stg_full[i], 1); mbar_init(&sm.stg_empty[i], kNumUnpackWarps); }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.b_full[i], kNumUnpackWarps); mbar_init(&sm.b_empty[i], (p.n_beams >> 6) == 1 ? 1 : 2);   // one commit per MMA issuer
            mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], kNumEpiWarps);
        }
        for (int i = 0; i < kMaxMT; i++) { mbar_init(&sm.a_full[i], 1); mbar_init(&sm.a_empty[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(&sm.tmem_base, kTmemCols); tmem_relinquish(); }
    // constant operand of the bias MMA (read by the tensor core through the async proxy)
    for (int i = threadIdx.x; i < kBiasABytes / 16; i += kThreads) reinterpret_cast<uint4 *>(sm.bias_a)[i] = make_uint4(0x40404040u, 0x40404040u, 0x40404040u, 0x40404040u);
    for (int i = threadIdx.x; i < kBiasBBytes / 16; i += kThreads) {
        const uint32_t v = (i < kBiasBBytes / 32) ? 0x40404040u : 0u;      // 8-row groups are 256 B apart: first half = columns 0..127
        reinterpret_cast<uint4 *>(sm.bias_b)[i] = make_uint4(v, v, v, v);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    // 32-bit shared addresses, computed once (opaque so that they are not rematerialised in the loops)
    asm volatile("" : "+r"(sbase));
    const uint32_t a_smem = sbase + (uint32_t)offsetof(Smem, a), b_smem = sbase + (uint32_t)offsetof(Smem, b),
                   stg_smem = sbase + (uint32_t)offsetof(Smem, stg);
    const uint32_t bar_stg_full = sbase + (uint32_t)offsetof(Smem, stg_full), bar_stg_empty = sbase + (uint32_t)offsetof(Smem, stg_empty);
    const uint32_t bar_b_full = sbase + (uint32_t)offsetof(Smem, b_full), bar_b_empty = sbase + (uint32_t)offsetof(Smem, b_empty);
    const uint32_t bar_a_full = sbase + (uint32_t)offsetof(Smem, a_full), bar_a_empty = sbase + (uint32_t)offsetof(Smem, a_empty);
    const uint32_t bar_acc_full = sbase + (uint32_t)offsetof(Smem, acc_full), bar_acc_empty = sbase + (uint32_t)offsetof(Smem, acc_empty);

    if (warp < kUnpackWarp0) {
        setmaxnreg_dec<kRegsLight>();
        if (warp == 0 && lane == 0) {
            // ================= TMA producer: packed voltages =================
            const uint64_t pol = policy_evict_first();
            for (int item = item_begin, it = 0; item < item_end; item++, it++) {
                const int f = item / g.n_tiles, t = item - f * g.n_tiles;
                const int s = it % kStages;
                mbar_wait(bar_stg_empty + 8 * s, ((it / kStages) & 1) ^ 1, 1);
                const int col0 = t * kTileN;
                if (g.cpc >= kTileN) {
                    const int chunk = col0 / g.cpc, c_in = col0 - chunk * g.cpc;
                    const uint8_t *src = p.packed + (((size_t)chunk * g.n_chan + f) * g.cpc + c_in) * 64;
                    mbar_arrive_expect_tx(bar_stg_full + 8 * s, kStageBytes);
                    tma_bulk_g2s_hint(stg_smem + s * kStageBytes, src, kStageBytes, bar_stg_full + 8 * s, pol);
                } else {
                    const int per_tile = kTileN / g.cpc;
                    const int chunk0 = t * per_tile;
                    const int nseg = min(per_tile, g.n_chunks - chunk0);
                    const uint32_t seg_bytes = g.cpc * 64;
                    mbar_arrive_expect_tx(bar_stg_full + 8 * s, seg_bytes * nseg);
                    for (int k = 0; k < nseg; k++) {
                        const uint8_t *src = p.packed + (((size_t)(chunk0 + k) * g.n_chan + f) * g.cpc) * 64;
                        tma_bulk_g2s_hint(stg_smem + s * kStageBytes + k * seg_bytes, src, seg_bytes, bar_stg_full + 8 * s, pol);
                    }
                }
            }
        } else if (warp == 2 && lane == 0) {
            // ================= weights loader =================
            const uint64_t pol = policy_evict_last();
            int cur_f = -1, ch = 0;
            for (int item = item_begin; item < item_end; item++) {
                const int f = item / g.n_tiles;
                if (f == cur_f) continue;
                cur_f = f;
                for (int m = 0; m < mt; m++) {
                    mbar_wait(bar_a_empty + 8 * m, (ch & 1) ^ 1, 2);
                    mbar_arrive_expect_tx(bar_a_full + 8 * m, kATileBytes);
                    tma_bulk_g2s_hint(a_smem + m * kATileBytes, p.wprep + ((size_t)f * mt + m) * kATileBytes, kATileBytes, bar_a_full + 8 * m, pol);
                }
                ch++;
            }
        } else if ((warp == 1 || warp == 3) && lane == 0) {
            // ================= MMA issuers =================
            // Two issuing threads, one per TMEM accumulator stage (stage = parity of the M-tile sequence
            // number).  A single thread serialises  wait(acc_empty) -> issue (which blocks while the tensor
            // pipe drains its queue) -> commit  and left the tensor pipe idle ~half of the time (clock
            // trace, profiles/r01_mma_issue_trace.txt); with two threads one's waits overlap the other's MMAs.
            const int me = (warp == 3) ? 1 : 0;
            constexpr uint32_t idesc = make_idesc_i8(128, kTileN);
            const uint64_t bias_adesc = make_desc_k_nosw(sbase + (uint32_t)offsetof(Smem, bias_a));
            const uint64_t bias_bdesc = make_desc_k_nosw(sbase + (uint32_t)offsetof(Smem, bias_b));
            int cur_f = -1, ch = -1;
            uint32_t acc_cnt = 0, use = 0;          // use: how often this issuer's stage has been filled
            for (int item = item_begin, it = 0; item < item_end; item++, it++) {
                const int f = item / g.n_tiles;
                const bool new_ch = (f != cur_f);
                if (new_ch) { cur_f = f; ch++; }
                const bool last_of_ch = (item + 1 == item_end) || ((item + 1) / g.n_tiles != f);
                const int buf = it & 1;
                bool touched = false;
                uint64_t bdesc = 0;
                for (int m = 0; m < mt; m++, acc_cnt++) {
                    if ((int)(acc_cnt & 1) != me) continue;
                    if (!touched) {
                        mbar_wait(bar_b_full + 8 * buf, (it >> 1) & 1, 3);
                        bdesc = make_desc_k_sw128(b_smem + buf * kBTileBytes);
                        touched = true;
                    }
                    if (new_ch) mbar_wait(bar_a_full + 8 * m, ch & 1, 4);
                    const long long tr0 = p.trace ? clock64() : 0;
                    mbar_wait(bar_acc_empty + 8 * me, (use & 1) ^ 1, 5);
                    tc_fence_after();
                    if (p.trace && blockIdx.x == 0 && acc_cnt < 128) { p.trace[acc_cnt * 16 + 0] = tr0; p.trace[acc_cnt * 16 + 1] = clock64(); }
                    const uint64_t adesc = make_desc_k_sw128(a_smem + m * kATileBytes);
                    const uint32_t d = tmem_base + me * kAccCols;
                    // mixed: one constant MMA first (columns 0..127 := 2^17, columns 128..255 := 0), then the data
                    if (kMixed) mma_i8_ss(d, bias_adesc, bias_bdesc, idesc, 0);
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; k++)   // K = 32 bytes per instruction; +32 B = +2 in the descriptor
                        mma_i8_ss(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kMixed || k > 0) ? 1 : 0);
                    tc_commit(bar_acc_full + 8 * me);
                    if (p.trace && blockIdx.x == 0 && acc_cnt < 128) p.trace[acc_cnt * 16 + 2] = clock64();
                    if (last_of_ch) tc_commit(bar_a_empty + 8 * m);
                    use++;
                }
                if (touched) tc_commit(bar_b_empty + 8 * buf);
            }
        }
    } else if (warp < kEpiWarp0) {
        // ================= unpack: 4-bit+4-bit -> sign-extended int8, into SW128 K-major B' =================
        setmaxnreg_dec<kRegsLight>();
        const int uw = warp - kUnpackWarp0;
        // lane -> (row r in an 8-column group, 16-byte antenna chunk c): chosen so that both the
        // 16 B loads (64 B rows) and the swizzled 16 B stores (128 B rows) are bank-conflict free.
        const int q = lane >> 3, j = lane & 7, c = j & 3;
        const int r = (j < 4) ? q : (q ^ 5);
        const uint32_t ld_off = r * 64 + c * 16;
        const uint32_t st_re = r * 128 + ((c ^ r) << 4);
        const uint32_t st_im = r * 128 + (((c + 4) ^ r) << 4);
        for (int item = item_begin, it = 0; item < item_end; item++, it++) {
            const int s = it % kStages, buf = it & 1;
            const bool utr = p.trace && blockIdx.x == 0 && uw == 0 && lane == 0 && it < 32;
            if (utr) p.trace[it * 64 + 12] = clock64();
            mbar_wait(bar_stg_full + 8 * s, (it / kStages) & 1, 6);
            if (utr) p.trace[it * 64 + 13] = clock64();
            mbar_wait(bar_b_empty + 8 * buf, ((it >> 1) & 1) ^ 1, 7);
            if (utr) p.trace[it * 64 + 14] = clock64();
            const uint32_t src = stg_smem + s * kStageBytes + ld_off;
            const uint32_t dst = b_smem + buf * kBTileBytes;
#pragma unroll
            for (int i = 0; i < (kTileN / 8) / kNumUnpackWarps; i++) {
                const int grp = uw + i * kNumUnpackWarps;      // 8-column group
                const uint4 v = lds128(src + grp * 512);
                uint4 re, im;      // per byte: nibble n -> two's complement int8 ((n ^ 0x88) - 8) ^ 0x80, no carry between bytes
                re.x = sext4(v.x >> 4); im.x = sext4(v.x);
                re.y = sext4(v.y >> 4); im.y = sext4(v.y);
                re.z = sext4(v.z >> 4); im.z = sext4(v.z);
                re.w = sext4(v.w >> 4); im.w = sext4(v.w);
                sts128(dst + grp * 1024 + st_re, re);
                sts128(dst + grp * 1024 + st_im, im);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(bar_b_full + 8 * buf); mbar_arrive(bar_stg_empty + 8 * s); }
            if (utr) p.trace[it * 64 + 15] = clock64();
        }
    } else {
        // ================= epilogue: TMEM -> |V|^2 summed over G columns -> float power =================
        setmaxnreg_inc<kRegsEpi>();
        const int ew = warp - kEpiWarp0;
        const int quarter = warp & 3;              // TMEM lane quarter this warp may access
        const int wc = ew >> 2;                    // 0..3: which columns of the tile
        const int row = quarter * 32 + lane;       // A' row: beam = row/2, component = row&1
        // this warp's two 32-column blocks: mixed -> [wc*32, +32) (biased half) and [128 + wc*32, +32);
        // otherwise 64 contiguous columns [wc*64, +64)
        const int col_b0 = kMixed ? wc * 32 : wc * 64;
        const int col_b1 = kMixed ? 128 + wc * 32 : wc * 64 + 32;
        uint32_t tbase_w = tmem_base + ((uint32_t)(quarter * 32) << 16);
        asm volatile("" : "+r"(tbase_w));
        constexpr int kGpb = (G <= 32) ? 32 / G : 1;               // groups per 32-column block
        constexpr bool kLaneOff = G < 32;                          // odd lanes store the next output time (pairs inside a block)
        EpiCtx<EXACT> ctx;
        ctx.lane_odd = lane & 1;
        ctx.tstride = (size_t)g.n_chan * p.n_beams;                // floats between consecutive output times
        ctx.scale[0] = (kMixed && !EXACT) ? p.scale * 70368744177664.0f : p.scale;      // biased block carries power * 2^-46
        ctx.scale[1] = p.scale;
        uint32_t acc_cnt = 0;
#pragma unroll 1
        for (int item = item_begin; item < item_end; item++) {
            const int f = item / g.n_tiles, t = item - f * g.n_tiles;
            const int c0 = t * kTileN + col_b0, c1 = t * kTileN + col_b1;   // first global column of the two blocks
            // valid groups per block (ragged last tile), pre-biased by the lane's own offset
            const int left0 = g.n_cols_total - c0, left1 = g.n_cols_total - c1;
            if (G == 64) {
                ctx.vg[0] = (left0 >= 64) ? 1 : 0;
                ctx.vg[1] = 0;
            } else {
                ctx.vg[0] = (left0 >= 32 ? kGpb : (left0 > 0 ? left0 / G : 0)) - (kLaneOff ? ctx.lane_odd : 0);
                ctx.vg[1] = (left1 >= 32 ? kGpb : (left1 > 0 ? left1 / G : 0)) - (kLaneOff ? ctx.lane_odd : 0);
            }
            const size_t beam_off = (size_t)f * p.n_beams + (row >> 1);
            size_t off0 = ((size_t)(c0 / G) + (kLaneOff ? ctx.lane_odd : 0)) * ctx.tstride + beam_off;
            size_t off1 = ((size_t)(c1 / G) + (kLaneOff ? ctx.lane_odd : 0)) * ctx.tstride + beam_off;
            asm volatile("" : "+l"(off0), "+l"(off1));               // keep live; do not recompute per M-tile
#pragma unroll 1
            for (int m = 0; m < mt; m++, acc_cnt++) {
                const uint32_t as = acc_cnt & 1;
                const bool trc = p.trace && blockIdx.x == 0 && (ew == 0 || ew == 15) && lane == 0 && acc_cnt < 128;
                const int trb = acc_cnt * 16 + (ew == 0 ? 4 : 8);
                if (trc) p.trace[trb + 0] = clock64();
                mbar_wait(bar_acc_full + 8 * as, (acc_cnt >> 1) & 1, 8);
                tc_fence_after();
                if (trc) p.trace[trb + 1] = clock64();
                const uint32_t taddr = tbase_w + as * kAccCols;
                uint32_t va[32], vb[32];
                tmem_ld_32x32(taddr + col_b0, va);
                tmem_ld_32x32(taddr + col_b1, vb);
                tmem_ld_wait(va);
                tmem_ld_wait(vb);
                // both blocks are in registers: hand the TMEM stage back to the MMA right away
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
                // ptxas sinks the arrive below all of the math that follows (it has no consumers), which
                // keeps the TMEM stage -- and so the tensor pipe -- waiting.  A never-taken, runtime-opaque
                // branch ends the basic block here and pins it (measured: 0.825 -> 0.462 ms on the DEBUG shape).
                if (acc_cnt == 0xFFFFFFFFu) continue;
                if (trc) p.trace[trb + 2] = clock64();

                if (DUMP) {
                    // debug tap: int32 voltages, [f][col][beam][2]; this thread's (beam, component) = row
                    int32_t *vd = p.vdump + ((size_t)f * g.n_cols_total * p.n_beams + m * 64) * 2 + row;
                    const size_t cs = (size_t)p.n_beams * 2;
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) {
                        if (c0 + jj < g.n_cols_total) vd[(size_t)(c0 + jj) * cs] = (int)va[jj] - (kMixed ? kBias : 0);
                        if (c1 + jj < g.n_cols_total) vd[(size_t)(c1 + jj) * cs] = (int)vb[jj];
                    }
                }
                ctx.acc.zero();
                ctx.pend = 0;
                ctx.op[0] = p.out + off0 + m * 64;                   // M-tile m = beams [64m, 64m+64)
                ctx.op[1] = p.out + off1 + m * 64;
                ctx.pp[0] = (EXACT && p.pdump) ? p.pdump + off0 + m * 64 : nullptr;
                ctx.pp[1] = (EXACT && p.pdump) ? p.pdump + off1 + m * 64 : nullptr;
                epi_block<G, EXACT, 0, kMixed>(ctx, va);
                epi_block<G, EXACT, 1, false>(ctx, vb);
                if (trc) p.trace[trb + 3] = clock64();
            }
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

template <int G, int MODE>
cudaError_t launch_one(const FusedParams &p, int grid, cudaStream_t stream)
{
    const size_t smem = sizeof(Smem) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(fused_beamform_kernel<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    fused_beamform_kernel<G, MODE><<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

template <int MODE>
cudaError_t launch_g(const FusedParams &p, int grid, cudaStream_t stream)
{
    switch (p.group) {
    case 2: return launch_one<2, MODE>(p, grid, stream);
    case 4: return launch_one<4, MODE>(p, grid, stream);
    case 8: return launch_one<8, MODE>(p, grid, stream);
    case 16: return launch_one<16, MODE>(p, grid, stream);
    case 32: return launch_one<32, MODE>(p, grid, stream);
    case 64: return launch_one<64, MODE>(p, grid, stream);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace

cudaError_t launch_fused_beamform(const FusedParams &p, bool exact, int num_sms, cudaStream_t stream)
{
    if (p.n_beams % 64 || p.n_beams < 64 || p.n_beams > 64 * kMaxMT) return cudaErrorInvalidValue;
    if (p.cols_per_chunk % p.group) return cudaErrorInvalidValue;
    if (!((p.cols_per_chunk >= kTileN && p.cols_per_chunk % kTileN == 0) || (kTileN % p.cols_per_chunk == 0)))
        return cudaErrorInvalidValue;
    const long long items = (long long)p.n_chan * p.n_tiles;
    if (items <= 0) return cudaSuccess;
    const int grid = (int)(items < num_sms ? items : num_sms);
    if (p.vdump) return launch_g<2>(p, grid, stream);
    return exact ? launch_g<1>(p, grid, stream) : launch_g<0>(p, grid, stream);
}

// ---------------------------------------------------------------------------
// Weights re-layout: reference layout [chan][ant][beam][2] int8  ->  per (chan,
// M-tile) a 16 KB K-major SWIZZLE_128B image of A' (row 2i = [Wre | -Wim] of beam
// i, row 2i+1 = [Wim | Wre]), ready for a plain bulk copy into shared memory.
__global__ void prep_weights_kernel(const int8_t *__restrict__ w, uint8_t *__restrict__ wprep, int n_chan, int n_ant, int n_beams,
                                    int *__restrict__ bad)
{
    const int mt = n_beams >> 6;
    const size_t total = (size_t)n_chan * mt * 128 * 128;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int kb = idx & 127;                  // logical K byte
        const int r = (idx >> 7) & 127;            // row in the M tile
        const size_t fm = idx >> 14;
        const int m = fm % mt;
        const int f = fm / mt;
        const int beam = m * 64 + (r >> 1);
        const int comp = r & 1;                    // 0: re row, 1: im row
        const int a = kb & 63;
        const int hi = kb >> 6;                    // 0: multiplies Xre, 1: multiplies Xim
        const int8_t *wp = w + (((size_t)f * n_ant + a) * n_beams + beam) * 2;
        const int wre = wp[0], wim = wp[1];
        if (wre == -128 || wim == -128) *bad = 1;
        int val;
        if (comp == 0) val = hi ? -wim : wre;
        else           val = hi ? wre : wim;
        const size_t off = fm * kATileBytes + (size_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((kb >> 4) ^ (r & 7))) << 4) + (kb & 15);
        wprep[off] = (uint8_t)(int8_t)val;
    }
}

cudaError_t launch_prep_weights(const int8_t *d_w, uint8_t *d_wprep, int n_chan, int n_ant, int n_beams, int *d_bad, cudaStream_t stream)
{
    if (n_ant != 64) return cudaErrorInvalidValue;
    prep_weights_kernel<<<1184, 256, 0, stream>>>(d_w, d_wprep, n_chan, n_ant, n_beams, d_bad);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// DM-0 "dedispersion" of output time 0 of every chunk (the DEBUG cublasSgemv,
// /root/reference/src/beamformer.cu:498-504): ded[chunk][b] = sum_f out[chunk*n_out_per_chunk][f][b].
__global__ void dm0_kernel(const float *__restrict__ out, float *__restrict__ ded, int n_chunks, int n_out_per_chunk, int n_chan, int n_beams)
{
    const int chunk = blockIdx.x;
    const float *base = out + (size_t)chunk * n_out_per_chunk * n_chan * n_beams;
    for (int b = threadIdx.x; b < n_beams; b += blockDim.x) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int f = 0;
        for (; f + 3 < n_chan; f += 4) {
            s0 += base[(size_t)f * n_beams + b];
            s1 += base[(size_t)(f + 1) * n_beams + b];
            s2 += base[(size_t)(f + 2) * n_beams + b];
            s3 += base[(size_t)(f + 3) * n_beams + b];
        }
        for (; f < n_chan; f++) s0 += base[(size_t)f * n_beams + b];
        ded[(size_t)chunk * n_beams + b] = (s0 + s1) + (s2 + s3);
    }
}

cudaError_t launch_dm0(const float *d_out, float *d_ded, int n_chunks, int n_out_per_chunk, int n_chan, int n_beams, cudaStream_t stream)
{
    if (n_chunks <= 0) return cudaSuccess;
    dm0_kernel<<<n_chunks, 256, 0, stream>>>(d_out, d_ded, n_chunks, n_out_per_chunk, n_chan, n_beams);
    return cudaGetLastError();
}

}  // namespace dsabf

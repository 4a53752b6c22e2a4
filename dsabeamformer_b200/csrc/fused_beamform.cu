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
// 64 beams and a 256-beam channel is 4 resident 16 KB tiles.  X is NOT shifted
// down when unpacked: the nibbles are left in the high half of each byte
// (x' = 16 x), so unpacking is 2 LOP3 + 1 SHF per 4 samples; V' = 16 V stays
// < 2^22, and the factor 256 in the power is folded into the output scale.
//
// Data flow per CTA (persistent, one CTA per SM, contiguous range of work items;
// item = (channel, 256-column tile), 4 M-tiles per item):
//   warp 0      TMA producer: cp.async.bulk packed voltages -> staging ring (4 x 16 KB)
//   warps 1, 3  MMA issuers (one thread each, one per TMEM accumulator stage): 4 x tcgen05.mma
//               M128 N256 K32 per M-tile, accumulators double-buffered (2 x 256 TMEM columns)
//   warp 2      weights loader (cp.async.bulk of the pre-swizzled A' tiles) + TMEM alloc
//   warps 4-11  epilogue: tcgen05.ld -> square/accumulate over G columns -> one
//               shuffle to add re^2+im^2 rows -> coalesced float stores
//   warps 12-15 unpack: staging -> sign-preserving nibble split -> SW128 K-major B' tile
// Voltages are read from HBM once, weights once per (CTA, channel); nothing but
// the final float power is written.
#include "fused_epilogue.cuh"

namespace dsabf {

using namespace ptx;
using namespace fused;

namespace {

// In-kernel clock64 trace (dsabf_debug_trace, tools/trace_one.py): compiled in only with -DDSABF_TRACE; the
// predicated stamps cost ~15 of the epilogue's 165 instructions per accumulator stage even when switched off.
#ifdef DSABF_TRACE
constexpr bool kTrace = true;
constexpr int kTraceMode = (DSABF_TRACE + 0) == 2 ? 2 : 1;     // 2: eight stamps per accumulator stage of ONE epilogue warp (tools/trace_fine.py)
#else
constexpr bool kTrace = false;
constexpr int kTraceMode = 0;
#endif
#ifndef DSABF_TRACE_EW
#define DSABF_TRACE_EW 0
#endif

// Epilogue variants that were built, measured and dropped (prefetching the next tile's TMEM loads, ping-pong between the
// warps of a sub-partition, early barrier probes, producer roles on the highest warp ids, XU / int64 / mixed conversions,
// stage-owning warps, half-width pipelines ...) are kept as patches under tools/experiments/ with their numbers in
// profiles/r02_epilogue_experiments.txt; this file holds the product path only.

struct __align__(1024) Smem {
    uint8_t a[kMaxMT][kATileBytes];      // 64 KB  resident weights of the current channel
    uint8_t b[2][kBTileBytes];           // 64 KB  unpacked voltages, double buffered
    uint8_t stg[kStages][kStageBytes];   // 64 KB  packed voltages ring
    uint64_t stg_full[kStages], stg_empty[kStages];
    uint64_t b_full[2], b_empty[2];
    uint64_t a_full[kMaxMT], a_empty[kMaxMT];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    // all-0xFF operand of the bias MMA (A: 128 rows, B: kBiasCols rows of one 128-byte swizzle row each; the content
    // is constant, so the swizzle pattern does not matter and A and B share the bytes)
    alignas(1024) uint8_t bias_ff[kBiasCols > 0 ? (kBiasCols > 128 ? kBiasCols : 128) * kRowBytes : 1024];
};

// MODE: 0 = fast fp32 power, 1 = exact int64 power (+ optional integer-power dump),
//       2 = exact + int32 voltage dump (debug / parity tap)
template <int G, int MODE>
__global__ void __launch_bounds__(kThreads, 1) fused_beamform_kernel(const FusedParams p)
{
    constexpr bool EXACT = MODE != 0;
    constexpr bool DUMP = MODE == 2;
    extern __shared__ uint8_t smem_raw[];
    // align in the shared address space (keeps every later address a cheap 32-bit add)
    uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw + (sbase - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;             // role: 0 TMA, 1/3 MMA issuers, 2 weights, 4-7 unpack, 8-23 epilogue
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
        for (int i = 0; i < kMaxMT; i++) { mbar_init(&sm.a_full[i], 1); mbar_init(&sm.a_empty[i], 2); }   // both issuers release a weight tile
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(&sm.tmem_base, kTmemCols); tmem_relinquish(); }
    if (kBiasCols > 0 && !EXACT) {
        for (uint32_t i = threadIdx.x; i < sizeof(sm.bias_ff) / 16; i += kThreads) reinterpret_cast<uint4 *>(sm.bias_ff)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        fence_proxy_async_smem();
    }
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
                mbar_wait<1000>(bar_stg_empty + 8 * s, ((it / kStages) & 1) ^ 1, 1);
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
                    mbar_wait<2000>(bar_a_empty + 8 * m, (ch & 1) ^ 1, 2);
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
            int cur_f = -1, ch = -1;
            uint32_t acc_cnt = 0, use = 0;          // use: how often this issuer's stage has been filled
            uint32_t seen = 0, released = 0;        // weight tiles of the current channel this issuer has seen arrive / has released
            for (int item = item_begin, it = 0; item < item_end; item++, it++) {
                const int f = item / g.n_tiles;
                const bool new_ch = (f != cur_f);
                if (new_ch) { cur_f = f; ch++; seen = 0; released = 0; }
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
                    if (!((seen >> m) & 1u)) { mbar_wait(bar_a_full + 8 * m, ch & 1, 4); seen |= 1u << m; }
                    const long long tr0 = (kTrace && p.trace) ? clock64() : 0;
                    mbar_wait(bar_acc_empty + 8 * me, (use & 1) ^ 1, 5);
                    tc_fence_after();
                    if (kTrace && p.trace && blockIdx.x == 0 && acc_cnt < 128) { p.trace[acc_cnt * 16 + 0] = tr0; p.trace[acc_cnt * 16 + 1] = clock64(); }
                    const uint64_t adesc = make_desc_k_sw128(a_smem + m * kATileBytes);
                    const uint32_t d = tmem_base + me * kAccCols;
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; k++)   // K = 32 bytes per instruction; +32 B = +2 in the descriptor
                        mma_i8_ss(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k > 0);
                    if (kBiasCols > 0 && !EXACT) {             // + 32 * 255 * 255 on columns [0, kBiasCols): see Acc<false>::add2b
                        const uint64_t ffdesc = make_desc_k_sw128(sbase + (uint32_t)offsetof(Smem, bias_ff));
                        mma_i8_ss(d, ffdesc, ffdesc, make_idesc_u8(128, kBiasCols > 0 ? kBiasCols : 64), 1);
                    }
                    tc_commit(bar_acc_full + 8 * me);
                    if (kTrace && p.trace && blockIdx.x == 0 && acc_cnt < 128) p.trace[acc_cnt * 16 + 2] = clock64();
                    if (last_of_ch) { tc_commit(bar_a_empty + 8 * m); released |= 1u << m; }
                    use++;
                }
                if (touched) tc_commit(bar_b_empty + 8 * buf);
                // A parity wait cannot tell "this phase is done" from "the previous one is not yet", so a waiter has
                // to observe every phase of a barrier, at most one phase late.  Both issuers therefore wait for and
                // release every weight tile of every channel -- also the tiles the other issuer multiplied -- and
                // the loader refills a tile only after both releases (a_empty counts 2).
                if (last_of_ch) {
                    for (int m = 0; m < mt; m++) {
                        if ((released >> m) & 1u) continue;
                        if (!((seen >> m) & 1u)) mbar_wait(bar_a_full + 8 * m, ch & 1, 4);
                        tc_commit(bar_a_empty + 8 * m);
                    }
                }
            }
        }
    } else if (warp < kEpiWarp0) {
        // ================= unpack: 4-bit+4-bit -> int8 (x16), into SW128 K-major B' =================
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
            const bool utr = kTraceMode == 1 && p.trace && blockIdx.x == 0 && uw == 0 && lane == 0 && it < 32;
            if (utr) p.trace[it * 64 + 12] = clock64();
            mbar_wait<200>(bar_stg_full + 8 * s, (it / kStages) & 1, 6);
            if (utr) p.trace[it * 64 + 13] = clock64();
            mbar_wait<500>(bar_b_empty + 8 * buf, ((it >> 1) & 1) ^ 1, 7);
            if (utr) p.trace[it * 64 + 14] = clock64();
            const uint32_t src = stg_smem + s * kStageBytes + ld_off;
            const uint32_t dst = b_smem + buf * kBTileBytes;
#pragma unroll
            for (int i = 0; i < (kTileN / 8) / kNumUnpackWarps; i++) {
                const int grp = uw + i * kNumUnpackWarps;      // 8-column group
                const uint4 v = lds128(src + grp * 512);
                uint4 re, im;
                re.x = v.x & 0xF0F0F0F0u; im.x = (v.x << 4) & 0xF0F0F0F0u;
                re.y = v.y & 0xF0F0F0F0u; im.y = (v.y << 4) & 0xF0F0F0F0u;
                re.z = v.z & 0xF0F0F0F0u; im.z = (v.z << 4) & 0xF0F0F0F0u;
                re.w = v.w & 0xF0F0F0F0u; im.w = (v.w << 4) & 0xF0F0F0F0u;
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
        const int wc = ew >> 2;                    // which 64 columns of the tile
        const int row = quarter * 32 + lane;       // A' row: beam = row/2, component = row&1
        uint32_t tbase_w = tmem_base + ((uint32_t)(quarter * 32) << 16) + wc * kEpiCols;
        asm volatile("" : "+r"(tbase_w));
        constexpr int kGroups = kEpiCols / G;
        EpiCtx<EXACT> ctx;
        ctx.lane_odd = lane & 1;
        const size_t tstride = (size_t)g.n_chan * p.n_beams;      // floats between consecutive output times
        ctx.pair_stride = 2 * tstride;
        const bool biased = kBiasCols > 0 && !EXACT && wc * kEpiCols < kBiasCols;     // warp-uniform: this slice converts on the FMA pipe
        ctx.scale = biased ? p.scale * kBiasUnscale : p.scale;
        uint32_t acc_cnt = 0;
        int f = item_begin / g.n_tiles, t = item_begin - f * g.n_tiles - 1;   // advanced incrementally: no division per item
#pragma unroll 1
        for (int item = item_begin; item < item_end; item++) {
            if (++t == g.n_tiles) { t = 0; f++; }
            const int colbase = t * kTileN + wc * kEpiCols;          // first global column of this warp
            const int cols_left = g.n_cols_total - colbase;
            ctx.vg_lane = (cols_left >= kEpiCols ? kGroups : (cols_left > 0 ? cols_left / G : 0)) - ctx.lane_odd;
            // even lanes (re rows) store the even group of each pair, odd lanes the odd group
            const size_t out_off = ((size_t)(colbase / G) + ((kGroups > 1) ? ctx.lane_odd : 0)) * tstride + (size_t)f * p.n_beams + (row >> 1);
            float *out_item = p.out + out_off;
            asm volatile("" : "+l"(out_item));                       // keep the pointer live; do not recompute per M-tile
            long long *pp_item = (EXACT && p.pdump) ? p.pdump + out_off : nullptr;
#pragma unroll 1
            for (int m = 0; m < mt; m++, acc_cnt++) {
                const uint32_t as = acc_cnt & 1;
                uint32_t va[32], vb[32];
                const bool trc = kTraceMode == 1 && p.trace && blockIdx.x == 0 && (ew == 0 || ew == 15) && lane == 0 && acc_cnt < 128;
                const bool ftr = kTraceMode == 2 && p.trace && blockIdx.x == 0 && ew == DSABF_TRACE_EW && lane == 0 && acc_cnt < 128;
#define FSTAMP(k) do { if (ftr) p.trace[acc_cnt * 16 + 4 + (k)] = clock64(); } while (0)
                const int trb = acc_cnt * 16 + (ew == 0 ? 4 : 8);
                if (trc) p.trace[trb + 0] = clock64();
                FSTAMP(0);
                // first probe without a suspend-time hint: one SYNCS instruction instead of three, and the stage is
                // practically always full already
                if (!mbar_try_wait_nohint(bar_acc_full + 8 * as, (acc_cnt >> 1) & 1)) mbar_wait(bar_acc_full + 8 * as, (acc_cnt >> 1) & 1, 8);
                tc_fence_after();
                FSTAMP(1);
                tmem_ld_2x32_wait(tbase_w + as * kAccCols, va, vb);
                if (trc) p.trace[trb + 1] = clock64();               // accumulators are in registers
                FSTAMP(2);
                // both halves are in registers: hand the TMEM stage back to the MMA right away
                tc_fence_before();
                __syncwarp();
                FSTAMP(3);
                if (elect_one()) mbar_arrive(bar_acc_empty + 8 * as);
                FSTAMP(4);
                // ptxas sinks the arrive below all of the math that follows (it has no consumers), which
                // keeps the TMEM stage -- and so the tensor pipe -- waiting.  A never-taken, runtime-opaque
                // branch ends the basic block here and pins it (measured: 0.825 -> 0.462 ms on the DEBUG shape).
                if (acc_cnt == 0xFFFFFFFFu) continue;

                if (DUMP) {
                    // debug tap: int32 voltages, [f][col][beam][2]; this thread's (beam, component) = row
                    int32_t *vd = p.vdump + ((size_t)f * g.n_cols_total * p.n_beams + m * 64) * 2 + row;
                    const size_t cs = (size_t)p.n_beams * 2;
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) {
                        if (colbase + jj < g.n_cols_total) vd[(size_t)(colbase + jj) * cs] = ((int)va[jj]) >> 4;
                        if (colbase + 32 + jj < g.n_cols_total) vd[(size_t)(colbase + 32 + jj) * cs] = ((int)vb[jj]) >> 4;
                    }
                }
                if (trc) p.trace[trb + 2] = clock64();               // arithmetic starts
                FSTAMP(5);
                ctx.acc.zero();
                ctx.pend = 0;
                ctx.op = out_item;
                ctx.pp = pp_item;
                if (biased) epi_block<G, EXACT, 0, true>(ctx, va); else epi_block<G, EXACT, 0>(ctx, va);
                FSTAMP(6);
                if (biased) epi_block<G, EXACT, 1, true>(ctx, vb); else epi_block<G, EXACT, 1>(ctx, vb);
                if (trc) p.trace[trb + 3] = clock64();               // arithmetic issued
                FSTAMP(7);
                out_item += 64;                                      // next M-tile = next 64 beams
                if (EXACT) { if (pp_item) pp_item += 64; }
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
    // The attribute belongs to the (function, device) pair and a process may hold handles on several GPUs from
    // several threads: set it on every launch for the current device (a cheap driver call) instead of caching it.
    cudaError_t e = cudaFuncSetAttribute(fused_beamform_kernel<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
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
    if (p.n_ant != 64 || p.n_beams % 64 || p.n_beams < 64 || p.n_beams > 64 * kMaxMT) return cudaErrorInvalidValue;
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
    const int mt = n_beams >> 6, kbn = n_ant >> 6;
    const size_t total = (size_t)n_chan * mt * kbn * 128 * 128;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int kbyte = idx & 127;               // logical K byte inside the k-block
        const int r = (idx >> 7) & 127;            // row in the M tile
        const size_t tile = idx >> 14;             // ((f * mt) + m) * kbn + kb
        const int kb = tile % kbn;
        const size_t fm = tile / kbn;
        const int m = fm % mt;
        const int f = fm / mt;
        const int beam = m * 64 + (r >> 1);
        const int comp = r & 1;                    // 0: re row, 1: im row
        const int a = kb * 64 + (kbyte & 63);      // antenna
        const int hi = kbyte >> 6;                 // 0: multiplies Xre, 1: multiplies Xim
        const int8_t *wp = w + (((size_t)f * n_ant + a) * n_beams + beam) * 2;
        const int wre = wp[0], wim = wp[1];
        if (wre == -128 || wim == -128) *bad = 1;
        int val;
        if (comp == 0) val = hi ? -wim : wre;
        else           val = hi ? wre : wim;
        const size_t off = tile * kATileBytes + (size_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((kbyte >> 4) ^ (r & 7))) << 4) + (kbyte & 15);
        wprep[off] = (uint8_t)(int8_t)val;
    }
}

cudaError_t launch_prep_weights(const int8_t *d_w, uint8_t *d_wprep, int n_chan, int n_ant, int n_beams, int *d_bad, cudaStream_t stream)
{
    if (n_ant % 64 || n_ant < 64 || n_ant > 256 || n_beams % 64) return cudaErrorInvalidValue;
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

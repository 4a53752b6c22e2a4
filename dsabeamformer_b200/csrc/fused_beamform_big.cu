// fused_beamform_big.cu -- the fused unpack -> beamform -> detect kernel for arrays that do not fit
// the resident-weights kernel of fused_beamform.cu: up to 256 antennas (K' = up to 4 swizzle rows of
// 128 bytes, "k-blocks" of 64 antennas) and up to 4096 beams (BASELINE.json configs[3]: 256 antennas x
// 2048 beams x 2048 channels sharded over the GPUs).  Same math, same epilogue, same output as
// fused_beamform.cu (see there); what changes is what stays in shared memory:
//
//   * the UNPACKED voltages of one work item (channel, 256 columns) stay resident: KB tiles of 32 KB,
//     written once per item by the unpack warps straight from global memory (each 16-byte load is a
//     quarter of one k-block of one column; 64-byte runs, all sectors fully used);
//   * the steering weights stream: the weights loader walks (M-tile, k-block) in order through a ring
//     of five 16 KB slots (cp.async.bulk from L2; the whole sub-band's weights are L2-resident between
//     items), ~32 B/clk/SM at the tensor rate of this shape;
//   * per M-tile the two MMA issuers accumulate KB x 4 tcgen05.mma (M128 N256 K32) into their TMEM stage.
// Each accumulator now receives 2*n_ant MACs, so for 256 antennas the epilogue load per MMA is 4x lower
// than in the 64-antenna kernel and the kernel is tensor-bound.
#include "fused_epilogue.cuh"

namespace dsabf {

using namespace ptx;
using namespace fused;

namespace {

constexpr int kMaxKB = 4;      // k-blocks of 64 antennas
// Weight tiles in flight: one ring, filled in stream order by the loader and drained by the two MMA
// issuers, which alternate M-tile by M-tile.  TMA loads complete out of order and an mbarrier parity wait
// cannot tell "round R is done" from "round R-1 is not", so a waiter must never ask for round R of a slot
// before round R-1 has completed.  With two consumers that needs the baton below (turn[]): an issuer starts
// waiting for the tiles of M-tile a only after the other issuer has seen every tile of M-tile a-1 arrive.
// (Without it issuer B could run ahead on a stale slot, release it, and the loader would re-arm a barrier
// whose transaction is still pending -- an illegal-instruction trap under load.)
constexpr int kARing = 6;
// Register budgets (setmaxnreg).  The CTA owns 768 x 80 registers; after redistribution the sum must not
// exceed that, or the last setmaxnreg.inc waits forever:  4 light warps x 32 + 4 unpack warps x 64 +
// 16 epilogue warps x 96 = 480 x 128 = 768 x 80 exactly.
constexpr int kRegsLightBig = 32;
constexpr int kRegsUnpackBig = 64;
static_assert(4 * kRegsLightBig + 4 * kRegsUnpackBig + 16 * kRegsEpi <= 24 * 80, "setmaxnreg budget exceeds the CTA's allocation");

struct __align__(1024) SmemBig {
    uint8_t a[kARing][kATileBytes];      //  96 KB  streamed weight tiles
    uint8_t b[kMaxKB][kBTileBytes];      // 128 KB  unpacked voltages of the current item, one tile per k-block
    uint64_t a_full[kARing], a_empty[kARing];
    uint64_t b_full[2], b_empty[2];      // B' is double-buffered when it fits (<= 2 k-blocks), else one buffer
    uint64_t acc_full[2], acc_empty[2];
    uint64_t turn[2];                    // baton between the two MMA issuers
    uint32_t tmem_base;
};
static_assert(sizeof(SmemBig) + 1024 <= 232448, "SmemBig exceeds the 227 KB a CTA can opt in to");

__device__ __forceinline__ uint4 ldg128_stream(const void *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int G, int MODE>
__global__ void __launch_bounds__(kThreads, 1) fused_beamform_big_kernel(const FusedParams p)
{
    constexpr bool EXACT = MODE != 0;
    constexpr bool DUMP = MODE == 2;
    extern __shared__ uint8_t smem_raw[];
    uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    SmemBig &sm = *reinterpret_cast<SmemBig *>(smem_raw + (sbase - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const Geom g(p);
    const int mt = p.n_beams >> 6;
    const int kb_n = p.n_ant >> 6;

    const long long items_total = (long long)g.n_chan * g.n_tiles;
    // Work items (channel f, column tile t) are dealt round-robin: at any time the resident CTAs work on ~gridDim.x
    // consecutive items = gridDim.x / n_tiles channels, whose weights (n_beams * n_ant * 4 bytes each) then stay in L2
    // while every column tile of the channel re-streams them.  (Contiguous ranges per CTA had 148 channels in flight:
    // 296 MB of weights at the scaled-array shape, every re-read a DRAM read -- 4.4 GB per launch instead of 0.65.)
    const int item_begin = blockIdx.x, item_end = (int)items_total, item_step = gridDim.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kARing; i++) { mbar_init(&sm.a_full[i], 1); mbar_init(&sm.a_empty[i], 1); }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.b_full[i], kNumUnpackWarps);
            mbar_init(&sm.b_empty[i], mt == 1 ? 1 : 2);          // one commit per MMA issuer that multiplied the item
        }
        for (int i = 0; i < 2; i++) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], kNumEpiWarps); mbar_init(&sm.turn[i], 1); }
        fence_mbar_init();
    }
    if (warp == 2) { tmem_alloc(&sm.tmem_base, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = sm.tmem_base;
    asm volatile("" : "+r"(sbase));
    const uint32_t a_smem = sbase + (uint32_t)offsetof(SmemBig, a), b_smem = sbase + (uint32_t)offsetof(SmemBig, b);
    const uint32_t bar_a_full = sbase + (uint32_t)offsetof(SmemBig, a_full), bar_a_empty = sbase + (uint32_t)offsetof(SmemBig, a_empty);
    const uint32_t bar_b_full = sbase + (uint32_t)offsetof(SmemBig, b_full), bar_b_empty = sbase + (uint32_t)offsetof(SmemBig, b_empty);
    // B' buffers: item `it` uses buffer it % nbuf, fill number it / nbuf of that buffer
    const int nbuf_log2 = (kb_n <= 2) ? 1 : 0;
    const uint32_t b_buf_bytes = (uint32_t)kb_n * kBTileBytes;
    const uint32_t bar_acc_full = sbase + (uint32_t)offsetof(SmemBig, acc_full), bar_acc_empty = sbase + (uint32_t)offsetof(SmemBig, acc_empty);
    const uint32_t bar_turn = sbase + (uint32_t)offsetof(SmemBig, turn);

    if (warp < kUnpackWarp0) {
        setmaxnreg_dec<kRegsLightBig>();
        if (warp == 0 && lane == 0) {
            // ================= weights loader: stream (item, M-tile, k-block) into the issuers' rings =================
            const uint64_t pol = policy_evict_last();
            uint32_t q = 0;                                      // position in the weight stream
            for (int item = item_begin; item < item_end; item += item_step) {
                const int f = item / g.n_tiles;
                const uint8_t *wf = p.wprep + (size_t)f * mt * kb_n * kATileBytes;
                for (int mk = 0; mk < mt * kb_n; mk++, q++, wf += kATileBytes) {
                    const uint32_t slot = q % kARing;
                    mbar_wait<100>(bar_a_empty + 8 * slot, ((q / kARing) & 1) ^ 1, 2);
                    mbar_arrive_expect_tx(bar_a_full + 8 * slot, kATileBytes);
                    tma_bulk_g2s_hint(a_smem + slot * kATileBytes, wf, kATileBytes, bar_a_full + 8 * slot, pol);
                }
            }
        } else if ((warp == 1 || warp == 3) && lane == 0) {
            // ================= MMA issuers (one per TMEM accumulator stage) =================
            const int me = (warp == 3) ? 1 : 0;
            constexpr uint32_t idesc = make_idesc_i8(128, kTileN);
            uint32_t acc_cnt = 0, use = 0;                       // use: M-tiles this issuer has multiplied
            for (int item = item_begin, it = 0; item < item_end; item += item_step, it++) {
                // every issuer observes every phase of b_full, also for an item it has no M-tile of (n_beams = 64)
                const uint32_t bbuf = it & ((1 << nbuf_log2) - 1);
                mbar_wait(bar_b_full + 8 * bbuf, (it >> nbuf_log2) & 1, 3);
                bool touched = false;
                for (int m = 0; m < mt; m++, acc_cnt++) {
                    if ((int)(acc_cnt & 1) != me) continue;
                    touched = true;
                    mbar_wait(bar_acc_empty + 8 * me, (use & 1) ^ 1, 5);
                    // baton: the other issuer has seen all tiles of M-tile acc_cnt-1 land
                    if (acc_cnt > 0) mbar_wait(bar_turn + 8 * me, (me ? use : use - 1) & 1, 6);
                    tc_fence_after();
                    const uint32_t d = tmem_base + me * kAccCols;
                    for (int kb = 0; kb < kb_n; kb++) {
                        const uint32_t q = acc_cnt * kb_n + kb;
                        const uint32_t slot = q % kARing;
                        mbar_wait(bar_a_full + 8 * slot, (q / kARing) & 1, 4);
                        if (kb == kb_n - 1) mbar_arrive(bar_turn + 8 * (me ^ 1));      // pass the baton
                        tc_fence_after();
                        const uint64_t adesc = make_desc_k_sw128(a_smem + slot * kATileBytes);
                        const uint64_t bdesc = make_desc_k_sw128(b_smem + bbuf * b_buf_bytes + kb * kBTileBytes);
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; k++)
                            mma_i8_ss(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1 : 0);
                        tc_commit(bar_a_empty + 8 * slot);
                    }
                    tc_commit(bar_acc_full + 8 * me);
                    use++;
                }
                if (touched) tc_commit(bar_b_empty + 8 * bbuf);
            }
        }
    } else if (warp < kEpiWarp0) {
        // ================= unpack: global -> int8 (x16) -> SW128 K-major B' tiles, one per k-block =================
        setmaxnreg_dec<kRegsUnpackBig>();
        const int uw = warp - kUnpackWarp0;
        const int q4 = lane >> 3, j = lane & 7, c = j & 3;
        const int r = (j < 4) ? q4 : (q4 ^ 5);                       // same conflict-free (row, chunk) mapping as fused_beamform.cu
        const uint32_t st_re = r * 128 + ((c ^ r) << 4);
        const uint32_t st_im = r * 128 + (((c + 4) ^ r) << 4);
        for (int item = item_begin, it = 0; item < item_end; item += item_step, it++) {
            const int f = item / g.n_tiles, t = item - f * g.n_tiles;
            const uint32_t bbuf = it & ((1 << nbuf_log2) - 1);
            mbar_wait<200>(bar_b_empty + 8 * bbuf, ((it >> nbuf_log2) & 1) ^ 1, 7);
#pragma unroll 1
            for (int i = 0; i < (kTileN / 8) / kNumUnpackWarps; i += 2) {
                uint4 v[2][kMaxKB];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int grp = uw + (i + u) * kNumUnpackWarps;     // 8-column group of the tile
                    const int cg = t * kTileN + grp * 8 + r;            // global column of this lane's row
                    const int chunk = cg / g.cpc, c_in = cg - chunk * g.cpc;
                    const uint8_t *src = p.packed + ((((size_t)chunk * g.n_chan + f) * g.cpc + c_in) * p.n_ant) + c * 16;
                    const bool ok = cg < g.n_cols_total;
#pragma unroll
                    for (int kb = 0; kb < kMaxKB; kb++)
                        if (kb < kb_n) v[u][kb] = ok ? ldg128_stream(src + kb * 64) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int grp = uw + (i + u) * kNumUnpackWarps;
#pragma unroll
                    for (int kb = 0; kb < kMaxKB; kb++) {
                        if (kb < kb_n) {
                            const uint4 w = v[u][kb];
                            uint4 re, im;
                            re.x = w.x & 0xF0F0F0F0u; im.x = (w.x << 4) & 0xF0F0F0F0u;
                            re.y = w.y & 0xF0F0F0F0u; im.y = (w.y << 4) & 0xF0F0F0F0u;
                            re.z = w.z & 0xF0F0F0F0u; im.z = (w.z << 4) & 0xF0F0F0F0u;
                            re.w = w.w & 0xF0F0F0F0u; im.w = (w.w << 4) & 0xF0F0F0F0u;
                            const uint32_t dst = b_smem + bbuf * b_buf_bytes + kb * kBTileBytes + grp * 1024;
                            sts128(dst + st_re, re);
                            sts128(dst + st_im, im);
                        }
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_b_full + 8 * bbuf);
            // The next item's columns are loaded only after the MMAs of an item in between have drained their B'
            // buffer; pull them into L2 now so that those loads -- the bubble of a single-buffered B' -- do not
            // pay DRAM latency.
            const int item2 = item + item_step;
            if (item2 < item_end) {
                const int f2 = item2 / g.n_tiles, t2 = item2 - f2 * g.n_tiles;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int cg = t2 * kTileN + h * (kTileN / 2) + uw * 32 + lane;
                    if (cg < g.n_cols_total) {
                        const int chunk = cg / g.cpc, c_in = cg - chunk * g.cpc;
                        const uint8_t *row = p.packed + ((((size_t)chunk * g.n_chan + f2) * g.cpc + c_in) * p.n_ant);
                        for (int o = 0; o < p.n_ant; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
                    }
                }
            }
        }
    } else {
        // ================= epilogue (identical to fused_beamform.cu) =================
        setmaxnreg_inc<kRegsEpi>();
        const int ew = warp - kEpiWarp0;
        const int quarter = warp & 3;
        const int wc = ew >> 2;
        const int row = quarter * 32 + lane;
        uint32_t tbase_w = tmem_base + ((uint32_t)(quarter * 32) << 16) + wc * kEpiCols;
        asm volatile("" : "+r"(tbase_w));
        constexpr int kGroups = kEpiCols / G;
        EpiCtx<EXACT> ctx;
        ctx.lane_odd = lane & 1;
        const size_t tstride = (size_t)g.n_chan * p.n_beams;
        ctx.pair_stride = 2 * tstride;
        ctx.scale = p.scale;
        uint32_t acc_cnt = 0;
#pragma unroll 1
        for (int item = item_begin; item < item_end; item += item_step) {
            const int f = item / g.n_tiles, t = item - f * g.n_tiles;
            const int colbase = t * kTileN + wc * kEpiCols;
            const int cols_left = g.n_cols_total - colbase;
            ctx.vg_lane = (cols_left >= kEpiCols ? kGroups : (cols_left > 0 ? cols_left / G : 0)) - ctx.lane_odd;
            const size_t out_off = ((size_t)(colbase / G) + ((kGroups > 1) ? ctx.lane_odd : 0)) * tstride + (size_t)f * p.n_beams + (row >> 1);
            float *out_item = p.out + out_off;
            asm volatile("" : "+l"(out_item));
            long long *pp_item = (EXACT && p.pdump) ? p.pdump + out_off : nullptr;
#pragma unroll 1
            for (int m = 0; m < mt; m++, acc_cnt++) {
                const uint32_t as = acc_cnt & 1;
                mbar_wait(bar_acc_full + 8 * as, (acc_cnt >> 1) & 1, 8);
                tc_fence_after();
                const uint32_t taddr = tbase_w + as * kAccCols;
                uint32_t va[32], vb[32];
                tmem_ld_2x32_wait(taddr, va, vb);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
                if (acc_cnt == 0xFFFFFFFFu) continue;            // pins the arrive above the arithmetic (see fused_beamform.cu)

                if (DUMP) {
                    int32_t *vd = p.vdump + ((size_t)f * g.n_cols_total * p.n_beams + m * 64) * 2 + row;
                    const size_t cs = (size_t)p.n_beams * 2;
#pragma unroll
                    for (int jj = 0; jj < 32; jj++) {
                        if (colbase + jj < g.n_cols_total) vd[(size_t)(colbase + jj) * cs] = ((int)va[jj]) >> 4;
                        if (colbase + 32 + jj < g.n_cols_total) vd[(size_t)(colbase + 32 + jj) * cs] = ((int)vb[jj]) >> 4;
                    }
                }
                ctx.acc.zero();
                ctx.pend = 0;
                ctx.op = out_item;
                ctx.pp = pp_item;
                epi_block<G, EXACT, 0>(ctx, va);
                epi_block<G, EXACT, 1>(ctx, vb);
                out_item += 64;
                if (EXACT) { if (pp_item) pp_item += 64; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

template <int G, int MODE>
cudaError_t launch_one(const FusedParams &p, int grid, cudaStream_t stream)
{
    const size_t smem = sizeof(SmemBig) + 1024;
    // The attribute belongs to the (function, device) pair and a process may hold handles on several GPUs from
    // several threads: set it on every launch for the current device (a cheap driver call) instead of caching it.
    cudaError_t e = cudaFuncSetAttribute(fused_beamform_big_kernel<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    fused_beamform_big_kernel<G, MODE><<<grid, kThreads, smem, stream>>>(p);
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

cudaError_t launch_fused_beamform_big(const FusedParams &p, bool exact, int num_sms, cudaStream_t stream)
{
    if (p.n_beams % 64 || p.n_beams < 64 || p.n_ant % 64 || p.n_ant < 64 || p.n_ant > 64 * kMaxKB) return cudaErrorInvalidValue;
    if (p.cols_per_chunk % p.group || p.cols_per_chunk % 8) return cudaErrorInvalidValue;
    const long long items = (long long)p.n_chan * p.n_tiles;
    if (items <= 0) return cudaSuccess;
    const int grid = (int)(items < num_sms ? items : num_sms);
    if (p.vdump) return launch_g<2>(p, grid, stream);
    return exact ? launch_g<1>(p, grid, stream) : launch_g<0>(p, grid, stream);
}

}  // namespace dsabf

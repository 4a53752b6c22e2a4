// exchange.cu -- the frequency -> beam transpose in front of full-band dedispersion, as ONE kernel over
// NVLink peer memory.
//
// Every GPU g holds the detected power of its own sub-band, [n_t][c_local][n_beams] (the fused kernel's
// output layout).  Full-band dedispersion needs, on GPU r, all world*c_local channels of r's n_beams/world
// beams: full[t][g*c_local + c][b] = power_g[t][c][r*Bs + b]  (the one exchange step of the
// path; the reference itself never exchanges anything -- it sums only its own 256 channels,
// /root/reference/src/beamformer.cu:498-504).
//
// Instead of permute-copy + NCCL all_to_all_single + permute-copy (three passes over the data and a
// staging buffer), the receiving GPU PULLS: its threads load 16-byte pieces straight from the peers'
// buffers (peer pointers of a symmetric allocation, mapped over NVLink / NVSwitch) and store them at their
// final place.  One pass, no staging; consecutive CTAs read from different peers, so all links are busy.
#include <cstdint>

#include "fused_beamform.cuh"

namespace dsabf {
namespace {

constexpr int kMaxPeers = 16;
struct PeerPtrs { const float *p[kMaxPeers]; };

__global__ void __launch_bounds__(256)
transpose_pull_kernel(PeerPtrs peers, int world, int rank, int n_t, int c_local, int n_beams, float *__restrict__ full)
{
    const int bs = n_beams / world;                // beams per GPU after the exchange
    const int q4 = bs >> 2;                        // 16-byte pieces per (time, channel) row segment
    const long long total = (long long)n_t * world * c_local * q4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(idx % q4);
        long long r = idx / q4;
        const int c = (int)(r % c_local);
        r /= c_local;
        const int g = (int)(r % world);
        const int t = (int)(r / world);
        const float4 *src = reinterpret_cast<const float4 *>(peers.p[g] + ((size_t)t * c_local + c) * n_beams + (size_t)rank * bs) + q;
        float4 *dst = reinterpret_cast<float4 *>(full + ((size_t)t * ((size_t)world * c_local) + (size_t)g * c_local + c) * bs) + q;
        float4 v;
        // plain (coherent) load: the peer wrote this buffer in an earlier kernel; no read-only path for remote memory
        asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src));
        *dst = v;
    }
}

}  // namespace

cudaError_t launch_transpose_pull(const float *const *peer_power, int world, int rank, int n_t, int c_local, int n_beams, float *d_full,
                                  int num_sms, cudaStream_t stream)
{
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_beams % world || (n_beams / world) % 4) return cudaErrorInvalidValue;
    if (n_t <= 0 || c_local <= 0) return cudaSuccess;
    PeerPtrs pp{};
    for (int g = 0; g < world; g++) {
        if (!peer_power[g] || ((uintptr_t)peer_power[g] & 15)) return cudaErrorInvalidValue;
        pp.p[g] = peer_power[g];
    }
    if ((uintptr_t)d_full & 15) return cudaErrorInvalidValue;
    const long long total = (long long)n_t * world * c_local * (n_beams / world / 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)num_sms * 16;             // grid-stride: enough CTAs to keep every link busy
    if (blocks > cap) blocks = cap;
    transpose_pull_kernel<<<(unsigned)blocks, 256, 0, stream>>>(pp, world, rank, n_t, c_local, n_beams, d_full);
    return cudaGetLastError();
}

}  // namespace dsabf

// a2a.cu -- the data movement of an all-to-all over NVLink peer memory (one process per GPU, symmetric buffers).
// The sequence-parallel transformer block (tblock.py "sp" mode) exchanges token-sharded <-> head-sharded activations four
// times per layer (reference analogue: none -- the reference has sample-level DDP only, src/trainer/stat.py:431-436;
// SURVEY.md section 8e / 8f4).  As ncclSend/ncclRecv groups those 40 exchanges of ~3 MB cost 24-29 us each at 8 ranks
// (profiles/r02d_trace_shard8m_8_rank*.txt): latency, not bandwidth.  Here ONE kernel stores block j of the send buffer
// straight into slot `rank` of peer j's receive buffer (16-byte stores over NVLink / NVSwitch, every peer in parallel);
// the cross-rank barrier that follows (signal pads of the same symmetric allocation) orders them before the consumer.
#include "common.cuh"
#include <algorithm>

namespace gaot {

struct PeerPtrs { uint4* p[16]; };

// bcast = 0: all-to-all (block j of `send` goes to rank j); bcast = 1: all-gather (the one block of `send` goes to every rank)
__global__ void __launch_bounds__(256)
a2a_put_kernel(const uint4* __restrict__ send, const PeerPtrs peers, int rank, int64_t block16, int bcast) {
    const int j = blockIdx.y;                                   // destination rank
    const uint4* s = send + (bcast ? (size_t)0 : (size_t)j * block16);
    uint4* d = peers.p[j] + (size_t)rank * block16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < block16; i += stride) d[i] = s[i];
}

// out[i] = sum over ranks r = 0..world-1 (fixed order: deterministic) of peer r's float buffer at element offset + i: the
// reduce half of a reduce-scatter / two-shot all-reduce, pulled over NVLink (16-byte loads, every peer's stream in flight)
__global__ void __launch_bounds__(256)
p2p_reduce_kernel(const PeerPtrs peers, int world, int64_t offset4, int64_t n4, float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (r < world) v[r] = *reinterpret_cast<const float4*>(reinterpret_cast<const float4*>(peers.p[r]) + offset4 + i);
        float4 acc = v[0];
#pragma unroll
        for (int r = 1; r < 16; ++r)
            if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
        out[i] = acc;
    }
}

}  // namespace gaot

using namespace gaot;

extern "C" {

static int fill_peers(PeerPtrs& pp, const void* const* peer, int world, const char* what) {
    for (int j = 0; j < 16; ++j) pp.p[j] = j < world ? (uint4*)peer[j] : nullptr;
    for (int j = 0; j < world; ++j)
        if (pp.p[j] == nullptr || ((uintptr_t)pp.p[j] % 16) != 0) { set_error("%s: bad peer pointer %d", what, j); return GAOT_ERR_INVALID; }
    return GAOT_OK;
}

int gaot_p2p_reduce(const void* const* peer_src, int32_t world, int64_t offset_floats, int64_t n_floats, float* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(world >= 1 && world <= 16 && peer_src != nullptr, "p2p_reduce: 1..16 ranks");
    GAOT_CHECK_ARG(offset_floats >= 0 && offset_floats % 4 == 0 && n_floats >= 0 && n_floats % 4 == 0 && ((uintptr_t)out % 16) == 0,
                   "p2p_reduce: offset / count must be multiples of 4 floats, output 16-byte aligned");
    if (n_floats == 0) return GAOT_OK;
    GAOT_CHECK_ARG(out != nullptr, "p2p_reduce: null output");
    PeerPtrs pp;
    int rc = fill_peers(pp, peer_src, world, "p2p_reduce");
    if (rc) return rc;
    const int64_t n4 = n_floats / 4;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n4 + 255) / 256, 4 * kNumSMs));
    p2p_reduce_kernel<<<grid, 256, 0, st>>>(pp, world, offset_floats / 4, n4, (float4*)out);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

int gaot_a2a_put(const void* send, const void* const* peer_recv, int32_t rank, int32_t world, int64_t block_bytes, void* stream) {
    return gaot_p2p_put(send, peer_recv, rank, world, block_bytes, 0, stream);
}

int gaot_p2p_put(const void* send, const void* const* peer_recv, int32_t rank, int32_t world, int64_t block_bytes, int32_t bcast,
                 void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GAOT_CHECK_ARG(world >= 1 && world <= 16 && rank >= 0 && rank < world, "a2a_put: 1..16 ranks");
    GAOT_CHECK_ARG(block_bytes >= 0 && block_bytes % 16 == 0 && ((uintptr_t)send % 16) == 0, "a2a_put: blocks must be 16-byte multiples, 16-byte aligned");
    GAOT_CHECK_ARG(send != nullptr && peer_recv != nullptr, "a2a_put: null pointer");
    if (block_bytes == 0) return GAOT_OK;
    PeerPtrs pp;
    int rc = fill_peers(pp, peer_recv, world, "p2p_put");
    if (rc) return rc;
    const int64_t block16 = block_bytes / 16;
    // enough CTAs per destination to keep every NVLink port busy, not more than the SMs can hold
    const int per_peer = (int)std::max<int64_t>(1, std::min<int64_t>((block16 + 1023) / 1024, (2 * kNumSMs + world - 1) / world));
    a2a_put_kernel<<<dim3((unsigned)per_peer, (unsigned)world), 256, 0, st>>>((const uint4*)send, pp, rank, block16, bcast);
    GAOT_LAUNCH_CHECK();
    return GAOT_OK;
}

}  // extern "C"

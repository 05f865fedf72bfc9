// Compaction of the rows of finished environments for the host boundary (VecMarineNavEnv.step_host).
//
// After an auto-reset step only the environments whose episode ended carry a NEW observation (the first one of their
// next episode, marinenav_env.py:186); everything else the host needs is already final when the step kernel ends.  The
// host boundary therefore ships the step's own observation block to the host while the masked reset + re-observe run,
// and afterwards only this compact list of re-observed rows (+ their environment indices), instead of waiting for the
// reset before any byte can move.
#include "mnv_common.cuh"

namespace {

constexpr int kGatherWarps = 8;

__global__ void __launch_bounds__(kGatherWarps * 32)
mnv_gather_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ rows, long long E, int row_len, int cap,
                  float* __restrict__ compact, int* __restrict__ index, int* __restrict__ count)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
    const long long e = warp * 32 + lane;
    const bool m = e < E && mask[e] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, m);
    if (bal == 0u) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(bal));          // order between warps is arbitrary: the index list travels along
    base = __shfl_sync(0xffffffffu, base, 0);
    const int slot = base + __popc(bal & ((1u << lane) - 1u));
    if (m && slot < cap) index[slot] = (int)e;
    // the warp copies its selected rows one after the other, 32 floats per trip
    unsigned left = bal;
    while (left != 0u) {
        const int src_lane = __ffs(left) - 1;
        left &= left - 1u;
        const int s = __shfl_sync(0xffffffffu, slot, src_lane);
        if (s >= cap) continue;
        const float* src = rows + (warp * 32 + src_lane) * row_len;
        float* dst = compact + (long long)s * row_len;
        for (int i = lane; i < row_len; i += 32) dst[i] = src[i];
    }
}

}  // namespace

extern "C" int mnv_gather_rows(const uint8_t* d_mask, const float* d_rows, int64_t E, int32_t row_len, int32_t cap,
                               float* d_compact, int32_t* d_index, int32_t* d_count, void* stream)
{
    if (d_mask == nullptr || d_count == nullptr) { mnv_set_error("mnv_gather_rows: null mask / count"); return MNV_E_NULL; }
    MNV_CHECK_PTR(d_rows); MNV_CHECK_PTR(d_compact); MNV_CHECK_PTR(d_index);
    if (E <= 0 || row_len <= 0 || cap < 0) { mnv_set_error("mnv_gather_rows: bad sizes (E=%lld row_len=%d cap=%d)", (long long)E, row_len, cap); return MNV_E_SIZE; }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t err = cudaMemsetAsync(d_count, 0, sizeof(int32_t), st);
    if (err != cudaSuccess) { mnv_set_error("mnv_gather_rows: cudaMemsetAsync: %s", cudaGetErrorString(err)); return (int)err; }
    const long long warps = (E + 31) / 32;
    const unsigned grid = (unsigned)((warps + kGatherWarps - 1) / kGatherWarps);
    mnv_gather_kernel<<<grid, kGatherWarps * 32, 0, st>>>(d_mask, d_rows, E, row_len, cap, d_compact, d_index, d_count);
    return mnv_launch_status("mnv_gather_rows");
}

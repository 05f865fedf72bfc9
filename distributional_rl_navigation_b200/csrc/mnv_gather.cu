// Host-boundary helper of the vectorised env (VecMarineNavEnv.step_host).
//
// After an auto-reset step only the environments whose episode ended carry a NEW observation (the first one of their
// next episode, marinenav_env.py:186); everything else the host needs is already final when the step kernel ends.  The
// host boundary therefore ships the step's own observation block to the host with one bulk copy WHILE the masked reset +
// re-observe run, and afterwards this kernel overwrites just the rows of the re-observed environments -- straight into the
// pinned host array, which is mapped into the device address space (zero-copy stores: a row is 104 contiguous bytes, a few
// hundred rows per step).  No compaction, no index list, no host-side patching.
#include "mnv_common.cuh"

namespace {

constexpr int kScatterWarps = 8;

__global__ void __launch_bounds__(kScatterWarps * 32)
mnv_scatter_rows_kernel(const uint8_t* __restrict__ mask, const float* __restrict__ rows, float* __restrict__ host_rows,
                        long long E, int row_len)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * kScatterWarps + (threadIdx.x >> 5);
    const long long e = warp * 32 + lane;
    unsigned left = __ballot_sync(0xffffffffu, e < E && mask[e] != 0);
    while (left != 0u) {                                           // the warp copies its selected rows one after the other
        const int src_lane = __ffs(left) - 1;
        left &= left - 1u;
        const long long off = (warp * 32 + src_lane) * row_len;
        for (int i = lane; i < row_len; i += 32) host_rows[off + i] = rows[off + i];
    }
}

}  // namespace

extern "C" int mnv_scatter_rows_host(const uint8_t* d_mask, const float* d_rows, float* h_rows_mapped, int64_t E, int32_t row_len,
                                     void* stream)
{
    if (d_mask == nullptr) { mnv_set_error("mnv_scatter_rows_host: null mask"); return MNV_E_NULL; }
    MNV_CHECK_PTR(d_rows); MNV_CHECK_PTR(h_rows_mapped);
    if (E <= 0 || row_len <= 0) { mnv_set_error("mnv_scatter_rows_host: bad sizes (E=%lld row_len=%d)", (long long)E, row_len); return MNV_E_SIZE; }
    const long long warps = (E + 31) / 32;
    const unsigned grid = (unsigned)((warps + kScatterWarps - 1) / kScatterWarps);
    mnv_scatter_rows_kernel<<<grid, kScatterWarps * 32, 0, (cudaStream_t)stream>>>(d_mask, d_rows, h_rows_mapped, E, row_len);
    return mnv_launch_status("mnv_scatter_rows_host");
}

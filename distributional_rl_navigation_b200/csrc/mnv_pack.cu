// Host-boundary helper of the vectorised env (VecMarineNavEnv.step_host): compact transport of the observation block.
//
// An observation row (marinenav_env.py:273-326) is 4 head values (velocity, goal in the robot frame) + 2 floats per sonar
// beam, and a beam without a return is exactly (0, 0) (:318-320): in a rollout ~4 % of the beam slots carry a return.  The
// device -> host link (~52 GB/s, shared by the GPUs of a node) is the floor of the host API, so instead of the dense 6.8 MB
// block (65 536 x 26 floats) this kernel writes
//     head  f32 [E][4]                     the head of every row
//     mask  u32 [E][W]                     W = ceil(n_beams / 32): bit b of environment e <=> beam b has a return
//     dir   u32 [ceil(E / 32)]             where the returns of environments 32 g .. 32 g + 31 start in `vals`
//     count u32 [4]                        count[0] = number of returns (may exceed `capacity`: the tail is then not written
//                                          and the host falls back to the dense block for that step); count[3] = launch
//                                          sequence number (+1 per launch: a host that polls a copy of it in pinned memory
//                                          knows that the packet of THIS launch has landed)
//     vals  f32 [capacity][2]              (x, y) of the returns of a group, environment by environment, beam by beam
// into ONE contiguous buffer that ships with one copy (~1.9 MB); a native multi-threaded helper (csrc_host/mnv_host.c)
// expands it into the dense [E][obs_dim] array on the host with one sequential sweep per thread.  One warp = one group of
// 32 rows, staged through shared memory with coalesced 16-byte loads; a group's slots are contiguous (one atomicAdd per warp).
#include "mnv_common.cuh"

namespace {

constexpr int kPackWarps = 4;

__global__ void __launch_bounds__(kPackWarps * 32)
mnv_pack_obs_kernel(const float* __restrict__ obs, long long E, int D, float4* __restrict__ head, unsigned* __restrict__ mask,
                    unsigned* __restrict__ dir, unsigned* __restrict__ count, float2* __restrict__ vals, unsigned capacity)
{
    extern __shared__ __align__(16) float s_rows[];                // [kPackWarps][32 * D]
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long g = (long long)blockIdx.x * kPackWarps + w, e0 = g * 32;
    if (blockIdx.x == 0 && threadIdx.x == 0) count[3] += 1u;      // launch sequence number (not cleared by the memset)
    if (e0 >= E) return;
    float* rows = s_rows + w * 32 * D;
    const long long left = E - e0;
    const int n_rows = left < 32 ? (int)left : 32, n = n_rows * D;
    const float* src = obs + e0 * D;                               // 32 * D * 4 bytes per warp: 16-byte aligned
    const int n4 = n >> 2;
    for (int i = lane; i < n4; i += 32) reinterpret_cast<float4*>(rows)[i] = reinterpret_cast<const float4*>(src)[i];
    for (int i = (n4 << 2) + lane; i < n; i += 32) rows[i] = src[i];
    __syncwarp();
    const bool live = lane < n_rows;
    const float* r = rows + lane * D;
    const int n_beams = (D - 4) >> 1, W = (n_beams + 31) >> 5;
    int cnt = 0;
    if (live) {
        head[e0 + lane] = make_float4(r[0], r[1], r[2], r[3]);
        for (int wd = 0; wd < W; ++wd) {
            unsigned m = 0u;
            const int b1 = min(n_beams, 32 * wd + 32);
            for (int b = 32 * wd; b < b1; ++b) m |= (r[4 + 2 * b] != 0.0f || r[5 + 2 * b] != 0.0f) ? 1u << (b & 31) : 0u;
            mask[(e0 + lane) * W + wd] = m;
            cnt += __popc(m);
        }
    }
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    unsigned base = 0;
    if (lane == 0) { base = total > 0 ? atomicAdd(count, (unsigned)total) : 0u; dir[g] = base; }
    if (total == 0) return;
    base = __shfl_sync(0xffffffffu, base, 0);
    unsigned slot = base + (unsigned)(incl - cnt);
    if (live && cnt > 0) {
        for (int b = 0; b < n_beams; ++b) {
            const float x = r[4 + 2 * b], y = r[5 + 2 * b];
            if (x != 0.0f || y != 0.0f) {
                if (slot < capacity) vals[slot] = make_float2(x, y);
                ++slot;
            }
        }
    }
}

}  // namespace

extern "C" int mnv_pack_obs(const float* d_obs, int64_t E, int32_t obs_dim, float* d_head, uint32_t* d_mask, uint32_t* d_dir,
                            uint32_t* d_count, float* d_vals, int64_t capacity, void* stream)
{
    MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_head); MNV_CHECK_PTR(d_mask); MNV_CHECK_PTR(d_dir); MNV_CHECK_PTR(d_count);
    if (d_vals == nullptr || (reinterpret_cast<uintptr_t>(d_vals) & 7u)) { mnv_set_error("mnv_pack_obs: bad value list pointer"); return MNV_E_NULL; }
    if (E <= 0 || obs_dim < 6 || (obs_dim & 1) || (obs_dim - 4) / 2 > MNV_MAX_BEAMS || capacity < 0 || capacity > 0xffffffffll) {
        mnv_set_error("mnv_pack_obs: bad sizes (E=%lld, obs_dim=%d, capacity=%lld)", (long long)E, obs_dim, (long long)capacity);
        return MNV_E_SIZE;
    }
    cudaError_t err = cudaMemsetAsync(d_count, 0, 3 * sizeof(uint32_t), (cudaStream_t)stream);     // count[3] = sequence number: kept
    if (err != cudaSuccess) { mnv_set_error("mnv_pack_obs: memset: %s", cudaGetErrorString(err)); return (int)err; }
    const long long warps = (E + 31) / 32;
    const unsigned grid = (unsigned)((warps + kPackWarps - 1) / kPackWarps);
    const size_t smem = (size_t)kPackWarps * 32 * obs_dim * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t a = cudaFuncSetAttribute(mnv_pack_obs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (a != cudaSuccess) { mnv_set_error("mnv_pack_obs: cudaFuncSetAttribute: %s", cudaGetErrorString(a)); return (int)a; }
    }
    mnv_pack_obs_kernel<<<grid, kPackWarps * 32, smem, (cudaStream_t)stream>>>(d_obs, E, obs_dim, reinterpret_cast<float4*>(d_head), d_mask, d_dir,
                                                                             d_count, reinterpret_cast<float2*>(d_vals), (unsigned)capacity);
    return mnv_launch_status("mnv_pack_obs");
}

// Everything of IQNAgent.train behind the per-tile backward in ONE launch (agent.py:298-300 + the data-parallel exchange):
//   1. sum of the tile partials of iqn_train_kernel in a fixed order (deterministic),
//   2. (data-parallel replicas) one-shot all-reduce of the flat gradient through PEER MEMORY over NVLink: every CTA publishes
//      its 256-parameter slice in this rank's exchange buffer, raises a per-(rank, slice) flag on every peer with a
//      system-scope release store, waits for the peers' flags of the same slice and sums the peers' slices in rank order --
//      the same order on every rank, so the replicas stay bit-identical.  No NCCL call, no extra launch, and a slice is
//      exchanged as soon as it is reduced (the transfer overlaps the reduction of the other slices),
//   3. clip_grad_norm_: per-CTA sums of squares -> grid barrier -> every CTA adds the 140 partials in the same order,
//   4. Adam.step + refresh of the kernel-side weight copies (fp32 transposes, bf16 tensor-core tiles).
// Replaces iqn_reduce_kernel + ncclAllReduce + iqn_clip_adam_kernel (11 + 13..30 + 8 us) with one ~8 us launch.
//
// Grid barrier and flags use monotonically increasing epochs kept on the device (d_sync), so the launch carries no
// host-side counter and may be captured in a CUDA graph.  All 140 CTAs (1 024 threads, < 32 registers, 5 KB of shared
// memory) are co-resident on a B200 by construction (148 SMs); CTAs of other streams can only delay them, not block them.
#include <math.h>
#include <string.h>

#include <cuda_bf16.h>

#include "iqn_common.cuh"

namespace {

using namespace iqn;

constexpr int kTailThreads = 1024;
constexpr int kSlice = 256;                                     // parameters per CTA
constexpr int kTailBlocks = (kParams + kSlice - 1) / kSlice;    // 140
constexpr int kGroups = kTailThreads / kSlice;                  // 4 groups of tiles summed side by side
constexpr int kSlotFloats = kTailBlocks * kSlice;               // one exchange slot: 35 840 floats
constexpr int kMaxWorld = 8;
constexpr size_t kFlagsOffset = 2 * (size_t)kSlotFloats * sizeof(float);           // [2 slots][35 840] f32 | flags u64 [8][160]
constexpr int kFlagStride = 160;
constexpr size_t kXchgBytes = kFlagsOffset + (size_t)kMaxWorld * kFlagStride * sizeof(unsigned long long);
constexpr size_t kSyncBytes = 1024;                             // u64 epoch | u64 arrivals | u64 error | pad to 64 B | f32 ss_part[160]
constexpr unsigned long long kSpinLimit = 1ull << 24;           // x >= 64 ns sleeps: seconds, far beyond any legitimate wait

struct TailPeers {
    float* buf[kMaxWorld];                                      // exchange buffer of every rank (own one included), peer-mapped
    int rank, world;
};

struct TailHyper { float max_norm, step_size, beta1, beta2, inv_sqrt_bc2, eps; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kTailThreads, 1)
iqn_tail_kernel(float* __restrict__ P, float* __restrict__ m, float* __restrict__ v, float* __restrict__ PT,
                __nv_bfloat16* __restrict__ Wtc, const float* __restrict__ gpart, const float* __restrict__ loss_part, int n_tiles,
                float* __restrict__ loss, float* __restrict__ grad_out, float* __restrict__ grad_norm,
                unsigned long long* __restrict__ sync, const __grid_constant__ TailPeers peers, const __grid_constant__ TailHyper H,
                const mnv_vstep_ctl* __restrict__ ctl)
{
    __shared__ float s_part[kGroups][kSlice];
    __shared__ float s_red[32];
    asm volatile("griddepcontrol.wait;" ::: "memory");            // the partials of the train kernel are complete and visible
    const int t = threadIdx.x, p = t & (kSlice - 1), q = t >> 8, blk = blockIdx.x;
    const int i = blk * kSlice + p;
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(sync);   // launches completed so far
    float* ss_part = reinterpret_cast<float*>(sync + 8);

    // ---- 1. tile partials -> gradient slice.  Group q sums the tiles [q per, (q + 1) per) of parameter i with eight
    //         independent accumulators (eight L2 loads in flight per thread); fixed order -> deterministic. ----
    {
        const int per = (n_tiles + kGroups - 1) / kGroups;
        const int t0 = q * per, t1 = min(n_tiles, t0 + per);
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (i < kParams) {
            const float* g = gpart + (long long)t0 * kPartStride + i;
            int tile = t0;
            for (; tile + 8 <= t1; tile += 8, g += 8ll * kPartStride) {
#pragma unroll
                for (int k = 0; k < 8; ++k) a[k] += __ldcg(g + (long long)k * kPartStride);
            }
            for (; tile < t1; ++tile, g += kPartStride) a[0] += __ldcg(g);
        }
        s_part[q][p] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    }
    if (blk == 0 && (t >> 5) == 31 && loss != nullptr) {           // loss = sum of the tile partials (fixed order): one warp of CTA 0
        const int lane = t & 31;
        float acc = 0.f;
        for (int tile = lane; tile < n_tiles; tile += 32) acc += __ldcg(loss_part + tile);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) *loss = acc;
    }
    __syncthreads();
    float g = 0.f;
    if (q == 0) g = (s_part[0][p] + s_part[1][p]) + (s_part[2][p] + s_part[3][p]);

    // ---- 2. one-shot all-reduce over peer memory (data-parallel replicas) ----
    if (peers.world > 1) {
        const int slot = (int)(epoch & 1ull);                      // double-buffered: a rank cannot run two launches ahead of a peer
        const size_t off = (size_t)slot * kSlotFloats + blk * kSlice + p;
        float* mine = peers.buf[peers.rank];                       // (__grid_constant__: indexed straight in the constant bank)
        if (q == 0) mine[off] = g;
        __syncthreads();
        if (t < peers.world && t != peers.rank) {
            float* peer = peers.buf[t];
            __threadfence_system();                                // the slice (written by other threads, ordered by the barrier) before the flag
            unsigned long long* theirs = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(peer) + kFlagsOffset);
            st_release_sys(theirs + peers.rank * kFlagStride + blk, epoch + 1ull);
            const unsigned long long* own = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const char*>(mine) + kFlagsOffset);
            // bounded spin (~2 s): a peer that never launches (a rank died) must not hang this GPU; the error word is sticky
            unsigned long long spins = 0;
            while (ld_acquire_sys(own + t * kFlagStride + blk) < epoch + 1ull) {
                if (++spins > kSpinLimit) { atomicExch(reinterpret_cast<unsigned long long*>(sync) + 2, 1ull); break; }
                if (spins > 4096) __nanosleep(64);
            }
        }
        __syncthreads();
        if (q == 0) {
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < kMaxWorld; ++r) {                  // rank order: identical on every replica
                if (r < peers.world) {
                    const float x = (r == peers.rank) ? g : ld_relaxed_sys(peers.buf[r] + off);
                    sum = r == 0 ? x : sum + x;
                }
            }
            g = sum * (1.f / (float)peers.world);
        }
    }

    // ---- 3. clip_grad_norm_: total norm over all parameters ----
    {
        float ss = (q == 0 && i < kParams) ? g * g : 0.f;
        if (t < kSlice) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
            if ((t & 31) == 0) s_red[t >> 5] = ss;
        }
        __syncthreads();
        if (t == 0) {
            float x = 0.f;
#pragma unroll
            for (int w = 0; w < kSlice / 32; ++w) x += s_red[w];
            __stcg(ss_part + blk, x);
            __threadfence();
            atomicAdd(sync + 1, 1ull);
            const unsigned long long target = (epoch + 1ull) * (unsigned long long)gridDim.x;
            unsigned long long spins = 0;
            while (ld_acquire_gpu(sync + 1) < target) {
                if (++spins > kSpinLimit) { atomicExch(sync + 2, 2ull); break; }
                if (spins > 4096) __nanosleep(64);
            }
        }
        __syncthreads();
        // every CTA adds the per-CTA partials in the same order: 160 slots (zero beyond gridDim.x) over five warps, then 5 adds
        float x = 0.f;
        if (t < kFlagStride) {
            x = t < (int)gridDim.x ? __ldcg(ss_part + t) : 0.f;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            if ((t & 31) == 0) s_red[8 + (t >> 5)] = x;
        }
        __syncthreads();
    }
    const float total = sqrtf((((s_red[8] + s_red[9]) + s_red[10]) + s_red[11]) + s_red[12]);
    const float coef = fminf(H.max_norm / (total + 1e-6f), 1.f);                    // torch.nn.utils.clip_grad_norm_
    if (blk == 0 && t == 0) {
        if (grad_norm != nullptr) *grad_norm = total;
        *reinterpret_cast<volatile unsigned long long*>(sync) = epoch + 1ull;        // every CTA read the epoch before it arrived at the barrier
    }

    // ---- 4. Adam.step (torch defaults: no amsgrad, no weight decay) + the kernel-side copies of the parameters ----
    if (q == 0 && i < kParams) {
        if (grad_out != nullptr) grad_out[i] = g;                                  // the averaged, unclipped gradient
        const float gi = g * coef;
        const float mi = H.beta1 * m[i] + (1.f - H.beta1) * gi;
        const float vi = H.beta2 * v[i] + (1.f - H.beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        // graph replays: Adam's bias corrections (host-computed per step, like H's) come from the control block
        const float inv_sqrt_bc2 = ctl != nullptr ? ctl->adam_inv_sqrt_bc2 : H.inv_sqrt_bc2;
        const float step_size = ctl != nullptr ? ctl->adam_step_size : H.step_size;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + H.eps;
        const float pn = P[i] - step_size * (mi / denom);
        P[i] = pn;
        int pt, tc;
        packed_slots(i, pt, tc);
        if (PT != nullptr && pt >= 0) PT[pt] = pn;
        if (Wtc != nullptr && tc >= 0) Wtc[tc] = __float2bfloat16(pn);
    }
}

}  // namespace

extern "C" int64_t iqn_tail_sync_bytes(void) { return (int64_t)kSyncBytes; }
extern "C" int64_t iqn_xchg_bytes(void) { return (int64_t)kXchgBytes; }
extern "C" int32_t iqn_xchg_handle_bytes(void) { return (int32_t)sizeof(cudaIpcMemHandle_t); }

extern "C" int iqn_xchg_alloc(void** d_ptr, unsigned char* handle)
{
    if (d_ptr == nullptr || handle == nullptr) { mnv_set_error("iqn_xchg_alloc: null argument"); return MNV_E_NULL; }
    cudaError_t e = cudaMalloc(d_ptr, kXchgBytes);
    if (e == cudaSuccess) e = cudaMemset(*d_ptr, 0, kXchgBytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *d_ptr);
    if (e != cudaSuccess) { mnv_set_error("iqn_xchg_alloc: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    memcpy(handle, &h, sizeof(h));
    return 0;
}

extern "C" int iqn_xchg_open(const unsigned char* handle, void** d_ptr)
{
    if (d_ptr == nullptr || handle == nullptr) { mnv_set_error("iqn_xchg_open: null argument"); return MNV_E_NULL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { mnv_set_error("iqn_xchg_open: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    return 0;
}

extern "C" int iqn_xchg_close(void* d_ptr)
{
    cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
    if (e != cudaSuccess) { mnv_set_error("iqn_xchg_close: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    return 0;
}

extern "C" int iqn_xchg_free(void* d_ptr)
{
    cudaError_t e = cudaFree(d_ptr);
    if (e != cudaSuccess) { mnv_set_error("iqn_xchg_free: %s", cudaGetErrorString(e)); cudaGetLastError(); return (int)e; }
    return 0;
}

static int tail_impl(float* d_params, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                     const float* d_scratch, int64_t B, float* d_loss, float* d_grad, float* d_grad_norm,
                     void* d_sync, void* const* peer_xchg, int32_t rank, int32_t world,
                     float max_norm, float lr, float beta1, float beta2, float eps, int64_t step, const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_update_tail: B must be > 0"); return MNV_E_SIZE; }
    if (step < 1) { mnv_set_error("iqn_update_tail: step must be >= 1"); return MNV_E_PARAM; }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_m); MNV_CHECK_PTR(d_v); MNV_CHECK_PTR(d_scratch); MNV_CHECK_PTR(d_sync);
    MNV_CHECK_PTR_OPT(d_packed); MNV_CHECK_PTR_OPT(d_packed_tc); MNV_CHECK_PTR_OPT(d_grad);
    TailPeers peers;
    memset(&peers, 0, sizeof(peers));
    peers.rank = 0; peers.world = 1;
    if (world > 1) {
        if (world > kMaxWorld || rank < 0 || rank >= world || peer_xchg == nullptr) {
            mnv_set_error("iqn_update_tail: world=%d (max %d), rank=%d, peer_xchg=%p", world, kMaxWorld, rank, (const void*)peer_xchg);
            return MNV_E_PARAM;
        }
        for (int r = 0; r < world; ++r) {
            if (peer_xchg[r] == nullptr) { mnv_set_error("iqn_update_tail: peer_xchg[%d] is null", r); return MNV_E_NULL; }
            peers.buf[r] = (float*)peer_xchg[r];
        }
        peers.rank = rank; peers.world = world;
    }
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    TailHyper H{max_norm, (float)((double)lr / bc1), beta1, beta2, (float)(1.0 / sqrt(bc2)), eps};
    const long long tiles = (B + 7) / 8;
    const float* gpart = d_scratch;
    const float* lpart = d_scratch + tiles * (long long)kPartStride;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(kTailBlocks); cfg.blockDim = dim3(kTailThreads); cfg.dynamicSmemBytes = 0; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // CTAs are placed while the train kernel drains; they wait at
    attr[0].val.programmaticStreamSerializationAllowed = 1;           // griddepcontrol.wait until its partials are complete
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, iqn_tail_kernel, d_params, d_m, d_v, d_packed, (__nv_bfloat16*)d_packed_tc, gpart, lpart, (int)tiles,
                       d_loss, d_grad, d_grad_norm, (unsigned long long*)d_sync, peers, H, d_ctl);
    return mnv_launch_status("iqn_update_tail");
}

extern "C" int iqn_update_tail(float* d_params, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                               const float* d_scratch, int64_t B, float* d_loss, float* d_grad, float* d_grad_norm,
                               void* d_sync, void* const* peer_xchg, int32_t rank, int32_t world,
                               float max_norm, float lr, float beta1, float beta2, float eps, int64_t step, void* stream)
{
    return tail_impl(d_params, d_m, d_v, d_packed, d_packed_tc, d_scratch, B, d_loss, d_grad, d_grad_norm, d_sync, peer_xchg, rank, world,
                     max_norm, lr, beta1, beta2, eps, step, nullptr, stream);
}

extern "C" int iqn_update_tail_ctl(float* d_params, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                                   const float* d_scratch, int64_t B, float* d_loss, float* d_grad, float* d_grad_norm,
                                   void* d_sync, void* const* peer_xchg, int32_t rank, int32_t world,
                                   float max_norm, float beta1, float beta2, float eps, const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (d_ctl == nullptr) { mnv_set_error("iqn_update_tail_ctl: null control block"); return MNV_E_NULL; }
    return tail_impl(d_params, d_m, d_v, d_packed, d_packed_tc, d_scratch, B, d_loss, d_grad, d_grad_norm, d_sync, peer_xchg, rank, world,
                     max_norm, 0.f, beta1, beta2, eps, 1, d_ctl, stream);
}

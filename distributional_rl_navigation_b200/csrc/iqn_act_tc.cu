// IQN acting forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   IQNAgent.act / ObsEncoder.get_qvals (thirdparty/IQN/agent.py:186-205, model.py:160-191) for a whole env batch:
//   K = 32 quantile samples per environment, mean over the samples, argmax over the 9 actions.
//
// This is the one real dense GEMM of the path: 65 536 envs x 32 taus = 2.1 M rows through 64->208->64->64->9
// (132 GFLOP per vector step).  Only the argmax of the mean is consumed, so operands are bf16 with fp32 accumulation
// (the fp32 FFMA kernel in iqn.cu stays the parity path for training and for act_eval's quantile outputs).
//
// One persistent CTA per SM, 256 threads, a tile = 128 rows = 4 environments x 32 taus:
//   * all four weight matrices live in shared memory for the whole kernel as bf16 K-major core-matrix tiles
//     (pre-packed by iqn_pack_tc, 63.5 KB);
//   * A operands are produced in-kernel and written straight into the same UMMA canonical layout (no swizzle):
//       A0 = cos(pi i tau)            (rotation recurrence from one sincospif per row)
//       A1 = relu(D1 + b_c) * feat    A2 = relu(D2 + b_1)    A3 = relu(D3 + b_2)
//   * one elected thread issues tcgen05.mma (M = 128, N = 208 / 64 / 64 / 16, K = 16 per instruction), accumulators in
//     TMEM (208 + 64 + 64 + 16 columns), completion through tcgen05.commit -> mbarrier;
//   * epilogues read TMEM with tcgen05.ld.32x32b (thread = row), apply bias / relu / the feature product, convert to
//     bf16 and store 16-byte chunks for the next layer; the last epilogue averages the 32 rows of an environment with
//     warp shuffles (one warp == one environment) and takes the argmax.
#include <cuda_bf16.h>
#include <math.h>

#include "iqn_common.cuh"

namespace {

using namespace iqn;

constexpr int kThreads = 256;
constexpr int kRows = 128;              // rows per tile (UMMA M)
constexpr int kTaus = 32;               // quantile samples per environment (ObsEncoder.K)
constexpr int kEnvsPerTile = kRows / kTaus;
constexpr int kN4 = 16;                 // output layer padded 9 -> 16 (UMMA N granularity at M = 128)
constexpr int kTmemCols = 512;

// bf16 element counts of the packed weight tiles (core-matrix layout, see tile_offset)
constexpr int kWcEl = kFeat * kCos, kW1El = kHid * kFeat, kW2El = kHid * kHid, kW3El = kN4 * kHid;
constexpr int kPackedTcEl = kWcEl + kW1El + kW2El + kW3El;       // 31 744 bf16 = 63 488 bytes

// TMEM column bases of the four accumulators
constexpr uint32_t kD1 = 0, kD2 = 208, kD3 = 272, kD4 = 336;

// UMMA canonical K-major layout without swizzle: 8 x 8 (bf16) core matrices of 128 contiguous bytes; the core matrices of
// one 8-row group are contiguous along K (LBO = 128 B) and row groups follow each other (SBO = (K/8) * 128 B).
__host__ __device__ constexpr int tile_offset(int r, int k, int K)       // in elements
{
    return (r >> 3) * (K * 8) + (k >> 3) * 64 + (r & 7) * 8 + (k & 7);
}

struct __align__(128) Smem {
    __nv_bfloat16 wc[kWcEl], w1[kW1El], w2[kW2El], w3[kW3El];
    __nv_bfloat16 a0[kRows * kCos], a1[kRows * kFeat], a2[kRows * kHid], a3[kRows * kHid];
    float feat[kEnvsPerTile * kFeat];
    float bc[kFeat], b1[kHid], b2[kHid], b3[kN4];
    float x[kEnvsPerTile * 28];
    float tau[kRows];
    unsigned long long bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (SWIZZLE_NONE, K-major), cute::UMMA::SmemDescriptor bit layout
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int K)
{
    const uint64_t lbo = 128 >> 4, sbo = (uint64_t)(K * 16) >> 4;
    return (uint64_t)((saddr >> 4) & 0x3fff) | (lbo << 16) | (sbo << 32) | (1ull << 46);      // version = 1, layout_type = 0
}

// instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void store_chunk(__nv_bfloat16* base, int r, int kc, int K, const float (&v)[8])
{
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(base + tile_offset(r, kc * 8, K)) = u;
}

// issue the K/16 MMAs of one layer (one elected thread), then commit to the mbarrier
__device__ __forceinline__ void issue_layer(const __nv_bfloat16* A, const __nv_bfloat16* B, int K, int N, uint32_t tmem_d,
                                            unsigned long long* bar)
{
    const uint32_t idesc = make_idesc(kRows, N);
    const uint32_t a0 = smem_u32(A), b0 = smem_u32(B);
    for (int k = 0; k < K / 16; ++k)                       // one instruction consumes K = 16 = two 128-byte core matrices
        umma(tmem_d, make_desc(a0 + k * 256, K), make_desc(b0 + k * 256, K), idesc, k > 0 ? 1u : 0u);
    umma_commit(bar);
}

__global__ void __launch_bounds__(kThreads, 1)
iqn_act_tc_kernel(const float* __restrict__ P, const __nv_bfloat16* __restrict__ Wp, const float* __restrict__ obs,
                  const float* __restrict__ taus, const float* __restrict__ cvar, float cvar_scalar,
                  float* __restrict__ qmean, int32_t* __restrict__ greedy, float* __restrict__ debug, long long B)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    // ---- one-time setup: weights + biases to shared memory, TMEM allocation, mbarrier ----
    {
        const uint4* src = reinterpret_cast<const uint4*>(Wp);
        uint4* dst = reinterpret_cast<uint4*>(s.wc);        // wc, w1, w2, w3 are contiguous in Smem and in the packed buffer
        for (int i = t; i < kPackedTcEl / 8; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = t; i < kFeat; i += kThreads) s.bc[i] = P[oCB + i];
        if (t < kHid) { s.b1[t] = P[oH1B + t]; s.b2[t] = P[oH2B + t]; }
        if (t < kN4) s.b3[t] = t < kAct ? P[oOB + t] : 0.f;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (t == 32) mbar_init(&s.bar, 1);
    fence_async_smem();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;     // this warp's TMEM lane quadrant
    const int half = warp >> 2;                                       // column half handled by this warpgroup
    const int row = (warp & 3) * 32 + lane;                           // TMEM lane == tile row of this thread
    uint32_t phase = 0;

    const long long n_tiles = (B + kEnvsPerTile - 1) / kEnvsPerTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long env0 = tile * kEnvsPerTile;
        // ---- inputs ----
        for (int idx = t; idx < kEnvsPerTile * 28; idx += kThreads) {
            const int e = idx / 28, k = idx % 28;
            const long long b = env0 + e;
            s.x[idx] = (b < B && k < kObs) ? obs[b * kObs + k] : 0.f;
        }
        if (t < kRows) {
            const long long b = env0 + t / kTaus;
            float v = 0.f;
            if (b < B) v = taus[b * kTaus + (t % kTaus)] * (cvar != nullptr ? cvar[b] : cvar_scalar);    // model.py:153
            s.tau[t] = v;
        }
        __syncthreads();
        // ---- observation encoders (fp32, model.py:169-172) ----
        for (int idx = t; idx < kEnvsPerTile * kFeat; idx += kThreads) {
            const int e = idx / kFeat, f = idx % kFeat;
            const float* x = s.x + e * 28;
            float v;
            if (f < 16) v = fmaf(__ldg(P + oVW + f * 2 + 1), x[1], fmaf(__ldg(P + oVW + f * 2), x[0], __ldg(P + oVB + f)));
            else if (f < 32) {
                const int g = f - 16;
                v = fmaf(__ldg(P + oGW + g * 2 + 1), x[3], fmaf(__ldg(P + oGW + g * 2), x[2], __ldg(P + oGB + g)));
            } else {
                const int g = f - 32;
                v = __ldg(P + oSB + g);
#pragma unroll
                for (int k = 0; k < 22; ++k) v = fmaf(__ldg(P + oSW + g * 22 + k), x[4 + k], v);
            }
            s.feat[idx] = v;
        }
        // ---- A0 = cos(pi * i * tau), i = 0..63: thread (row, half) fills i in [32 half, 32 half + 32) by rotating
        //      (cos, sin)(i0 * pi * tau) with (cos, sin)(pi * tau) ----
        {
            const int r = t & 127, h = t >> 7;
            const float tau = s.tau[r];
            float c1, s1, c, sn;
            sincospif(tau, &s1, &c1);
            sincospif(32.f * (float)h * tau, &sn, &c);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    v[j] = c;
                    const float c2 = fmaf(c, c1, -sn * s1), s2 = fmaf(sn, c1, c * s1);
                    c = c2; sn = s2;
                }
                store_chunk(s.a0, r, h * 4 + kc, kCos, v);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- layer 1: D1[128 x 208] = A0 . Wc^T ----
        if (t == 0) { tc_fence_after(); issue_layer(s.a0, s.wc, kCos, kFeat, tmem + kD1, &s.bar); }
        mbar_wait(&s.bar, phase); phase ^= 1;
        tc_fence_after();
        {
            const float* feat = s.feat + (row / kTaus) * kFeat;
            for (int ch = half * 13; ch < half * 13 + 13; ++ch) {           // 26 chunks of 8 columns, 13 per warpgroup
                float v[8];
                tmem_ld8(tmem + lane_base + kD1 + ch * 8, v);
                if (debug != nullptr && tile == 0)
                    for (int j = 0; j < 8; ++j) debug[row * kFeat + ch * 8 + j] = v[j];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + s.bc[ch * 8 + j], 0.f) * feat[ch * 8 + j];   // model.py:177-180
                store_chunk(s.a1, row, ch, kFeat, v);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- layer 2: D2[128 x 64] = A1 . W1^T ----
        if (t == 0) { tc_fence_after(); issue_layer(s.a1, s.w1, kFeat, kHid, tmem + kD2, &s.bar); }
        mbar_wait(&s.bar, phase); phase ^= 1;
        tc_fence_after();
        for (int ch = half * 4; ch < half * 4 + 4; ++ch) {
            float v[8];
            tmem_ld8(tmem + lane_base + kD2 + ch * 8, v);
            if (debug != nullptr && tile == 0)
                for (int j = 0; j < 8; ++j) debug[kRows * kFeat + row * kHid + ch * 8 + j] = v[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + s.b1[ch * 8 + j], 0.f);
            store_chunk(s.a2, row, ch, kHid, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- layer 3: D3[128 x 64] = A2 . W2^T ----
        if (t == 0) { tc_fence_after(); issue_layer(s.a2, s.w2, kHid, kHid, tmem + kD3, &s.bar); }
        mbar_wait(&s.bar, phase); phase ^= 1;
        tc_fence_after();
        for (int ch = half * 4; ch < half * 4 + 4; ++ch) {
            float v[8];
            tmem_ld8(tmem + lane_base + kD3 + ch * 8, v);
            if (debug != nullptr && tile == 0)
                for (int j = 0; j < 8; ++j) debug[kRows * (kFeat + kHid) + row * kHid + ch * 8 + j] = v[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j] + s.b2[ch * 8 + j], 0.f);
            store_chunk(s.a3, row, ch, kHid, v);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- output layer: D4[128 x 16] = A3 . W3^T, then mean over the 32 taus of each env (one warp) + argmax ----
        if (t == 0) { tc_fence_after(); issue_layer(s.a3, s.w3, kHid, kN4, tmem + kD4, &s.bar); }
        mbar_wait(&s.bar, phase); phase ^= 1;
        tc_fence_after();
        if (half == 0) {
            float q[kN4];
            {
                float v[8];
                tmem_ld8(tmem + lane_base + kD4, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) q[j] = v[j];
                tmem_ld8(tmem + lane_base + kD4 + 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) q[8 + j] = v[j];
            }
            if (debug != nullptr && tile == 0)
                for (int j = 0; j < kN4; ++j) debug[kRows * (kFeat + 2 * kHid) + row * kN4 + j] = q[j];
#pragma unroll
            for (int a = 0; a < kAct; ++a) {
                float v = q[a];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                q[a] = v * (1.f / kTaus) + s.b3[a];                                      // get_qvals: mean over taus (model.py:190)
            }
            const long long b = env0 + (warp & 3);
            if (lane == 0 && b < B) {
                int best = 0; float bv = q[0];
#pragma unroll
                for (int a = 0; a < kAct; ++a) {
                    if (qmean != nullptr) qmean[b * kAct + a] = q[a];
                    if (q[a] > bv) { bv = q[a]; best = a; }                             // np.argmax: first maximum (agent.py:201)
                }
                if (greedy != nullptr) greedy[b] = best;
            }
        }
        tc_fence_before();
        __syncthreads();
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// fp32 parameters -> bf16 K-major core-matrix tiles: Wc [208][64], W1 [64][208], W2 [64][64], W3 [16][64] (rows >= 9 zero)
__global__ void __launch_bounds__(256) iqn_pack_tc_kernel(const float* __restrict__ P, __nv_bfloat16* __restrict__ W)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < kWcEl) { const int n = i / kCos, k = i % kCos; W[tile_offset(n, k, kCos)] = __float2bfloat16(P[oCW + i]); }
    else if (i < kWcEl + kW1El) { const int j = i - kWcEl, n = j / kFeat, k = j % kFeat; W[kWcEl + tile_offset(n, k, kFeat)] = __float2bfloat16(P[oH1W + j]); }
    else if (i < kWcEl + kW1El + kW2El) { const int j = i - kWcEl - kW1El, n = j / kHid, k = j % kHid; W[kWcEl + kW1El + tile_offset(n, k, kHid)] = __float2bfloat16(P[oH2W + j]); }
    else if (i < kPackedTcEl) {
        const int j = i - kWcEl - kW1El - kW2El, n = j / kHid, k = j % kHid;
        W[kWcEl + kW1El + kW2El + tile_offset(n, k, kHid)] = __float2bfloat16(n < kAct ? P[oOW + n * kHid + k] : 0.f);
    }
}

}  // namespace

extern "C" int iqn_packed_tc_bytes(void) { return kPackedTcEl * 2; }

extern "C" int iqn_pack_tc(const float* d_params, void* d_packed_tc, void* stream)
{
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed_tc);
    iqn_pack_tc_kernel<<<(kPackedTcEl + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_params, (__nv_bfloat16*)d_packed_tc);
    return mnv_launch_status("iqn_pack_tc");
}

extern "C" int iqn_act_tc(const float* d_params, const void* d_packed_tc, const float* d_obs, const float* d_taus,
                          const float* d_cvar, float cvar_scalar, float* d_qmean, int32_t* d_greedy, float* d_debug,
                          int64_t B, int32_t n_tau, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_act_tc: B must be > 0"); return MNV_E_SIZE; }
    if (n_tau != kTaus) { mnv_set_error("iqn_act_tc: n_tau must be 32 (ObsEncoder.K), got %d", n_tau); return MNV_E_CAPACITY; }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed_tc); MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_taus);
    if (d_qmean == nullptr && d_greedy == nullptr) { mnv_set_error("iqn_act_tc: no output"); return MNV_E_NULL; }
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(iqn_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) { mnv_set_error("cudaFuncSetAttribute(iqn_act_tc): %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long n_tiles = (B + kEnvsPerTile - 1) / kEnvsPerTile;
    const int grid = (int)(n_tiles < sms ? n_tiles : sms);
    iqn_act_tc_kernel<<<grid, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(
        d_params, (const __nv_bfloat16*)d_packed_tc, d_obs, d_taus, d_cvar, cvar_scalar, d_qmean, d_greedy, d_debug, B);
    return mnv_launch_status("iqn_act_tc");
}

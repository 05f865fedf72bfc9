// IQN acting forward on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   IQNAgent.act / ObsEncoder.get_qvals (thirdparty/IQN/agent.py:186-205, model.py:160-191) for a whole env batch:
//   K = 32 quantile samples per environment, mean over the samples, argmax over the 9 actions.
//
// This is the one real dense GEMM of the path: 65 536 envs x 32 taus = 2.1 M rows through 64->208->64->64->9
// (132 GFLOP per vector step).  Only the argmax of the mean is consumed, so operands are bf16 with fp32 accumulation
// (the fp32 FFMA kernel in iqn.cu stays the parity path for training and for act_eval's quantile outputs).
//
// A pre-pass (iqn_encode_kernel) computes what depends on the environment alone: the three observation encoders
// (fp32 FFMA, 208 features, stored as bf16 -- they only ever multiply a bf16 operand) and, for the adaptive policy, the
// CVaR level of agent.py:249-267.  The quantile samples come either from the caller (parity harnesses) or from the
// counter-based Philox stream (seed, step) inside the kernel, which then also applies the epsilon-greedy rule of
// agent.py:200-203 -- no tau tensor in HBM, no separate random / select launches.
//
// One persistent CTA per SM, 512 threads = two groups of 256, each group owns one tile (128 rows = 4 environments x 32 taus)
// at a time, so that one group's epilogue overlaps the other group's MMAs; inside a group the FIRST layer of the next tile is
// software-pipelined under the later layers of the current one:
//   * all four weight matrices live in shared memory for the whole kernel as bf16 K-major core-matrix tiles (pre-packed by
//     iqn_pack_tc with the bias as one extra reduction column; layer 1's bias is folded into its cos_0 = 1 column when the
//     tile is copied in, so its reduction length is exactly 64);
//   * A operands are produced in-kernel and written straight into the same UMMA canonical layout (no swizzle):
//       A0 = cos(pi i tau)            (Chebyshev recurrence c_{i+1} = 2 c_1 c_i - c_{i-1}: one FFMA per feature; own 16 KB buffer)
//       A1 = bf16(relu(D1)) * bf16(feat)    A2 = relu(D2)    A3 = relu(D3)     (cvt.rn.relu.bf16x2 + mul.bf16x2; one 56 KB region)
//   * one elected thread issues tcgen05.mma (M = 128, N = 192 + 16 / 64 / 64 / 16, K = 16 per instruction), accumulators in
//     TMEM, completion through tcgen05.commit -> two mbarriers per group (layer 1 | layers 2-4);
//   * per tile i:  wait D1(i) -> epilogue 1 -> issue layer 2 -> [while it runs: A0(i+1), issue layer 1a(i+1)] -> epilogue 2
//     -> layer 3 -> epilogue 3 -> issue layer 4 (+ layer 1b(i+1), whose 16 TMEM columns are free only now) -> mean over the 32
//     rows of an environment with warp shuffles (one warp == one environment), argmax, epsilon-greedy.
//     TMEM columns of a group (256): D1a (features 0..191) at [64, 256), D1b (192..207) at [0, 16), D2 = D3 at [0, 64),
//     D4 at [32, 48);
//   * epilogues read TMEM with tcgen05.ld.32x32b (thread = row), convert to bf16 and store 16-byte chunks for the next layer.
#include <cuda_bf16.h>
#include <math.h>
#include <type_traits>

#include "iqn_common.cuh"
#include "philox.cuh"

namespace {

using namespace iqn;

constexpr int kComputeThreads = 512;     // two tile groups x 256 threads (8 warps: 4 TMEM lane quadrants x 2 column halves)
constexpr int kThreads = kComputeThreads + 64;   // + one MMA-issuing warp per tile group (warps 16 and 17)
constexpr int kGroupThreads = 256;
constexpr int kRows = 128;              // rows per tile (UMMA M)
constexpr int kTaus = 32;               // quantile samples per environment (ObsEncoder.K)
constexpr int kEnvsPerTile = kRows / kTaus;
constexpr int kTmemCols = 512;

// TMEM column bases inside a group's 256 columns (see the header comment)
constexpr uint32_t kD1a = 64, kD1b = 0, kD2 = 0, kD3 = 0, kD4 = 32;
constexpr int kN1a = 192, kN1b = kFeat - kN1a;      // layer 1 is issued as two MMAs: features [0, 192) and [192, 208)
constexpr int kK0s = kCos;                          // layer 1's reduction length in shared memory: bias folded into column 0

// per tile-group buffers (two groups of 256 threads keep two tiles in flight per CTA)
struct __align__(128) GroupSmem {
    // A1 (52 KB + bias step) is written by the first epilogue; A2 / A3 (16 KB each) overwrite it after layers 2 / 3 consumed it
    __nv_bfloat16 a1[kRows * kK1];
    __nv_bfloat16 a0[kRows * kK0s];                      // cos features of the NEXT tile while this tile is in layers 2-4
    __nv_bfloat16 feat[kEnvsPerTile * kFeat];            // observation-encoder output of the tile's environments (from the pre-pass)
    unsigned long long bar_a, bar_b;                     // layer 1 | layers 2-4
};

struct __align__(128) Smem {
    __nv_bfloat16 wc[kFeat * kK0s], w1[kW1El], w2[kW2El], w3[kW3El];
    GroupSmem g[2];
    uint32_t tmem_base;
    float dyn_eps;                                       // eps / Philox step of this launch: by-value arguments, or (graph replays)
    unsigned long long dyn_step;                         // the device control block -- read once, kept out of the register budget
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
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kGroupThreads) : "memory"); }
// hand-off points between a group's 8 compute warps (arrive, never wait) and its MMA-issuing warp (sync): one named barrier
// per point (0: A1 ready, 1: A0 of the next tile ready, 2: A2 ready, 3: A3 ready), 256 + 32 participants
__device__ __forceinline__ void handoff_arrive(int g, int k) { asm volatile("bar.arrive %0, %1;" ::"r"(3 + 4 * g + k), "r"(kGroupThreads + 32) : "memory"); }
__device__ __forceinline__ void handoff_wait(int g, int k) { asm volatile("bar.sync %0, %1;" ::"r"(3 + 4 * g + k), "r"(kGroupThreads + 32) : "memory"); }

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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Split form for software-pipelined epilogues: the load is asynchronous until tcgen05.wait::ld, so the next block's load can be
// in flight while this block is converted and stored.  The wait names the registers as in/out operands: that is what orders
// the consumers behind it for the compiler (a plain "memory" clobber does not order register reads).
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

// 32 consecutive columns of this thread's row with ONE round trip to TMEM
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
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

// relu + round to bf16 of two accumulator columns in ONE instruction (cvt.rn.relu.bf16x2.f32: upper half <- a, lower half <- b)
__device__ __forceinline__ uint32_t relu_pack(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// one 8-column chunk of the next layer's A operand: relu(v) in bf16 (x feat in bf16 when f != nullptr), 16-byte store
__device__ __forceinline__ void store_relu_chunk(__nv_bfloat16* base, int r, int kc, int K, const float* v, const __nv_bfloat16* f)
{
    uint4 u;
    u.x = relu_pack(v[0], v[1]); u.y = relu_pack(v[2], v[3]); u.z = relu_pack(v[4], v[5]); u.w = relu_pack(v[6], v[7]);
    if (f != nullptr) {
        const uint4 fv = *reinterpret_cast<const uint4*>(f);           // 8 bf16 features, 16-byte aligned
        u.x = mul_bf16x2(u.x, fv.x); u.y = mul_bf16x2(u.y, fv.y); u.z = mul_bf16x2(u.z, fv.z); u.w = mul_bf16x2(u.w, fv.w);
    }
    *reinterpret_cast<uint4*>(base + tile_offset(r, kc * 8, K)) = u;
}

// the bias K-step of an A operand: chunk kc = (1, 0, ..., 0), chunk kc + 1 = 0
__device__ __forceinline__ void store_bias_step(__nv_bfloat16* base, int r, int kc, int K)
{
    *reinterpret_cast<uint4*>(base + tile_offset(r, kc * 8, K)) = make_uint4(0x00003f80u, 0u, 0u, 0u);      // bf16(1.0) = 0x3f80
    *reinterpret_cast<uint4*>(base + tile_offset(r, kc * 8 + 8, K)) = make_uint4(0u, 0u, 0u, 0u);
}

// issue the K/16 MMAs of one layer (one elected thread of a warp whose operand addresses are warp-uniform FOR THE COMPILER --
// see the *_u values in the kernel -- so that the descriptors sit in uniform registers and every MMA is one UTCHMMA;
// thread-dependent addresses make ptxas wrap each MMA in an ELECT / R2UR.BROADCAST waterfall, ~100 cycles per instruction),
// then commit to the mbarrier (if given).  One instruction consumes K = 16 = two 128-byte core matrices: the descriptors'
// address fields advance by 256 B >> 4.
template <int K, int N, int K0 = 0, int K1 = K / 16>
__device__ __forceinline__ void issue_layer(uint32_t a_saddr, uint32_t b_saddr, uint32_t tmem_d, unsigned long long* bar)
{
    constexpr uint32_t idesc = make_idesc(kRows, N);
    const uint64_t da = make_desc(a_saddr, K), db = make_desc(b_saddr, K);
#pragma unroll
    for (int k = K0; k < K1; ++k) umma(tmem_d, da + (uint64_t)(k * 16), db + (uint64_t)(k * 16), idesc, k > 0 ? 1u : 0u);   // K-steps [K0, K1)
    if (bar != nullptr) umma_commit(bar);
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// ---- pre-pass: everything that depends on the environment alone -------------------------------------------------------
// ObsEncoder's three encoders (model.py:125-127,169-172; no activation), 128 environments per CTA of 4 warps.  A lane owns TWO
// environments (their 26 observation values in registers); a warp owns 64 environments and HALF of the 13 groups of 16
// features (group 0 = velocity encoder, 1 = goal encoder, 2..12 = sensor encoder).  The weights sit in shared memory as
// [group][input][16 features]: the 16 weights of one input are four broadcast 16-byte loads feeding sixteen packed FFMA2 (two
// features per instruction, two environments per load).  A broadcast 16-byte load still moves 512 bytes into the register
// file (4 cycles of the shared-memory pipe), so with one environment per lane the kernel was bound by that pipe (ncu: l1tex
// 79 %, mio_throttle the top stall); two environments per lane put it back on the FMA pipe.  Every feature is accumulated
// bias-first in ascending input order (bit-identical to a scalar fmaf chain).  Output: bf16 [B][208] rows, 32 contiguous
// bytes per (environment, group).  Optionally the adaptive CVaR level of IQNAgent.adjust_cvar (agent.py:249-267): closest
// sonar return / 10 if closer than 10 m, else 1; a beam with |x|, |y| < 1e-3 is "no return".
constexpr int kEncEnvs = 128, kEncThreads = 128, kEncGroups = kFeat / 16;     // 13 groups of 16 features
constexpr int kEncSensorIn = kObs - 4;                                        // 22 sonar inputs
// shared-memory weight layout: group g < 2: [2 inputs][16] then bias [16]; sensor group: [22 inputs][16] then bias [16]
constexpr int kEncSmallFloats = (2 + 1) * 16, kEncSensorFloats = (kEncSensorIn + 1) * 16;
constexpr int kEncWFloats = 2 * kEncSmallFloats + (kEncGroups - 2) * kEncSensorFloats;     // 4 144: every encoder parameter once

__device__ __forceinline__ void enc_store16(const float2 (&acc)[8], __nv_bfloat16* __restrict__ out)
{
    uint4 u0, u1;
    auto pk = [](float2 v) { __nv_bfloat162 p = __floats2bfloat162_rn(v.x, v.y); return *reinterpret_cast<uint32_t*>(&p); };
    u0.x = pk(acc[0]); u0.y = pk(acc[1]); u0.z = pk(acc[2]); u0.w = pk(acc[3]);
    u1.x = pk(acc[4]); u1.y = pk(acc[5]); u1.z = pk(acc[6]); u1.w = pk(acc[7]);
    reinterpret_cast<uint4*>(out)[0] = u0;                                  // 32-byte aligned: rows are 416 = 13 x 32 bytes
    reinterpret_cast<uint4*>(out)[1] = u1;
}

// one group of 16 features for the lane's two environments (xa, xb: their inputs; outa / outb may be null for a missing env)
template <int NIN>
__device__ __forceinline__ void enc_group(const float* __restrict__ wg, const float* xa, const float* xb,
                                          __nv_bfloat16* __restrict__ outa, __nv_bfloat16* __restrict__ outb)
{
    float2 acca[8], accb[8];
    const float4* b4 = reinterpret_cast<const float4*>(wg + NIN * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = b4[q];
        acca[2 * q] = make_float2(b.x, b.y); acca[2 * q + 1] = make_float2(b.z, b.w);
        accb[2 * q] = acca[2 * q]; accb[2 * q + 1] = acca[2 * q + 1];
    }
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
        const float4* w4 = reinterpret_cast<const float4*>(wg + k * 16);
        const float2 xxa = make_float2(xa[k], xa[k]), xxb = make_float2(xb[k], xb[k]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = w4[q];
            const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
            acca[2 * q] = __ffma2_rn(w01, xxa, acca[2 * q]); acca[2 * q + 1] = __ffma2_rn(w23, xxa, acca[2 * q + 1]);
            accb[2 * q] = __ffma2_rn(w01, xxb, accb[2 * q]); accb[2 * q + 1] = __ffma2_rn(w23, xxb, accb[2 * q + 1]);
        }
    }
    if (outa != nullptr) enc_store16(acca, outa);
    if (outb != nullptr) enc_store16(accb, outb);
}

__global__ void __launch_bounds__(kEncThreads, 4)         // 4 CTAs per SM: the 512 CTAs of a 65 536-env batch are ONE wave
iqn_encode_kernel(const float* __restrict__ P, const float* __restrict__ obs, __nv_bfloat16* __restrict__ feat,
                  float* __restrict__ cvar_out, long long B)
{
    __shared__ __align__(16) float s_w[kEncWFloats];
    __shared__ __align__(16) float s_raw[oCW];                         // the encoder parameters as stored (coalesced 16-byte copy)
    __shared__ __align__(16) float s_x[kEncEnvs * kObs];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const long long e0 = (long long)blockIdx.x * kEncEnvs;
    const int n_env = (B - e0) < kEncEnvs ? (int)(B - e0) : kEncEnvs;
    for (int i = t; i < oCW / 4; i += kEncThreads) reinterpret_cast<float4*>(s_raw)[i] = __ldg(reinterpret_cast<const float4*>(P) + i);
    for (int i = t; i < n_env * kObs; i += kEncThreads) s_x[i] = obs[e0 * kObs + i];
    __syncthreads();
    // parameters (state_dict order: oVW [16][2], oVB, oGW [16][2], oGB, oSW [176][22], oSB) -> [group][input][16] + bias rows
    for (int i = t; i < kEncWFloats; i += kEncThreads) {
        float v;
        if (i < 2 * kEncSmallFloats) {
            const int g = i / kEncSmallFloats, r = i % kEncSmallFloats, k = r / 16, j = r % 16;
            const int ow = g == 0 ? oVW : oGW, ob = g == 0 ? oVB : oGB;
            v = k < 2 ? s_raw[ow + j * 2 + k] : s_raw[ob + j];
        } else {
            const int r0 = i - 2 * kEncSmallFloats, g = r0 / kEncSensorFloats, r = r0 % kEncSensorFloats, k = r / 16, j = r % 16;
            v = k < kEncSensorIn ? s_raw[oSW + (g * 16 + j) * kEncSensorIn + k] : s_raw[oSB + g * 16 + j];
        }
        s_w[i] = v;
    }
    __syncthreads();
    // warp w: environments [64 (w & 1), 64 (w & 1) + 64) of the CTA (lane: el and el + 32), feature half w >> 1
    const int ela = (w & 1) * 64 + lane, elb = ela + 32;
    const bool has_a = ela < n_env, has_b = elb < n_env;
    float xa[kObs], xb[kObs];
#pragma unroll
    for (int k = 0; k < kObs; ++k) { xa[k] = has_a ? s_x[ela * kObs + k] : 0.f; xb[k] = has_b ? s_x[elb * kObs + k] : 0.f; }
    __nv_bfloat16* outa = has_a ? feat + (e0 + ela) * kFeat : nullptr;
    __nv_bfloat16* outb = has_b ? feat + (e0 + elb) * kFeat : nullptr;
    auto at = [](__nv_bfloat16* p, int off) { return p != nullptr ? p + off : nullptr; };
    if ((w >> 1) == 0) {                                                  // warp-uniform: groups 0..6
        enc_group<2>(s_w, xa, xb, outa, outb);
        enc_group<2>(s_w + kEncSmallFloats, xa + 2, xb + 2, at(outa, 16), at(outb, 16));
#pragma unroll 1
        for (int g = 2; g < 7; ++g)
            enc_group<kEncSensorIn>(s_w + 2 * kEncSmallFloats + (g - 2) * kEncSensorFloats, xa + 4, xb + 4, at(outa, g * 16), at(outb, g * 16));
        if (cvar_out != nullptr) {
            auto level = [](const float* x) {
                float closest = INFINITY;
#pragma unroll
                for (int b = 0; b < kEncSensorIn / 2; ++b) {
                    const float px = x[4 + 2 * b], py = x[5 + 2 * b];
                    if (fabsf(px) < 1e-3f && fabsf(py) < 1e-3f) continue;     // agent.py:256-258
                    closest = fminf(closest, sqrtf(px * px + py * py));
                }
                return closest < 10.0f ? closest / 10.0f : 1.0f;               // agent.py:262-265 (sonar range 10 m)
            };
            if (has_a) cvar_out[e0 + ela] = level(xa);
            if (has_b) cvar_out[e0 + elb] = level(xb);
        }
    } else {                                                              // groups 7..12
#pragma unroll 1
        for (int g = 7; g < kEncGroups; ++g)
            enc_group<kEncSensorIn>(s_w + 2 * kEncSmallFloats + (g - 2) * kEncSensorFloats, xa + 4, xb + 4, at(outa, g * 16), at(outb, g * 16));
    }
}

struct ActArgs {
    const __nv_bfloat16* Wp; const __nv_bfloat16* feat; const float* taus; const float* cvar; float cvar_scalar;
    float* qmean; int32_t* greedy; int32_t* action; float* debug; long long B;
    unsigned long long seed, step; float eps; int sample;             // sample != 0: taus / epsilon-greedy from Philox(seed, step)
    const mnv_vstep_ctl* ctl;                                         // != nullptr: eps / step from the device control block
    long long* timing;                                                // lab ("act_timing" option): clock64 stamps of CTA 0's phases
};
constexpr int kStamps = 12, kStampTiles = 64;
#ifndef MNV_E1_PIPELINED
#define MNV_E1_PIPELINED 1
#endif
constexpr bool mnv_e1_pipelined = MNV_E1_PIPELINED != 0;

// Random streams of the sampling mode, all from Philox4x32-10 keyed by `seed`, counter = (env, sub-stream, step):
//   sub-stream 0..7: the 32 taus of the environment (4 per draw, torch.rand-style 24-bit uniforms, model.py:149)
//   sub-stream 8:    .x -> the epsilon-greedy coin (greedy iff u > eps, agent.py:200), .y -> the random action (:203)
__device__ __forceinline__ philox::u4 act_draw(const ActArgs& A, unsigned long long step, long long env, unsigned sub)
{
    return philox::philox4x32_10(philox::u4{(uint32_t)env, (uint32_t)((unsigned long long)env >> 32) ^ (sub << 24), (uint32_t)step, (uint32_t)(step >> 32)},
                                  (uint32_t)A.seed, (uint32_t)(A.seed >> 32));
}

// Two groups of 256 threads per CTA, each owning one 128-row tile at a time (its own A buffers, TMEM columns and
// mbarriers): while one group runs an epilogue on the CUDA cores the other group's MMAs occupy the tensor core.
__global__ void __launch_bounds__(kThreads, 1)
iqn_act_tc_kernel(const __grid_constant__ ActArgs A)
{
    const __nv_bfloat16* __restrict__ Wp = A.Wp;
    const float* __restrict__ taus = A.taus; const float* __restrict__ cvar = A.cvar; const float cvar_scalar = A.cvar_scalar;
    float* __restrict__ qmean = A.qmean; int32_t* __restrict__ greedy = A.greedy; float* __restrict__ debug = A.debug;
    const long long B = A.B;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // warp index made warp-uniform FOR THE COMPILER (a shuffle from lane 0, like cutlass::canonical_warp_idx_sync): the
    // MMA-issuing warps address their operands through values derived from it, so the descriptors live in uniform registers
    const int warp_u = __shfl_sync(0xffffffffu, t >> 5, 0);
    const bool issuer = warp_u >= kComputeThreads / 32;       // warps 16, 17: MMA issue for tile group 0, 1 (one elected lane)
    const int g_u = issuer ? warp_u - kComputeThreads / 32 : warp_u >> 3;
    const int g = g_u, tg = issuer ? 256 + lane : (t & 255);  // tile group, thread index inside the group (issuer lanes: >= 256)
    const int half = (warp >> 2) & 1;                        // column half handled by this warp (warps q and q+4 share a lane quadrant)
    GroupSmem& gs = s.g[g];
    GroupSmem& gsu = s.g[g_u];
    if (A.timing != nullptr && tg == 0 && blockIdx.x < 4) A.timing[2 * kStampTiles * kStamps + blockIdx.x * 4 + g] = clock64();        // lab: kernel entry

    // ---- one-time setup: weights to shared memory, TMEM allocation, mbarriers ----
    {
        // layer 1: packed [208][80] (64 weights | bias | 15 zeros) -> shared [208][64] with the bias folded into column 0
        // (cos(pi 0 tau) = 1 for every row): 16-byte chunk (f, kc) keeps its place inside the row group, the row stride shrinks
        const uint4* src = reinterpret_cast<const uint4*>(Wp);
        uint4* dst = reinterpret_cast<uint4*>(s.wc);
        for (int i = t; i < kFeat * (kK0s / 8); i += kThreads) {
            const int f = i / (kK0s / 8), kc = i % (kK0s / 8);
            uint4 v = __ldg(src + tile_offset(f, kc * 8, kK0) / 8);
            if (kc == 0) {
                const __nv_bfloat16 b = Wp[tile_offset(f, kCos, kK0)];
                __nv_bfloat16 w0 = *reinterpret_cast<const __nv_bfloat16*>(&v.x);
                w0 = __float2bfloat16_rn(__bfloat162float(w0) + __bfloat162float(b));
                v.x = (v.x & 0xffff0000u) | (uint32_t)(*reinterpret_cast<const unsigned short*>(&w0));
            }
            dst[tile_offset(f, kc * 8, kK0s) / 8] = v;
        }
        // layers 2-4: contiguous in the packed buffer and in Smem
        const uint4* src2 = reinterpret_cast<const uint4*>(Wp + kWcEl);
        uint4* dst2 = reinterpret_cast<uint4*>(s.w1);
        for (int i = t; i < (kW1El + kW2El + kW3El) / 8; i += kThreads) dst2[i] = __ldg(src2 + i);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s.tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tg == 32) { mbar_init(&gs.bar_a, 1); mbar_init(&gs.bar_b, 1); }
    if (t == 64) {
        s.dyn_eps = A.ctl != nullptr ? A.ctl->act_eps : A.eps;
        s.dyn_step = A.ctl != nullptr ? A.ctl->act_step : A.step;
    }
    fence_async_smem();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base + (uint32_t)g * 256u;           // this group's 256 TMEM columns
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, s.tmem_base, 0) + (uint32_t)g_u * 256u;
    const uint32_t a0_u = smem_u32(gsu.a0), a1_u = smem_u32(gsu.a1);
    const uint32_t wc_u = smem_u32(s.wc), w1_u = smem_u32(s.w1), w2_u = smem_u32(s.w2), w3_u = smem_u32(s.w3);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;     // this warp's TMEM lane quadrant
    const int row = (warp & 3) * 32 + lane;                           // TMEM lane == tile row of this thread
    uint32_t phase_a = 0, phase_b = 0;

    const long long n_tiles = (B + kEnvsPerTile - 1) / kEnvsPerTile;
    const long long tile_step = (long long)gridDim.x * 2;
    long long tile = (long long)blockIdx.x * 2 + g;

    // inputs of a tile, prefetched one tile ahead: every thread carries tau of ITS row (from the caller's tensor, or drawn
    // from the Philox stream; both column halves of a row compute the same value), threads 128..231 one 16-byte chunk of
    // the tile's 4 x 208 bf16 encoder features
    auto load_tau = [&](long long tl) -> float {
        if (tl >= n_tiles) return 0.f;
        const long long b = tl * kEnvsPerTile + row / kTaus;
        if (b >= B) return 0.f;
        const int k = row % kTaus;
        float u;
        if (A.sample) {
            const philox::u4 r = act_draw(A, s.dyn_step, b, (unsigned)(k >> 2));
            const uint32_t rk = (k & 3) == 0 ? r.x : ((k & 3) == 1 ? r.y : ((k & 3) == 2 ? r.z : r.w));
            u = philox::u01(rk);
        } else u = taus[b * kTaus + k];
        return u * (cvar != nullptr ? cvar[b] : cvar_scalar);          // model.py:153
    };
    auto load_feat = [&](long long tl) -> uint4 {
        const int i = tg - kRows;                                       // chunk i = (env i / 26, 8 features (i % 26) * 8 ...)
        if (tl >= n_tiles || i < 0 || i >= kEnvsPerTile * (kFeat / 8)) return make_uint4(0u, 0u, 0u, 0u);
        const long long b = tl * kEnvsPerTile + i / (kFeat / 8);
        if (b >= B) return make_uint4(0u, 0u, 0u, 0u);
        return __ldg(reinterpret_cast<const uint4*>(A.feat + b * kFeat) + (i % (kFeat / 8)));
    };
    // A0 = cos(pi i tau), i = 0..63 (model.py:130,155): thread (row, half) fills i in [32 half, 32 half + 32) with the
    // Chebyshev recurrence c_{i+1} = 2 c_1 c_i - c_{i-1} (one FFMA per feature), started from cos.approx of the argument
    // reduced to [-pi, pi].  The values are rounded to bf16 (2^-9 relative) right away; the recurrence's error over 32
    // steps stays below 1e-4.  Column 0 is exactly 1: it carries layer 1's bias (folded into the weight tile).
    float a0_two_c1 = 0.f, a0_cm = 0.f, a0_c = 1.f;          // recurrence state between the chunks of one row
    auto a0_begin = [&](float tau) {
        auto cospi = [](float x) { x = x - 2.f * rintf(0.5f * x); return __cosf(3.14159265358979f * x); };
        const float c1 = cospi(tau);
        a0_two_c1 = 2.f * c1;
        a0_cm = half == 0 ? c1 : cospi(31.f * tau);           // c_{i-1} at i = 32 half  (c_{-1} = c_1)
        a0_c = half == 0 ? 1.f : cospi(32.f * tau);           // c_i
    };
    auto a0_chunk = [&](int kc) {                             // 8 features: one 16-byte chunk of the operand tile
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = a0_c;
            const float cn = fmaf(a0_two_c1, a0_c, -a0_cm);
            a0_cm = a0_c; a0_c = cn;
        }
        store_chunk(gs.a0, row, half * 4 + kc, kK0s, v);
    };
    auto produce_a0 = [&](float tau) {
        a0_begin(tau);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) a0_chunk(kc);
    };
    // layer 1 of a tile as two MMAs: features [0, 192) -> D1a, [192, 208) -> D1b
    auto issue_l1a = [&]() { issue_layer<kK0s, kN1a>(a0_u, wc_u, tmem_u + kD1a, nullptr); };
    auto issue_l1b = [&]() { issue_layer<kK0s, kN1b>(a0_u, wc_u + 2u * tile_offset(kN1a, 0, kK0s), tmem_u + kD1b, nullptr); };
    // one block of W (32 | 16 | 8) D1 columns starting at feature n0: relu(D1) * feat in bf16 -> A1 chunks (model.py:177-180)
    auto e1_block = [&](auto width, int n0, const __nv_bfloat16* feat) {
        constexpr int W = decltype(width)::value;
        float v[W];
        const uint32_t taddr = tmem + lane_base + (n0 < kN1a ? kD1a + (uint32_t)n0 : kD1b + (uint32_t)(n0 - kN1a));
        if constexpr (W == 32) tmem_ld32(taddr, v); else if constexpr (W == 16) tmem_ld16(taddr, v); else tmem_ld8(taddr, v);
        if (debug != nullptr && tile == (long long)blockIdx.x * 2 + g && blockIdx.x == 0 && g == 0)
            for (int j = 0; j < W; ++j) debug[row * kFeat + n0 + j] = v[j];
#pragma unroll
        for (int q = 0; q < W / 8; ++q) store_relu_chunk(gs.a1, row, (n0 >> 3) + q, kK1, v + q * 8, feat + n0 + q * 8);
    };

    // epilogue 1, software-pipelined: the thread's 104 D1 columns as blocks of 16 (8) columns through two register sets; the
    // TMEM load of block k + 1 is in flight while block k is converted (relu, x feat, bf16) and stored
    auto e1_taddr = [&](int n0) { return tmem + lane_base + (n0 < kN1a ? kD1a + (uint32_t)n0 : kD1b + (uint32_t)(n0 - kN1a)); };
    auto e1_store = [&](const uint32_t (&r)[16], int n0, int width, const __nv_bfloat16* feat) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (debug != nullptr && tile == (long long)blockIdx.x * 2 + g && blockIdx.x == 0 && g == 0)
            for (int j = 0; j < width; ++j) debug[row * kFeat + n0 + j] = v[j];
        store_relu_chunk(gs.a1, row, n0 >> 3, kK1, v, feat + n0);
        if (width == 16) store_relu_chunk(gs.a1, row, (n0 >> 3) + 1, kK1, v + 8, feat + n0 + 8);
    };
    auto epilogue1 = [&](const __nv_bfloat16* feat) {
        // half 0: features [0, 104) = 6 x 16 + 8; half 1: [104, 208) = 5 x 16 + 8 (-> 192: the D1a | D1b boundary) + 16
        uint32_t ra[16], rb[16];
        const int base = half * 104;
        tmem_ld16_async(e1_taddr(base), ra);
        tmem_wait_ld(ra);
        tmem_ld16_async(e1_taddr(base + 16), rb);
        e1_store(ra, base, 16, feat);
        tmem_wait_ld(rb);
        tmem_ld16_async(e1_taddr(base + 32), ra);
        e1_store(rb, base + 16, 16, feat);
        tmem_wait_ld(ra);
        tmem_ld16_async(e1_taddr(base + 48), rb);
        e1_store(ra, base + 32, 16, feat);
        tmem_wait_ld(rb);
        tmem_ld16_async(e1_taddr(base + 64), ra);
        e1_store(rb, base + 48, 16, feat);
        tmem_wait_ld(ra);
        if (half == 0) {
            tmem_ld16_async(e1_taddr(80), rb);
            e1_store(ra, 64, 16, feat);
            tmem_wait_ld(rb);
            tmem_ld8_async(e1_taddr(96), ra);
            e1_store(rb, 80, 16, feat);
            tmem_wait_ld(ra);
            e1_store(ra, 96, 8, feat);
        } else {
            tmem_ld8_async(e1_taddr(184), rb);
            e1_store(ra, 168, 16, feat);
            tmem_wait_ld(rb);
            tmem_ld16_async(e1_taddr(192), ra);
            e1_store(rb, 184, 8, feat);
            tmem_wait_ld(ra);
            e1_store(ra, 192, 16, feat);
        }
    };

    float n_tau = load_tau(tile);
    uint4 n_feat = load_feat(tile);
    // ================= MMA-issuing warp of the group: waits at the hand-off points, one elected lane issues =================
    // The tensor core takes an MMA from a thread only when the previous one has started, so the issuing thread is held for
    // the whole execution time of a layer (~60 cycles per N = 64 K-step); a warp of its own keeps that off the compute warps.
    if (issuer) {
        const bool lead = elect_one();
        if (tile < n_tiles) {
            handoff_wait(g_u, 1);                            // A0 of the first tile
            if (lead) { tc_fence_after(); issue_l1a(); issue_l1b(); umma_commit(&gsu.bar_a); }
            __syncwarp();
        }
        for (; tile < n_tiles; tile += tile_step) {
            const bool has_next = tile + tile_step < n_tiles;
            handoff_wait(g_u, 0);                            // A1 ready, D1 drained
            if (lead) { tc_fence_after(); issue_layer<kK1, kHid>(a1_u, w1_u, tmem_u + kD2, &gsu.bar_b); }
            __syncwarp();
            if (has_next) {
                handoff_wait(g_u, 1);                        // A0 of the next tile ready (layer 1 of this tile completed long ago)
                if (lead) { tc_fence_after(); issue_l1a(); umma_commit(&gsu.bar_a); }
                __syncwarp();
            }
            handoff_wait(g_u, 2);                            // A2 ready, D2 drained
            if (lead) { tc_fence_after(); issue_layer<kK2, kHid>(a1_u, w2_u, tmem_u + kD3, &gsu.bar_b); }
            __syncwarp();
            handoff_wait(g_u, 3);                            // A3 ready, D3 drained
            if (lead) {
                tc_fence_after();
                issue_layer<kK3, kN4>(a1_u, w3_u, tmem_u + kD4, nullptr);
                if (has_next) issue_l1b();                   // its 16 TMEM columns [0, 16) are free only now
                umma_commit(&gsu.bar_b);
            }
            __syncwarp();
        }
    } else {
    // ================= compute warps =================
    if (tile < n_tiles) {                                     // prologue: A0 of this group's first tile
        produce_a0(n_tau);
        fence_async_smem();
        tc_fence_before();
        handoff_arrive(g, 1);
        n_tau = load_tau(tile + tile_step);
    }

    for (; tile < n_tiles; tile += tile_step) {
        const long long env0 = tile * kEnvsPerTile;
        const bool has_next = tile + tile_step < n_tiles;    // group-uniform
        const bool dbg = debug != nullptr && blockIdx.x == 0 && g == 0 && tile == 0;
        const long long it = (tile - ((long long)blockIdx.x * 2 + g)) / tile_step;
        long long* stamps = (A.timing != nullptr && blockIdx.x == 0 && tg == 0 && it < kStampTiles) ? A.timing + (g * kStampTiles + it) * kStamps : nullptr;
        auto stamp = [&](int k) { if (stamps != nullptr) stamps[k] = clock64(); };
        stamp(0);
        if (tg >= kRows && tg < kRows + kEnvsPerTile * (kFeat / 8)) reinterpret_cast<uint4*>(gs.feat)[tg - kRows] = n_feat;
        n_feat = load_feat(tile + tile_step);                // prefetch the next tile's features: consumed one iteration later
        group_sync(g);

        // ---- D1 = A0 . Wc'^T (issued one tile ahead) -> epilogue 1: this warp's 104 features [104 half, 104 half + 104) ----
        mbar_wait(&gs.bar_a, phase_a);                       // every thread of the group sleeps on the MMA's mbarrier
        phase_a ^= 1;
        tc_fence_after();
        stamp(1);
        {
            const __nv_bfloat16* feat = gs.feat + (row / kTaus) * kFeat;
            if (mnv_e1_pipelined) {
                epilogue1(feat);
                if (half == 1) store_bias_step(gs.a1, row, kFeat / 8, kK1);
            } else if (half == 0) {
                e1_block(std::integral_constant<int, 32>{}, 0, feat); e1_block(std::integral_constant<int, 32>{}, 32, feat);
                e1_block(std::integral_constant<int, 32>{}, 64, feat); e1_block(std::integral_constant<int, 8>{}, 96, feat);
            } else {
                e1_block(std::integral_constant<int, 32>{}, 104, feat); e1_block(std::integral_constant<int, 32>{}, 136, feat);
                e1_block(std::integral_constant<int, 16>{}, 168, feat); e1_block(std::integral_constant<int, 8>{}, 184, feat);
                e1_block(std::integral_constant<int, 16>{}, 192, feat);
                store_bias_step(gs.a1, row, kFeat / 8, kK1);
            }
        }
        fence_async_smem();
        tc_fence_before();
        handoff_arrive(g, 0);

        // ---- layer 2: D2[128 x 64] = A1 . W1^T (issuing warp); while it runs: A0 of the next tile, then its layer 1a ----
        stamp(2);
        stamp(3);
        if (has_next) {                                      // (D1a and A0 of this tile were consumed: layer 1 completed above)
            produce_a0(n_tau);
            fence_async_smem();
            tc_fence_before();
            handoff_arrive(g, 1);
        }
        stamp(4);
        mbar_wait(&gs.bar_b, phase_b);
        phase_b ^= 1;
        tc_fence_after();
        stamp(5);
        {
            float v[32];
            const int col = half * 32;
            tmem_ld32(tmem + lane_base + kD2 + col, v);
            if (dbg)
                for (int j = 0; j < 32; ++j) debug[kRows * kFeat + row * kHid + col + j] = v[j];
#pragma unroll
            for (int q = 0; q < 4; ++q) store_relu_chunk(gs.a1, row, (col >> 3) + q, kK2, v + q * 8, nullptr);   // A2 over the (consumed) A1
            if (half == 0) store_bias_step(gs.a1, row, kHid / 8, kK2);
        }
        fence_async_smem();
        tc_fence_before();
        handoff_arrive(g, 2);

        // ---- layer 3: D3[128 x 64] = A2 . W2^T (issuing warp) ----
        // the tau of the tile after next (a Philox draw in the sampling mode) is computed HERE, where the compute warps would
        // only wait for layer 3 -- not in the A0 phase above, which is on the tile's critical path
        if (has_next) n_tau = load_tau(tile + 2 * tile_step);
        stamp(6);
        mbar_wait(&gs.bar_b, phase_b);
        phase_b ^= 1;
        tc_fence_after();
        stamp(7);
        {
            float v[32];
            const int col = half * 32;
            tmem_ld32(tmem + lane_base + kD3 + col, v);
            if (dbg)
                for (int j = 0; j < 32; ++j) debug[kRows * (kFeat + kHid) + row * kHid + col + j] = v[j];
#pragma unroll
            for (int q = 0; q < 4; ++q) store_relu_chunk(gs.a1, row, (col >> 3) + q, kK3, v + q * 8, nullptr);   // A3 over the (consumed) A2
            if (half == 0) store_bias_step(gs.a1, row, kHid / 8, kK3);
        }
        fence_async_smem();
        tc_fence_before();
        handoff_arrive(g, 3);

        // ---- output layer: D4[128 x 16] = A3 . W3^T (+ layer 1b of the next tile; issuing warp), then mean over the 32 taus of
        //      each env (one warp) + argmax ----
        // the epsilon-greedy draw of this warp's environment does not depend on D4: drawn while layer 4 runs
        uint32_t eg_coin = 0u, eg_action = 0u;
        if (half == 0 && A.sample && A.action != nullptr) {
            const philox::u4 r = act_draw(A, s.dyn_step, env0 + (warp & 3), 8u);
            eg_coin = r.x; eg_action = r.y;
        }
        stamp(8);
        mbar_wait(&gs.bar_b, phase_b);
        phase_b ^= 1;
        tc_fence_after();
        stamp(9);
        if (half == 0) {
            float q[kN4];
            {
                float v[16];
                tmem_ld16(tmem + lane_base + kD4, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) q[j] = v[j];
            }
            if (dbg)
                for (int j = 0; j < kN4; ++j) debug[kRows * (kFeat + 2 * kHid) + row * kN4 + j] = q[j];
            // sums over the 32 rows (= lanes) of the 16 columns: transposed butterfly -- at every step a lane keeps one half of its
            // columns and hands the other half to its partner (8 + 4 + 2 + 1 + 1 shuffles instead of 9 x 5); lane L ends with
            // column L >> 1, lane 0 then collects the 9 actions (9 shuffles)
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
                const bool up = (lane & (2 * m)) != 0;
#pragma unroll
                for (int j = 0; j < m; ++j) {
                    const float send = up ? q[j] : q[j + m], keep = up ? q[j + m] : q[j];
                    q[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * m);
                }
            }
            const float qsum = (q[0] + __shfl_xor_sync(0xffffffffu, q[0], 1)) * (1.f / kTaus);   // get_qvals: mean over taus (model.py:190); bias inside D4
#pragma unroll
            for (int a = 0; a < kAct; ++a) q[a] = __shfl_sync(0xffffffffu, qsum, 2 * a);
            const long long b = env0 + (warp & 3);
            if (lane == 0 && b < B) {
                int best = 0; float bv = q[0];
#pragma unroll
                for (int a = 0; a < kAct; ++a) {
                    if (qmean != nullptr) qmean[b * kAct + a] = q[a];
                    if (q[a] > bv) { bv = q[a]; best = a; }                             // np.argmax: first maximum (agent.py:201)
                }
                if (greedy != nullptr) greedy[b] = best;
                if (A.action != nullptr) {                                                  // agent.py:200-203
                    int act = best;
                    const float eps = s.dyn_eps;
                    if (A.sample && eps > 0.f && !(philox::u01(eg_coin) > eps)) act = (int)__umulhi(eg_action, (uint32_t)kAct);
                    A.action[b] = act;
                }
            }
        }
        tc_fence_before();                                   // (the next hand-off orders these TMEM reads before the MMAs that overwrite them)
        stamp(10);
    }
    }   // compute warps

    // ---- teardown ----
    if (A.timing != nullptr && tg == 0 && blockIdx.x < 4) A.timing[2 * kStampTiles * kStamps + blockIdx.x * 4 + 2 + g] = clock64();    // lab: loop end
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s.tmem_base), "r"(kTmemCols) : "memory");
    }
}

// fp32 parameters -> bf16 K-major core-matrix tiles with the bias column appended:
//   Wc [208][80], W1 [64][224], W2 [64][80], W3 [16][80] (rows >= 9 zero); column K_orig = bias, the other pad columns = 0
__global__ void __launch_bounds__(256) iqn_pack_tc_kernel(const float* __restrict__ P, __nv_bfloat16* __restrict__ W)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    int base, n, k, K, Korig, rows, ow, ob;
    if (i < kWcEl) { base = 0; K = kK0; Korig = kCos; rows = kFeat; ow = oCW; ob = oCB; n = i / K; k = i % K; }
    else if (i < kWcEl + kW1El) { base = kWcEl; K = kK1; Korig = kFeat; rows = kHid; ow = oH1W; ob = oH1B; n = (i - base) / K; k = (i - base) % K; }
    else if (i < kWcEl + kW1El + kW2El) { base = kWcEl + kW1El; K = kK2; Korig = kHid; rows = kHid; ow = oH2W; ob = oH2B; n = (i - base) / K; k = (i - base) % K; }
    else if (i < kPackedTcEl) { base = kWcEl + kW1El + kW2El; K = kK3; Korig = kHid; rows = kAct; ow = oOW; ob = oOB; n = (i - base) / K; k = (i - base) % K; }
    else return;
    float v = 0.f;
    if (n < rows) v = k < Korig ? P[ow + n * Korig + k] : (k == Korig ? P[ob + n] : 0.f);
    W[base + tile_offset(n, k, K)] = __float2bfloat16(v);
}

}  // namespace

extern "C" int iqn_packed_tc_bytes(void) { return kPackedTcEl * 2; }

extern "C" int iqn_pack_tc(const float* d_params, void* d_packed_tc, void* stream)
{
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed_tc);
    iqn_pack_tc_kernel<<<(kPackedTcEl + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_params, (__nv_bfloat16*)d_packed_tc);
    return mnv_launch_status("iqn_pack_tc");
}

extern "C" int64_t iqn_act_scratch_bytes(int64_t B) { return B > 0 ? ((B + kEncEnvs - 1) / kEncEnvs) * kEncEnvs * kFeat * 2 : 0; }

namespace {

int launch_act(const float* d_params, const void* d_packed_tc, const float* d_obs, float* d_cvar_adaptive, void* d_scratch, ActArgs A,
               cudaStream_t st, const char* what)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static unsigned long long attr_mask = 0;                       // the attribute is per device
    if (dev >= 64 || !((attr_mask >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(iqn_act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) { mnv_set_error("cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(e)); return (int)e; }
        if (dev < 64) attr_mask |= 1ull << dev;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long B = A.B;
    iqn_encode_kernel<<<(unsigned)((B + kEncEnvs - 1) / kEncEnvs), kEncThreads, 0, st>>>(d_params, d_obs, (__nv_bfloat16*)d_scratch, d_cvar_adaptive, B);
    A.Wp = (const __nv_bfloat16*)d_packed_tc; A.feat = (const __nv_bfloat16*)d_scratch;
    const long long n_tiles = (B + kEnvsPerTile - 1) / kEnvsPerTile;
    const long long pairs = (n_tiles + 1) / 2;                     // two tile groups per CTA
    const int grid = (int)(pairs < sms ? pairs : sms);
    iqn_act_tc_kernel<<<grid, kThreads, sizeof(Smem), st>>>(A);
    return mnv_launch_status(what);
}

}  // namespace

extern "C" int iqn_act_tc(const float* d_params, const void* d_packed_tc, const float* d_obs, const float* d_taus,
                          const float* d_cvar, float cvar_scalar, float* d_qmean, int32_t* d_greedy, float* d_debug,
                          void* d_scratch, int64_t B, int32_t n_tau, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_act_tc: B must be > 0"); return MNV_E_SIZE; }
    if (n_tau != kTaus) { mnv_set_error("iqn_act_tc: n_tau must be 32 (ObsEncoder.K), got %d", n_tau); return MNV_E_CAPACITY; }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed_tc); MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_taus); MNV_CHECK_PTR(d_scratch);
    if (d_qmean == nullptr && d_greedy == nullptr) { mnv_set_error("iqn_act_tc: no output"); return MNV_E_NULL; }
    ActArgs A{};
    A.taus = d_taus; A.cvar = d_cvar; A.cvar_scalar = cvar_scalar; A.qmean = d_qmean; A.greedy = d_greedy; A.debug = d_debug; A.B = B;
    if (mnv_option(MNV_OPT_ACT_TIMING) && d_debug != nullptr) { A.timing = (long long*)d_debug; A.debug = nullptr; }   // lab: phase stamps instead of accumulators
    return launch_act(d_params, d_packed_tc, d_obs, nullptr, d_scratch, A, (cudaStream_t)stream, "iqn_act_tc");
}

extern "C" int iqn_act_tc_sample(const float* d_params, const void* d_packed_tc, const float* d_obs, int32_t adaptive_cvar,
                                 float* d_cvar, float cvar_scalar, float eps, uint64_t seed, uint64_t step,
                                 int32_t* d_action, int32_t* d_greedy, float* d_qmean, void* d_scratch, int64_t B, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_act_tc_sample: B must be > 0"); return MNV_E_SIZE; }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed_tc); MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_scratch);
    if (d_action == nullptr) { mnv_set_error("iqn_act_tc_sample: null action output"); return MNV_E_NULL; }
    if (adaptive_cvar && d_cvar == nullptr) { mnv_set_error("iqn_act_tc_sample: adaptive CVaR needs the d_cvar buffer"); return MNV_E_NULL; }
    ActArgs A{};
    A.cvar = adaptive_cvar ? d_cvar : nullptr; A.cvar_scalar = cvar_scalar; A.qmean = d_qmean; A.greedy = d_greedy; A.action = d_action; A.B = B;
    A.seed = seed; A.step = step; A.eps = eps; A.sample = 1;
    return launch_act(d_params, d_packed_tc, d_obs, adaptive_cvar ? d_cvar : nullptr, d_scratch, A, (cudaStream_t)stream, "iqn_act_tc_sample");
}

extern "C" int iqn_act_tc_sample_ctl(const float* d_params, const void* d_packed_tc, const float* d_obs, int32_t adaptive_cvar,
                                     float* d_cvar, float cvar_scalar, uint64_t seed, int32_t* d_action, int32_t* d_greedy,
                                     float* d_qmean, void* d_scratch, int64_t B, const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_act_tc_sample_ctl: B must be > 0"); return MNV_E_SIZE; }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed_tc); MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_scratch);
    if (d_ctl == nullptr) { mnv_set_error("iqn_act_tc_sample_ctl: null control block"); return MNV_E_NULL; }
    if (d_action == nullptr) { mnv_set_error("iqn_act_tc_sample_ctl: null action output"); return MNV_E_NULL; }
    if (adaptive_cvar && d_cvar == nullptr) { mnv_set_error("iqn_act_tc_sample_ctl: adaptive CVaR needs the d_cvar buffer"); return MNV_E_NULL; }
    ActArgs A{};
    A.cvar = adaptive_cvar ? d_cvar : nullptr; A.cvar_scalar = cvar_scalar; A.qmean = d_qmean; A.greedy = d_greedy; A.action = d_action; A.B = B;
    A.seed = seed; A.sample = 1; A.ctl = d_ctl;
    return launch_act(d_params, d_packed_tc, d_obs, adaptive_cvar ? d_cvar : nullptr, d_scratch, A, (cudaStream_t)stream, "iqn_act_tc_sample_ctl");
}

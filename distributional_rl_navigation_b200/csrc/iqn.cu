// IQN (ObsEncoder) forward / act / train kernels in fp32 for sm_100a -- the parity path of
//   thirdparty/IQN/model.py:111-191 (ObsEncoder.calc_cos / forward / get_qvals)
//   thirdparty/IQN/agent.py:269-304,401-407 (IQNAgent.train, calculate_huber_loss), clip_grad_norm_ 0.5 + Adam (agent.py:66,299-300)
//
// Work decomposition: one CTA (256 threads) owns a tile of 64 "rows" = (sample, quantile) pairs (8 samples x 8 taus
// when training, 2 samples x 32 taus when acting).  All activations of the tile live in shared memory; each weight
// matrix is staged into shared memory once per layer.  The training kernel runs target forward -> local forward ->
// pairwise quantile Huber loss (warp-shuffle reduction) -> full backward inside the SAME CTA and writes one partial
// gradient per tile; iqn_reduce sums the partials in a fixed order (deterministic), iqn_clip_adam applies the global-norm
// clip + Adam.
//
// GEMM core: warp-level tensor-core MMAs (mma.sync.m16n8k8, TF32 operands, FP32 accumulation) with the error-compensated
// 3xTF32 split -- every fp32 operand x is used as x_hi + x_lo (x_hi = the top 19 bits of x, x_lo = x - x_hi, exact) and a
// product is accumulated as a_lo b_hi + a_hi b_lo + a_hi b_hi.  What is dropped is a_lo b_lo and the truncation of the lo
// parts, both <= 2^-20 relative per product: the results stay at fp32 level (loss within ~1e-6 of the PyTorch fp32 path,
// tests/test_iqn_parity.py; north_star asks for 1e-4), while a warp issues ~4x fewer instructions than the register-tiled
// FFMA loop it replaces (which ran at 28 % of the FMA pipe: shared-memory-operand bound).  Plain TF32 / bf16 operands
// (one MMA per product) would NOT meet the loss tolerance -- those are for the acting path (iqn_act_tc.cu), where only
// the argmax is consumed.
#include <math.h>

#include <cuda_bf16.h>

#include "iqn_common.cuh"

namespace {

using namespace iqn;

constexpr int kThreads = 256;     // 8 warps per 64-row tile (the warp tiling of MmaAcc assumes exactly 8)
constexpr int R = 64;              // rows per tile
constexpr int LD64 = 68;           // padded leading dimension of [64][64] activation tiles (== 4 mod 32, multiple of 4)
constexpr int LD208 = 212;         // padded leading dimension of [64][208] activation tiles

struct Smem {
    float cos[R * LD64];
    float c[R * LD208];            // relu(cos_embedding(cos)); the backward pass overwrites it with dzc
    float h1[R * LD64];            // ... overwritten with dz1
    float h2[R * LD64];            // ... overwritten with dz2
    float w[2 * 7488];             // weight stage: two halves (reduction rows [0,RED/2) and [RED/2,RED), rows padded by 8), filled by cp.async
    float w3[kAct * LD64 + 12];    // output layer of the network being evaluated: rows padded to LD64 (conflict-free 16-byte loads), then the bias
    float dfp[128];                // small scratch (mean over taus of the acting forward)
    float feat[8 * kFeat];
    float dfeat[8 * kFeat];
    float x[8 * 28];
    float tau[R];
    float q[R * 12];
    float T[R], E[R], dE[R];
    float red[32];
    int   act[8];
};

// ---- 3xTF32 tensor-core GEMM:  out(m, n) = sum_{r < RED} X(r, m) * Y(r, n),  m in [0, M), n in [0, N) ----------------
// xf(r, m) / yf(r, n) return single fp32 elements (shared memory).  Fragment layout of mma.m16n8k8 (g = lane >> 2,
// t = lane & 3):  A (16 x 8, rows = m): a0 (g, t) a1 (g + 8, t) a2 (g, t + 4) a3 (g + 8, t + 4);  B (8 x 8, cols = n):
// b0 (t, g) b1 (t + 4, g);  C (16 x 8): c0 (g, 2t) c1 (g, 2t + 1) c2 (g + 8, 2t) c3 (g + 8, 2t + 1).
// Warp tiling over the CTA's 8 warps:
//   M == 64  ("row split"):  warp w owns the 16-row block w & 3 and the N / 16 column blocks of half w >> 2
//                            (A fragment loaded once per 8 reduction rows, reused for every column block);
//   M == 208 ("col split"):  warp w owns the 8-column block w (N == 64) and all 13 row blocks (B fragment reused).
__device__ __forceinline__ void split3(float v, uint32_t& hi, uint32_t& lo)
{
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;                // v rounded to the 19 bits a TF32 operand keeps
    lo = __float_as_uint(v - __uint_as_float(hi));                    // exact; the MMA truncates it to TF32 itself
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1,
                                     uint32_t bl0, uint32_t bl1)
{
    mma_tf32(c, al, bh0, bh1);                                        // small terms first
    mma_tf32(c, ah, bl0, bl1);
    mma_tf32(c, ah, bh0, bh1);
}

template <int M, int N>
struct MmaAcc {
    static constexpr bool kRowSplit = (M == 64);
    static_assert((kRowSplit && N % 16 == 0) || (M == 208 && N == 64), "supported tile shapes");
    static constexpr int kBlocks = kRowSplit ? N / 16 : M / 16;       // accumulator blocks per warp
    float acc[kBlocks][4];
    int m0, n0, g, t;                                                 // first row / column of this warp's blocks
    __device__ __forceinline__ MmaAcc()
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        g = lane >> 2; t = lane & 3;
        m0 = kRowSplit ? (w & 3) * 16 : 0;
        n0 = kRowSplit ? (w >> 2) * (N / 2) : w * 8;
#pragma unroll
        for (int i = 0; i < kBlocks; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
    }
    // accumulate reduction rows [r0, r1) (a multiple of 8); yf sees the row index relative to yrow0.
    // The tensor core adds into its FP32 accumulator with truncation: chaining all RED / 8 x 3 MMAs on one accumulator
    // biases the sum by ~n 2^-24 (measured 3e-6 on the loss).  So the MMAs of TWO reduction steps (16 rows, 6 MMAs) go
    // into a zeroed scratch accumulator, which is then added to the running sum with round-to-nearest FADDs: the error
    // returns to the level of the FFMA kernel (~2e-7) for 4 extra FADDs per block and 16 rows.
    template <class XF, class YF>
    __device__ __forceinline__ void run(int r0, int r1, int yrow0, XF xf, YF yf)
    {
        float tmp[kBlocks][4];
        auto zero = [&]() {
#pragma unroll
            for (int i = 0; i < kBlocks; ++i) { tmp[i][0] = 0.f; tmp[i][1] = 0.f; tmp[i][2] = 0.f; tmp[i][3] = 0.f; }
        };
        auto flush = [&]() {
#pragma unroll
            for (int i = 0; i < kBlocks; ++i) { acc[i][0] += tmp[i][0]; acc[i][1] += tmp[i][1]; acc[i][2] += tmp[i][2]; acc[i][3] += tmp[i][3]; }
        };
        zero();
        int pending = 0;
#pragma unroll 1
        for (int k0 = r0; k0 < r1; k0 += 8) {
            if constexpr (kRowSplit) {
                uint32_t ah[4], al[4];
                split3(xf(k0 + t, m0 + g), ah[0], al[0]); split3(xf(k0 + t, m0 + g + 8), ah[1], al[1]);
                split3(xf(k0 + t + 4, m0 + g), ah[2], al[2]); split3(xf(k0 + t + 4, m0 + g + 8), ah[3], al[3]);
#pragma unroll
                for (int i = 0; i < kBlocks; ++i) {
                    uint32_t bh0, bl0, bh1, bl1;
                    split3(yf(k0 + t - yrow0, n0 + i * 8 + g), bh0, bl0);
                    split3(yf(k0 + t + 4 - yrow0, n0 + i * 8 + g), bh1, bl1);
                    mma3(tmp[i], ah, al, bh0, bh1, bl0, bl1);
                }
            } else {
                uint32_t bh0, bl0, bh1, bl1;
                split3(yf(k0 + t - yrow0, n0 + g), bh0, bl0);
                split3(yf(k0 + t + 4 - yrow0, n0 + g), bh1, bl1);
#pragma unroll
                for (int i = 0; i < kBlocks; ++i) {
                    uint32_t ah[4], al[4];
                    const int mb = i * 16;
                    split3(xf(k0 + t, mb + g), ah[0], al[0]); split3(xf(k0 + t, mb + g + 8), ah[1], al[1]);
                    split3(xf(k0 + t + 4, mb + g), ah[2], al[2]); split3(xf(k0 + t + 4, mb + g + 8), ah[3], al[3]);
                    mma3(tmp[i], ah, al, bh0, bh1, bl0, bl1);
                }
            }
            if (++pending == 2) { flush(); zero(); pending = 0; }
        }
        if (pending) flush();
    }
    // epi4(m, n, c0, c1, c2, c3): (m, n) (m, n + 1) (m + 8, n) (m + 8, n + 1); every lane calls it for every block (uniform)
    template <class Epi4>
    __device__ __forceinline__ void finish4(Epi4 epi4)
    {
#pragma unroll
        for (int i = 0; i < kBlocks; ++i) {
            const int m = (kRowSplit ? m0 : i * 16) + g, n = (kRowSplit ? n0 + i * 8 : n0) + 2 * t;
            epi4(m, n, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
    }
    template <class Epi>
    __device__ __forceinline__ void finish(Epi epi)
    {
        finish4([&](int m, int n, float c0, float c1, float c2, float c3) { epi(m, n, c0); epi(m, n + 1, c1); epi(m + 8, n, c2); epi(m + 8, n + 1, c3); });
    }
};

template <int RED, int M, int N, class XF, class YF, class Epi>
__device__ __forceinline__ void tile_mm(XF xf, YF yf, Epi epi)
{
    static_assert(RED % 8 == 0, "reduction length");
    MmaAcc<M, N> t;
    t.run(0, RED, 0, xf, yf);
    t.finish(epi);
}

// out(m, n) -> dst[m * ld + n] in global memory with 8-byte stores (an accumulator block holds the column pairs (n, n + 1),
// n even; dst and ld even -> aligned): half the store instructions of the scalar epilogue
template <int RED, int M, int N, class XF, class YF>
__device__ __forceinline__ void tile_mm_store2(XF xf, YF yf, float* __restrict__ dst, int ld)
{
    static_assert(RED % 8 == 0, "reduction length");
    MmaAcc<M, N> t;
    t.run(0, RED, 0, xf, yf);
    t.finish4([&](int m, int n, float c0, float c1, float c2, float c3) {
        *reinterpret_cast<float2*>(dst + m * ld + n) = make_float2(c0, c1);
        *reinterpret_cast<float2*>(dst + (m + 8) * ld + n) = make_float2(c2, c3);
    });
}

// ---- weight staging: global [RED][N] fp32 -> shared memory rows of N + 8 floats (the pad makes the B-fragment loads of a
// warp -- 4 reduction rows x 8 columns -- hit 32 different banks) with cp.async, in two halves of the reduction dimension.
// While a GEMM consumes half 0 its half 1 is in flight, and while it consumes half 1 the NEXT GEMM's half 0 is in flight,
// so only the very first half of a kernel is an exposed L2 round trip.
constexpr int kStagePad = 8;
constexpr int kHalfStage = 104 * (64 + kStagePad);                 // floats per half buffer: the largest half, W1^T [104][64 + 8]

__device__ __forceinline__ void stage_async(float* dst, const float* __restrict__ src, int rows, int n)
{
    const int chunks_per_row = n >> 2, ld = n + kStagePad;          // n % 4 == 0; rows are 16-byte aligned in both spaces
    for (int i = threadIdx.x; i < rows * chunks_per_row; i += kThreads) {
        const int r = i / chunks_per_row, c = i - r * chunks_per_row;
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + r * ld + c * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + r * n + c * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// request half 0 of a [RED][N] matrix into buffer A (= s.w)
__device__ __forceinline__ void request_first_half(Smem& s, const float* __restrict__ W, int red, int n)
{
    stage_async(s.w, W, red / 2, n);
}

// out = xf x W with W [RED][N] in global memory whose half 0 has already been requested into s.w; after half 1 landed the
// first half of `next` ([next_red][next_n], may be null) is requested.  Ends with the epilogue, no trailing barrier.
template <int RED, int M, int N, class XF, class Fin>
__device__ __forceinline__ void tile_mm_staged(Smem& s, const float* __restrict__ W, const float* __restrict__ next, int next_red,
                                               int next_n, XF xf, Fin fin)
{
    static_assert(RED % 16 == 0 && (RED / 2) * (N + kStagePad) <= kHalfStage, "stage halves");
    float* A = s.w;
    float* B = s.w + kHalfStage;
    constexpr int LDW = N + kStagePad;
    MmaAcc<M, N> t;
    stage_wait();
    __syncthreads();                                               // half 0 visible; nobody still reads buffer B
    stage_async(B, W + (RED / 2) * N, RED / 2, N);
    t.run(0, RED / 2, 0, xf, [&](int r, int n) { return A[r * LDW + n]; });
    stage_wait();
    __syncthreads();                                               // half 1 visible; nobody still reads buffer A
    if (next != nullptr) stage_async(A, next, next_red / 2, next_n);
    t.run(RED / 2, RED, RED / 2, xf, [&](int r, int n) { return B[r * LDW + n]; });
    fin(t);
}

// ObsEncoder.forward (model.py:160-186) for one 64-row tile; row r belongs to sample r / n_tau of the tile.
// In: s.x (8 x 28, zero padded), s.tau (already multiplied by cvar); half 0 of PT's WcT ALREADY requested into s.w.
// Out: s.feat, s.cos, s.c, s.h1, s.h2, s.q.  P: flat parameters (torch layout); PT: packed transposes (iqn_pack).
// After the last staged GEMM the first half of `next` ([next_red][next_n]) is requested (the following pass's first matrix).
__device__ __noinline__ void forward_tile(Smem& s, const float* __restrict__ P, const float* __restrict__ PT, int n_tau,
                             const float* __restrict__ next, int next_red, int next_n)
{
    const int t = threadIdx.x;
    __syncthreads();
    // observation encoders, no activation (model.py:169-172)
    if (t < kFeat) {                                               // thread = feature: its weights are loaded ONCE (independent loads, one round trip)
        const int f = t;
        if (f < 32) {                                              // velocity (f < 16) / goal encoder: two inputs
            const int j = f & 15, ow = f < 16 ? oVW : oGW, ob = f < 16 ? oVB : oGB, xi = f < 16 ? 0 : 2;
            const float w0 = __ldg(P + ow + j * 2), w1 = __ldg(P + ow + j * 2 + 1), b = __ldg(P + ob + j);
#pragma unroll
            for (int smp = 0; smp < 8; ++smp) s.feat[smp * kFeat + f] = fmaf(w1, s.x[smp * 28 + xi + 1], fmaf(w0, s.x[smp * 28 + xi], b));
        } else {
            const int g = f - 32;
            float w[22];
#pragma unroll
            for (int k = 0; k < 22; ++k) w[k] = __ldg(P + oSW + g * 22 + k);
            const float b = __ldg(P + oSB + g);
#pragma unroll
            for (int smp = 0; smp < 8; ++smp) {
                float v = b;
#pragma unroll
                for (int k = 0; k < 22; ++k) v = fmaf(w[k], s.x[smp * 28 + 4 + k], v);
                s.feat[smp * kFeat + f] = v;
            }
        }
    }
    // cos(tau * pi*i): fp32 product with fp32(pi*i), then fp32 cos (model.py:130,155)
    for (int idx = t; idx < R * kCos; idx += kThreads) {
        const int r = idx / kCos, i = idx % kCos;
        const float pis = (float)(MNV_PI_D * (double)i);
        s.cos[r * LD64 + i] = cosf(s.tau[r] * pis);
    }
    for (int idx = t; idx < kAct * kHid + kAct; idx += kThreads) {   // output layer of this network -> shared memory
        const float v = __ldg(P + oOW + idx);                      // (output_layer.weight and .bias are contiguous)
        if (idx < kAct * kHid) s.w3[(idx / kHid) * LD64 + (idx % kHid)] = v;
        else s.w3[kAct * LD64 + (idx - kAct * kHid)] = v;
    }
    // c = relu(cos_embedding(cos))  (model.py:177)      [the staged GEMM starts with a barrier: cos / feat are visible]
    tile_mm_staged<kCos, R, kFeat>(s, PT + ptWc, PT + ptW1, kFeat, kHid,
                                   [&](int k, int row) { return s.cos[row * LD64 + k]; },
                                   [&](auto& acc) { acc.finish([&](int row, int f, float a) { s.c[row * LD208 + f] = fmaxf(a + __ldg(P + oCB + f), 0.f); }); });
    // h1 = relu(hidden_layer(feat * c))  (model.py:180-182)
    tile_mm_staged<kFeat, R, kHid>(s, PT + ptW1, PT + ptW2, kHid, kHid,
                                   [&](int k, int row) { return s.c[row * LD208 + k] * s.feat[(row / n_tau) * kFeat + k]; },
                                   [&](auto& acc) { acc.finish([&](int row, int o, float a) { s.h1[row * LD64 + o] = fmaxf(a + __ldg(P + oH1B + o), 0.f); }); });
    // h2 = relu(hidden_layer_2(h1))  (model.py:183)
    tile_mm_staged<kHid, R, kHid>(s, PT + ptW2, next, next_red, next_n,
                                  [&](int k, int row) { return s.h1[row * LD64 + k]; },
                                  [&](auto& acc) { acc.finish([&](int row, int o, float a) { s.h2[row * LD64 + o] = fmaxf(a + __ldg(P + oH2B + o), 0.f); }); });
    __syncthreads();
    // q = output_layer(h2)  (model.py:184)
    for (int idx = t; idx < R * kAct; idx += kThreads) {
        const int row = idx / kAct, a = idx % kAct;
        float v = s.w3[kAct * LD64 + a];
        const float4* h4 = reinterpret_cast<const float4*>(s.h2 + row * LD64);
        const float4* w4 = reinterpret_cast<const float4*>(s.w3 + a * LD64);
#pragma unroll 4
        for (int k4 = 0; k4 < kHid / 4; ++k4) {                    // same summation order as a scalar k loop
            const float4 h = h4[k4], w = w4[k4];
            v = fmaf(h.x, w.x, v); v = fmaf(h.y, w.y, v); v = fmaf(h.z, w.z, v); v = fmaf(h.w, w.w, v);
        }
        s.q[row * 12 + a] = v;
    }
    __syncthreads();
}

__device__ __forceinline__ void load_inputs(Smem& s, const float* __restrict__ obs, const float* __restrict__ taus,
                                            const float* __restrict__ cvar, float cvar_scalar, long long B, int n_tau,
                                            long long tile)
{
    const int t = threadIdx.x;
    const int S = R / n_tau;
    const long long s0 = tile * S;
    __syncthreads();
    for (int idx = t; idx < 8 * 28; idx += kThreads) {
        const int smp = idx / 28, k = idx % 28;
        const long long b = s0 + smp;
        s.x[idx] = (smp < S && b < B && k < kObs) ? obs[b * kObs + k] : 0.f;
    }
    if (t < R) {
        const long long b = s0 + t / n_tau;
        float v = 0.f;
        if (b < B) {
            const float cv = cvar != nullptr ? cvar[b] : cvar_scalar;
            v = taus[b * n_tau + t % n_tau] * cv;                       // distorted quantile sampling (model.py:153)
        }
        s.tau[t] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// forward / act
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
iqn_forward_kernel(const float* __restrict__ P, const float* __restrict__ PT, const float* __restrict__ obs,
                   const float* __restrict__ taus, const float* __restrict__ cvar, float cvar_scalar,
                   float* __restrict__ quantiles, float* __restrict__ qmean, int32_t* __restrict__ greedy,
                   long long B, int n_tau)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const long long tile = blockIdx.x;
    const int S = R / n_tau, t = threadIdx.x;
    request_first_half(s, PT + ptWc, kCos, kFeat);
    load_inputs(s, obs, taus, cvar, cvar_scalar, B, n_tau, tile);
    forward_tile(s, P, PT, n_tau, nullptr, 0, 0);
    const long long s0 = tile * S;
    if (quantiles != nullptr) {
        for (int idx = t; idx < R * kAct; idx += kThreads) {
            const int row = idx / kAct, a = idx % kAct;
            const long long b = s0 + row / n_tau;
            if (b < B) quantiles[(b * n_tau + row % n_tau) * kAct + a] = s.q[row * 12 + a];
        }
    }
    if (qmean != nullptr || greedy != nullptr) {
        // get_qvals: mean over the quantile samples (model.py:188-191), fixed summation order
        if (t < S * kAct) {
            const int smp = t / kAct, a = t % kAct;
            float acc = 0.f;
            for (int n = 0; n < n_tau; ++n) acc += s.q[(smp * n_tau + n) * 12 + a];
            s.dfp[t] = acc / (float)n_tau;
        }
        __syncthreads();
        if (t < S) {
            const long long b = s0 + t;
            if (b < B) {
                int best = 0; float bv = s.dfp[t * kAct];
                for (int a = 0; a < kAct; ++a) {
                    const float v = s.dfp[t * kAct + a];
                    if (qmean != nullptr) qmean[b * kAct + a] = v;
                    if (v > bv) { bv = v; best = a; }                     // np.argmax: first maximum (agent.py:201)
                }
                if (greedy != nullptr) greedy[b] = best;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// train: loss + gradient of one tile (8 samples x 8 taus)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
iqn_train_kernel(const float* __restrict__ PL, const float* __restrict__ PTL, const float* __restrict__ PTG,
                 const float* __restrict__ PTTG, const float* __restrict__ states, const long long* __restrict__ actions,
                 const float* __restrict__ rewards, const float* __restrict__ next_states, const float* __restrict__ dones,
                 const float* __restrict__ taus_t, const float* __restrict__ taus_l, float gamma_n,
                 float* __restrict__ gpart_all, float* __restrict__ loss_part, long long B)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a dependent launch (iqn_update_tail) may be placed early; it waits for this grid's completion itself
    const int t = threadIdx.x;
    const long long tile = blockIdx.x, s0 = tile * 8;
    float* __restrict__ g = gpart_all + tile * (long long)kPartStride;
    constexpr int NT8 = kTrainTaus;

    // ---- target network on next_states (taus drawn first: Q9) -> T_j = r + gamma^n (1 - done) max_a Q'(s', tau_j) ----
    request_first_half(s, PTTG + ptWc, kCos, kFeat);
    load_inputs(s, next_states, taus_t, nullptr, 1.f, B, NT8, tile);
    forward_tile(s, PTG, PTTG, NT8, PTL + ptWc, kCos, kFeat);          // ... then prefetch the local network's first matrix
    if (t < R) {
        const long long b = s0 + t / NT8;
        float tv = 0.f;
        if (b < B) {
            float mx = s.q[t * 12];
#pragma unroll
            for (int a = 1; a < kAct; ++a) mx = fmaxf(mx, s.q[t * 12 + a]);          // agent.py:280
            tv = rewards[b] + gamma_n * mx * (1.f - dones[b]);                        // agent.py:283
        }
        s.T[t] = tv;
    }
    if (t < 8) { const long long b = s0 + t; s.act[t] = b < B ? (int)actions[b] : 0; }

    // ---- local network on states ----
    load_inputs(s, states, taus_l, nullptr, 1.f, B, NT8, tile);
    forward_tile(s, PL, PTL, NT8, PL + oH2W, kHid, kHid);              // ... then prefetch W2 (torch layout) for the backward pass

    // ---- pairwise quantile Huber loss (agent.py:289-295) and dL/dE ----
    float li = 0.f;
    if (t < R) {
        const int smp = t / NT8;
        const bool valid = (s0 + smp) < B;
        const float e = s.q[t * 12 + s.act[smp]];                                    // gather, agent.py:286
        const float tau = s.tau[t];
        float gsum = 0.f;
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            const float td = s.T[smp * NT8 + j] - e;                                 // td[b,i,j] = T_j - E_i
            const float ad = fabsf(td);
            const float hub = ad <= 1.f ? 0.5f * td * td : ad - 0.5f;                // agent.py:401-407 (k = 1)
            const float w = fabsf(tau - (td < 0.f ? 1.f : 0.f));                     // agent.py:292
            li = fmaf(w, hub, li);
            gsum = fmaf(w, fminf(fmaxf(td, -1.f), 1.f), gsum);
        }
        const float inv = 1.f / (8.f * (float)B);                                    // mean_j (1/8), mean_b (1/B)
        s.dE[t] = valid ? -gsum * inv : 0.f;
        li = valid ? li * inv : 0.f;
    }
    // warp-shuffle reduction of the loss over the tile
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) li += __shfl_xor_sync(0xffffffffu, li, off);
    if ((t & 31) == 0) s.red[t >> 5] = li;
    __syncthreads();
    if (t == 0) loss_part[tile] = s.red[0] + s.red[1];

    // ================= backward =================
    // output layer: only the taken action's row sees a gradient.  Per sample first: v[smp][o] = sum over its 8 rows of
    // dE[r] h2[r][o] (and the sum of dE for the bias) into s.q (free after the loss), then row a of dW3 adds the samples that
    // took action a in sample order (fixed order; no divergent 64-row loops)
    for (int idx = t; idx < 8 * (kHid + 1); idx += kThreads) {
        const int smp = idx / (kHid + 1), o = idx % (kHid + 1);
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            const int r = smp * NT8 + j;
            acc = o < kHid ? fmaf(s.dE[r], s.h2[r * LD64 + o], acc) : acc + s.dE[r];
        }
        s.q[smp * (kHid + 1) + o] = acc;                           // 8 x 65 floats <= R * 12
    }
    __syncthreads();
    for (int idx = t; idx < kAct * kHid + kAct; idx += kThreads) {
        const int a = idx < kAct * kHid ? idx / kHid : idx - kAct * kHid, o = idx < kAct * kHid ? idx % kHid : kHid;
        float acc = 0.f;
#pragma unroll
        for (int smp = 0; smp < 8; ++smp)
            if (s.act[smp] == a) acc += s.q[smp * (kHid + 1) + o];
        g[oOW + idx] = acc;                                        // (output_layer.weight and .bias are contiguous: oOB = oOW + 576)
    }
    __syncthreads();
    // dz2 = dE * W3[a,:] * (h2 > 0), in place over h2
    for (int idx = t; idx < R * kHid; idx += kThreads) {
        const int r = idx / kHid, o = idx % kHid;
        const float h = s.h2[r * LD64 + o];
        s.h2[r * LD64 + o] = h > 0.f ? s.dE[r] * __ldg(PL + oOW + s.act[r / NT8] * kHid + o) : 0.f;
    }
    __syncthreads();
    // dW2[o][k] = sum_r dz2[r][o] h1[r][k] ; db2
    tile_mm_store2<R, kHid, kHid>([&](int r, int o) { return s.h2[r * LD64 + o]; }, [&](int r, int k) { return s.h1[r * LD64 + k]; },
                                  g + oH2W, kHid);
    if (t < kHid) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += s.h2[r * LD64 + t];
        g[oH2B + t] = acc;
    }
    __syncthreads();
    // dz1 = (dz2 W2) * (h1 > 0), in place over h1  (W2 in torch layout [o][k], prefetched during the forward pass;
    // W1 in torch layout is prefetched for dh0 meanwhile)
    tile_mm_staged<kHid, R, kHid>(s, PL + oH2W, PL + oH1W, kHid, kFeat,
                                  [&](int o, int r) { return s.h2[r * LD64 + o]; },
                                  [&](auto& acc) { acc.finish([&](int r, int k, float a) { float& h = s.h1[r * LD64 + k]; h = h > 0.f ? a : 0.f; }); });
    __syncthreads();
    // dW1[o][k] = sum_r dz1[r][o] h0[r][k], h0 = feat * c ; db1
    tile_mm_store2<R, kHid, kFeat>([&](int r, int o) { return s.h1[r * LD64 + o]; },
                                   [&](int r, int k) { return s.c[r * LD208 + k] * s.feat[(r / NT8) * kFeat + k]; },
                                   g + oH1W, kFeat);
    if (t < kHid) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += s.h1[r * LD64 + t];
        g[oH1B + t] = acc;
    }
    __syncthreads();
    // dh0 = dz1 W1 ; dzc = dh0 * feat * (c > 0) in place over c ; d(feat)[sample] = sum over the sample's 8 rows of dh0 * c.
    // In the accumulator layout a warp's 16-row block is exactly two samples (rows g and g + 8 of lane group g), so the sum
    // over a sample's rows is a butterfly over the lane bits 2..4 (fixed order: deterministic) and lane group 0 owns the result.
    tile_mm_staged<kHid, R, kFeat>(s, PL + oH1W, nullptr, 0, 0, [&](int o, int r) { return s.h1[r * LD64 + o]; },
                                   [&](auto& acc) {
                                       acc.finish4([&](int r, int k, float a0, float a1, float a2, float a3) {
                                           const int smp = r >> 3;                       // r = 16 mb + g, g < 8: sample 2 mb; r + 8: sample 2 mb + 1
                                           float* c_top = s.c + r * LD208 + k;
                                           float* c_bot = s.c + (r + 8) * LD208 + k;
                                           const float ct0 = c_top[0], ct1 = c_top[1], cb0 = c_bot[0], cb1 = c_bot[1];
                                           float d0 = a0 * ct0, d1 = a1 * ct1, d2 = a2 * cb0, d3 = a3 * cb1;
#pragma unroll
                                           for (int off = 4; off < 32; off <<= 1) {
                                               d0 += __shfl_xor_sync(0xffffffffu, d0, off); d1 += __shfl_xor_sync(0xffffffffu, d1, off);
                                               d2 += __shfl_xor_sync(0xffffffffu, d2, off); d3 += __shfl_xor_sync(0xffffffffu, d3, off);
                                           }
                                           if ((r & 7) == 0) {
                                               s.dfeat[smp * kFeat + k] = d0; s.dfeat[smp * kFeat + k + 1] = d1;
                                               s.dfeat[(smp + 1) * kFeat + k] = d2; s.dfeat[(smp + 1) * kFeat + k + 1] = d3;
                                           }
                                           const float* f_top = s.feat + smp * kFeat + k;
                                           const float* f_bot = s.feat + (smp + 1) * kFeat + k;
                                           c_top[0] = ct0 > 0.f ? a0 * f_top[0] : 0.f; c_top[1] = ct1 > 0.f ? a1 * f_top[1] : 0.f;
                                           c_bot[0] = cb0 > 0.f ? a2 * f_bot[0] : 0.f; c_bot[1] = cb1 > 0.f ? a3 * f_bot[1] : 0.f;
                                       });
                                   });
    __syncthreads();
    // dWc[f][i] = sum_r dzc[r][f] cos[r][i] ; dbc
    tile_mm_store2<R, kFeat, kCos>([&](int r, int f) { return s.c[r * LD208 + f]; }, [&](int r, int i) { return s.cos[r * LD64 + i]; },
                                   g + oCW, kCos);
    if (t < kFeat) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += s.c[r * LD208 + t];
        g[oCB + t] = acc;
    }
    __syncthreads();
    // encoders: dW[f][k] = sum_s dfeat[s][f] x[s][k]
    for (int idx = t; idx < oCW; idx += kThreads) {
        float acc = 0.f;
        int f, k;                       // feature index, input index (-1 = bias)
        if (idx < oVB) { f = idx / 2; k = idx % 2; }
        else if (idx < oGW) { f = idx - oVB; k = -1; }
        else if (idx < oGB) { f = 16 + (idx - oGW) / 2; k = 2 + (idx - oGW) % 2; }
        else if (idx < oSW) { f = 16 + idx - oGB; k = -1; }
        else if (idx < oSB) { f = 32 + (idx - oSW) / 22; k = 4 + (idx - oSW) % 22; }
        else { f = 32 + idx - oSB; k = -1; }
#pragma unroll
        for (int smp = 0; smp < 8; ++smp) acc = fmaf(s.dfeat[smp * kFeat + f], k < 0 ? 1.f : s.x[smp * 28 + k], acc);
        g[idx] = acc;
    }
}

// grad[i] = sum over tiles (fixed order) ; loss = sum of the tile partials
__global__ void __launch_bounds__(256)
iqn_reduce_kernel(const float* __restrict__ gpart, const float* __restrict__ loss_part, int n_tiles,
                  float* __restrict__ grad, float* __restrict__ loss)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < kParams) {
        // four independent partial sums (tiles = 0,1,2,3 mod 4) keep several loads in flight; fixed order -> deterministic
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* p = gpart + i;
        int tile = 0;
        for (; tile + 4 <= n_tiles; tile += 4, p += 4ll * kPartStride) {
            a0 += p[0]; a1 += p[kPartStride]; a2 += p[2ll * kPartStride]; a3 += p[3ll * kPartStride];
        }
        for (; tile < n_tiles; ++tile, p += kPartStride) a0 += p[0];
        grad[i] = (a0 + a1) + (a2 + a3);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float acc = 0.f;
        for (int tile = 0; tile < n_tiles; ++tile) acc += loss_part[tile];
        *loss = acc;
    }
}

// clip_grad_norm_(params, max_norm) + Adam.step (torch defaults), every CTA recomputes the same global norm
__global__ void __launch_bounds__(1024)
iqn_clip_adam_kernel(float* __restrict__ P, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
                     float grad_scale, float max_norm, float step_size, float beta1, float beta2, float inv_sqrt_bc2, float eps,
                     float* __restrict__ grad_norm, float* __restrict__ PT, __nv_bfloat16* __restrict__ Wtc)
{
    __shared__ float red[32];
    const int t = threadIdx.x;
    float ss = 0.f;
    for (int i = t; i < kParams; i += 1024) { const float gi = grad[i] * grad_scale; ss = fmaf(gi, gi, ss); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((t & 31) == 0) red[t >> 5] = ss;
    __syncthreads();
    if (t < 32) {
        float x = red[t];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        if (t == 0) red[0] = x;
    }
    __syncthreads();
    const float total = sqrtf(red[0]);
    const float coef = fminf(max_norm / (total + 1e-6f), 1.f);                      // torch.nn.utils.clip_grad_norm_
    if (blockIdx.x == 0 && t == 0 && grad_norm != nullptr) *grad_norm = total;
    const int i = blockIdx.x * 1024 + t;
    if (i < kParams) {
        const float gi = grad[i] * grad_scale * coef;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;                         // torch.optim.Adam (no amsgrad / decay)
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        const float pn = P[i] - step_size * (mi / denom);
        P[i] = pn;
        // keep the kernel-side copies of the parameters current: fp32 transposes (iqn_pack layout) and the bf16 tensor-core
        // tiles (iqn_pack_tc layout; their zero / one padding never changes)
        int pt, tc;
        packed_slots(i, pt, tc);
        if (PT != nullptr && pt >= 0) PT[pt] = pn;
        if (Wtc != nullptr && tc >= 0) Wtc[tc] = __float2bfloat16(pn);
    }
}

__global__ void __launch_bounds__(256) iqn_pack_kernel(const float* __restrict__ P, float* __restrict__ PT)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= kPacked) return;
    if (i < ptW1) { const int k = i / kFeat, f = i % kFeat; PT[i] = P[oCW + f * kCos + k]; }               // WcT [64][208]
    else if (i < ptW2) { const int j = i - ptW1, k = j / kHid, o = j % kHid; PT[i] = P[oH1W + o * kFeat + k]; }   // W1T [208][64]
    else { const int j = i - ptW2, k = j / kHid, o = j % kHid; PT[i] = P[oH2W + o * kHid + k]; }            // W2T [64][64]
}

template <class K>
int set_smem(K kern)
{
    static unsigned long long done_mask = 0;      // per kernel instantiation, one bit per device (the attribute is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && ((done_mask >> dev) & 1ull)) return 0;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e != cudaSuccess) { mnv_set_error("cudaFuncSetAttribute(iqn): %s", cudaGetErrorString(e)); return (int)e; }
    if (dev < 64) done_mask |= 1ull << dev;
    return 0;
}

}  // namespace

extern "C" int iqn_param_count(void) { return iqn::kParams; }
extern "C" int iqn_packed_count(void) { return iqn::kPacked; }

extern "C" int64_t iqn_train_scratch_floats(int64_t B)
{
    const int64_t tiles = (B + 7) / 8;
    return tiles * (iqn::kPartStride + 1);
}

extern "C" int iqn_pack(const float* d_params, float* d_packed, void* stream)
{
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed);
    iqn_pack_kernel<<<(kPacked + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_params, d_packed);
    return mnv_launch_status("iqn_pack");
}

extern "C" int iqn_forward(const float* d_params, const float* d_packed, const float* d_obs, const float* d_taus,
                           const float* d_cvar, float cvar_scalar, float* d_quantiles, float* d_qmean, int32_t* d_greedy,
                           int64_t B, int32_t n_tau, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_forward: B must be > 0"); return MNV_E_SIZE; }
    if (!(n_tau == 8 || n_tau == 16 || n_tau == 32 || n_tau == 64)) {
        mnv_set_error("iqn_forward: n_tau=%d not in {8,16,32,64}", n_tau); return MNV_E_CAPACITY;
    }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_packed); MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_taus);
    if (d_quantiles == nullptr && d_qmean == nullptr && d_greedy == nullptr) { mnv_set_error("iqn_forward: no output"); return MNV_E_NULL; }
    int rc = set_smem(iqn_forward_kernel);
    if (rc) return rc;
    const int S = R / n_tau;
    const long long tiles = (B + S - 1) / S;
    iqn_forward_kernel<<<(unsigned)tiles, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(
        d_params, d_packed, d_obs, d_taus, d_cvar, cvar_scalar, d_quantiles, d_qmean, d_greedy, B, n_tau);
    return mnv_launch_status("iqn_forward");
}

namespace {

// the fused per-tile kernel alone: tile partials of the gradient + loss into d_scratch
int launch_train(const float* d_params_local, const float* d_packed_local, const float* d_params_target,
                 const float* d_packed_target, const float* d_states, const int64_t* d_actions,
                 const float* d_rewards, const float* d_next_states, const float* d_dones,
                 const float* d_taus_target, const float* d_taus_local, float gamma_n, float* d_scratch, int64_t B, void* stream,
                 const char* what)
{
    if (B <= 0) { mnv_set_error("%s: B must be > 0", what); return MNV_E_SIZE; }
    MNV_CHECK_PTR(d_params_local); MNV_CHECK_PTR(d_packed_local); MNV_CHECK_PTR(d_params_target); MNV_CHECK_PTR(d_packed_target);
    MNV_CHECK_PTR(d_states); MNV_CHECK_PTR(d_actions); MNV_CHECK_PTR(d_rewards); MNV_CHECK_PTR(d_next_states); MNV_CHECK_PTR(d_dones);
    MNV_CHECK_PTR(d_taus_target); MNV_CHECK_PTR(d_taus_local); MNV_CHECK_PTR(d_scratch);
    int rc = set_smem(iqn_train_kernel);
    if (rc) return rc;
    const long long tiles = (B + 7) / 8;
    float* gpart = d_scratch;
    float* lpart = d_scratch + tiles * (long long)kPartStride;
    iqn_train_kernel<<<(unsigned)tiles, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(
        d_params_local, d_packed_local, d_params_target, d_packed_target, d_states, (const long long*)d_actions, d_rewards,
        d_next_states, d_dones, d_taus_target, d_taus_local, gamma_n, gpart, lpart, B);
    return mnv_launch_status(what);
}

}  // namespace

extern "C" int iqn_loss_partials(const float* d_params_local, const float* d_packed_local, const float* d_params_target,
                                 const float* d_packed_target, const float* d_states, const int64_t* d_actions,
                                 const float* d_rewards, const float* d_next_states, const float* d_dones,
                                 const float* d_taus_target, const float* d_taus_local, float gamma_n,
                                 float* d_scratch, int64_t B, void* stream)
{
    return launch_train(d_params_local, d_packed_local, d_params_target, d_packed_target, d_states, d_actions, d_rewards, d_next_states,
                        d_dones, d_taus_target, d_taus_local, gamma_n, d_scratch, B, stream, "iqn_loss_partials");
}

extern "C" int iqn_loss_grad(const float* d_params_local, const float* d_packed_local, const float* d_params_target,
                             const float* d_packed_target, const float* d_states, const int64_t* d_actions,
                             const float* d_rewards, const float* d_next_states, const float* d_dones,
                             const float* d_taus_target, const float* d_taus_local, float gamma_n,
                             float* d_scratch, float* d_loss, float* d_grad, int64_t B, void* stream)
{
    MNV_CHECK_PTR(d_grad);
    if (d_loss == nullptr) { mnv_set_error("iqn_loss_grad: null loss"); return MNV_E_NULL; }
    int rc = launch_train(d_params_local, d_packed_local, d_params_target, d_packed_target, d_states, d_actions, d_rewards, d_next_states,
                          d_dones, d_taus_target, d_taus_local, gamma_n, d_scratch, B, stream, "iqn_loss_grad(train)");
    if (rc) return rc;
    const long long tiles = (B + 7) / 8;
    iqn_reduce_kernel<<<(kParams + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_scratch, d_scratch + tiles * (long long)kPartStride, (int)tiles,
                                                                                 d_grad, d_loss);
    return mnv_launch_status("iqn_loss_grad(reduce)");
}

extern "C" int iqn_clip_adam(float* d_params, const float* d_grad, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                             float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                             float* d_grad_norm, void* stream)
{
    if (step < 1) { mnv_set_error("iqn_clip_adam: step must be >= 1"); return MNV_E_PARAM; }
    MNV_CHECK_PTR(d_params); MNV_CHECK_PTR(d_grad); MNV_CHECK_PTR(d_m); MNV_CHECK_PTR(d_v);
    MNV_CHECK_PTR_OPT(d_packed); MNV_CHECK_PTR_OPT(d_packed_tc);
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    iqn_clip_adam_kernel<<<(kParams + 1023) / 1024, 1024, 0, (cudaStream_t)stream>>>(
        d_params, d_grad, d_m, d_v, grad_scale, max_norm, step_size, beta1, beta2, inv_sqrt_bc2, eps, d_grad_norm, d_packed,
        (__nv_bfloat16*)d_packed_tc);
    return mnv_launch_status("iqn_clip_adam");
}

// IQN (ObsEncoder) forward / act / train kernels in fp32 for sm_100a -- the parity path of
//   thirdparty/IQN/model.py:111-191 (ObsEncoder.calc_cos / forward / get_qvals)
//   thirdparty/IQN/agent.py:269-304,401-407 (IQNAgent.train, calculate_huber_loss), clip_grad_norm_ 0.5 + Adam (agent.py:66,299-300)
//
// Work decomposition: one CTA (256 threads) owns a tile of 64 "rows" = (sample, quantile) pairs (8 samples x 8 taus
// when training, 2 samples x 32 taus when acting).  All activations of the tile live in shared memory; each weight
// matrix is staged into shared memory once per layer and consumed by a register-tiled rank-1-update loop
// (4x13 / 4x4 / 13x4 outputs per thread).  The training kernel runs target forward -> local forward -> pairwise quantile
// Huber loss (warp-shuffle reduction) -> full backward inside the SAME CTA and writes one partial gradient per tile;
// iqn_reduce sums the partials in a fixed order (deterministic), iqn_clip_adam applies the global-norm clip + Adam.
// FP32 FFMA on purpose: north_star asks for the loss within 1e-4 of the PyTorch fp32 path, which bf16/tf32 tensor-core
// operands do not give; tensor cores are for the acting path where only the argmax is consumed.
#include <math.h>

#include <cuda_bf16.h>

#include "iqn_common.cuh"

namespace {

using namespace iqn;

constexpr int kThreads = 256;     // 8 warps per 64-row tile, thread tiles 4x13 / 4x4 / 13x4 (512 threads with 2x13 / 2x4 tiles measured slower: 145 vs 135 us per update, shared-memory-load bound)
constexpr int kMT = 64 * 16 / kThreads;   // rows per thread tile: (64 / kMT) * 16 column groups == kThreads
constexpr int kGroups = 64 / kMT;  // row groups of a tile (d(feat) partial sums)
constexpr int R = 64;              // rows per tile
constexpr int LD64 = 68;           // padded leading dimension of [64][64] activation tiles (== 4 mod 32, multiple of 4)
constexpr int LD208 = 212;         // padded leading dimension of [64][208] activation tiles

struct Smem {
    float cos[R * LD64];
    float c[R * LD208];            // relu(cos_embedding(cos)); the backward pass overwrites it with dzc
    float h1[R * LD64];            // ... overwritten with dz1
    float h2[R * LD64];            // ... overwritten with dz2
    float w[13312];                // weight stage: two halves (reduction rows [0,RED/2) and [RED/2,RED)), filled by cp.async
    float w3[kAct * kHid + 12];    // output layer weights + bias of the network being evaluated
    float dfp[kGroups * kFeat];    // per row-group partial sums of d(feat)
    float feat[8 * kFeat];
    float dfeat[8 * kFeat];
    float x[8 * 28];
    float tau[R];
    float q[R * 12];
    float T[R], E[R], dE[R];
    float red[32];
    int   act[8];
};

// out(m, n) = sum_{r < RED} xf(r, m) * Y(r, n)     m in [0, M), n in [0, N)
// thread tile MT x NT with (M/MT)*(N/NT) == 256 threads.  yf(r, n0, yv) loads NT values of row r starting at column n0.
template <int M, int N, int MT, int NT>
struct TileAcc {
    static_assert((M / MT) * (N / NT) == kThreads && M % MT == 0 && N % NT == 0, "tile shape");
    float acc[MT][NT];
    int m0, n0;
    __device__ __forceinline__ TileAcc()
    {
        const int tn = threadIdx.x % (N / NT), tm = threadIdx.x / (N / NT);
        m0 = tm * MT; n0 = tn * NT;
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
    }
    // accumulate reduction rows [r0, r1); yf sees the row index relative to yrow0
    template <class XF, class YF>
    __device__ __forceinline__ void run(int r0, int r1, int yrow0, XF xf, YF yf)
    {
#pragma unroll 2
        for (int r = r0; r < r1; ++r) {
            float xv[MT], yv[NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) xv[i] = xf(r, m0 + i);
            yf(r - yrow0, n0, yv);
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
        }
    }
    template <class Epi>
    __device__ __forceinline__ void finish(Epi epi)
    {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) epi(m0 + i, n0 + j, acc[i][j]);
    }
};

template <int RED, int M, int N, int MT, int NT, class XF, class YF, class Epi>
__device__ __forceinline__ void tile_mm(XF xf, YF yf, Epi epi)
{
    TileAcc<M, N, MT, NT> t;
    t.run(0, RED, 0, xf, yf);
    t.finish(epi);
}

// Y loaders for a shared-memory matrix [RED][ld]
template <int NT>
struct YMat {
    const float* p; int ld;
    __device__ __forceinline__ void operator()(int r, int n0, float (&yv)[NT]) const
    {
        if constexpr (NT % 4 == 0) {
#pragma unroll
            for (int j = 0; j < NT; j += 4) {
                const float4 v = *reinterpret_cast<const float4*>(p + r * ld + n0 + j);
                yv[j] = v.x; yv[j + 1] = v.y; yv[j + 2] = v.z; yv[j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NT; ++j) yv[j] = p[r * ld + n0 + j];
        }
    }
};

// ---- weight staging: global [RED][N] fp32 -> shared memory with cp.async, in two halves of the reduction dimension.
// While a GEMM consumes half 0 its half 1 is in flight, and while it consumes half 1 the NEXT GEMM's half 0 is in flight,
// so only the very first half of a kernel is an exposed L2 round trip.
constexpr int kHalfStage = 13312 / 2;                              // floats per half buffer (largest matrix / 2)

__device__ __forceinline__ void stage_async(float* dst, const float* __restrict__ src, int n)
{
    for (int i = threadIdx.x * 4; i < n; i += kThreads * 4) {       // n % 4 == 0, both 16-byte aligned
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// request half 0 of a [RED][N] matrix into buffer A (= s.w)
__device__ __forceinline__ void request_first_half(Smem& s, const float* __restrict__ W, int red, int n)
{
    stage_async(s.w, W, (red / 2) * n);
}

// out = xf x W with W [RED][N] in global memory whose half 0 has already been requested into s.w; after half 1 landed the
// first half of `next` ([next_red][next_n], may be null) is requested.  Ends with the epilogue, no trailing barrier.
template <int RED, int M, int N, int MT, int NT, class XF, class Epi>
__device__ __forceinline__ void tile_mm_staged(Smem& s, const float* __restrict__ W, const float* __restrict__ next, int next_red,
                                               int next_n, XF xf, Epi epi)
{
    static_assert(RED % 2 == 0 && ((RED / 2) * N) % 4 == 0 && (RED / 2) * N <= kHalfStage, "stage halves");
    float* A = s.w;
    float* B = s.w + kHalfStage;
    TileAcc<M, N, MT, NT> t;
    stage_wait();
    __syncthreads();                                               // half 0 visible; nobody still reads buffer B
    stage_async(B, W + (RED / 2) * N, (RED / 2) * N);
    t.run(0, RED / 2, 0, xf, YMat<NT>{A, N});
    stage_wait();
    __syncthreads();                                               // half 1 visible; nobody still reads buffer A
    if (next != nullptr) stage_async(A, next, (next_red / 2) * next_n);
    t.run(RED / 2, RED, RED / 2, xf, YMat<NT>{B, N});
    t.finish(epi);
}

// ObsEncoder.forward (model.py:160-186) for one 64-row tile; row r belongs to sample r / n_tau of the tile.
// In: s.x (8 x 28, zero padded), s.tau (already multiplied by cvar); half 0 of PT's WcT ALREADY requested into s.w.
// Out: s.feat, s.cos, s.c, s.h1, s.h2, s.q.  P: flat parameters (torch layout); PT: packed transposes (iqn_pack).
// After the last staged GEMM the first half of `next` ([next_red][next_n]) is requested (the following pass's first matrix).
__device__ void forward_tile(Smem& s, const float* __restrict__ P, const float* __restrict__ PT, int n_tau,
                             const float* __restrict__ next, int next_red, int next_n)
{
    const int t = threadIdx.x;
    __syncthreads();
    // observation encoders, no activation (model.py:169-172)
    for (int idx = t; idx < 8 * kFeat; idx += kThreads) {
        const int smp = idx / kFeat, f = idx % kFeat;
        const float* x = s.x + smp * 28;
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
    // cos(tau * pi*i): fp32 product with fp32(pi*i), then fp32 cos (model.py:130,155)
    for (int idx = t; idx < R * kCos; idx += kThreads) {
        const int r = idx / kCos, i = idx % kCos;
        const float pis = (float)(MNV_PI_D * (double)i);
        s.cos[r * LD64 + i] = cosf(s.tau[r] * pis);
    }
    for (int idx = t; idx < kAct * kHid + kAct; idx += kThreads)     // output layer of this network -> shared memory
        s.w3[idx] = __ldg(P + oOW + idx);                          // (output_layer.weight and .bias are contiguous)
    // c = relu(cos_embedding(cos))  (model.py:177)      [the staged GEMM starts with a barrier: cos / feat are visible]
    tile_mm_staged<kCos, R, kFeat, kMT, 13>(s, PT + ptWc, PT + ptW1, kFeat, kHid,
                                          [&](int k, int row) { return s.cos[row * LD64 + k]; },
                                          [&](int row, int f, float a) { s.c[row * LD208 + f] = fmaxf(a + __ldg(P + oCB + f), 0.f); });
    // h1 = relu(hidden_layer(feat * c))  (model.py:180-182)
    tile_mm_staged<kFeat, R, kHid, kMT, 4>(s, PT + ptW1, PT + ptW2, kHid, kHid,
                                         [&](int k, int row) { return s.c[row * LD208 + k] * s.feat[(row / n_tau) * kFeat + k]; },
                                         [&](int row, int o, float a) { s.h1[row * LD64 + o] = fmaxf(a + __ldg(P + oH1B + o), 0.f); });
    // h2 = relu(hidden_layer_2(h1))  (model.py:183)
    tile_mm_staged<kHid, R, kHid, kMT, 4>(s, PT + ptW2, next, next_red, next_n,
                                        [&](int k, int row) { return s.h1[row * LD64 + k]; },
                                        [&](int row, int o, float a) { s.h2[row * LD64 + o] = fmaxf(a + __ldg(P + oH2B + o), 0.f); });
    __syncthreads();
    // q = output_layer(h2)  (model.py:184)
    for (int idx = t; idx < R * kAct; idx += kThreads) {
        const int row = idx / kAct, a = idx % kAct;
        float v = s.w3[kAct * kHid + a];
#pragma unroll 8
        for (int k = 0; k < kHid; ++k) v = fmaf(s.h2[row * LD64 + k], s.w3[a * kHid + k], v);
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
    const int t = threadIdx.x;
    const long long tile = blockIdx.x, s0 = tile * 8;
    float* __restrict__ g = gpart_all + tile * (long long)kParams;
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
    for (int idx = t; idx < kGroups * kFeat; idx += kThreads) s.dfp[idx] = 0.f;
    __syncthreads();
    if (t == 0) loss_part[tile] = s.red[0] + s.red[1];

    // ================= backward =================
    // output layer: only the taken action's row sees a gradient
    for (int idx = t; idx < kAct * kHid + kAct; idx += kThreads) {
        float acc = 0.f;
        if (idx < kAct * kHid) {
            const int a = idx / kHid, o = idx % kHid;
            for (int r = 0; r < R; ++r)
                if (s.act[r / NT8] == a) acc = fmaf(s.dE[r], s.h2[r * LD64 + o], acc);
            g[oOW + idx] = acc;
        } else {
            const int a = idx - kAct * kHid;
            for (int r = 0; r < R; ++r)
                if (s.act[r / NT8] == a) acc += s.dE[r];
            g[oOB + a] = acc;
        }
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
    tile_mm<R, kHid, kHid, kMT, 4>([&](int r, int o) { return s.h2[r * LD64 + o]; }, YMat<4>{s.h1, LD64},
                                 [&](int o, int k, float a) { g[oH2W + o * kHid + k] = a; });
    if (t < kHid) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += s.h2[r * LD64 + t];
        g[oH2B + t] = acc;
    }
    __syncthreads();
    // dz1 = (dz2 W2) * (h1 > 0), in place over h1  (W2 in torch layout [o][k], prefetched during the forward pass;
    // W1 in torch layout is prefetched for dh0 meanwhile)
    tile_mm_staged<kHid, R, kHid, kMT, 4>(s, PL + oH2W, PL + oH1W, kHid, kFeat,
                                        [&](int o, int r) { return s.h2[r * LD64 + o]; },
                                        [&](int r, int k, float a) { float& h = s.h1[r * LD64 + k]; h = h > 0.f ? a : 0.f; });
    __syncthreads();
    // dW1[o][k] = sum_r dz1[r][o] h0[r][k], h0 = feat * c ; db1
    tile_mm<R, kHid, kFeat, kMT, 13>([&](int r, int o) { return s.h1[r * LD64 + o]; },
                                   [&](int r, int n0, float (&yv)[13]) {
#pragma unroll
                                       for (int j = 0; j < 13; ++j)
                                           yv[j] = s.c[r * LD208 + n0 + j] * s.feat[(r / NT8) * kFeat + n0 + j];
                                   },
                                   [&](int o, int k, float a) { g[oH1W + o * kFeat + k] = a; });
    if (t < kHid) {
        float acc = 0.f;
        for (int r = 0; r < R; ++r) acc += s.h1[r * LD64 + t];
        g[oH1B + t] = acc;
    }
    __syncthreads();
    // dh0 = dz1 W1 ; dzc = dh0 * feat * (c > 0) in place over c ; d(feat) += dh0 * c
    {
        const int tm = threadIdx.x / 16;                                             // row group of kMT rows (one sample = 8 / kMT groups)
        tile_mm_staged<kHid, R, kFeat, kMT, 13>(s, PL + oH1W, nullptr, 0, 0, [&](int o, int r) { return s.h1[r * LD64 + o]; },
                                       [&](int r, int k, float a) {
                                           float& cv = s.c[r * LD208 + k];
                                           const float c0 = cv;
                                           s.dfp[tm * kFeat + k] += a * c0;             // (tm, k) is private to this thread
                                           cv = c0 > 0.f ? a * s.feat[(r / NT8) * kFeat + k] : 0.f;
                                       });
    }
    __syncthreads();
    for (int idx = t; idx < 8 * kFeat; idx += kThreads) {
        const int smp = idx / kFeat, k = idx % kFeat;
        float acc = 0.f;
#pragma unroll
        for (int gq = 0; gq < 8 / kMT; ++gq) acc += s.dfp[((8 / kMT) * smp + gq) * kFeat + k];      // fixed order: deterministic
        s.dfeat[idx] = acc;
    }
    // dWc[f][i] = sum_r dzc[r][f] cos[r][i] ; dbc
    tile_mm<R, kFeat, kCos, 13, kMT>([&](int r, int f) { return s.c[r * LD208 + f]; }, YMat<kMT>{s.cos, LD64},
                                   [&](int f, int i, float a) { g[oCW + f * kCos + i] = a; });
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
        for (; tile + 4 <= n_tiles; tile += 4, p += 4ll * kParams) {
            a0 += p[0]; a1 += p[kParams]; a2 += p[2ll * kParams]; a3 += p[3ll * kParams];
        }
        for (; tile < n_tiles; ++tile, p += kParams) a0 += p[0];
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
    return tiles * (iqn::kParams + 1);
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

extern "C" int iqn_loss_grad(const float* d_params_local, const float* d_packed_local, const float* d_params_target,
                             const float* d_packed_target, const float* d_states, const int64_t* d_actions,
                             const float* d_rewards, const float* d_next_states, const float* d_dones,
                             const float* d_taus_target, const float* d_taus_local, float gamma_n,
                             float* d_scratch, float* d_loss, float* d_grad, int64_t B, void* stream)
{
    if (B <= 0) { mnv_set_error("iqn_loss_grad: B must be > 0"); return MNV_E_SIZE; }
    MNV_CHECK_PTR(d_params_local); MNV_CHECK_PTR(d_packed_local); MNV_CHECK_PTR(d_params_target); MNV_CHECK_PTR(d_packed_target);
    MNV_CHECK_PTR(d_states); MNV_CHECK_PTR(d_actions); MNV_CHECK_PTR(d_rewards); MNV_CHECK_PTR(d_next_states); MNV_CHECK_PTR(d_dones);
    MNV_CHECK_PTR(d_taus_target); MNV_CHECK_PTR(d_taus_local); MNV_CHECK_PTR(d_scratch); MNV_CHECK_PTR(d_grad);
    if (d_loss == nullptr) { mnv_set_error("iqn_loss_grad: null loss"); return MNV_E_NULL; }
    int rc = set_smem(iqn_train_kernel);
    if (rc) return rc;
    const long long tiles = (B + 7) / 8;
    float* gpart = d_scratch;
    float* lpart = d_scratch + tiles * (long long)kParams;
    iqn_train_kernel<<<(unsigned)tiles, kThreads, sizeof(Smem), (cudaStream_t)stream>>>(
        d_params_local, d_packed_local, d_params_target, d_packed_target, d_states, (const long long*)d_actions, d_rewards,
        d_next_states, d_dones, d_taus_target, d_taus_local, gamma_n, gpart, lpart, B);
    rc = mnv_launch_status("iqn_loss_grad(train)");
    if (rc) return rc;
    iqn_reduce_kernel<<<(kParams + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gpart, lpart, (int)tiles, d_grad, d_loss);
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

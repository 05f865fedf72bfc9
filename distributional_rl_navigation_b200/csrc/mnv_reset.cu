// Device-side MarineNavEnv.seed / MarineNavEnv.reset (marinenav_env.py:75-78, 86-197, 344-420) for sm_100a.
//
// One WARP per (finished) environment, each environment with its own numpy-compatible legacy MT19937 stream kept in HBM
// (key u32 [E][624]: one contiguous 2.5 KB record per environment, staged through shared memory; pos i32 [E]).  The draw
// order and the arithmetic (separately rounded * and +, IEEE sqrt and /) follow the reference exactly, so environment e
// reproduces MarineNavEnv(seed=s_e).reset() bit-for-bit: this file MUST be compiled with -fmad=false.
// A persistent grid of warps scans the done flags in contiguous chunks, so a masked launch with a few finished
// environments costs one flag scan plus ~one reset latency.
#include <math.h>

#include "mnv_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kBlock = kWarpsPerCta * 32;
constexpr int kMtN = 624, kMtM = 397;
constexpr double kTwoPi = 2 * MNV_PI;

// One warp owns one environment at a time.  The 624-word MT19937 state of that environment is staged in shared memory
// (coalesced 2.5 KB load / store); every lane then runs the SAME sampling logic redundantly (warp-uniform control flow,
// shared-memory broadcasts), which lets the block regeneration be executed by all 32 lanes in parallel.
struct Mt {
    uint32_t* mt;      // shared memory, this warp's 624 words
    int pos;
    int lane;

    // numpy / reference genrand: in-place regeneration.  Elements 0..226 depend only on old values, 227..453 on the new
    // 0..226, 454..622 on the new 227..395, and 623 on the new 0 and 396 -> three data-parallel sweeps + one element.
    __device__ void regenerate()
    {
        const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAG = 0x9908b0dfu;
        auto sweep = [&](int k0, int k1, int off) {
            for (int base = k0; base < k1; base += 32) {
                const int k = base + lane;
                uint32_t v = 0;
                const bool on = k < k1;
                if (on) {
                    const uint32_t y = (mt[k] & UP) | (mt[k + 1] & LO);
                    v = mt[k + off] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
                }
                __syncwarp();
                if (on) mt[k] = v;
                __syncwarp();
            }
        };
        sweep(0, kMtN - kMtM, kMtM);                       // 0 .. 226   (mt[k + 397], old)
        sweep(kMtN - kMtM, 2 * (kMtN - kMtM), kMtM - kMtN); // 227 .. 453 (mt[k - 227], new)
        sweep(2 * (kMtN - kMtM), kMtN - 1, kMtM - kMtN);    // 454 .. 622
        if (lane == 0) {
            const uint32_t y = (mt[kMtN - 1] & UP) | (mt[0] & LO);
            mt[kMtN - 1] = mt[kMtM - 1] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        __syncwarp();
        pos = 0;
    }
    __device__ uint32_t next32()
    {
        if (pos >= kMtN) regenerate();                       // warp-uniform
        uint32_t y = mt[pos++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    // RandomState.random_sample: 53-bit double
    __device__ double sample()
    {
        const uint32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    // RandomState.uniform(lo, hi) = lo + (hi - lo) * U
    __device__ double uniform(double lo, double hi) { return lo + (hi - lo) * sample(); }
};

__global__ void __launch_bounds__(kBlock) mnv_seed_kernel(uint32_t* key, int32_t* pos, const uint32_t* seeds, long long E)
{
    __shared__ uint32_t s_mt[kWarpsPerCta][kMtN];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long n_warps = (long long)gridDim.x * kWarpsPerCta;
    for (long long e = (long long)blockIdx.x * kWarpsPerCta + w; e < E; e += n_warps) {
        if (lane == 0) {                                     // init_genrand: a serial recurrence
            uint32_t prev = seeds[e];
            s_mt[w][0] = prev;
            for (int i = 1; i < kMtN; ++i) {
                prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
                s_mt[w][i] = prev;
            }
            pos[e] = kMtN;
        }
        __syncwarp();
        for (int i = lane; i < kMtN; i += 32) key[e * kMtN + i] = s_mt[w][i];
        __syncwarp();
    }
}

__device__ double dist2d(double ax, double ay, double bx, double by)
{
    const double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}

struct ResetPtrs {
    uint32_t* key; int32_t* pos; const uint8_t* mask;
    double* state; double* goal; double* cores; double* obst; double* start_pose;
    int32_t* ep_step; uint8_t* n_placed;
};

struct WarpScratch {
    uint32_t mt[kMtN];
    double cx[MNV_MAX_CORES], cy[MNV_MAX_CORES], cG[MNV_MAX_CORES];
    double ox[MNV_MAX_OBSTACLES], oy[MNV_MAX_OBSTACLES], orr[MNV_MAX_OBSTACLES];
    int ccw[MNV_MAX_CORES];
};

__global__ void __launch_bounds__(kBlock)
mnv_reset_kernel(const ResetPtrs P, const mnv_reset_params R, long long E, int max_c, int max_o)
{
    __shared__ WarpScratch s_ws[kWarpsPerCta];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    WarpScratch& S = s_ws[w];
    // contiguous chunk of environments per warp: a warp whose chunk has no finished environment only scans its flags
    const long long n_warps = (long long)gridDim.x * kWarpsPerCta;
    const long long gw = (long long)blockIdx.x * kWarpsPerCta + w;
    const long long chunk = (E + n_warps - 1) / n_warps;
    const long long lo = gw * chunk, hi = (lo + chunk < E) ? lo + chunk : E;

    for (long long base = lo; base < hi; base += 32) {
        unsigned todo = 0xffffffffu;
        if (P.mask != nullptr) {
            const long long i = base + lane;
            todo = __ballot_sync(0xffffffffu, i < hi && P.mask[i] != 0);
        } else if (hi - base < 32) todo = (1u << (int)(hi - base)) - 1u;
        while (todo) {
            const int bit = __ffs(todo) - 1;
            todo &= todo - 1;
            const long long e = base + bit;

            for (int i = lane; i < kMtN; i += 32) S.mt[i] = P.key[e * kMtN + i];
            __syncwarp();
            Mt rd{S.mt, P.pos[e], lane};

            double sx = R.start[0], sy = R.start[1], gx = R.goal[0], gy = R.goal[1];
            if (!R.reset_start_and_goal && P.start_pose != nullptr && R.start[0] != R.start[0]) {
                // NaN start in the parameter block = "keep the per-environment start / goal already in the tables"
                sx = P.start_pose[e]; sy = P.start_pose[E + e]; gx = P.goal[e]; gy = P.goal[E + e];
            }
            if (R.reset_start_and_goal) {                                    // marinenav_env.py:112-127
                int iteration = 500; double max_dist = 0.0;
                for (;;) {
                    const double s0 = rd.uniform(2.0, R.width - 2.0), s1 = rd.uniform(2.0, R.height - 2.0);
                    const double g0 = rd.uniform(2.0, R.width - 2.0), g1 = rd.uniform(2.0, R.height - 2.0);
                    iteration -= 1;
                    const double d = dist2d(g0, g1, s0, s1);
                    if (d > max_dist) { max_dist = d; sx = s0; sy = s1; gx = g0; gy = g1; }
                    if (max_dist > R.min_start_goal_dis || iteration == 0) break;
                }
            }

            int nc = 0;
            int num_cores = R.num_cores;
            if (num_cores > 0) {                                             // marinenav_env.py:130-143
                int iteration = 500;
                for (;;) {
                    const double x = rd.uniform(0.0, R.width), y = rd.uniform(0.0, R.height);
                    const int clockwise = rd.sample() > 0.5 ? 1 : 0;         // binomial(1, 0.5): inversion, one draw
                    const double v_edge = rd.uniform(R.v_range[0], R.v_range[1]);
                    const double Gamma = kTwoPi * R.core_r * v_edge;
                    iteration -= 1;
                    // check_core, marinenav_env.py:344-383 (the y test uses width, sic)
                    bool ok = !(x - R.core_r < 0.0 || x + R.core_r > R.width) && !(y - R.core_r < 0.0 || y + R.core_r > R.width);
                    ok = ok && !(dist2d(x, y, sx, sy) < R.core_r + R.clear_r) && !(dist2d(x, y, gx, gy) < R.core_r + R.clear_r);
                    // pairwise constraints against the accepted cores: lane i checks core i (same arithmetic per pair as the
                    // reference's loop, marinenav_env.py:361-381), the verdict is a warp vote
                    {
                        bool fail = false;
                        if (ok && lane < nc) {
                            const int i = lane;
                            const double dx = S.cx[i] - x, dy = S.cy[i] - y;
                            const double dis = sqrt(dx * dx + dy * dy);
                            if (S.ccw[i] == clockwise) {
                                const double bi = S.cG[i] / (kTwoPi * R.v_rel_max), bj = Gamma / (kTwoPi * R.v_rel_max);
                                if (dis < bi + bj) fail = true;
                            } else {
                                const double Gl = S.cG[i] > Gamma ? S.cG[i] : Gamma, Gs = S.cG[i] < Gamma ? S.cG[i] : Gamma;
                                const double v1 = Gl / (kTwoPi * (dis - 2 * R.core_r));      // Q7: negative when dis < 2r -> accepted
                                const double v2 = Gs / (kTwoPi * R.core_r);
                                if (v1 > R.p * v2) fail = true;
                            }
                        }
                        ok = ok && !__any_sync(0xffffffffu, fail);
                    }
                    if (ok) {
                        __syncwarp();
                        if (lane == 0) { S.cx[nc] = x; S.cy[nc] = y; S.cG[nc] = Gamma; S.ccw[nc] = clockwise; }
                        __syncwarp();
                        ++nc; num_cores -= 1;
                    }
                    if (iteration == 0 || num_cores == 0) break;
                }
            }

            int no = 0;
            int num_obs = R.num_obs;
            if (num_obs > 0) {                                               // marinenav_env.py:158-169
                int iteration = 500;
                for (;;) {
                    const double x = rd.uniform(5.0, R.width - 5.0), y = rd.uniform(5.0, R.height - 5.0);
                    const double r = rd.uniform(R.obs_r_range[0], R.obs_r_range[1]);
                    iteration -= 1;
                    // check_obstacle, marinenav_env.py:385-420
                    bool ok = !(x - r < 0.0 || x + r > R.width) && !(y - r < 0.0 || y + r > R.height);
                    ok = ok && !(dist2d(x, y, sx, sy) < r + R.clear_r) && !(dist2d(x, y, gx, gy) < r + R.clear_r);
                    {   // lane i checks vortex core i and obstacle i (marinenav_env.py:402-418), warp vote
                        bool fail = false;
                        if (ok && lane < nc) {
                            const double dx = S.cx[lane] - x, dy = S.cy[lane] - y;
                            if (sqrt(dx * dx + dy * dy) <= R.core_r + r) fail = true;
                        }
                        if (ok && lane < no) {
                            const double dx = S.ox[lane] - x, dy = S.oy[lane] - y;
                            if (sqrt(dx * dx + dy * dy) <= S.orr[lane] + r) fail = true;
                        }
                        ok = ok && !__any_sync(0xffffffffu, fail);
                    }
                    if (ok) {
                        __syncwarp();
                        if (lane == 0) { S.ox[no] = x; S.oy[no] = y; S.orr[no] = r; }
                        __syncwarp();
                        ++no; num_obs -= 1;
                    }
                    if (iteration == 0 || num_obs == 0) break;
                }
            }

            double th0 = R.init_theta, sp0 = R.init_speed;                    // reset_robot, marinenav_env.py:188-197
            if (R.random_reset_state) {
                th0 = rd.uniform(0.0, kTwoPi);
                sp0 = rd.uniform(0.0, R.max_speed);
            }
            __syncwarp();
            for (int i = lane; i < kMtN; i += 32) P.key[e * kMtN + i] = S.mt[i];
            if (lane == 0) {
                P.pos[e] = rd.pos;
                P.state[e] = sx; P.state[E + e] = sy; P.state[2 * E + e] = th0; P.state[3 * E + e] = sp0;
                P.goal[e] = gx; P.goal[E + e] = gy;
                if (P.start_pose != nullptr) {
                    P.start_pose[e] = sx; P.start_pose[E + e] = sy; P.start_pose[2 * E + e] = th0; P.start_pose[3 * E + e] = sp0;
                }
                if (P.ep_step != nullptr) P.ep_step[e] = 0;                  // marinenav_env.py:106
                if (P.n_placed != nullptr) { P.n_placed[e] = (uint8_t)nc; P.n_placed[E + e] = (uint8_t)no; }   // Q8
            }
            for (int i = lane; i < max_c; i += 32) {
                const bool on = i < nc;
                P.cores[(long long)i * E + e] = on ? S.cx[i] : 0.0;
                P.cores[(long long)(max_c + i) * E + e] = on ? S.cy[i] : 0.0;
                P.cores[(long long)(2 * max_c + i) * E + e] = on ? (S.ccw[i] ? S.cG[i] : -S.cG[i]) : 0.0;
            }
            for (int j = lane; j < max_o; j += 32) {
                const bool on = j < no;
                P.obst[(long long)j * E + e] = on ? S.ox[j] : 0.0;
                P.obst[(long long)(max_o + j) * E + e] = on ? S.oy[j] : 0.0;
                P.obst[(long long)(2 * max_o + j) * E + e] = on ? S.orr[j] : 0.0;
            }
            __syncwarp();
        }
    }
}

int reset_grid(long long E)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (E + kWarpsPerCta - 1) / kWarpsPerCta;        // at most one environment per warp ...
    long long cap = (long long)sms * 2;                               // ... but never more than 2 CTAs per SM: persistent warps
    return (int)(want < cap ? want : cap);
}

}  // namespace

extern "C" int mnv_seed(uint32_t* d_rng_key, int32_t* d_rng_pos, const uint32_t* d_seeds, int64_t E, void* stream)
{
    if (E <= 0) { mnv_set_error("mnv_seed: E must be > 0"); return MNV_E_SIZE; }
    MNV_CHECK_PTR(d_rng_key); MNV_CHECK_PTR(d_rng_pos); MNV_CHECK_PTR(d_seeds);
    mnv_seed_kernel<<<reset_grid(E), kBlock, 0, (cudaStream_t)stream>>>(d_rng_key, d_rng_pos, d_seeds, E);
    return mnv_launch_status("mnv_seed");
}

extern "C" int mnv_reset(uint32_t* d_rng_key, int32_t* d_rng_pos, const uint8_t* d_mask,
                         double* d_state, double* d_goal, double* d_cores, double* d_obstacles,
                         double* d_start_pose, int32_t* d_episode_step, uint8_t* d_n_placed,
                         int64_t E, int32_t max_c, int32_t max_o, const mnv_reset_params* rp, void* stream)
{
    if (rp == nullptr) { mnv_set_error("mnv_reset: params is null"); return MNV_E_NULL; }
    if (E <= 0) { mnv_set_error("mnv_reset: E must be > 0"); return MNV_E_SIZE; }
    if (max_c < 0 || max_c > MNV_MAX_CORES || max_o < 0 || max_o > MNV_MAX_OBSTACLES ||
        rp->num_cores < 0 || rp->num_cores > max_c || rp->num_obs < 0 || rp->num_obs > max_o) {
        mnv_set_error("mnv_reset: num_cores=%d/max_c=%d, num_obs=%d/max_o=%d exceed capacity (%d/%d)", rp->num_cores, max_c,
                      rp->num_obs, max_o, MNV_MAX_CORES, MNV_MAX_OBSTACLES);
        return MNV_E_CAPACITY;
    }
    MNV_CHECK_PTR(d_rng_key); MNV_CHECK_PTR(d_rng_pos); MNV_CHECK_PTR(d_state); MNV_CHECK_PTR(d_goal);
    if (max_c > 0) MNV_CHECK_PTR(d_cores);
    if (max_o > 0) MNV_CHECK_PTR(d_obstacles);
    MNV_CHECK_PTR_OPT(d_start_pose);
    ResetPtrs P{d_rng_key, d_rng_pos, d_mask, d_state, d_goal, d_cores, d_obstacles, d_start_pose, d_episode_step, d_n_placed};
    mnv_reset_kernel<<<reset_grid(E), kBlock, 0, (cudaStream_t)stream>>>(P, *rp, E, max_c, max_o);
    return mnv_launch_status("mnv_reset");
}

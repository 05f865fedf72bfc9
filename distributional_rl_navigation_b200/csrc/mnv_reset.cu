// Device-side MarineNavEnv.seed / MarineNavEnv.reset (marinenav_env.py:75-78, 86-197, 344-420) for sm_100a.
//
// One thread per environment, each with its own numpy-compatible legacy MT19937 stream kept in HBM
// (key u32 [624][E] so that the block regeneration is coalesced across a warp, pos i32 [E]).  The draw order and the
// arithmetic (separately rounded * and +, IEEE sqrt and /) follow the reference exactly, so environment e reproduces
// MarineNavEnv(seed=s_e).reset() bit-for-bit: this file MUST be compiled with -fmad=false.
// Cold path (episodes last hundreds of steps): clarity over speed.
#include <math.h>

#include "mnv_common.cuh"

namespace {

constexpr int kBlock = 64;
constexpr double kTwoPi = 2 * MNV_PI;

struct Mt {
    uint32_t* key; long long E, e; int pos;
    __device__ uint32_t& at(int i) { return key[(long long)i * E + e]; }
    __device__ void regenerate()
    {
        const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAG = 0x9908b0dfu;
        int kk; uint32_t y;
        for (kk = 0; kk < 624 - 397; ++kk) {
            y = (at(kk) & UP) | (at(kk + 1) & LO);
            at(kk) = at(kk + 397) ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        for (; kk < 623; ++kk) {
            y = (at(kk) & UP) | (at(kk + 1) & LO);
            at(kk) = at(kk + (397 - 624)) ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        }
        y = (at(623) & UP) | (at(0) & LO);
        at(623) = at(396) ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
        pos = 0;
    }
    __device__ uint32_t next32()
    {
        if (pos >= 624) regenerate();
        uint32_t y = at(pos++);
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
    const long long e = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (e >= E) return;
    uint32_t prev = seeds[e];                                       // init_genrand
    key[e] = prev;
    for (int i = 1; i < 624; ++i) {
        prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
        key[(long long)i * E + e] = prev;
    }
    pos[e] = 624;
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

__global__ void __launch_bounds__(kBlock)
mnv_reset_kernel(const ResetPtrs P, const mnv_reset_params R, long long E, int max_c, int max_o)
{
    const long long e = (long long)blockIdx.x * kBlock + threadIdx.x;
    if (e >= E) return;
    if (P.mask != nullptr && P.mask[e] == 0) return;
    Mt rd{P.key, E, e, P.pos[e]};

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

    double cxs[MNV_MAX_CORES], cys[MNV_MAX_CORES], cG[MNV_MAX_CORES]; int ccw[MNV_MAX_CORES];
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
            for (int i = 0; ok && i < nc; ++i) {
                const double dx = cxs[i] - x, dy = cys[i] - y;
                const double dis = sqrt(dx * dx + dy * dy);
                if (ccw[i] == clockwise) {
                    const double bi = cG[i] / (kTwoPi * R.v_rel_max), bj = Gamma / (kTwoPi * R.v_rel_max);
                    if (dis < bi + bj) ok = false;
                } else {
                    const double Gl = cG[i] > Gamma ? cG[i] : Gamma, Gs = cG[i] < Gamma ? cG[i] : Gamma;
                    const double v1 = Gl / (kTwoPi * (dis - 2 * R.core_r));      // Q7: negative when dis < 2r -> accepted
                    const double v2 = Gs / (kTwoPi * R.core_r);
                    if (v1 > R.p * v2) ok = false;
                }
            }
            if (ok) { cxs[nc] = x; cys[nc] = y; cG[nc] = Gamma; ccw[nc] = clockwise; ++nc; num_cores -= 1; }
            if (iteration == 0 || num_cores == 0) break;
        }
    }

    double oxs[MNV_MAX_OBSTACLES], oys[MNV_MAX_OBSTACLES], ors[MNV_MAX_OBSTACLES];
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
            for (int i = 0; ok && i < nc; ++i) {
                const double dx = cxs[i] - x, dy = cys[i] - y;
                if (sqrt(dx * dx + dy * dy) <= R.core_r + r) ok = false;
            }
            for (int i = 0; ok && i < no; ++i) {
                const double dx = oxs[i] - x, dy = oys[i] - y;
                if (sqrt(dx * dx + dy * dy) <= ors[i] + r) ok = false;
            }
            if (ok) { oxs[no] = x; oys[no] = y; ors[no] = r; ++no; num_obs -= 1; }
            if (iteration == 0 || num_obs == 0) break;
        }
    }

    double th0 = R.init_theta, sp0 = R.init_speed;                    // reset_robot, marinenav_env.py:188-197
    if (R.random_reset_state) {
        th0 = rd.uniform(0.0, kTwoPi);
        sp0 = rd.uniform(0.0, R.max_speed);
    }

    P.pos[e] = rd.pos;
    P.state[e] = sx; P.state[E + e] = sy; P.state[2 * E + e] = th0; P.state[3 * E + e] = sp0;
    P.goal[e] = gx; P.goal[E + e] = gy;
    if (P.start_pose != nullptr) {
        P.start_pose[e] = sx; P.start_pose[E + e] = sy; P.start_pose[2 * E + e] = th0; P.start_pose[3 * E + e] = sp0;
    }
    for (int i = 0; i < max_c; ++i) {
        const bool on = i < nc;
        P.cores[(long long)i * E + e] = on ? cxs[i] : 0.0;
        P.cores[(long long)(max_c + i) * E + e] = on ? cys[i] : 0.0;
        P.cores[(long long)(2 * max_c + i) * E + e] = on ? (ccw[i] ? cG[i] : -cG[i]) : 0.0;
    }
    for (int j = 0; j < max_o; ++j) {
        const bool on = j < no;
        P.obst[(long long)j * E + e] = on ? oxs[j] : 0.0;
        P.obst[(long long)(max_o + j) * E + e] = on ? oys[j] : 0.0;
        P.obst[(long long)(2 * max_o + j) * E + e] = on ? ors[j] : 0.0;
    }
    if (P.ep_step != nullptr) P.ep_step[e] = 0;                      // marinenav_env.py:106
    if (P.n_placed != nullptr) { P.n_placed[e] = (uint8_t)nc; P.n_placed[E + e] = (uint8_t)no; }   // Q8
}

}  // namespace

extern "C" int mnv_seed(uint32_t* d_rng_key, int32_t* d_rng_pos, const uint32_t* d_seeds, int64_t E, void* stream)
{
    if (E <= 0) { mnv_set_error("mnv_seed: E must be > 0"); return MNV_E_SIZE; }
    MNV_CHECK_PTR(d_rng_key); MNV_CHECK_PTR(d_rng_pos); MNV_CHECK_PTR(d_seeds);
    mnv_seed_kernel<<<(unsigned)((E + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(d_rng_key, d_rng_pos, d_seeds, E);
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
    mnv_reset_kernel<<<(unsigned)((E + kBlock - 1) / kBlock), kBlock, 0, (cudaStream_t)stream>>>(P, *rp, E, max_c, max_o);
    return mnv_launch_status("mnv_reset");
}

// Device-side MarineNavEnv.seed / MarineNavEnv.reset (marinenav_env.py:75-78, 86-197, 344-420) for sm_100a.
//
// One WARP per (finished) environment, each environment with its own numpy-compatible legacy MT19937 stream kept in HBM
// (key u32 [E][624]: one contiguous 2.5 KB record per environment, staged through shared memory; pos i32 [E]).  The draw
// order and the arithmetic (separately rounded * and +, IEEE sqrt and /) follow the reference exactly, so environment e
// reproduces MarineNavEnv(seed=s_e).reset() bit-for-bit: this file MUST be compiled with -fmad=false.
//
// SPECULATIVE REJECTION SAMPLING.  The reference's three sampling loops (start/goal :114-127, cores :130-143, obstacles
// :158-169) draw a FIXED number of stream words per candidate (8 / 8 / 6) whether or not the candidate is accepted, so
// candidate k of a loop is a pure function of the stream words [pos + k w, pos + (k + 1) w).  The warp therefore evaluates
// up to 32 consecutive candidates at once, one per lane: draws, map-bound / clearance tests and the tests against
// everything accepted before the batch run in parallel; only the dependence on candidates accepted INSIDE the batch is
// resolved in order (first surviving lane is accepted, later lanes test against it, repeat).  The stream position then
// advances by exactly the candidates the sequential loop would have consumed.  A batch never reads past the end of the
// current MT19937 block; the one candidate that straddles a regeneration is evaluated on its own.  Same candidates, same
// arithmetic per candidate, same accept decisions -> bit-identical maps, ~12 dependent rounds instead of ~30 serial
// candidates with several square-root chains each.
// A persistent grid scans the done flags; each CTA spreads the finished environments of its range over its warps
// through a shared-memory list, so a masked launch with a few finished environments costs one flag scan plus ~one
// reset latency.
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
    __device__ static uint32_t temper(uint32_t y)
    {
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    // RandomState.random_sample from two consecutive outputs: (a * 2^26 + b) / 2^53 -- the division by a power of two is
    // exact, so multiplying by 2^-53 gives the identical double
    __device__ static double to_sample(uint32_t t0, uint32_t t1)
    {
        const uint32_t a = t0 >> 5, b = t1 >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    __device__ double sample() { const uint32_t t0 = next32(), t1 = next32(); return to_sample(t0, t1); }
    // RandomState.uniform(lo, hi) = lo + (hi - lo) * U
    __device__ static double uni(double lo, double hi, double u) { return lo + (hi - lo) * u; }
    // sample j of the candidate that starts `word` words after the current position (no regeneration: caller stays in the block)
    __device__ double peek(int word, int j) const { return to_sample(temper(mt[pos + word + 2 * j]), temper(mt[pos + word + 2 * j + 1])); }
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
    // Work distribution: the CTA owns a contiguous range of environments and walks it in tiles of kBlock.  Per tile every
    // thread tests ONE flag and the finished environments go to a shared-memory list (slot = position among the tile's
    // finished environments: ballot + per-warp prefix, no atomics, deterministic); after a barrier the CTA's warps take the
    // list entries round robin, one environment per warp at a time.  A reset is a ~8 us dependent chain, so what matters is
    // that no warp gets two while another has none: with 1 % of 65 536 environments finished a CTA's 222 environments
    // hold ~2 resets for its 8 warps (a contiguous chunk per WARP left the unluckiest warp with 3-4).
    __shared__ int s_list[kBlock];
    __shared__ int s_warp_cnt[kWarpsPerCta];
    const long long per_cta = (E + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per_cta, hi = (lo + per_cta < E) ? lo + per_cta : E;

    for (long long tile = lo; tile < hi; tile += kBlock) {
        const long long i = tile + threadIdx.x;
        const bool fin = i < hi && (P.mask == nullptr || P.mask[i] != 0);
        const unsigned bal = __ballot_sync(0xffffffffu, fin);
        if (lane == 0) s_warp_cnt[w] = __popc(bal);
        __syncthreads();
        int before = 0, n_fin = 0;
#pragma unroll
        for (int k = 0; k < kWarpsPerCta; ++k) { const int c = s_warp_cnt[k]; before += k < w ? c : 0; n_fin += c; }
        if (fin) s_list[before + __popc(bal & ((1u << lane) - 1u))] = threadIdx.x;
        __syncthreads();
        for (int q = w; q < n_fin; q += kWarpsPerCta) {
            const long long e = tile + s_list[q];

            for (int i = lane; i < kMtN; i += 32) S.mt[i] = P.key[e * kMtN + i];
            __syncwarp();
            Mt rd{S.mt, P.pos[e], lane};

            double sx = R.start[0], sy = R.start[1], gx = R.goal[0], gy = R.goal[1];
            if (!R.reset_start_and_goal && P.start_pose != nullptr && R.start[0] != R.start[0]) {
                // NaN start in the parameter block = "keep the per-environment start / goal already in the tables"
                sx = P.start_pose[e]; sy = P.start_pose[E + e]; gx = P.goal[e]; gy = P.goal[E + e];
            }
            // One batch of candidates: lane k < n holds candidate k of the loop, u[] = its uniform samples in draw order.
            // n = how many whole candidates the current MT19937 block still holds (<= 32, <= iterations left); when the next
            // candidate straddles a regeneration it is drawn word by word (all lanes, uniform) and forms a batch of one whose
            // words are already consumed (returns true).
            auto draw_batch = [&](int words, int iterations_left, int& n, double (&u)[4]) -> bool {
                if (rd.pos >= kMtN) rd.regenerate();                         // block exhausted: the next one (warp-uniform)
                n = (kMtN - rd.pos) / words;
                n = n < 32 ? n : 32;
                n = n < iterations_left ? n : iterations_left;
                if (n == 0) {
                    for (int j = 0; j < words / 2; ++j) u[j] = rd.sample();
                    n = 1;
                    return true;
                }
                if (lane < n)
                    for (int j = 0; j < words / 2; ++j) u[j] = rd.peek(lane * words, j);
                return false;
            };

            if (R.reset_start_and_goal) {                                    // marinenav_env.py:112-127
                // the loop keeps the farthest pair seen and stops at the first candidate farther apart than
                // min_start_goal_dis (every earlier one was not) or after 500 candidates
                int iteration = 500; double max_dist = 0.0;
                while (iteration > 0) {
                    int n; double u[4];
                    const bool consumed = draw_batch(8, iteration, n, u);
                    double s0 = 0, s1 = 0, g0 = 0, g1 = 0, d = -1.0;
                    if (lane < n) {
                        s0 = Mt::uni(2.0, R.width - 2.0, u[0]); s1 = Mt::uni(2.0, R.height - 2.0, u[1]);
                        g0 = Mt::uni(2.0, R.width - 2.0, u[2]); g1 = Mt::uni(2.0, R.height - 2.0, u[3]);
                        d = dist2d(g0, g1, s0, s1);
                    }
                    // the loop ends at the first candidate k whose running maximum max(max_dist, d_0 .. d_k) exceeds
                    // min_start_goal_dis; m = candidates consumed
                    unsigned far = __ballot_sync(0xffffffffu, lane < n && d > R.min_start_goal_dis);
                    if (max_dist > R.min_start_goal_dis) far |= 1u;          // (only reachable with a negative threshold)
                    const int m = far != 0u ? __ffs(far) : n;
                    // running maximum over the consumed candidates: the FIRST one holding the largest distance, if it
                    // exceeds the carried maximum (strict >, like the reference's update)
                    double best = lane < m ? d : -1.0; int who = lane;
                    for (int off = 16; off > 0; off >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, best, off);
                        const int ow = __shfl_xor_sync(0xffffffffu, who, off);
                        if (od > best || (od == best && ow < who)) { best = od; who = ow; }
                    }
                    if (best > max_dist) {
                        max_dist = best;
                        sx = __shfl_sync(0xffffffffu, s0, who); sy = __shfl_sync(0xffffffffu, s1, who);
                        gx = __shfl_sync(0xffffffffu, g0, who); gy = __shfl_sync(0xffffffffu, g1, who);
                    }
                    if (!consumed) rd.pos += m * 8;
                    iteration -= m;
                    if (far != 0u) break;
                }
            }

            int nc = 0;
            if (R.num_cores > 0) {                                           // marinenav_env.py:130-143
                int iteration = 500, need = R.num_cores;
                // check_core's pairwise rule (marinenav_env.py:361-381): existing core i against the new core (x, y, cw, Gamma)
                auto core_conflict = [&](int i, double x, double y, int clockwise, double Gamma) -> bool {
                    const double dx = S.cx[i] - x, dy = S.cy[i] - y;
                    const double dis = sqrt(dx * dx + dy * dy);
                    if (S.ccw[i] == clockwise) {
                        const double bi = S.cG[i] / (kTwoPi * R.v_rel_max), bj = Gamma / (kTwoPi * R.v_rel_max);
                        return dis < bi + bj;
                    }
                    const double Gl = S.cG[i] > Gamma ? S.cG[i] : Gamma, Gs = S.cG[i] < Gamma ? S.cG[i] : Gamma;
                    const double v1 = Gl / (kTwoPi * (dis - 2 * R.core_r));          // Q7: negative when dis < 2r -> accepted
                    const double v2 = Gs / (kTwoPi * R.core_r);
                    return v1 > R.p * v2;
                };
                while (iteration > 0 && need > 0) {
                    int n; double u[4];
                    const bool consumed = draw_batch(8, iteration, n, u);
                    double x = 0, y = 0, Gamma = 0; int clockwise = 0;
                    bool alive = lane < n;
                    if (alive) {
                        x = Mt::uni(0.0, R.width, u[0]); y = Mt::uni(0.0, R.height, u[1]);
                        clockwise = u[2] > 0.5 ? 1 : 0;                      // binomial(1, 0.5): inversion, one draw
                        const double v_edge = Mt::uni(R.v_range[0], R.v_range[1], u[3]);
                        Gamma = kTwoPi * R.core_r * v_edge;
                        // check_core, marinenav_env.py:344-383 (the y test uses width, sic)
                        alive = !(x - R.core_r < 0.0 || x + R.core_r > R.width) && !(y - R.core_r < 0.0 || y + R.core_r > R.width);
                        alive = alive && !(dist2d(x, y, sx, sy) < R.core_r + R.clear_r) && !(dist2d(x, y, gx, gy) < R.core_r + R.clear_r);
                        for (int i = 0; alive && i < nc; ++i) alive = !core_conflict(i, x, y, clockwise, Gamma);   // cores of earlier batches
                    }
                    int m = n;
                    for (;;) {                                               // resolve the batch in candidate order
                        const unsigned bal = __ballot_sync(0xffffffffu, alive);
                        if (bal == 0u) break;
                        const int f = __ffs(bal) - 1;                        // the next candidate the sequential loop accepts
                        const double ax = __shfl_sync(0xffffffffu, x, f), ay = __shfl_sync(0xffffffffu, y, f);
                        const double aG = __shfl_sync(0xffffffffu, Gamma, f);
                        const int acw = __shfl_sync(0xffffffffu, clockwise, f);
                        __syncwarp();
                        if (lane == 0) { S.cx[nc] = ax; S.cy[nc] = ay; S.cG[nc] = aG; S.ccw[nc] = acw; }
                        __syncwarp();
                        ++nc; --need;
                        if (need == 0) { m = f + 1; break; }
                        alive = alive && lane > f && !core_conflict(nc - 1, x, y, clockwise, Gamma);
                    }
                    if (!consumed) rd.pos += m * 8;
                    iteration -= m;
                }
            }

            int no = 0;
            if (R.num_obs > 0) {                                             // marinenav_env.py:158-169
                int iteration = 500, need = R.num_obs;
                while (iteration > 0 && need > 0) {
                    int n; double u[4];
                    const bool consumed = draw_batch(6, iteration, n, u);
                    double x = 0, y = 0, r = 0;
                    bool alive = lane < n;
                    if (alive) {
                        x = Mt::uni(5.0, R.width - 5.0, u[0]); y = Mt::uni(5.0, R.height - 5.0, u[1]);
                        r = Mt::uni(R.obs_r_range[0], R.obs_r_range[1], u[2]);
                        // check_obstacle, marinenav_env.py:385-420
                        alive = !(x - r < 0.0 || x + r > R.width) && !(y - r < 0.0 || y + r > R.height);
                        alive = alive && !(dist2d(x, y, sx, sy) < r + R.clear_r) && !(dist2d(x, y, gx, gy) < r + R.clear_r);
                        for (int i = 0; alive && i < nc; ++i) {              // every vortex core (marinenav_env.py:402-408)
                            const double dx = S.cx[i] - x, dy = S.cy[i] - y;
                            alive = !(sqrt(dx * dx + dy * dy) <= R.core_r + r);
                        }
                        for (int i = 0; alive && i < no; ++i) {              // obstacles of earlier batches (:410-418)
                            const double dx = S.ox[i] - x, dy = S.oy[i] - y;
                            alive = !(sqrt(dx * dx + dy * dy) <= S.orr[i] + r);
                        }
                    }
                    int m = n;
                    for (;;) {
                        const unsigned bal = __ballot_sync(0xffffffffu, alive);
                        if (bal == 0u) break;
                        const int f = __ffs(bal) - 1;
                        const double ax = __shfl_sync(0xffffffffu, x, f), ay = __shfl_sync(0xffffffffu, y, f), ar = __shfl_sync(0xffffffffu, r, f);
                        __syncwarp();
                        if (lane == 0) { S.ox[no] = ax; S.oy[no] = ay; S.orr[no] = ar; }
                        __syncwarp();
                        ++no; --need;
                        if (need == 0) { m = f + 1; break; }
                        const double dx = ax - x, dy = ay - y;               // existing obstacle (ax, ay, ar) against the candidate
                        alive = alive && lane > f && !(sqrt(dx * dx + dy * dy) <= ar + r);
                    }
                    if (!consumed) rd.pos += m * 6;
                    iteration -= m;
                }
            }

            double th0 = R.init_theta, sp0 = R.init_speed;                    // reset_robot, marinenav_env.py:188-197
            if (R.random_reset_state) {
                th0 = Mt::uni(0.0, kTwoPi, rd.sample());
                sp0 = Mt::uni(0.0, R.max_speed, rd.sample());
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
        __syncthreads();                                    // the list is rebuilt by the next tile
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

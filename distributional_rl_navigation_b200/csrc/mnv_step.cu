// Fused batched MarineNavEnv.step / get_observation for sm_100a.
//
// One thread integrates one environment: fp64 state, N sub-steps of Rankine-vortex current + USV kinematics
// (marinenav_env.py:422-465, robot.py:95-123), then an fp64 robot-frame ray cast of the sonar beams
// (robot.py:125-198), the observation (marinenav_env.py:273-326), reward and termination priority
// (marinenav_env.py:199-262).  Everything an environment needs (4 state + 2 goal + 3*max_c + 3*max_o doubles) is
// read once with coalesced 8-byte loads from the SoA tables and kept in registers; observation rows are staged in
// shared memory and written back as one contiguous float4 stream per CTA.
//
// Formulation differences against the reference (all below 1e-12 except where the reference itself is
// ill-conditioned; tests/test_env_parity.py states the tolerances):
//   * vortex velocity: tangent*speed = k*(-dy,dx)/d^2 outside the core and k*(-dy,dx)/r^2 inside (k = +-Gamma/2pi),
//     i.e. no sqrt and one Newton reciprocal per core instead of normalise-then-scale (same value, Q1 kept: every
//     core contributes; Q2: summation order is table order, ulp-level only);
//     d == 0 (the robot exactly on a core centre): the reference divides 0/0 there (marinenav_env.py:447-449 -> NaN state
//     for the rest of the episode); this kernel takes the solid-body branch, whose contribution at d = 0 is exactly zero.
//     Unreachable in practice (a measure-zero point the fp64 trajectory would have to hit exactly) and deliberately NOT
//     reproduced: a NaN pose has no defined observation, reward or flag to be in parity with;
//   * heading: sincos(theta) once, then the sub-steps rotate (cos,sin) by the fixed yaw increment w*dt;
//   * sonar: obstacle centres are rotated into the robot frame once, beams are the constant directions
//     (cos b, sin b) there; ray/circle in direction form t = t_ca -+ sqrt(r^2 - cross^2) picking the reference's
//     "nearer root first" (robot.py:184), range and behind tests (robot.py:185-190), the ordered break Q3
//     (robot.py:192-195), and the vertical snap Q10 (robot.py:134-162: a beam within 1e-3 rad of +-pi/2 is
//     intersected along exactly (0,+-1)).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "mnv_common.cuh"

namespace {

constexpr int kBlock = 64;      // E = 65536 -> 1024 CTAs = 6.9 per SM with 8 resident (<= 128 registers): one balanced wave

struct KParams {
    double dt;
    double accel[3], wdt[3], cos_wdt[3], sin_wdt[3];
    double k_drag, max_speed, robot_r, core_r2, inv_core_r2, goal_dis;
    double pen_step, pen_coll, rew_goal;
    double range, range_slack;
    double width, height;
    int n_substeps, n_beams, max_ep_steps, set_boundary;
    int max_c, max_o, obs_dim, velocity_from_state, allow_tma, pdl_prefetch;
    long long E;
    double snap_t1, snap_t2;        // pi/2 - beam_angle[0], 3pi/2 - beam_angle[0]   (Q10 pre-test, see snap_beam)
    float inv_phi, snap_tol;        // 1 / beam spacing, (1e-3 / spacing) + 1e-4 in beam-index units
    float beam0f; int precise_bins; // beam_angle[0]; 1: per-obstacle beam intervals (<= 16 beams, field of view < pi), 0: every beam of every obstacle within reach
    int dense_bins;                 // dense kernel: per-obstacle beam intervals (<= 64 beams, field of view < pi)
    double beam_angle[MNV_MAX_BEAMS];
    double beam_cos[MNV_MAX_BEAMS];
    double beam_sin[MNV_MAX_BEAMS];
};

struct EnvPtrs {
    double* state; double* velocity; const double* goal; const double* cores; const double* obst;
    const int32_t* action; int32_t* ep_step; const uint8_t* mask; double* traj;
    float* obs; float* reward; uint8_t* done; uint8_t* info;
};

// 1/a to ~1 ulp: MUFU.RCP64H seed (>= 20 bits) + one cubic step r(1 + e + e^2), e = 1 - a r  (e^3 <= 2^-60);
// a is a squared distance in [1e-300, 1e300] here.
__device__ __forceinline__ double fast_rcp(double a)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    const double e = fma(-a, r, 1.0);
    return fma(r, fma(e, e, e), r);
}

// sqrt(a) to <= 1 ulp for a in [1e-300, 1e300]: MUFU.RSQ64H seed (>= 20 bits) + two coupled Newton steps.  Used where the
// reference's value passes through further roundings anyway (reward differences, ray parameters of the reformulated
// sonar); threshold DECISIONS on a square root are taken on the squares (collides(), reaches()).
__device__ __forceinline__ double fast_sqrt(double a)
{
    if (a <= 0.0) return a == 0.0 ? 0.0 : sqrt(a);           // 0 and the NaN of a negative argument, like sqrt
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y;                                       // ~ sqrt(a)
    double h = 0.5 * y;
    double r = fma(-g, h, 0.5);                             // 0.5 - g h
    g = fma(g, r, g); h = fma(h, r, h);                     // one Goldschmidt step: ~40 bits
    const double e = fma(-g, g, a);                         // residual
    return fma(e, h, g);                                    // final Newton correction: <= 1 ulp
}

// sin / cos for theta in [0, 2 pi) (the heading after Robot.update_state's wrap, robot.py:120-123): quadrant by one
// multiply, a two-term Cody-Waite reduction (k <= 4, so k * pio2_hi is exact) and the fdlibm kernel polynomials on
// [-pi/4, pi/4] (< 1 ulp).  Anything else (a caller poked theta out of range) takes the general sincos().
__device__ __forceinline__ void sincos_heading(double th, double* sn, double* cs)
{
    if (!(th >= 0.0 && th < 6.5)) { sincos(th, sn, cs); return; }
    const int k = __double2int_rn(th * 0.63661977236758134308);        // 2 / pi
    const double kd = (double)k;
    double r = fma(-kd, 1.57079632673412561417e+00, th);               // pio2_1  (33 bits)
    r = fma(-kd, 6.07710050650619224932e-11, r);                       // pio2_1t
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    ps = fma(z, ps, 2.75573137070700676789e-06); pc = fma(z, pc, -2.75573143513906633035e-07);
    ps = fma(z, ps, -1.98412698298579493134e-04); pc = fma(z, pc, 2.48015872894767294178e-05);
    ps = fma(z, ps, 8.33333333332248946124e-03); pc = fma(z, pc, -1.38888888888741095749e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01); pc = fma(z, pc, 4.16666666666666019037e-02);
    const double sr = fma(r * z, ps, r);                               // sin r
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));               // cos r
    const double a = (k & 1) ? cr : sr, b = (k & 1) ? sr : cr;         // sin(r + k pi/2), cos(r + k pi/2)
    *sn = (k & 2) ? -a : a;
    *cs = ((k + 1) & 2) ? -b : b;
}

// ---- conservative angular bounds in fp32 for the sonar binning (errors are covered by the margins at the use site) ----
// atan2(y, x) with |error| <= 1.2e-5 rad (Abramowitz & Stegun 4.4.47 on min/max); (0, 0) -> 0
__device__ __forceinline__ float atan2_approx(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);
    const float t = __fdividef(mn, mx), t2 = t * t;
    float a = fmaf(t2, 0.0208351f, -0.0851330f);
    a = fmaf(t2, a, 0.1801410f); a = fmaf(t2, a, -0.3302995f); a = fmaf(t2, a, 0.9998660f);
    a *= t;
    a = ay > ax ? 1.57079632679f - a : a;
    a = x < 0.0f ? 3.14159265359f - a : a;
    return copysignf(a, y);
}
// asin(x), 0 <= x <= 1, |error| <= 6e-5 rad (Abramowitz & Stegun 4.4.45)
__device__ __forceinline__ float asin_approx(float x)
{
    float p = fmaf(x, -0.0187293f, 0.0742610f);
    p = fmaf(x, p, -0.2121144f); p = fmaf(x, p, 1.5707288f);
    return fmaf(-sqrtf(1.0f - x), p, 1.57079632679f);
}

// check_collision (marinenav_env.py:329-336): sqrt(d2) <= thr, decided on the squares unless d2 is within 1e-15 (relative)
// of thr^2 -- only there can the rounding of the square root matter, and only there is it evaluated.
__device__ __forceinline__ bool collides(double d2, double thr)
{
    const double t2 = thr * thr;
    if (d2 <= t2 * (1.0 - 1e-15)) return true;
    if (!(d2 <= t2 * (1.0 + 1e-15))) return false;                 // also d2 = inf (no obstacle) and NaN
    return sqrt(d2) <= thr;
}

// check_reach_goal (marinenav_env.py:338-342): sqrt(d2) <= goal_dis, decided like collides()
__device__ __forceinline__ bool reaches(double d2, double thr) { return collides(d2, thr); }

// programmatic dependent launch: no-ops unless the launch carries the programmatic-stream-serialization attribute
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
// TMA (bulk async copy engine): one thread moves a whole 16-byte-aligned row slice global -> shared and signals an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(d), "l"(gmem_src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Dynamic shared memory: [kBlock * obs_dim floats, padded to 16 B][3 * max_o rows x kBlock doubles][mbarriers]
//                        [per-warp lists of pending exact sonar tests][per-env obstacle beam masks][per-env relevant-obstacle masks]
constexpr int kBeamWord = 16;   // beams handled per pass of the sonar phase (the default 11 beams: one pass)
constexpr int kRing = 32 * kBeamWord;   // per-warp list of pending exact tests: every (lane, beam of the pass) at most once

#ifndef MNV_MIN_CTAS
#define MNV_MIN_CTAS 8
#endif
template <int MAXC, int MAXO, bool STEP, bool PREFETCH>
__global__ void __launch_bounds__(kBlock, MNV_MIN_CTAS)
mnv_env_kernel(const EnvPtrs P, const __grid_constant__ KParams K)
{
    static_assert(STEP || !PREFETCH, "only the step launch fetches the map tables ahead of its grid dependency");
    static_assert((MAXO & 1) == 0 && MAXO <= 16, "per-env obstacle masks are 16-bit words");
    extern __shared__ __align__(16) float s_obs[];           // [kBlock][obs_dim]
    const long long E = K.E;
    const long long e0 = (long long)blockIdx.x * kBlock;
    const int tid = threadIdx.x;
    const int lane = tid & 31, w0 = tid & ~31;               // w0 = first row of this warp inside the CTA
    const long long e = e0 + tid;
    const int D = K.obs_dim;
    float* my_obs = s_obs + tid * D;
    double* s_ob = reinterpret_cast<double*>(s_obs + ((kBlock * D + 3) & ~3));   // obstacle rows of this CTA
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_ob + 3 * K.max_o * kBlock);
    unsigned short* ring = reinterpret_cast<unsigned short*>(s_bar + kBlock / 32) + (tid >> 5) * kRing;
    constexpr int kBmStride = (MAXO + 7) & ~7;               // 16-byte rows
    unsigned short* s_bm = reinterpret_cast<unsigned short*>(s_bar + kBlock / 32) + (kBlock / 32) * kRing;   // [kBlock][kBmStride]: beams obstacle j may return on
    unsigned* s_rel = reinterpret_cast<unsigned*>(s_bm + kBlock * kBmStride);                                // [kBlock]: obstacles within sonar reach
    const int max_o = K.max_o;
    // vortex cores (registers), goal and obstacle rows (shared memory, per-thread cp.async: each thread copies its own
    // column; first needed after the sub-step loop, so that DRAM round trip hides under the integration)
    double cx[MAXC], cy[MAXC], ck[MAXC], gx = 0.0, gy = 0.0;
    const bool coop = (e0 + w0 + 32 <= E) && ((E & 1) == 0) && (STEP || P.mask == nullptr);   // warp-uniform
    auto load_tables = [&]() {
        gx = P.goal[e]; gy = P.goal[E + e];
        const double* pc = P.cores + e;
        const long long rowstride = (long long)K.max_c * E;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            if (i < K.max_c) { cx[i] = __ldg(pc); cy[i] = __ldg(pc + rowstride); ck[i] = __ldg(pc + 2 * rowstride); }
            else { cx[i] = 0.0; cy[i] = 0.0; ck[i] = 0.0; }
            pc += E;
        }
        if (coop) {
            // a full warp: row r of the warp's 32 environments is 256 contiguous, 16-byte aligned bytes -> 16 lanes x 16 B, two
            // rows per trip (half the cp.async instructions of the per-thread copy; the data is shared after a __syncwarp)
            const double* pw = P.obst + (e0 + w0) + (lane & 15) * 2;
            double* ps = s_ob + w0 + (lane & 15) * 2;
            for (int row = lane >> 4; row < 3 * max_o; row += 2) cp_async16(ps + row * kBlock, pw + (long long)row * E);
        } else {
            const double* po = P.obst + e;
            for (int row = 0; row < 3 * max_o; ++row, po += E) cp_async8(s_ob + row * kBlock + tid, po);
        }
        cp_async_commit();
    };
    // "pdl" = 2: the map tables (goal, cores, obstacles) are not written by the launch right before this one (the caller's
    // contract; true for step -> step and for step after anything but mnv_reset / a table upload), so they are fetched
    // while that launch still drains; everything else waits for it.
    if (PREFETCH && e < E) load_tables();                     // (PREFETCH implies STEP: a full warp is all live)
    pdl_wait();                                               // everything below reads what earlier launches wrote
    pdl_launch_dependents();                                  // the next launch may be scheduled while this one runs (it waits like this one)
    const bool live = (e < E) && (STEP || P.mask == nullptr || P.mask[e] != 0);

    // ---- (optional, "tma" = 1) obstacle table slice of every warp -> shared memory with the TMA bulk-copy engine: 3*max_o rows of 32 consecutive
    //      environments (256 contiguous bytes each), issued by the warp's lane 0 and tracked by the warp's own mbarrier.  The rows are first
    //      needed after the sub-step loop, so this DRAM round trip overlaps the integration.  Ragged / masked / odd-E
    //      launches (rows not 16-byte aligned or not full) use per-thread cp.async instead. ----
    const bool use_tma = !PREFETCH && K.allow_tma && (e0 + kBlock <= E) && ((E & 1) == 0) && (STEP || P.mask == nullptr) && max_o > 0;
    const int warp_in_cta = tid >> 5;
    unsigned long long* my_bar = s_bar + warp_in_cta;             // one mbarrier per warp: no CTA-wide sync needed
    if (use_tma) {
        if (lane == 0) {
            mbar_init(my_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(my_bar, (unsigned)(3 * max_o * 32 * sizeof(double)));
            const double* po = P.obst + e0 + warp_in_cta * 32;
            double* ps = s_ob + warp_in_cta * 32;
            for (int row = 0; row < 3 * max_o; ++row, po += E, ps += kBlock)
                tma_bulk_g2s(ps, po, (unsigned)(32 * sizeof(double)), my_bar);
        }
        __syncwarp();                                              // the warp's barrier is initialised before its lanes poll it
    }

    // state that crosses the three phases (per-lane integration / warp-wide sonar / per-lane termination); lanes without a
    // live environment carry inert values through the warp-wide phase (no candidates, nothing pushed)
    double x = 0.0, y = 0.0, th = 0.0, sp = 0.0, c = 1.0, s = 0.0;
    double vx = 0.0, vy = 0.0, reward = 0.0, dis_after = 0.0;
    int ep = 0, bs1 = -1, bs2 = -1;
    unsigned rel = 0u, beams = 0u;                          // obstacles within sonar reach; beams with a candidate obstacle
    double best_d2 = INFINITY, best_r = 0.0;               // Q4: nearest CENTRE only (marinenav_env.py:329-336)
    double dis_after2 = 0.0;

    if (live) {
        // ---- every global load of this environment is issued before the first use of any of them, so that the kernel entry
        //      costs ONE DRAM round trip (the ncu source page of the previous version showed three in a row: state -> sincos,
        //      cores -> first sub-step, action -> accel / yaw tables) ----
        x = P.state[e]; y = P.state[E + e]; th = P.state[2 * E + e]; sp = P.state[3 * E + e];
        int action = 0;
        if (STEP) { action = P.action[e]; ep = P.ep_step[e]; }
        if (!PREFETCH) {
            if (!use_tma) load_tables();
            else {
                gx = P.goal[e]; gy = P.goal[E + e];
                const double* pc = P.cores + e;
                const long long rowstride = (long long)K.max_c * E;
#pragma unroll
                for (int i = 0; i < MAXC; ++i) {
                    if (i < K.max_c) { cx[i] = __ldg(pc); cy[i] = __ldg(pc + rowstride); ck[i] = __ldg(pc + 2 * rowstride); }
                    else { cx[i] = 0.0; cy[i] = 0.0; ck[i] = 0.0; }
                    pc += E;
                }
            }
        }
        sincos_heading(th, &s, &c);
#pragma unroll
        for (int i = 0; i < MAXC; ++i) ck[i] *= (1.0 / (2.0 * MNV_PI));
        auto current = [&](double px, double py, double& ux, double& uy) {
            ux = 0.0; uy = 0.0;
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const double dx = cx[i] - px, dy = cy[i] - py;
                const double d2 = fma(dx, dx, dy * dy);
                // compute_speed (marinenav_env.py:461-465): solid-body rotation inside the core radius
                const double f = ck[i] * (d2 <= K.core_r2 ? K.inv_core_r2 : fast_rcp(d2));
                ux = fma(-dy, f, ux);
                uy = fma(dx, f, uy);
            }
        };

        if (STEP) {
            const int ai = action / 3, wi = action - 3 * ai;
            const double acc = K.accel[ai], wdt = K.wdt[wi], cw = K.cos_wdt[wi], sw = K.sin_wdt[wi];
            const double dis_before = fast_sqrt(fma(gx - x, gx - x, (gy - y) * (gy - y)));
            const bool wdt_neg = wdt < 0.0;
            const double wrap_adj = wdt_neg ? 2.0 * MNV_PI : -2.0 * MNV_PI;
            auto substeps = [&](auto with_traj) {
                double* ptraj = decltype(with_traj)::value ? P.traj + e : nullptr;
                for (int it = 0; it < K.n_substeps; ++it) {
                    double ux, uy;
                    current(x, y, ux, uy);
                    vx = fma(sp, c, ux);                          // robot.py:98-100 (pre-update speed / heading: Q6)
                    vy = fma(sp, s, uy);
                    x = fma(vx, K.dt, x);                         // robot.py:105-107
                    y = fma(vy, K.dt, y);
                    sp = __dadd_rn(sp, __dmul_rn(__dsub_rn(acc, __dmul_rn(K.k_drag, sp)), K.dt));   // robot.py:113
                    sp = sp < 0.0 ? 0.0 : sp;                     // robot.py:114 (np.clip)
                    sp = sp > K.max_speed ? K.max_speed : sp;
                    th = __dadd_rn(th, wdt);                      // robot.py:117
                    // robot.py:120-123, the two while loops.  theta was in [0, 2 pi) and |w dt| << 2 pi, so at most ONE
                    // correction is due and the sign of the yaw increment says which: + 2 pi if w dt < 0 and theta < 0,
                    // - 2 pi if w dt > 0 and theta >= 2 pi.  Anything that does not land strictly inside (0, 2 pi) -- a
                    // theta that came in out of range, or a sum that rounded onto 0 / 2 pi -- runs the reference's own
                    // loops on the uncorrected sum (|theta - pi| < pi has no false negatives).
                    {
                        const double th_c = __dadd_rn(th, wrap_adj);
                        const double th_f = (wdt_neg ? th < 0.0 : th >= 2.0 * MNV_PI) ? th_c : th;
                        if (fabs(th_f - MNV_PI) < MNV_PI) th = th_f;
                        else {
#pragma unroll 1
                            while (th < 0.0) th += 2.0 * MNV_PI;
#pragma unroll 1
                            while (th >= 2.0 * MNV_PI) th -= 2.0 * MNV_PI;
                        }
                    }
                    const double c2 = fma(c, cw, -s * sw), s2 = fma(s, cw, c * sw);
                    c = c2; s = s2;
                    if (decltype(with_traj)::value) {             // Robot.trajectory (marinenav_env.py:212), optional
                        ptraj[0] = x; ptraj[E] = y; ptraj += 2 * E;
                    }
                }
            };
            if (P.traj != nullptr) substeps(std::true_type{}); else substeps(std::false_type{});
            dis_after2 = fma(gx - x, gx - x, (gy - y) * (gy - y));
            dis_after = fast_sqrt(dis_after2);
            reward = K.pen_step + (dis_before - dis_after);   // marinenav_env.py:220,229
        } else {
            if (K.velocity_from_state) {
                double ux, uy;
                current(x, y, ux, uy);
                vx = fma(sp, c, ux); vy = fma(sp, s, uy);
                P.velocity[e] = vx; P.velocity[E + e] = vy;
            } else { vx = P.velocity[e]; vy = P.velocity[E + e]; }
        }

        // ---- observation head: R^T v, R^T (goal - pos)  (marinenav_env.py:278-293) ----
        my_obs[0] = (float)fma(c, vx, s * vy);
        my_obs[1] = (float)fma(c, vy, -s * vx);
        my_obs[2] = (float)fma(c, gx - x, s * (gy - y));
        my_obs[3] = (float)fma(c, gy - y, -s * (gx - x));

        // ---- Q10 pre-test (robot.py:134-162): beam b is snapped to the vertical iff |theta + beam_angle[b] - T| < 1e-3,
        //      T = pi/2 or 3pi/2, i.e. iff u = (T - beam_angle[0] - theta) / spacing is within 1e-3 / spacing of the
        //      integer b.  One fp32 evaluation of u per T (error < 1e-5 index units, tolerance widened by 1e-4) names the
        //      only beam that can be snapped; the exact fp64 test runs for that beam alone. ----
        {
            const float u1 = (float)(K.snap_t1 - th) * K.inv_phi, u2 = (float)(K.snap_t2 - th) * K.inv_phi;
            const float n1 = rintf(u1), n2 = rintf(u2);
            bs1 = fabsf(u1 - n1) < K.snap_tol ? (int)n1 : -1;
            bs2 = fabsf(u2 - n2) < K.snap_tol ? (int)n2 : -1;
        }

        // ---- obstacles: nearest centre (Q4) and the set within sonar reach.  Only an obstacle with |centre - pos| <=
        //      range + r can return a hit (the hit distance is >= d - r), so everything below looks at those alone. ----
        if (use_tma) mbar_wait(my_bar, 0); else cp_async_wait_all();
        if (coop) __syncwarp();                               // rows were copied by other lanes of this (fully live) warp
        const double* ob = s_ob + tid;
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            double r = -1.0, ox = 0.0, oy = 0.0;
            if (j < max_o) { ox = ob[j * kBlock]; oy = ob[(max_o + j) * kBlock]; r = ob[(2 * max_o + j) * kBlock]; }
            const double dx = ox - x, dy = oy - y;
            const double d2 = fma(dx, dx, dy * dy);
            const bool on = r > 0.0;
            if (on && d2 < best_d2) { best_d2 = d2; best_r = r; }
            const double lim = K.range_slack + r;
            if (on && d2 <= lim * lim) rel |= 1u << j;
        }

        // ---- sonar binning, obstacle-major (robot.py:125-198).  In the robot frame the beams are the fixed directions
        //      beta_b = beta_0 + b * spacing, and the beams on which obstacle j can return a hit form ONE interval of beam
        //      indices: with q = R^T (centre - pos), d = |q|,
        //        robot outside the circle: real roots and the nearer root in front  <=>  |beta_b - atan2(q)| < asin(r / d);
        //        robot inside:  the nearer root is in front iff q . dir <= 0        <=>  |beta_b - atan2(-q)| <= pi / 2.
        //      The interval is computed conservatively in fp32 (radius inflated by 1e-3 m, half-width by 2e-4 rad + the
        //      approximation errors of atan2_approx / asin_approx, bounds rounded outwards by 1e-3 index units), so it can
        //      only add beams, never lose one; every flagged (env, beam) is then decided EXACTLY in fp64 below.  A robot
        //      within fp32 resolution of the circle line, or within 1 m of the centre of a circle it is inside of, flags
        //      every beam.  Geometries whose intervals could wrap around (field of view >= pi, or more than 16 beams)
        //      flag every beam of every obstacle within reach (K.precise_bins = 0). ----
        {
            unsigned short* my_bm = s_bm + tid * kBmStride;
#pragma unroll
            for (int j4 = 0; j4 < kBmStride / 8; ++j4) reinterpret_cast<uint4*>(my_bm)[j4] = make_uint4(0u, 0u, 0u, 0u);
            // "no return" everywhere (marinenav_env.py:318-320); the exact tests overwrite the beams that hit
#pragma unroll 4
            for (int b = 0; b < K.n_beams; ++b) *reinterpret_cast<float2*>(my_obs + 4 + 2 * b) = make_float2(0.0f, 0.0f);
        }
    }

    // ---- the intervals, warp-wide: an environment has ~1.3 obstacles within reach (at most 3-4 in a warp), so the
    //      (environment, obstacle) pairs of the warp are compacted into a list (slots from a prefix sum of the per-lane
    //      counts) and evaluated 32 pairs at a time, every lane busy; the evaluator fetches the owner's pose with shuffles
    //      and leaves the beam mask in the owner's row of s_bm. ----
    const unsigned sb1 = (live && bs1 >= 0 && bs1 < K.n_beams) ? (unsigned)(bs1 + 1) : 0u, sb2 = (live && bs2 >= 0 && bs2 < K.n_beams) ? (unsigned)(bs2 + 1) : 0u;
    s_rel[tid] = rel | (sb1 << 16) | (sb2 << 24);           // [15:0] obstacles within reach, [23:16] / [31:24] Q10 candidate beams + 1
    if (!K.precise_bins) beams = rel != 0u ? 0xffffffffu : 0u;
    else {
        const float cf = (float)c, sf = (float)s;
        const unsigned all_beams = (1u << K.n_beams) - 1u;  // precise_bins implies <= 16 beams
        const int cnt = __popc(rel);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        {
            int slot = incl - cnt;
            unsigned todo = rel;
            while (todo != 0u) {                            // pair: [4:0] owner lane, [8:5] obstacle
                const int j = __ffs(todo) - 1;
                todo &= todo - 1u;
                ring[slot++] = (unsigned short)((unsigned)lane | ((unsigned)j << 5));
            }
        }
        __syncwarp();
        const double* obw = s_ob + w0;
        for (int base = 0; base < total; base += 32) {
            const bool mine = base + lane < total;
            const unsigned ent = mine ? ring[base + lane] : (unsigned)lane;
            const int owner = ent & 31, j = ent >> 5;
            const double ox_ = __shfl_sync(0xffffffffu, x, owner), oy_ = __shfl_sync(0xffffffffu, y, owner);
            const float ocf = __shfl_sync(0xffffffffu, cf, owner), osf = __shfl_sync(0xffffffffu, sf, owner);
            if (mine) {
                const double* ob = obw + owner;
                const float dxf = (float)(ob[j * kBlock] - ox_), dyf = (float)(ob[(max_o + j) * kBlock] - oy_);
                const float rf = (float)ob[(2 * max_o + j) * kBlock];
                float qx = fmaf(ocf, dxf, osf * dyf), qy = fmaf(ocf, dyf, -osf * dxf);      // R^T (centre - pos)
                const float d2f = fmaf(qx, qx, qy * qy), rrf = rf * rf;
                unsigned bm = all_beams;
                const bool inside = d2f < rrf;
                if (fabsf(d2f - rrf) > 2e-3f * fmaxf(1.0f, rrf) && !(inside && d2f < 1.0f)) {
                    float alpha = 1.57079632679f + 2e-3f;                               // inside: the half plane q . dir <= 1e-3
                    if (inside) { qx = -qx; qy = -qy; }
                    else alpha = asin_approx(fminf(1.0f, (rf + 1e-3f) * rsqrtf(d2f))) + 4e-4f;
                    const float ci = (atan2_approx(qy, qx) - K.beam0f) * K.inv_phi;
                    const float wi = fmaf(alpha, K.inv_phi, 1e-3f);
                    const int lo = max(0, (int)ceilf(ci - wi)), hi = min(K.n_beams - 1, (int)floorf(ci + wi));
                    bm = lo <= hi ? (((2u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
                }
                s_bm[(w0 + owner) * kBmStride + j] = (unsigned short)bm;
            }
        }
        __syncwarp();
        if (rel != 0u) {                                    // union over my obstacles; a beam that may be snapped (Q10) is decided exactly
            const unsigned short* my_bm = s_bm + tid * kBmStride;
#pragma unroll
            for (int j4 = 0; j4 < kBmStride / 8; ++j4) {
                const uint4 v = reinterpret_cast<const uint4*>(my_bm)[j4];
                beams |= v.x | v.y | v.z | v.w;
            }
            beams = (beams | (beams >> 16)) & 0xffffu;
            if (sb1) beams |= 1u << (sb1 - 1u);
            if (sb2) beams |= 1u << (sb2 - 1u);
        }
        __syncwarp();                                       // the list buffer is reused for the exact tests below
    }

    // ================= sonar (robot.py:125-198), warp-wide =================
    // Few (env, beam) pairs are flagged (~1.5 lanes of 32 per beam), so the exact fp64 tests are not run in place: every lane
    // appends its flagged beams to a per-warp list in shared memory (slots from a warp prefix sum of the per-lane counts) and
    // the list is evaluated 32 entries at a time with every lane busy -- the evaluator fetches the owner's pose with
    // shuffles, the owner's obstacles and beam masks from shared memory, runs the reference's ordered scan (Q3) over the
    // obstacles whose interval contains the beam and writes the hit into the owner's observation row.
    {
        const double* obw = s_ob + w0;                          // obstacle columns of this warp's environments
        const int n_beams = K.n_beams;
        for (int b0 = 0; b0 < n_beams; b0 += kBeamWord) {       // one pass for <= 16 beams
            unsigned mine_beams = K.precise_bins ? beams : (beams != 0u ? (n_beams - b0 >= kBeamWord ? 0xffffu : (1u << (n_beams - b0)) - 1u) : 0u);
            const int cnt = __popc(mine_beams);
            int incl = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;                           // warp-uniform
            int slot = incl - cnt;
            const unsigned ent0 = (unsigned)lane | ((unsigned)b0 << 5);
            while (mine_beams != 0u) {                          // list entry: [4:0] owner lane, [12:5] beam
                const int bl = __ffs(mine_beams) - 1;
                mine_beams &= mine_beams - 1u;
                ring[slot++] = (unsigned short)(ent0 + ((unsigned)bl << 5));
            }
            __syncwarp();                                       // list entries, masks and zero-filled rows of other lanes are visible
            for (int base = 0; base < total; base += 32) {
                const bool mine = base + lane < total;
                const unsigned ent = mine ? ring[base + lane] : (unsigned)lane;
                const int owner = ent & 31, bb = (ent >> 5) & 0xff;
                // the owner's pose after the sub-steps
                const double ox_ = __shfl_sync(0xffffffffu, x, owner), oy_ = __shfl_sync(0xffffffffu, y, owner);
                const double oc = __shfl_sync(0xffffffffu, c, owner), os = __shfl_sync(0xffffffffu, s, owner);
                const double oth = __shfl_sync(0xffffffffu, th, owner);
                if (mine) {
                    const double* ob = obw + owner;
                    double bx = K.beam_cos[bb], by = K.beam_sin[bb];    // beam direction in the robot frame
                    bool scan_all = !K.precise_bins;
                    const unsigned orel = s_rel[w0 + owner];
                    if (bb + 1 == (int)((orel >> 16) & 0xffu) || bb + 1 == (int)(orel >> 24)) {
                        const double ang = oth + K.beam_angle[bb];       // robot.py:131 (not wrapped)
                        if (fabs(ang - 0.5 * MNV_PI) < 1e-03) { bx = os; by = oc; scan_all = true; }            // Q10: exactly (0,+1) in the world frame
                        else if (fabs(ang - 1.5 * MNV_PI) < 1e-03) { bx = -os; by = -oc; scan_all = true; }     // Q10: exactly (0,-1)
                    }
                    // obstacles whose beam interval contains this beam (all obstacles within reach for a snapped beam, whose
                    // direction is not the table's), in list order
                    unsigned mm = orel & 0xffffu;
                    if (!scan_all) {
                        const unsigned short* bm = s_bm + (w0 + owner) * kBmStride;
                        unsigned m = 0u;
#pragma unroll
                        for (int j4 = 0; j4 < kBmStride / 8; ++j4) {
                            const uint4 v = reinterpret_cast<const uint4*>(bm)[j4];
                            const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                m |= ((wv[q] >> bb) & 1u) << (8 * j4 + 2 * q);
                                m |= ((wv[q] >> (16 + bb)) & 1u) << (8 * j4 + 2 * q + 1);
                            }
                        }
                        mm = m;
                    }
                    // the reference's ordered scan (Q3, robot.py:149-195) over those obstacles; usually one pass
                    bool hit = false;
                    double best = INFINITY;
                    while (mm != 0u) {
                        const int j = __ffs(mm) - 1;
                        mm &= mm - 1u;
                        const double r = ob[(2 * max_o + j) * kBlock];
                        const double dx = ob[j * kBlock] - ox_, dy = ob[(max_o + j) * kBlock] - oy_;
                        const double ax = fma(oc, dx, os * dy), ay = fma(oc, dy, -os * dx);
                        const double tc = fma(ax, bx, ay * by);
                        const double cr = fma(ax, by, -ay * bx);
                        const double disc = fma(-cr, cr, r * r);
                        if (disc < 0.0) continue;                    // robot.py:172-174
                        const double hh = fast_sqrt(disc);
                        const double tj = tc > 0.0 ? tc - hh : tc + hh;   // nearer root first (robot.py:184)
                        if (fabs(tj) > K.range) continue;            // robot.py:185-187
                        if (tj < 0.0) continue;                      // robot.py:188-190
                        if (hit && tj >= best) break;                // robot.py:192-195
                        best = tj; hit = true;
                    }
                    if (hit)                                          // marinenav_env.py:314-317 (rows are 8-byte aligned: obs_dim is even)
                        *reinterpret_cast<float2*>(s_obs + (w0 + owner) * D + 4 + 2 * bb) = make_float2((float)(best * bx), (float)(best * by));
                }
            }
            __syncwarp();                                       // the list is rewritten by the next pass
        }
    }

    if (STEP && live) {
        // ---- termination priority (marinenav_env.py:240-257, Q5) ----
        int done = 0, info = MNV_INFO_NORMAL;
        const bool oob = (x < 0.0 || x > K.width) || (y < 0.0 || y > K.height);
        if (K.set_boundary && oob) { done = 1; info = MNV_INFO_OUT_OF_BOUNDARY; }
        else if (ep >= K.max_ep_steps) { done = 1; info = MNV_INFO_TOO_LONG; }
        else if (collides(best_d2, best_r + K.robot_r)) { reward += K.pen_coll; done = 1; info = MNV_INFO_COLLISION; }
        else if (reaches(dis_after2, K.goal_dis)) { reward += K.rew_goal; done = 1; info = MNV_INFO_REACH_GOAL; }   // marinenav_env.py:338-342, decided on the squares
        P.state[e] = x; P.state[E + e] = y; P.state[2 * E + e] = th; P.state[3 * E + e] = sp;
        P.velocity[e] = vx; P.velocity[E + e] = vy;
        P.ep_step[e] = ep + 1;                                // marinenav_env.py:259
        P.reward[e] = (float)reward;
        P.done[e] = (uint8_t)done;
        P.info[e] = (uint8_t)info;
    }

    // ---- observation rows: every warp streams out its own 32 rows (one contiguous 32*D*4-byte block, float4 stores);
    //      only a warp-level sync is needed because a warp's rows were written by that warp ----
    __syncwarp();
    const long long left = E - (e0 + w0);
    const int rows = left < 32 ? (left < 0 ? 0 : (int)left) : 32;
    const int n = rows * D;
    const float* src = s_obs + w0 * D;                            // w0*D*4 bytes: 16-byte aligned (32*4 % 16 == 0)
    float* dst = P.obs + (e0 + w0) * D;
    if (!STEP && P.mask != nullptr) {                             // masked observe: only the selected rows
        for (int i = lane; i < n; i += 32)
            if (P.mask[e0 + w0 + i / D] != 0) dst[i] = src[i];
        return;
    }
    const int n4 = n >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(src) + lane;
    float4* dst4 = reinterpret_cast<float4*>(dst) + lane;
    const int full = n4 >> 5;                                      // whole 32-lane trips (6 for 32 rows of 26 floats)
#pragma unroll 6
    for (int k = 0; k < full; ++k) dst4[k * 32] = src4[k * 32];
    if ((full << 5) + lane < n4) dst4[full * 32] = src4[full * 32];
    for (int i = (n4 << 2) + lane; i < n; i += 32) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------------------------
// Dense maps (BASELINE configs[4]: 32 obstacles x 64 beams, few thousand envs per GPU).  A thread-per-env mapping leaves
// a 16 384-env batch with 3.5 warps per SM and 2 048 serial ray/circle tests per thread; a warp per env makes the sonar
// parallel but would spend a whole warp on each environment's serial fp64 integration (40 % of the step).  So one CTA
// (4 warps) owns 8 environments and works in TWO PHASES inside the one launch:
//   phase 1  warp 0, one LANE per environment (8 lanes busy): loads, sincos, the N sub-steps (same code path as
//            mnv_env_kernel); the pose after the sub-steps goes to shared memory.  The other warps have their obstacle table
//            in flight meanwhile; CTAs of different waves are in different phases, so the SM stays busy.
//   phase 2  every warp takes 2 of the environments in turn, one WARP per environment, lane j owning obstacle j:
//   * candidate beams: every lane computes ONE conservative interval of beam indices for its obstacle (the obstacle-major
//     binning of mnv_env_kernel: fp32 atan2 / asin bounds with margins -- it can only add beams);
//   * the candidate (beam, obstacle) pairs go beam-major into batches of <= 32 for the exact fp64 tests: per pass of 32
//     beams, lane L collects the candidate obstacles of beam L (one ballot per beam), a prefix sum over the lanes' counts
//     gives every beam its slot range and a batch is the longest run of whole beams that fits 32 slots (one ballot per
//     batch); the batch is evaluated one pair per lane (the obstacle's robot-frame centre comes from its owner lane by shuffle);
//   * the reference's ordered first-hit scan (robot.py:192-195, Q3) is rebuilt per beam from warp ballots restricted to
//     the beam's SEGMENT of the batch (match.any on the beam index; a batch only holds complete beams):
//       V  = ballot(valid hit) & segment                          valid = real root, 0 <= t <= range (robot.py:172-190)
//       p(l) = highest set bit of V below lane l                  the previously recorded hit when the scan reaches l
//       Bk = ballot(valid_l && p(l) exists && t_l >= t_p(l)) & segment     the scan breaks at the first such l
//       result = Bk ? p(ffs(Bk)) : highest set bit of V           (the recorded hits form a strictly decreasing run)
//   * nearest-centre collision test (Q4) = warp arg-min.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDenseWarps = 4;
constexpr int kDenseEnvs = 8;                                      // environments per CTA (= active lanes of the integrating warp; 4 | 8 | 16 | 32 measured: 69.2 | 68.0 | 73.0 | 83.6 us)

struct DensePose { double x, y, th, sp, c, s, vx, vy, reward, dis_after, gx, gy; int ep, live, pad0, pad1; };

template <bool STEP, int MAXC>
__global__ void __launch_bounds__(kDenseWarps * 32)
mnv_env_dense_kernel(const EnvPtrs P, const __grid_constant__ KParams K)
{
    extern __shared__ __align__(16) unsigned char s_raw[];        // [kDenseEnvs] DensePose | [kDenseWarps][obs_dim] floats | [kDenseWarps][32] ring words
    pdl_wait();                                                    // everything below reads what earlier launches wrote
    pdl_launch_dependents();
    const long long E = K.E;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long e0 = (long long)blockIdx.x * kDenseEnvs;
    const int D = K.obs_dim;
    DensePose* s_pose = reinterpret_cast<DensePose*>(s_raw);
    float* my_obs = reinterpret_cast<float*>(s_pose + kDenseEnvs) + w * D;
    unsigned* ring = reinterpret_cast<unsigned*>(reinterpret_cast<float*>(s_pose + kDenseEnvs) + ((kDenseWarps * D + 3) & ~3)) + w * 32;
    const int max_o = K.max_o;

    // this warp's environments in phase 2: w, w + kDenseWarps, ...; lane j < max_o owns obstacle j
    auto load_obstacle = [&](long long e, double& ox, double& oy, double& orad) {
        ox = 0.0; oy = 0.0; orad = -1.0;
        if (e < E && lane < max_o) {
            ox = __ldg(P.obst + (long long)lane * E + e);
            oy = __ldg(P.obst + (long long)(max_o + lane) * E + e);
            orad = __ldg(P.obst + (long long)(2 * max_o + lane) * E + e);
        }
    };
    double nox, noy, norad;                                        // obstacle table of the next environment of this warp (prefetch)
    load_obstacle(e0 + w, nox, noy, norad);

    // ================= phase 1: warp 0, one lane per environment =================
    if (w == 0) {
        const long long e = e0 + lane;
        const bool live = (lane < kDenseEnvs) && (e < E) && (STEP || P.mask == nullptr || P.mask[e] != 0);
        DensePose ps;
        ps.live = live ? 1 : 0; ps.ep = 0;
        ps.x = ps.y = ps.th = ps.sp = ps.vx = ps.vy = ps.reward = ps.dis_after = ps.gx = ps.gy = 0.0; ps.c = 1.0; ps.s = 0.0;
        if (live) {
            double x = P.state[e], y = P.state[E + e], th = P.state[2 * E + e], sp = P.state[3 * E + e];
            const double gx = P.goal[e], gy = P.goal[E + e];
            int action = 0, ep = 0;
            if (STEP) { action = P.action[e]; ep = P.ep_step[e]; }
            double cx[MAXC], cy[MAXC], ck[MAXC];
            {
                const double* pc = P.cores + e;
                const long long rowstride = (long long)K.max_c * E;
#pragma unroll
                for (int i = 0; i < MAXC; ++i) {
                    if (i < K.max_c) { cx[i] = __ldg(pc); cy[i] = __ldg(pc + rowstride); ck[i] = __ldg(pc + 2 * rowstride); }
                    else { cx[i] = 0.0; cy[i] = 0.0; ck[i] = 0.0; }
                    pc += E;
                }
            }
            double c, s;
            sincos_heading(th, &s, &c);
#pragma unroll
            for (int i = 0; i < MAXC; ++i) ck[i] *= (1.0 / (2.0 * MNV_PI));
            auto current = [&](double px, double py, double& ux, double& uy) {
                ux = 0.0; uy = 0.0;
#pragma unroll
                for (int i = 0; i < MAXC; ++i) {
                    const double dx = cx[i] - px, dy = cy[i] - py;
                    const double d2 = fma(dx, dx, dy * dy);
                    const double f = ck[i] * (d2 <= K.core_r2 ? K.inv_core_r2 : fast_rcp(d2));   // marinenav_env.py:461-465
                    ux = fma(-dy, f, ux);
                    uy = fma(dx, f, uy);
                }
            };
            double vx = 0.0, vy = 0.0, reward = 0.0, dis_after = 0.0;
            if (STEP) {
                const int ai = action / 3, wi = action - 3 * ai;
                const double acc = K.accel[ai], wdt = K.wdt[wi], cw = K.cos_wdt[wi], sw = K.sin_wdt[wi];
                const double dis_before = fast_sqrt(fma(gx - x, gx - x, (gy - y) * (gy - y)));
                for (int it = 0; it < K.n_substeps; ++it) {
                    double ux, uy;
                    current(x, y, ux, uy);
                    vx = fma(sp, c, ux); vy = fma(sp, s, uy);              // robot.py:98-100 (Q6)
                    x = fma(vx, K.dt, x); y = fma(vy, K.dt, y);            // robot.py:105-107
                    sp = __dadd_rn(sp, __dmul_rn(__dsub_rn(acc, __dmul_rn(K.k_drag, sp)), K.dt));   // robot.py:113
                    sp = sp < 0.0 ? 0.0 : sp;                              // robot.py:114
                    sp = sp > K.max_speed ? K.max_speed : sp;
                    th = __dadd_rn(th, wdt);                               // robot.py:117
                    th = th < 0.0 ? __dadd_rn(th, 2.0 * MNV_PI) : th;      // robot.py:120-123 (see mnv_env_kernel)
                    th = th >= 2.0 * MNV_PI ? __dsub_rn(th, 2.0 * MNV_PI) : th;
                    if (th < 0.0 || th >= 2.0 * MNV_PI) {
#pragma unroll 1
                        while (th < 0.0) th += 2.0 * MNV_PI;
#pragma unroll 1
                        while (th >= 2.0 * MNV_PI) th -= 2.0 * MNV_PI;
                    }
                    const double c2 = fma(c, cw, -s * sw), s2 = fma(s, cw, c * sw);
                    c = c2; s = s2;
                    if (P.traj != nullptr) { P.traj[(long long)(2 * it) * E + e] = x; P.traj[(long long)(2 * it + 1) * E + e] = y; }
                }
                dis_after = fma(gx - x, gx - x, (gy - y) * (gy - y));        // kept SQUARED for the goal test (reaches())
                reward = K.pen_step + (dis_before - fast_sqrt(dis_after));  // marinenav_env.py:220,229
            } else {
                if (K.velocity_from_state) {
                    double ux, uy;
                    current(x, y, ux, uy);
                    vx = fma(sp, c, ux); vy = fma(sp, s, uy);
                    P.velocity[e] = vx; P.velocity[E + e] = vy;
                } else { vx = P.velocity[e]; vy = P.velocity[E + e]; }
            }
            ps.x = x; ps.y = y; ps.th = th; ps.sp = sp; ps.c = c; ps.s = s; ps.vx = vx; ps.vy = vy;
            ps.reward = reward; ps.dis_after = dis_after; ps.gx = gx; ps.gy = gy; ps.ep = ep;
        }
        if (lane < kDenseEnvs) s_pose[lane] = ps;
    }
    __syncthreads();

    // ================= phase 2: one warp per environment =================
    const int n_beams = K.n_beams;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int el = w; el < kDenseEnvs; el += kDenseWarps) {
        const long long e = e0 + el;
        const double ox = nox, oy = noy, orad = norad;
        load_obstacle(e + kDenseWarps < e0 + kDenseEnvs ? e + kDenseWarps : E, nox, noy, norad);   // prefetch the next one
        const DensePose& ps = s_pose[el];
        if (!ps.live) continue;                                    // warp-uniform
        const double x = ps.x, y = ps.y, th = ps.th, c = ps.c, s = ps.s;

        if (lane == 0) {                                           // observation head (marinenav_env.py:278-293)
            my_obs[0] = (float)fma(c, ps.vx, s * ps.vy);
            my_obs[1] = (float)fma(c, ps.vy, -s * ps.vx);
            my_obs[2] = (float)fma(c, ps.gx - x, s * (ps.gy - y));
            my_obs[3] = (float)fma(c, ps.gy - y, -s * (ps.gx - x));
        }
        for (int b = lane; b < n_beams; b += 32)                   // "no return" everywhere (marinenav_env.py:318-320); hits overwrite
            *reinterpret_cast<float2*>(my_obs + 4 + 2 * b) = make_float2(0.0f, 0.0f);

        // ---- my obstacle in the robot frame ----
        const bool on = orad > 0.0;
        const double dxo = ox - x, dyo = oy - y;
        const double d2o = fma(dxo, dxo, dyo * dyo);
        const double qx = fma(c, dxo, s * dyo), qy = fma(c, dyo, -s * dxo), rr = orad * orad;
        const double lim = K.range_slack + orad;
        const bool relevant = on && d2o <= lim * lim;              // reachable within the sonar range at all
        const bool borderline = relevant && fabs(d2o - rr) <= 1e-9 * rr;   // robot on the circle: never filter
        // ---- Q4: collision against the nearest CENTRE only (marinenav_env.py:329-336): warp arg-min ----
        double best_d2 = on ? d2o : INFINITY, best_r = orad;
        int best_i = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, best_d2, off), orr = __shfl_xor_sync(0xffffffffu, best_r, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
            // ties: the lowest obstacle index wins, like the first minimum of a sequential scan
            if (od < best_d2 || (od == best_d2 && oi < best_i)) { best_d2 = od; best_r = orr; best_i = oi; }
        }

        // ---- Q10 pre-test (see mnv_env_kernel): the only beams that can be snapped to the vertical ----
        int bs1, bs2;
        {
            const float u1 = (float)(K.snap_t1 - th) * K.inv_phi, u2 = (float)(K.snap_t2 - th) * K.inv_phi;
            const float n1 = rintf(u1), n2 = rintf(u2);
            bs1 = fabsf(u1 - n1) < K.snap_tol ? (int)n1 : -1;
            bs2 = fabsf(u2 - n2) < K.snap_tol ? (int)n2 : -1;
        }

        // ---- sonar ----
        int n_pend = 0;                                            // warp-uniform
        auto drain = [&]() {
            __syncwarp();
            const bool mine = lane < n_pend;
            const unsigned ent = mine ? ring[lane] : 0u;           // [4:0] obstacle (owner lane), [12:5] beam
            const int j = ent & 31, bb = ent >> 5;
            const double oqx = __shfl_sync(0xffffffffu, qx, j), oqy = __shfl_sync(0xffffffffu, qy, j), orr2 = __shfl_sync(0xffffffffu, rr, j);
            bool valid = false;
            double t = 0.0, bx = 0.0, by = 0.0;
            if (mine) {                                            // exact fp64 decision (robot.py:164-190)
                bx = K.beam_cos[bb]; by = K.beam_sin[bb];
                if (bb == bs1 || bb == bs2) {
                    const double ang = th + K.beam_angle[bb];      // robot.py:131 (not wrapped)
                    if (fabs(ang - 0.5 * MNV_PI) < 1e-03) { bx = s; by = c; }             // Q10
                    else if (fabs(ang - 1.5 * MNV_PI) < 1e-03) { bx = -s; by = -c; }
                }
                const double tc = fma(oqx, bx, oqy * by), cr = fma(oqx, by, -oqy * bx);
                const double disc = fma(-cr, cr, orr2);
                if (disc >= 0.0) {
                    const double h = fast_sqrt(disc);
                    t = tc > 0.0 ? tc - h : tc + h;                // nearer root first (robot.py:184)
                    valid = (t <= K.range) && (t >= 0.0);
                }
            }
            // ordered scan per beam = per segment of equal beam index (entries are in (beam, obstacle) order)
            const unsigned seg = __match_any_sync(0xffffffffu, mine ? bb : 0x1000 + lane);
            const unsigned V = __ballot_sync(0xffffffffu, valid) & seg;
            const unsigned below = V & lt_mask;
            const int pj = below ? 31 - __clz(below) : -1;         // previous valid hit of my beam in list order
            const double tp = __shfl_sync(0xffffffffu, t, pj < 0 ? 0 : pj);
            const unsigned Bk = __ballot_sync(0xffffffffu, valid && pj >= 0 && t >= tp) & seg;
            const int f = Bk ? __ffs(Bk) - 1 : 0;
            const int pf = __shfl_sync(0xffffffffu, pj, f);        // predecessor of the first breaking lane (>= 0 when Bk != 0)
            const int res = Bk ? pf : (V ? 31 - __clz(V) : -1);
            if (mine && lane == res)                               // marinenav_env.py:314-317
                *reinterpret_cast<float2*>(my_obs + 4 + 2 * bb) = make_float2((float)(t * bx), (float)(t * by));
            n_pend = 0;
            __syncwarp();
        };
        // ---- which beams can my obstacle return a hit on?  ONE conservative interval of beam indices per obstacle
        //      (the obstacle-major binning of mnv_env_kernel: robot outside the circle -> |beta_b - atan2(q)| < asin(r / d),
        //      inside -> the half plane q . dir <= 0; fp32 with margins, so it can only add beams); wrap-prone geometries
        //      (field of view >= pi) flag every beam of every obstacle within reach. ----
        int blo = 1, bhi = 0;                                      // candidate beams [blo, bhi] (empty unless within reach)
        if (relevant) {
            blo = 0; bhi = n_beams - 1;
            float fqx = (float)qx, fqy = (float)qy;
            const float rf = (float)orad, d2f = fmaf(fqx, fqx, fqy * fqy), rrf = rf * rf;
            const bool inside = d2f < rrf;
            if (K.dense_bins && !borderline && fabsf(d2f - rrf) > 2e-3f * fmaxf(1.0f, rrf) && !(inside && d2f < 1.0f)) {
                float alpha = 1.57079632679f + 2e-3f;
                if (inside) { fqx = -fqx; fqy = -fqy; }
                else alpha = asin_approx(fminf(1.0f, (rf + 1e-3f) * rsqrtf(d2f))) + 4e-4f;
                const float ci = (atan2_approx(fqy, fqx) - K.beam0f) * K.inv_phi;
                const float wi = fmaf(alpha, K.inv_phi, 1e-3f);
                blo = max(0, (int)ceilf(ci - wi)); bhi = min(n_beams - 1, (int)floorf(ci + wi));
            }
        }
        const int sn1 = (relevant && bs1 >= 0 && bs1 < n_beams) ? bs1 : -1000;   // possibly snapped beams (Q10): their direction is
        const int sn2 = (relevant && bs2 >= 0 && bs2 < n_beams) ? bs2 : -1000;   // not the table's -> every obstacle within reach
        // ---- (beam, obstacle) candidate pairs -> batches of <= 32 for the exact tests, beam-major, 32 beams per pass.
        //      Lane L owns beam 32 pass + L: its candidate obstacles are one ballot; a prefix sum over the lanes' counts gives
        //      every beam its slot range, and a batch = the longest run of whole beams that fits 32 slots (one ballot). ----
        for (int b0 = 0; b0 < n_beams; b0 += 32) {
            const int nb = n_beams - b0 < 32 ? n_beams - b0 : 32;
            unsigned mybits = 0u;                                  // my candidate beams of this pass
            {
                const int l = max(blo - b0, 0), h = min(bhi - b0, 31);
                if (l <= h) mybits = ((2u << h) - 1u) & ~((1u << l) - 1u);
                if (sn1 >= b0 && sn1 < b0 + 32) mybits |= 1u << (sn1 - b0);
                if (sn2 >= b0 && sn2 < b0 + 32) mybits |= 1u << (sn2 - b0);
            }
            unsigned cand = 0u;                                    // lane L: obstacles (lanes) that may return on beam b0 + L
            for (int b = 0; b < nb; ++b) {
                const unsigned bal = __ballot_sync(0xffffffffu, (mybits >> b) & 1u);
                if (lane == b) cand = bal;
            }
            const int cnt = __popc(cand);
            int incl = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += v;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            int base = 0, start = 0;                               // slots / beams already handed to earlier batches
            while (base < total) {
                const unsigned fits = __ballot_sync(0xffffffffu, lane >= start && incl - base <= 32);
                const int end = 32 - __clz(fits);                  // beams [start, end) form this batch (>= 1 beam: cnt <= 32)
                if (lane >= start && lane < end) {
                    int slot = incl - cnt - base;
                    unsigned m = cand;
                    while (m != 0u) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1u;
                        ring[slot++] = (unsigned)j | ((unsigned)(b0 + lane) << 5);
                    }
                }
                n_pend = __shfl_sync(0xffffffffu, incl, end - 1) - base;
                base += n_pend; start = end;
                if (n_pend > 0) drain();
            }
        }

        if (STEP && lane == 0) {
            double reward = ps.reward;
            int done = 0, info = MNV_INFO_NORMAL;                  // marinenav_env.py:240-257 (Q5)
            const bool oob = (x < 0.0 || x > K.width) || (y < 0.0 || y > K.height);
            if (K.set_boundary && oob) { done = 1; info = MNV_INFO_OUT_OF_BOUNDARY; }
            else if (ps.ep >= K.max_ep_steps) { done = 1; info = MNV_INFO_TOO_LONG; }
            else if (collides(best_d2, best_r + K.robot_r)) { reward += K.pen_coll; done = 1; info = MNV_INFO_COLLISION; }
            else if (reaches(ps.dis_after, K.goal_dis)) { reward += K.rew_goal; done = 1; info = MNV_INFO_REACH_GOAL; }   // ps.dis_after = squared distance
            P.state[e] = x; P.state[E + e] = y; P.state[2 * E + e] = th; P.state[3 * E + e] = ps.sp;
            P.velocity[e] = ps.vx; P.velocity[E + e] = ps.vy;
            P.ep_step[e] = ps.ep + 1;
            P.reward[e] = (float)reward;
            P.done[e] = (uint8_t)done;
            P.info[e] = (uint8_t)info;
        }
        __syncwarp();
        float* dst = P.obs + e * D;                                // 8-byte aligned rows (D even): float2 stores
        for (int i = lane; i < (D >> 1); i += 32)
            reinterpret_cast<float2*>(dst)[i] = reinterpret_cast<const float2*>(my_obs)[i];
        __syncwarp();                                              // the row buffer is reused by this warp's next environment
    }
}

// One launch, optionally with programmatic stream serialization (PDL): the kernels call griddepcontrol.launch_dependents
// at entry and griddepcontrol.wait before their first global access, so the next launch's CTAs are placed while this grid drains.
template <typename Kern>
cudaError_t launch_one(Kern kern, unsigned grid, unsigned block, size_t smem, cudaStream_t st, const EnvPtrs& P, const KParams& K)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (mnv_option(MNV_OPT_PDL) || K.pdl_prefetch) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, P, K);
}

template <bool STEP>
int launch_env(const EnvPtrs& P, const KParams& K, cudaStream_t st)
{
    if (K.max_o > 16) {                                            // dense maps: one warp per environment
        const unsigned dgrid = (unsigned)((K.E + kDenseEnvs - 1) / kDenseEnvs);
        const size_t dsmem = (size_t)kDenseEnvs * sizeof(DensePose) + (size_t)((kDenseWarps * K.obs_dim + 3) & ~3) * sizeof(float) +
                             (size_t)kDenseWarps * 32 * sizeof(unsigned);
        if (K.max_c <= 4) launch_one(mnv_env_dense_kernel<STEP, 4>, dgrid, kDenseWarps * 32, dsmem, st, P, K);
        else launch_one(mnv_env_dense_kernel<STEP, 8>, dgrid, kDenseWarps * 32, dsmem, st, P, K);
        return mnv_launch_status(STEP ? "mnv_step(dense)" : "mnv_observe(dense)");
    }
    const unsigned grid = (unsigned)((K.E + kBlock - 1) / kBlock);
#define MNV_LAUNCH(MC, MO)                                                                                   \
    do {                                                                                                     \
        const size_t smem = (size_t)((kBlock * K.obs_dim + 3) & ~3) * sizeof(float) + (size_t)3 * K.max_o * kBlock * sizeof(double) + \
                            8 * (kBlock / 32) + (size_t)(kBlock / 32) * kRing * sizeof(unsigned short) +    \
                            (size_t)kBlock * ((MO + 7) & ~7) * sizeof(unsigned short) + (size_t)kBlock * sizeof(unsigned); \
        auto kern = (STEP && K.pdl_prefetch) ? mnv_env_kernel<MC, MO, STEP, STEP> : mnv_env_kernel<MC, MO, STEP, false>; \
        if (smem > 48 * 1024) {                                                                              \
            cudaError_t a = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (a != cudaSuccess) { mnv_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(a)); return (int)a; } \
        }                                                                                                    \
        launch_one(kern, grid, kBlock, smem, st, P, K);                                                      \
    } while (0)
    if (K.max_c <= 4 && K.max_o <= 8) MNV_LAUNCH(4, 8);
    else if (K.max_c <= 8 && K.max_o <= 10) MNV_LAUNCH(8, 10);
    else MNV_LAUNCH(8, 16);
#undef MNV_LAUNCH
    return mnv_launch_status(STEP ? "mnv_step" : "mnv_observe");
}

int fill_kparams(KParams& K, const mnv_params* p, int64_t E, int max_c, int max_o)
{
    if (p == nullptr) { mnv_set_error("params is null"); return MNV_E_NULL; }
    if (E <= 0) { mnv_set_error("E must be > 0 (got %lld)", (long long)E); return MNV_E_SIZE; }
    if (max_c < 0 || max_c > MNV_MAX_CORES || max_o < 0 || max_o > MNV_MAX_OBSTACLES) {
        mnv_set_error("max_c=%d / max_o=%d exceed the compiled capacity (%d / %d)", max_c, max_o, MNV_MAX_CORES, MNV_MAX_OBSTACLES);
        return MNV_E_CAPACITY;
    }
    if (p->n_beams < 2 || p->n_beams > MNV_MAX_BEAMS) { mnv_set_error("n_beams=%d outside [2,%d]", p->n_beams, MNV_MAX_BEAMS); return MNV_E_CAPACITY; }
    if (p->n_substeps < 0 || !(p->dt > 0.0) || !(p->core_r > 0.0)) { mnv_set_error("bad dt / n_substeps / core_r"); return MNV_E_PARAM; }
    memset(&K, 0, sizeof(K));
    K.dt = p->dt; K.n_substeps = p->n_substeps;
    for (int i = 0; i < 3; ++i) {
        K.accel[i] = p->accel[i];
        K.wdt[i] = p->yaw_rate[i] * p->dt;
        K.cos_wdt[i] = cos(K.wdt[i]); K.sin_wdt[i] = sin(K.wdt[i]);
    }
    K.k_drag = p->k_drag; K.max_speed = p->max_speed; K.robot_r = p->robot_r;
    K.core_r2 = p->core_r * p->core_r; K.inv_core_r2 = 1.0 / K.core_r2; K.goal_dis = p->goal_dis;
    K.pen_step = p->timestep_penalty; K.pen_coll = p->collision_penalty; K.rew_goal = p->goal_reward;
    K.range = p->sonar_range; K.range_slack = p->sonar_range * (1.0 + 1e-9) + 1e-9;
    K.width = p->width; K.height = p->height;
    K.n_beams = p->n_beams; K.max_ep_steps = p->max_episode_steps; K.set_boundary = p->set_boundary;
    K.max_c = max_c; K.max_o = max_o; K.obs_dim = 4 + 2 * p->n_beams; K.E = E;
    K.allow_tma = mnv_option(MNV_OPT_TMA) ? 1 : 0;
    K.pdl_prefetch = (p->pdl_prefetch != 0 || mnv_option(MNV_OPT_PDL) >= 2) ? 1 : 0;
    // "tma" = 1 stages the obstacle rows with the TMA bulk-copy engine (UBLKCP) instead of per-thread cp.async (LDGSTS).
    // Measured A/B in one process (profiles/README.md): 25.03 us vs 24.18 us per step -> cp.async is the default.
    // Sonar.compute_phi / compute_beam_angles (robot.py:14-21)
    const double phi = p->sonar_angle / (p->n_beams - 1), a0 = -p->sonar_angle / 2;
    for (int i = 0; i < p->n_beams; ++i) {
        K.beam_angle[i] = a0 + i * phi;
        K.beam_cos[i] = cos(K.beam_angle[i]); K.beam_sin[i] = sin(K.beam_angle[i]);
    }
    K.snap_t1 = 0.5 * MNV_PI - a0; K.snap_t2 = 1.5 * MNV_PI - a0;      // Q10 pre-test (see the kernel)
    K.inv_phi = (float)(1.0 / phi); K.snap_tol = (float)(1e-3 / phi + 1e-4);
    K.beam0f = (float)a0;
    K.precise_bins = (p->n_beams <= 16 && p->sonar_angle < MNV_PI - 0.01 && p->sonar_angle > 0.0) ? 1 : 0;
    K.dense_bins = (p->sonar_angle < MNV_PI - 0.01 && p->sonar_angle > 0.0) ? 1 : 0;
    return 0;
}

}  // namespace

extern "C" int mnv_step(double* d_state, double* d_velocity, const double* d_goal, const double* d_cores,
                        const double* d_obstacles, const int32_t* d_action, int32_t* d_episode_step,
                        float* d_obs, float* d_reward, uint8_t* d_done, uint8_t* d_info, double* d_trajectory,
                        int64_t E, int32_t max_c, int32_t max_o, const mnv_params* p, void* stream)
{
    KParams K;
    int rc = fill_kparams(K, p, E, max_c, max_o);
    if (rc) return rc;
    MNV_CHECK_PTR(d_state); MNV_CHECK_PTR(d_velocity); MNV_CHECK_PTR(d_goal);
    if (max_c > 0) MNV_CHECK_PTR(d_cores);
    if (max_o > 0) MNV_CHECK_PTR(d_obstacles);
    MNV_CHECK_PTR(d_action); MNV_CHECK_PTR(d_episode_step); MNV_CHECK_PTR(d_obs); MNV_CHECK_PTR(d_reward);
    MNV_CHECK_PTR_OPT(d_trajectory);
    if (d_done == nullptr || d_info == nullptr) { mnv_set_error("mnv_step: null done/info"); return MNV_E_NULL; }
    if ((E & 1) != 0 && E != 1) { /* fp64 rows stay 8-byte aligned for any E; nothing to do */ }
    EnvPtrs P{d_state, d_velocity, d_goal, d_cores, d_obstacles, d_action, d_episode_step, nullptr, d_trajectory, d_obs, d_reward, d_done, d_info};
    return launch_env<true>(P, K, (cudaStream_t)stream);
}

extern "C" int mnv_observe(const double* d_state, double* d_velocity, const double* d_goal, const double* d_cores,
                           const double* d_obstacles, const uint8_t* d_mask, float* d_obs, int64_t E, int32_t max_c, int32_t max_o,
                           const mnv_params* p, int32_t velocity_from_state, void* stream)
{
    KParams K;
    int rc = fill_kparams(K, p, E, max_c, max_o);
    if (rc) return rc;
    K.velocity_from_state = velocity_from_state ? 1 : 0;
    MNV_CHECK_PTR(d_state); MNV_CHECK_PTR(d_velocity); MNV_CHECK_PTR(d_goal);
    if (max_c > 0) MNV_CHECK_PTR(d_cores);
    if (max_o > 0) MNV_CHECK_PTR(d_obstacles);
    MNV_CHECK_PTR(d_obs);
    EnvPtrs P{const_cast<double*>(d_state), d_velocity, d_goal, d_cores, d_obstacles, nullptr, nullptr, d_mask, nullptr, d_obs, nullptr, nullptr, nullptr};
    return launch_env<false>(P, K, (cudaStream_t)stream);
}

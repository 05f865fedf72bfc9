/*
 * marinenav_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle; never shipped, never on the product path).
 *
 * Scalar fp64 CPU restatement of the reference environment
 *   RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf
 *     marinenav_env/envs/marinenav_env.py (MarineNavEnv)   marinenav_env/envs/utils/robot.py (Robot, Sonar)
 * Each function cites the reference file:line it follows.  The arithmetic deliberately keeps the
 * reference's formulation (tan-slope ray/circle quadratic, normalise-then-scale vortex velocity,
 * ascending-distance summation, R^T p + t_rw frame change) so that it tracks the reference to ~1e-13;
 * the CUDA product path uses a different (robot-centred, division-light) formulation and is compared
 * against this file within the tolerance stated in the tests.
 *
 * PARITY PINNED (see marinenav_oracle.h): tests/test_oracle_golden.py (reference golden vectors, recorded episodes, eval_config KAT).
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fPIC -shared -pthread).
 */
#include "marinenav_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------
 * numpy.random.RandomState (legacy MT19937).  RandomState(int) -> init_genrand; random_sample() ->
 * 53-bit double (a>>5, b>>6); uniform(lo,hi) = lo + (hi-lo)*U; binomial(1,.5) = [U > 0.5] (inversion,
 * one draw).  Used by marinenav_env.py:75-78,114-139,158-162,190-192.
 * ---------------------------------------------------------------------------------------------- */
void orc_mt_seed(orc_mt19937* s, uint32_t seed)
{
    s->key[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->key[i] = 1812433253u * (s->key[i - 1] ^ (s->key[i - 1] >> 30)) + (uint32_t)i;
    s->pos = 624;
}

static void mt_generate(orc_mt19937* s)
{
    uint32_t* mt = s->key;
    const uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAG = 0x9908b0dfu;
    int kk;
    uint32_t y;
    for (kk = 0; kk < 624 - 397; ++kk) {
        y = (mt[kk] & UP) | (mt[kk + 1] & LO);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
    }
    for (; kk < 623; ++kk) {
        y = (mt[kk] & UP) | (mt[kk + 1] & LO);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
    }
    y = (mt[623] & UP) | (mt[0] & LO);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAG : 0u);
    s->pos = 0;
}

static uint32_t mt_next32(orc_mt19937* s)
{
    if (s->pos >= 624) mt_generate(s);
    uint32_t y = s->key[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

double orc_mt_random_sample(orc_mt19937* s)
{
    uint32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

static double rd_uniform(orc_mt19937* s, double lo, double hi) { return lo + (hi - lo) * orc_mt_random_sample(s); }
static int    rd_binomial_half(orc_mt19937* s) { return orc_mt_random_sample(s) > 0.5 ? 1 : 0; }

/* ------------------------------------------------------------------------------------------------
 * construction
 * ---------------------------------------------------------------------------------------------- */
unsigned long orc_sizeof_env(void) { return (unsigned long)sizeof(orc_env); }

/* Sonar.compute_phi / compute_beam_angles, robot.py:14-21 */
void orc_set_num_beams(orc_env* e, int n)
{
    e->num_beams = n;
    double phi = e->sonar_angle / (n - 1);
    double angle = -e->sonar_angle / 2;
    for (int i = 0; i < n; ++i) e->beam_angles[i] = angle + i * phi;
}

void orc_seed(orc_env* e, uint32_t seed) { orc_mt_seed(&e->rd, seed); }   /* marinenav_env.py:75-78 */

/* MarineNavEnv.__init__ marinenav_env.py:27-73 ; Robot.__init__ robot.py:25-49 ; Sonar.__init__ robot.py:5-12 */
void orc_env_init(orc_env* e, uint32_t seed)
{
    memset(e, 0, sizeof(*e));
    orc_seed(e, seed);
    e->width = 50; e->height = 50; e->r = 0.5; e->v_rel_max = 1.0; e->p = 0.8;
    e->v_range[0] = 5; e->v_range[1] = 10; e->obs_r_range[0] = 1; e->obs_r_range[1] = 3;
    e->clear_r = 10.0; e->reset_start_and_goal = 1;
    e->start[0] = 5.0; e->start[1] = 5.0; e->goal[0] = 45.0; e->goal[1] = 45.0;
    e->random_reset_state = 1; e->init_speed = 0.0; e->init_theta = M_PI / 4;
    e->goal_dis = 2.0; e->timestep_penalty = -1.0; e->collision_penalty = -50.0; e->goal_reward = 100.0;
    e->discount = 0.99; e->num_cores = 8; e->num_obs = 5; e->min_start_goal_dis = 25.0;
    e->set_boundary = 0; e->n_sched = 0;
    e->dt = 0.1; e->N = 10; e->robot_r = 0.8; e->max_speed = 2.0;
    e->a[0] = -0.4; e->a[1] = 0.0; e->a[2] = 0.4;
    e->w[0] = -M_PI / 6; e->w[1] = 0.0; e->w[2] = M_PI / 6;
    e->k = e->a[2] / e->max_speed;                       /* compute_k robot.py:51-52 (np.max(a)/max_speed) */
    e->sonar_range = 10.0; e->sonar_angle = 2 * M_PI / 3;
    orc_set_num_beams(e, 11);
    e->robot_init_theta = 0.0; e->robot_init_speed = 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * flow field: get_velocity marinenav_env.py:422-455, compute_speed :461-465
 * Q1: every core contributes (the `continue` at :437-442 only continues the inner loop).
 * Q2: summation in ascending-distance order (KDTree.query(k=n), :427).
 * ---------------------------------------------------------------------------------------------- */
static double compute_speed(double core_r, double Gamma, double d)
{
    if (d <= core_r) return Gamma / (2 * M_PI * core_r * core_r) * d;
    return Gamma / (2 * M_PI * d);
}

static void velocity_from_cores(const orc_core* cores, int n, double core_r, double x, double y, double out[2])
{
    out[0] = 0.0; out[1] = 0.0;
    if (n == 0) return;                                   /* :423-424 */
    int    idx[ORC_MAX_CORES];
    double dd[ORC_MAX_CORES];
    for (int i = 0; i < n; ++i) {                         /* stable insertion sort by distance */
        double dx = cores[i].x - x, dy = cores[i].y - y;
        double d = sqrt(dx * dx + dy * dy);
        int j = i;
        while (j > 0 && dd[j - 1] > d) { dd[j] = dd[j - 1]; idx[j] = idx[j - 1]; --j; }
        dd[j] = d; idx[j] = i;
    }
    double vx = 0.0, vy = 0.0;
    for (int q = 0; q < n; ++q) {
        const orc_core* c = &cores[idx[q]];
        double rx = c->x - x, ry = c->y - y;              /* v_radial :434 */
        double dis = sqrt(rx * rx + ry * ry);             /* np.linalg.norm :444 */
        rx /= dis; ry /= dis;                             /* :445 */
        double tx, ty;
        if (c->clockwise) { tx = -ry; ty = rx; }          /* [[0,-1],[1,0]] :446-447 */
        else              { tx = ry;  ty = -rx; }         /* [[0,1],[-1,0]] :448-449 */
        double speed = compute_speed(core_r, c->Gamma, dis);
        vx += tx * speed; vy += ty * speed;               /* :452 */
    }
    out[0] = vx; out[1] = vy;
}

void orc_get_velocity(const orc_env* e, double x, double y, double out[2])
{
    velocity_from_cores(e->cores, e->n_cores_placed, e->r, x, y, out);
}

/* ------------------------------------------------------------------------------------------------
 * Robot.update_state robot.py:102-123 (update_velocity :98-100, get_steer_velocity :95-96)
 * ---------------------------------------------------------------------------------------------- */
void orc_update_state(orc_env* e, int action, const double current[2])
{
    e->vx = e->speed * cos(e->theta) + current[0];        /* uses pre-update speed/theta (Q6) */
    e->vy = e->speed * sin(e->theta) + current[1];
    e->x += e->vx * e->dt;
    e->y += e->vy * e->dt;
    double a = e->a[action / 3], w = e->w[action % 3];    /* compute_actions :54-55: a outer, w inner */
    e->speed += (a - e->k * e->speed) * e->dt;
    if (e->speed < 0.0) e->speed = 0.0;                   /* np.clip :114 */
    if (e->speed > e->max_speed) e->speed = e->max_speed;
    e->theta += w * e->dt;
    while (e->theta < 0.0) e->theta += 2 * M_PI;          /* :120-123 */
    while (e->theta >= 2 * M_PI) e->theta -= 2 * M_PI;
}

/* ------------------------------------------------------------------------------------------------
 * Robot.sonar_reflection robot.py:125-198.  Q3 ordered break, Q10 vertical snap, nearer-root-first.
 * ---------------------------------------------------------------------------------------------- */
static void sonar_reflection(double x, double y, double theta, const double* beam_angles, int n_beams, double range,
                             const orc_obstacle* obstacles, int n_obs, double* refl_x, double* refl_y, int* refl_hit)
{
    for (int bi = 0; bi < n_beams; ++bi) {
        double angle = theta + beam_angles[bi];                                /* :131, not wrapped */
        int vert_up = fabs(angle - M_PI / 2) < 1e-03;
        int vert = vert_up || fabs(angle - 3 * M_PI / 2) < 1e-03;              /* :134-135 */
        double px, py;
        if (vert) {
            double d = vert_up ? 2.0 : -2.0;
            px = x; py = y + d * range;                                        /* :136-138 */
        } else {
            double d = 2.0;
            px = x + d * range * cos(angle);                                   /* :140-142 */
            py = y + d * range * sin(angle);
        }
        int hit = 0;
        double reflection_dist = INFINITY;                                     /* :147 */
        for (int j = 0; j < n_obs; ++j) {
            const orc_obstacle* ob = &obstacles[j];
            double x1, x2, y1, y2;
            if (vert) {
                double M = ob->r * ob->r - (x - ob->x) * (x - ob->x);          /* :152 */
                if (M < 0.0) continue;
                x1 = x; x2 = x;
                y1 = ob->y - sqrt(M); y2 = ob->y + sqrt(M);
            } else {
                double K = tan(angle);                                         /* :164 */
                double a = 1 + K * K;
                double b = 2 * K * (y - K * x - ob->y) - 2 * ob->x;
                double c = ob->x * ob->x + (y - K * x - ob->y) * (y - K * x - ob->y) - ob->r * ob->r;
                double delta = b * b - 4 * a * c;
                if (delta < 0.0) continue;                                     /* :172-174 */
                x1 = (-b - sqrt(delta)) / (2 * a);
                x2 = (-b + sqrt(delta)) / (2 * a);
                y1 = y + K * (x1 - x);
                y2 = y + K * (x2 - x);
            }
            double v1x = x1 - x, v1y = y1 - y, v2x = x2 - x, v2y = y2 - y;     /* :181-182 */
            double n1 = sqrt(v1x * v1x + v1y * v1y), n2 = sqrt(v2x * v2x + v2y * v2y);
            double vx, vy, nv;
            if (n1 < n2) { vx = v1x; vy = v1y; nv = n1; } else { vx = v2x; vy = v2y; nv = n2; }   /* :184 */
            if (nv > range) continue;                                          /* :185-187 */
            if (vx * cos(angle) + vy * sin(angle) < 0.0) continue;             /* :188-190 */
            if (hit) {
                if (nv >= reflection_dist) break;                              /* :192-195 (Q3) */
            }
            reflection_dist = nv;
            px = vx + x; py = vy + y; hit = 1;                                 /* :197-198 */
        }
        refl_x[bi] = px; refl_y[bi] = py; refl_hit[bi] = hit;
    }
}

void orc_sonar_reflection(orc_env* e)
{
    sonar_reflection(e->x, e->y, e->theta, e->beam_angles, e->num_beams, e->sonar_range,
                     e->obstacles, e->n_obs_placed, e->refl_x, e->refl_y, e->refl_hit);
}

/* ------------------------------------------------------------------------------------------------
 * get_observation marinenav_env.py:273-326 with get_robot_transform robot.py:89-93
 *   R_rw = R_wr^T = [[c, s], [-s, c]] ; t_rw = -R_rw t_wr ; p_r = R_rw p + t_rw
 * ---------------------------------------------------------------------------------------------- */
static void observation_from(double x, double y, double theta, double vx, double vy, double gx, double gy,
                             const double* refl_x, const double* refl_y, const int* refl_hit, int n_beams, double* obs)
{
    double c = cos(theta), s = sin(theta);
    double tx = -(c * x + s * y), ty = -(-s * x + c * y);                       /* :280-281 */
    obs[0] = c * vx + s * vy;                                                   /* :284 */
    obs[1] = -s * vx + c * vy;
    obs[2] = (c * gx + s * gy) + tx;                                            /* :290 */
    obs[3] = (-s * gx + c * gy) + ty;
    for (int b = 0; b < n_beams; ++b) {
        if (!refl_hit[b]) { obs[4 + 2 * b] = 0.0; obs[5 + 2 * b] = 0.0; }       /* :314-315 */
        else {
            obs[4 + 2 * b] = (c * refl_x[b] + s * refl_y[b]) + tx;              /* :317 */
            obs[5 + 2 * b] = (-s * refl_x[b] + c * refl_y[b]) + ty;
        }
    }
}

void orc_get_observation(orc_env* e, double* obs)
{
    orc_sonar_reflection(e);
    observation_from(e->x, e->y, e->theta, e->vx, e->vy, e->goal[0], e->goal[1],
                     e->refl_x, e->refl_y, e->refl_hit, e->num_beams, obs);
}

/* ------------------------------------------------------------------------------------------------
 * termination tests
 * ---------------------------------------------------------------------------------------------- */
/* check_collision :329-336 -- Q4: nearest obstacle CENTRE only */
static int check_collision(const orc_obstacle* obstacles, int n, double x, double y, double robot_r)
{
    if (n == 0) return 0;
    int best = 0; double bd = INFINITY;
    for (int j = 0; j < n; ++j) {
        double dx = obstacles[j].x - x, dy = obstacles[j].y - y;
        double d = sqrt(dx * dx + dy * dy);
        if (d < bd) { bd = d; best = j; }
    }
    return bd <= obstacles[best].r + robot_r;
}
int orc_check_collision(const orc_env* e) { return check_collision(e->obstacles, e->n_obs_placed, e->x, e->y, e->robot_r); }

static double dist2d(double ax, double ay, double bx, double by)
{
    double dx = ax - bx, dy = ay - by;
    return sqrt(dx * dx + dy * dy);
}

int orc_check_reach_goal(const orc_env* e) { return dist2d(e->x, e->y, e->goal[0], e->goal[1]) <= e->goal_dis; }   /* :338-342 */

int orc_out_of_boundary(const orc_env* e)                                                                        /* :264-268 */
{
    return (e->x < 0.0 || e->x > e->width) || (e->y < 0.0 || e->y > e->height);
}

/* ------------------------------------------------------------------------------------------------
 * step marinenav_env.py:199-262 (Q5 priority, Q6 velocity lag)
 * ---------------------------------------------------------------------------------------------- */
int orc_step(orc_env* e, int action, double* obs, double* reward, int* info)
{
    double dis_before = dist2d(e->goal[0], e->goal[1], e->x, e->y);            /* :205 */
    for (int i = 0; i < e->N; ++i) {                                           /* :208-212 */
        double cur[2];
        orc_get_velocity(e, e->x, e->y, cur);
        orc_update_state(e, action, cur);
    }
    double dis_after = dist2d(e->goal[0], e->goal[1], e->x, e->y);             /* :214 */
    orc_get_observation(e, obs);                                               /* :217 */
    double r = e->timestep_penalty;                                            /* :220 */
    r += dis_before - dis_after;                                               /* :229 */
    int done, st;
    if (e->set_boundary && orc_out_of_boundary(e)) { done = 1; st = ORC_OUT_OF_BOUNDARY; }       /* :240-243 */
    else if (e->episode_timesteps >= 1000)          { done = 1; st = ORC_TOO_LONG; }              /* :244-246 */
    else if (orc_check_collision(e))                { r += e->collision_penalty; done = 1; st = ORC_COLLISION; }
    else if (orc_check_reach_goal(e))               { r += e->goal_reward; done = 1; st = ORC_REACH_GOAL; }
    else                                            { done = 0; st = ORC_NORMAL; }
    e->episode_timesteps += 1;                                                 /* :259-260 */
    e->total_timesteps += 1;
    *reward = r; *info = st;
    return done;
}

/* ------------------------------------------------------------------------------------------------
 * reset marinenav_env.py:86-186 ; check_core :344-383 ; check_obstacle :385-420 ; reset_robot :188-197
 * ---------------------------------------------------------------------------------------------- */
static int check_core(const orc_env* e, const orc_core* cj)
{
    if (cj->x - e->r < 0.0 || cj->x + e->r > e->width) return 0;               /* :347-348 */
    if (cj->y - e->r < 0.0 || cj->y + e->r > e->width) return 0;               /* :349-350 (width, sic) */
    if (dist2d(cj->x, cj->y, e->start[0], e->start[1]) < e->r + e->clear_r) return 0;
    if (dist2d(cj->x, cj->y, e->goal[0], e->goal[1]) < e->r + e->clear_r) return 0;
    for (int i = 0; i < e->n_cores_placed; ++i) {
        const orc_core* ci = &e->cores[i];
        double dx = ci->x - cj->x, dy = ci->y - cj->y;
        double dis = sqrt(dx * dx + dy * dy);
        if (ci->clockwise == cj->clockwise) {
            double bi = ci->Gamma / (2 * M_PI * e->v_rel_max);                 /* :369-372 */
            double bj = cj->Gamma / (2 * M_PI * e->v_rel_max);
            if (dis < bi + bj) return 0;
        } else {
            double Gl = ci->Gamma > cj->Gamma ? ci->Gamma : cj->Gamma;         /* :376-381 (Q7) */
            double Gs = ci->Gamma < cj->Gamma ? ci->Gamma : cj->Gamma;
            double v1 = Gl / (2 * M_PI * (dis - 2 * e->r));
            double v2 = Gs / (2 * M_PI * e->r);
            if (v1 > e->p * v2) return 0;
        }
    }
    return 1;
}

static int check_obstacle(const orc_env* e, const orc_obstacle* ob)
{
    if (ob->x - ob->r < 0.0 || ob->x + ob->r > e->width) return 0;             /* :388-391 */
    if (ob->y - ob->r < 0.0 || ob->y + ob->r > e->height) return 0;
    if (dist2d(ob->x, ob->y, e->start[0], e->start[1]) < ob->r + e->clear_r) return 0;
    if (dist2d(ob->x, ob->y, e->goal[0], e->goal[1]) < ob->r + e->clear_r) return 0;
    for (int i = 0; i < e->n_cores_placed; ++i) {                              /* :402-408 */
        double dx = e->cores[i].x - ob->x, dy = e->cores[i].y - ob->y;
        if (sqrt(dx * dx + dy * dy) <= e->r + ob->r) return 0;
    }
    for (int i = 0; i < e->n_obs_placed; ++i) {                                /* :411-418 */
        double dx = e->obstacles[i].x - ob->x, dy = e->obstacles[i].y - ob->y;
        if (sqrt(dx * dx + dy * dy) <= e->obstacles[i].r + ob->r) return 0;
    }
    return 1;
}

/* robot.reset_state(x, y, current) robot.py:79-87 with init_theta/init_speed already chosen */
void orc_restart_episode(orc_env* e, double* obs)
{
    double cur[2];
    orc_get_velocity(e, e->start[0], e->start[1], cur);
    e->x = e->start[0]; e->y = e->start[1];
    e->theta = e->robot_init_theta; e->speed = e->robot_init_speed;
    e->vx = e->speed * cos(e->theta) + cur[0];
    e->vy = e->speed * sin(e->theta) + cur[1];
    e->episode_timesteps = 0;
    if (obs) orc_get_observation(e, obs);
}

void orc_reset_robot(orc_env* e)                                               /* :188-197 */
{
    if (e->random_reset_state) {
        e->robot_init_theta = rd_uniform(&e->rd, 0.0, 2 * M_PI);
        e->robot_init_speed = rd_uniform(&e->rd, 0.0, e->max_speed);
    } else {
        e->robot_init_theta = e->init_theta;
        e->robot_init_speed = e->init_speed;
    }
    int keep = e->episode_timesteps;
    orc_restart_episode(e, 0);
    e->episode_timesteps = keep;
}

void orc_reset(orc_env* e, double* obs)
{
    if (e->n_sched > 0) {                                                      /* :89-98 */
        int cnt = 0;
        for (int i = 0; i < e->n_sched; ++i) if (e->sched_timesteps[i] - e->total_timesteps <= 0) ++cnt;
        int idx = cnt - 1;
        if (idx < 0) idx = e->n_sched - 1;                                      /* python negative index */
        e->num_cores = e->sched_num_cores[idx];
        e->num_obs = e->sched_num_obs[idx];
        e->min_start_goal_dis = e->sched_min_start_goal_dis[idx];
    }
    e->episode_timesteps = 0;                                                  /* :106 */
    e->n_cores_placed = 0; e->n_obs_placed = 0;
    int num_cores = e->num_cores, num_obs = e->num_obs;

    if (e->reset_start_and_goal) {                                             /* :112-127 */
        int iteration = 500; double max_dist = 0.0;
        for (;;) {
            double s0 = rd_uniform(&e->rd, 2.0, e->width - 2.0), s1 = rd_uniform(&e->rd, 2.0, e->height - 2.0);
            double g0 = rd_uniform(&e->rd, 2.0, e->width - 2.0), g1 = rd_uniform(&e->rd, 2.0, e->height - 2.0);
            iteration -= 1;
            double d = dist2d(g0, g1, s0, s1);
            if (d > max_dist) { max_dist = d; e->start[0] = s0; e->start[1] = s1; e->goal[0] = g0; e->goal[1] = g1; }
            if (max_dist > e->min_start_goal_dis || iteration == 0) break;
        }
    }
    if (num_cores > 0) {                                                       /* :130-143 */
        int iteration = 500;
        for (;;) {
            orc_core c;
            c.x = rd_uniform(&e->rd, 0.0, e->width);
            c.y = rd_uniform(&e->rd, 0.0, e->height);
            c.clockwise = rd_binomial_half(&e->rd);
            double v_edge = rd_uniform(&e->rd, e->v_range[0], e->v_range[1]);
            c.Gamma = 2 * M_PI * e->r * v_edge;
            iteration -= 1;
            if (check_core(e, &c)) { e->cores[e->n_cores_placed++] = c; num_cores -= 1; }
            if (iteration == 0 || num_cores == 0) break;
        }
    }
    if (num_obs > 0) {                                                         /* :158-169 */
        int iteration = 500;
        for (;;) {
            orc_obstacle ob;
            ob.x = rd_uniform(&e->rd, 5.0, e->width - 5.0);
            ob.y = rd_uniform(&e->rd, 5.0, e->height - 5.0);
            ob.r = rd_uniform(&e->rd, e->obs_r_range[0], e->obs_r_range[1]);
            iteration -= 1;
            if (check_obstacle(e, &ob)) { e->obstacles[e->n_obs_placed++] = ob; num_obs -= 1; }
            if (iteration == 0 || num_obs == 0) break;
        }
    }
    orc_reset_robot(e);                                                        /* :184 */
    if (obs) orc_get_observation(e, obs);                                      /* :186 */
}

/* ------------------------------------------------------------------------------------------------
 * Batch forms over the SoA buffers of the CUDA C-ABI (layout in marinenav_oracle.h).
 * Signed-circulation convention of the tables: Gs = +Gamma if clockwise else -Gamma; Gs == 0 <=> empty slot.
 * ---------------------------------------------------------------------------------------------- */
void orc_default_params(orc_params* p)
{
    memset(p, 0, sizeof(*p));
    p->dt = 0.1; p->n_substeps = 10;
    p->accel[0] = -0.4; p->accel[1] = 0.0; p->accel[2] = 0.4;
    p->yaw_rate[0] = -M_PI / 6; p->yaw_rate[1] = 0.0; p->yaw_rate[2] = M_PI / 6;
    p->max_speed = 2.0; p->k_drag = p->accel[2] / p->max_speed;
    p->robot_r = 0.8; p->core_r = 0.5; p->goal_dis = 2.0;
    p->timestep_penalty = -1.0; p->collision_penalty = -50.0; p->goal_reward = 100.0;
    p->sonar_range = 10.0; p->sonar_angle = 2 * M_PI / 3; p->n_beams = 11;
    p->max_episode_steps = 1000; p->set_boundary = 0; p->width = 50; p->height = 50;
}

typedef struct {
    int kind;      /* 0 step, 1 observe, 2 reset */
    int64_t lo, hi, E; int max_c, max_o;
    const orc_params* p;
    double *state, *velocity, *obs_d, *reward, *start_pose;
    double *goal_w, *cores_w, *obst_w;
    const double *goal, *cores, *obst;
    const int32_t* action; int32_t* episode_step;
    uint8_t *done, *info, *ncp, *nop;
    int velocity_from_state;
    const uint32_t* seeds; int num_cores, num_obs; double min_start_goal_dis;
} batch_job;

static void load_map(const batch_job* J, int64_t i, orc_core* cores, int* nc, orc_obstacle* obst, int* no)
{
    int64_t E = J->E; *nc = 0; *no = 0;
    for (int c = 0; c < J->max_c; ++c) {
        double gs = J->cores[(2 * J->max_c + c) * E + i];
        if (gs == 0.0) continue;
        cores[*nc].x = J->cores[(c)*E + i];
        cores[*nc].y = J->cores[(J->max_c + c) * E + i];
        cores[*nc].clockwise = gs > 0.0; cores[*nc].Gamma = fabs(gs); ++*nc;
    }
    for (int o = 0; o < J->max_o; ++o) {
        double r = J->obst[(2 * J->max_o + o) * E + i];
        if (!(r > 0.0)) continue;
        obst[*no].x = J->obst[(o)*E + i];
        obst[*no].y = J->obst[(J->max_o + o) * E + i];
        obst[*no].r = r; ++*no;
    }
}

static void beam_table(const orc_params* p, double* beams)
{
    double phi = p->sonar_angle / (p->n_beams - 1), a0 = -p->sonar_angle / 2;
    for (int i = 0; i < p->n_beams; ++i) beams[i] = a0 + i * phi;
}

static void* batch_worker(void* arg)
{
    const batch_job* J = (const batch_job*)arg;
    const orc_params* p = J->p; int64_t E = J->E;
    int nobs = 4 + 2 * p->n_beams;
    double beams[ORC_MAX_BEAMS]; beam_table(p, beams);
    orc_core cores[ORC_MAX_CORES]; orc_obstacle obst[ORC_MAX_OBS];
    double rx[ORC_MAX_BEAMS], ry[ORC_MAX_BEAMS]; int rh[ORC_MAX_BEAMS];
    if (J->kind == 2) {
        orc_env* e = (orc_env*)malloc(sizeof(orc_env));
        for (int64_t i = J->lo; i < J->hi; ++i) {
            orc_env_init(e, J->seeds[i]);
            e->num_cores = J->num_cores; e->num_obs = J->num_obs; e->min_start_goal_dis = J->min_start_goal_dis;
            e->dt = p->dt; e->N = p->n_substeps; e->max_speed = p->max_speed; e->k = p->k_drag; e->robot_r = p->robot_r;
            e->r = p->core_r; e->width = p->width; e->height = p->height;
            for (int q = 0; q < 3; ++q) { e->a[q] = p->accel[q]; e->w[q] = p->yaw_rate[q]; }
            e->sonar_range = p->sonar_range; e->sonar_angle = p->sonar_angle; orc_set_num_beams(e, p->n_beams);
            orc_reset(e, J->obs_d + i * nobs);
            J->state[0 * E + i] = e->x; J->state[1 * E + i] = e->y; J->state[2 * E + i] = e->theta; J->state[3 * E + i] = e->speed;
            J->velocity[0 * E + i] = e->vx; J->velocity[1 * E + i] = e->vy;
            J->goal_w[0 * E + i] = e->goal[0]; J->goal_w[1 * E + i] = e->goal[1];
            for (int c = 0; c < J->max_c; ++c) {
                int on = c < e->n_cores_placed;
                J->cores_w[(c)*E + i] = on ? e->cores[c].x : 0.0;
                J->cores_w[(J->max_c + c) * E + i] = on ? e->cores[c].y : 0.0;
                J->cores_w[(2 * J->max_c + c) * E + i] = on ? (e->cores[c].clockwise ? e->cores[c].Gamma : -e->cores[c].Gamma) : 0.0;
            }
            for (int o = 0; o < J->max_o; ++o) {
                int on = o < e->n_obs_placed;
                J->obst_w[(o)*E + i] = on ? e->obstacles[o].x : 0.0;
                J->obst_w[(J->max_o + o) * E + i] = on ? e->obstacles[o].y : 0.0;
                J->obst_w[(2 * J->max_o + o) * E + i] = on ? e->obstacles[o].r : 0.0;
            }
            if (J->start_pose) {
                J->start_pose[0 * E + i] = e->start[0]; J->start_pose[1 * E + i] = e->start[1];
                J->start_pose[2 * E + i] = e->robot_init_theta; J->start_pose[3 * E + i] = e->robot_init_speed;
            }
            if (J->ncp) J->ncp[i] = (uint8_t)e->n_cores_placed;
            if (J->nop) J->nop[i] = (uint8_t)e->n_obs_placed;
        }
        free(e);
        return 0;
    }
    for (int64_t i = J->lo; i < J->hi; ++i) {
        int nc, no; load_map(J, i, cores, &nc, obst, &no);
        double x = J->state[0 * E + i], y = J->state[1 * E + i], th = J->state[2 * E + i], sp = J->state[3 * E + i];
        double gx = J->goal[0 * E + i], gy = J->goal[1 * E + i];
        double vx = J->velocity[0 * E + i], vy = J->velocity[1 * E + i];
        if (J->kind == 1) {
            if (J->velocity_from_state) {
                double cur[2]; velocity_from_cores(cores, nc, p->core_r, x, y, cur);
                vx = sp * cos(th) + cur[0]; vy = sp * sin(th) + cur[1];
                J->velocity[0 * E + i] = vx; J->velocity[1 * E + i] = vy;
            }
            sonar_reflection(x, y, th, beams, p->n_beams, p->sonar_range, obst, no, rx, ry, rh);
            observation_from(x, y, th, vx, vy, gx, gy, rx, ry, rh, p->n_beams, J->obs_d + i * nobs);
            continue;
        }
        int action = J->action[i];
        double dis_before = dist2d(gx, gy, x, y);
        for (int s = 0; s < p->n_substeps; ++s) {
            double cur[2]; velocity_from_cores(cores, nc, p->core_r, x, y, cur);
            vx = sp * cos(th) + cur[0]; vy = sp * sin(th) + cur[1];
            x += vx * p->dt; y += vy * p->dt;
            double a = p->accel[action / 3], w = p->yaw_rate[action % 3];
            sp += (a - p->k_drag * sp) * p->dt;
            if (sp < 0.0) sp = 0.0;
            if (sp > p->max_speed) sp = p->max_speed;
            th += w * p->dt;
            while (th < 0.0) th += 2 * M_PI;
            while (th >= 2 * M_PI) th -= 2 * M_PI;
        }
        double dis_after = dist2d(gx, gy, x, y);
        sonar_reflection(x, y, th, beams, p->n_beams, p->sonar_range, obst, no, rx, ry, rh);
        observation_from(x, y, th, vx, vy, gx, gy, rx, ry, rh, p->n_beams, J->obs_d + i * nobs);
        double r = p->timestep_penalty; r += dis_before - dis_after;
        int done, st;
        int oob = (x < 0.0 || x > p->width) || (y < 0.0 || y > p->height);
        if (p->set_boundary && oob) { done = 1; st = ORC_OUT_OF_BOUNDARY; }
        else if (J->episode_step[i] >= p->max_episode_steps) { done = 1; st = ORC_TOO_LONG; }
        else if (check_collision(obst, no, x, y, p->robot_r)) { r += p->collision_penalty; done = 1; st = ORC_COLLISION; }
        else if (dist2d(x, y, gx, gy) <= p->goal_dis) { r += p->goal_reward; done = 1; st = ORC_REACH_GOAL; }
        else { done = 0; st = ORC_NORMAL; }
        J->episode_step[i] += 1;
        J->state[0 * E + i] = x; J->state[1 * E + i] = y; J->state[2 * E + i] = th; J->state[3 * E + i] = sp;
        J->velocity[0 * E + i] = vx; J->velocity[1 * E + i] = vy;
        J->reward[i] = r; J->done[i] = (uint8_t)done; J->info[i] = (uint8_t)st;
    }
    return 0;
}

static void run_batch(batch_job* proto, int n_threads)
{
    int64_t E = proto->E;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > E) n_threads = (int)(E > 0 ? E : 1);
    if (n_threads == 1) { proto->lo = 0; proto->hi = E; batch_worker(proto); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    batch_job* jobs = (batch_job*)malloc(sizeof(batch_job) * n_threads);
    for (int t = 0; t < n_threads; ++t) {
        jobs[t] = *proto;
        jobs[t].lo = E * t / n_threads; jobs[t].hi = E * (t + 1) / n_threads;
        pthread_create(&th[t], 0, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
    free(th); free(jobs);
}

void orc_step_batch(double* state, double* velocity, const double* goal, const double* cores, const double* obstacles,
                    const int32_t* action, int32_t* episode_step,
                    double* obs, double* reward, uint8_t* done, uint8_t* info,
                    int64_t E, int max_c, int max_o, const orc_params* p, int n_threads)
{
    batch_job J; memset(&J, 0, sizeof(J));
    J.kind = 0; J.E = E; J.max_c = max_c; J.max_o = max_o; J.p = p;
    J.state = state; J.velocity = velocity; J.goal = goal; J.cores = cores; J.obst = obstacles;
    J.action = action; J.episode_step = episode_step; J.obs_d = obs; J.reward = reward; J.done = done; J.info = info;
    run_batch(&J, n_threads);
}

void orc_observe_batch(const double* state, double* velocity, const double* goal, const double* cores,
                       const double* obstacles, double* obs, int64_t E, int max_c, int max_o, const orc_params* p,
                       int velocity_from_state, int n_threads)
{
    batch_job J; memset(&J, 0, sizeof(J));
    J.kind = 1; J.E = E; J.max_c = max_c; J.max_o = max_o; J.p = p;
    J.state = (double*)state; J.velocity = velocity; J.goal = goal; J.cores = cores; J.obst = obstacles;
    J.obs_d = obs; J.velocity_from_state = velocity_from_state;
    run_batch(&J, n_threads);
}

void orc_reset_batch(const uint32_t* seeds, int num_cores, int num_obs, double min_start_goal_dis,
                     double* state, double* velocity, double* goal, double* cores, double* obstacles,
                     double* start_pose, uint8_t* n_cores_placed, uint8_t* n_obs_placed,
                     double* obs, int64_t E, int max_c, int max_o, const orc_params* p, int n_threads)
{
    batch_job J; memset(&J, 0, sizeof(J));
    J.kind = 2; J.E = E; J.max_c = max_c; J.max_o = max_o; J.p = p;
    J.state = state; J.velocity = velocity; J.goal_w = goal; J.cores_w = cores; J.obst_w = obstacles;
    J.start_pose = start_pose; J.ncp = n_cores_placed; J.nop = n_obs_placed; J.obs_d = obs;
    J.seeds = seeds; J.num_cores = num_cores; J.num_obs = num_obs; J.min_start_goal_dis = min_start_goal_dis;
    run_batch(&J, n_threads);
}

/*
 * marinenav_oracle.h -- TEST INFRASTRUCTURE ONLY (parity oracle; never shipped, never on the product path).
 *
 * Scalar fp64 CPU restatement of the reference environment
 *   RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf
 *     marinenav_env/envs/marinenav_env.py   (MarineNavEnv)
 *     marinenav_env/envs/utils/robot.py     (Robot, Sonar)
 * written from the algorithm description in SURVEY.md section 8(a) (quirks Q1..Q10), with every
 * function citing the reference file:line it follows.
 *
 * PARITY PINNED: tests/test_oracle_golden.py checks this restatement against
 *   (1) the reference itself imported in-process (teacher-forced step/observation/reset parity),
 *   (2) the reference's own golden vectors: pretrained_models/IQN/seed_3/eval_config.json (reset KAT,
 *       bit-for-bit) and the 27 000 recorded evaluation episodes (tests/golden/episodes_*.npz),
 *   (3) committed fixtures generated from the reference by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#ifndef MARINENAV_ORACLE_H
#define MARINENAV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_CORES 32
#define ORC_MAX_OBS   64
#define ORC_MAX_BEAMS 128
#define ORC_MAX_SCHED 8

/* info codes; strings at marinenav_env.py:243,246,250,254,257 */
enum { ORC_NORMAL = 0, ORC_TOO_LONG = 1, ORC_COLLISION = 2, ORC_REACH_GOAL = 3, ORC_OUT_OF_BOUNDARY = 4 };

typedef struct { double x, y; int clockwise; double Gamma; } orc_core;      /* marinenav_env.py:8-15  */
typedef struct { double x, y, r; } orc_obstacle;                            /* marinenav_env.py:17-23 */

typedef struct { uint32_t key[624]; int pos; } orc_mt19937;                 /* numpy RandomState (legacy MT19937) */

typedef struct {
    /* --- env parameters, defaults of marinenav_env.py:40-64,73 --- */
    double width, height, r, v_rel_max, p;
    double v_range[2], obs_r_range[2];
    double clear_r;
    int    reset_start_and_goal;
    double start[2], goal[2];
    int    random_reset_state;
    double init_speed, init_theta;
    double goal_dis, timestep_penalty, collision_penalty, goal_reward, discount;
    int    num_cores, num_obs;
    double min_start_goal_dis;
    int    set_boundary;
    /* --- curriculum schedule (marinenav_env.py:89-98); n_sched == 0 <=> schedule is None --- */
    int    n_sched;
    int64_t sched_timesteps[ORC_MAX_SCHED];
    int    sched_num_cores[ORC_MAX_SCHED], sched_num_obs[ORC_MAX_SCHED];
    double sched_min_start_goal_dis[ORC_MAX_SCHED];
    /* --- robot parameters, robot.py:28-37 ; sonar robot.py:7-9 --- */
    double dt; int N;
    double robot_r, max_speed, a[3], w[3], k;
    double sonar_range, sonar_angle; int num_beams;
    double beam_angles[ORC_MAX_BEAMS];
    /* --- dynamic state --- */
    double x, y, theta, speed, vx, vy;
    double robot_init_theta, robot_init_speed;     /* robot.init_theta / robot.init_speed */
    int     episode_timesteps;
    int64_t total_timesteps;
    int n_cores_placed, n_obs_placed;
    orc_core     cores[ORC_MAX_CORES];
    orc_obstacle obstacles[ORC_MAX_OBS];
    /* sonar.reflections: world-frame point + indicator, robot.py:146,198 */
    double refl_x[ORC_MAX_BEAMS], refl_y[ORC_MAX_BEAMS]; int refl_hit[ORC_MAX_BEAMS];
    orc_mt19937 rd;
} orc_env;

/* construction / RNG */
unsigned long orc_sizeof_env(void);
void   orc_env_init(orc_env* e, uint32_t seed);                 /* MarineNavEnv.__init__ :27-73 + Robot/Sonar ctor */
void   orc_seed(orc_env* e, uint32_t seed);                     /* :75-78 */
void   orc_set_num_beams(orc_env* e, int n);                    /* Sonar.compute_phi/compute_beam_angles robot.py:14-21 */
void   orc_mt_seed(orc_mt19937* s, uint32_t seed);
double orc_mt_random_sample(orc_mt19937* s);

/* the hot path */
void   orc_get_velocity(const orc_env* e, double x, double y, double out[2]);   /* :422-455 */
void   orc_update_state(orc_env* e, int action, const double current[2]);       /* robot.py:102-123 */
void   orc_sonar_reflection(orc_env* e);                                        /* robot.py:125-198 */
void   orc_get_observation(orc_env* e, double* obs);                            /* :273-326 */
int    orc_check_collision(const orc_env* e);                                   /* :329-336 */
int    orc_check_reach_goal(const orc_env* e);                                  /* :338-342 */
int    orc_out_of_boundary(const orc_env* e);                                   /* :264-268 */
/* returns done; writes obs[4+2*num_beams], *reward, *info (ORC_* code) */
int    orc_step(orc_env* e, int action, double* obs, double* reward, int* info);/* :199-262 */
void   orc_reset(orc_env* e, double* obs);                                      /* :86-197 */
void   orc_reset_robot(orc_env* e);                                             /* :188-197 */
/* reset_with_eval_config tail (:551-555): robot.reset_state(start, current(start)) using robot_init_* */
void   orc_restart_episode(orc_env* e, double* obs);

/*
 * Batch forms operating on the SAME structure-of-arrays buffers as the CUDA C-ABI (include/marinenav_b200.h),
 * so parity tests hand identical arrays to both sides.  All arrays are host memory.
 *   state      f64 [4][E]      x, y, theta, speed
 *   goal       f64 [2][E]
 *   cores      f64 [3*max_c][E]  rows x_0..x_{max_c-1}, y_0.., Gs_0..  with Gs = +Gamma if clockwise else -Gamma;
 *                               unused slots: Gs = 0 (skipped)
 *   obstacles  f64 [3*max_o][E]  rows x_j, y_j, r_j; unused slots: r <= 0 (skipped)
 *   velocity   f64 [2][E]      robot.velocity (world frame), in/out for observe, out for step
 */
typedef struct {
    double dt; int n_substeps;
    double accel[3], yaw_rate[3], k_drag, max_speed, robot_r, core_r, goal_dis;
    double timestep_penalty, collision_penalty, goal_reward;
    double sonar_range, sonar_angle; int n_beams;
    int max_episode_steps; int set_boundary; double width, height;
} orc_params;

void orc_default_params(orc_params* p);
void orc_step_batch(double* state, double* velocity, const double* goal, const double* cores, const double* obstacles,
                    const int32_t* action, int32_t* episode_step,
                    double* obs, double* reward, uint8_t* done, uint8_t* info,
                    int64_t E, int max_c, int max_o, const orc_params* p, int n_threads);
/* observation of the current state; if velocity_from_state != 0, velocity is first set to steer + current(pos)
 * (what reset_robot / reset_state leave behind, marinenav_env.py:196-197, robot.py:79-87) */
void orc_observe_batch(const double* state, double* velocity, const double* goal, const double* cores,
                       const double* obstacles, double* obs, int64_t E, int max_c, int max_o, const orc_params* p,
                       int velocity_from_state, int n_threads);
/* seeded reset of E independent envs: env i behaves like MarineNavEnv(seed=seeds[i]) with the given counts, first reset() */
void orc_reset_batch(const uint32_t* seeds, int num_cores, int num_obs, double min_start_goal_dis,
                     double* state, double* velocity, double* goal, double* cores, double* obstacles,
                     double* start_pose, uint8_t* n_cores_placed, uint8_t* n_obs_placed,
                     double* obs, int64_t E, int max_c, int max_o, const orc_params* p, int n_threads);

#ifdef __cplusplus
}
#endif
#endif

// Version, error string and default parameter blocks of libmarinenav_b200.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "mnv_common.cuh"

static thread_local char g_err[512] = "";

void mnv_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// tuning switches (mnv_set_option); MNV_TMA / MNV_PDL in the environment give the initial values
static int g_opt[MNV_OPT_COUNT] = {-1, -1, -1};
static const char* const g_opt_name[MNV_OPT_COUNT] = {"tma", "pdl", "act_timing"};
static const char* const g_opt_env[MNV_OPT_COUNT] = {"MNV_TMA", "MNV_PDL", "MNV_ACT_TIMING"};
static const int g_opt_default[MNV_OPT_COUNT] = {0, 0, 0};

int mnv_option(int which)
{
    if (g_opt[which] < 0) {
        const char* e = getenv(g_opt_env[which]);
        g_opt[which] = (e != nullptr && e[0] != 0) ? atoi(e) : g_opt_default[which];
    }
    return g_opt[which];
}

extern "C" int mnv_set_option(const char* key, int32_t value)
{
    for (int i = 0; key != nullptr && i < MNV_OPT_COUNT; ++i)
        if (strcmp(key, g_opt_name[i]) == 0) { g_opt[i] = value < 0 ? 0 : value; return 0; }
    mnv_set_error("mnv_set_option: unknown key '%s'", key ? key : "(null)");
    return MNV_E_PARAM;
}

extern "C" int mnv_get_option(const char* key)
{
    for (int i = 0; key != nullptr && i < MNV_OPT_COUNT; ++i)
        if (strcmp(key, g_opt_name[i]) == 0) return mnv_option(i);
    mnv_set_error("mnv_get_option: unknown key '%s'", key ? key : "(null)");
    return MNV_E_PARAM;
}

extern "C" int mnv_version(void) { return MNV_VERSION; }

extern "C" const char* mnv_last_error_string(void) { return g_err; }

// robot.py:7-9,28-37 ; marinenav_env.py:40-64,73,244
extern "C" void mnv_default_params(mnv_params* p)
{
    memset(p, 0, sizeof(*p));
    p->dt = 0.1; p->n_substeps = 10;
    p->accel[0] = -0.4; p->accel[1] = 0.0; p->accel[2] = 0.4;
    p->yaw_rate[0] = -MNV_PI / 6; p->yaw_rate[1] = 0.0; p->yaw_rate[2] = MNV_PI / 6;
    p->max_speed = 2.0; p->k_drag = p->accel[2] / p->max_speed;
    p->robot_r = 0.8; p->core_r = 0.5; p->goal_dis = 2.0;
    p->timestep_penalty = -1.0; p->collision_penalty = -50.0; p->goal_reward = 100.0;
    p->sonar_range = 10.0; p->sonar_angle = 2 * MNV_PI / 3; p->n_beams = 11;
    p->max_episode_steps = 1000; p->set_boundary = 0; p->width = 50.0; p->height = 50.0;
}

// marinenav_env.py:40-64
extern "C" void mnv_default_reset_params(mnv_reset_params* p)
{
    memset(p, 0, sizeof(*p));
    p->width = 50.0; p->height = 50.0; p->core_r = 0.5; p->v_rel_max = 1.0; p->p = 0.8;
    p->v_range[0] = 5.0; p->v_range[1] = 10.0; p->obs_r_range[0] = 1.0; p->obs_r_range[1] = 3.0; p->clear_r = 10.0;
    p->reset_start_and_goal = 1; p->start[0] = 5.0; p->start[1] = 5.0; p->goal[0] = 45.0; p->goal[1] = 45.0;
    p->random_reset_state = 1; p->init_theta = MNV_PI / 4; p->init_speed = 0.0; p->max_speed = 2.0;
    p->num_cores = 8; p->num_obs = 5; p->min_start_goal_dis = 25.0;
}

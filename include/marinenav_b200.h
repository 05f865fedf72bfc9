/*
 * marinenav_b200.h -- C-ABI of libmarinenav_b200.so: the B200 (sm_100a) implementation of the hot path of
 * RobustFieldAutonomyLab/Distributional_RL_Navigation (reference @ e77bbbf):
 *     marinenav_env/envs/marinenav_env.py  MarineNavEnv.step / reset / get_observation
 *     marinenav_env/envs/utils/robot.py    Robot.update_state / sonar_reflection
 *     thirdparty/IQN/model.py, agent.py    ObsEncoder.forward / get_qvals, IQNAgent.act / train
 *
 * The reference has no FFI of its own (pure Python); these entry points are what a ctypes binding inside the
 * reference's MarineNavEnv / IQNAgent would call (INTEGRATION.md shows that binding).  Conventions:
 *   - extern "C", plain pointers + sizes, no C++/torch types.  Pointers named d_* are DEVICE pointers
 *     (e.g. torch.Tensor.data_ptr()); the caller owns all memory, the library allocates nothing persistent.
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *     no host synchronisation inside.
 *   - return 0 = ok; > 0 = cudaError_t of the launch; < 0 = argument error (MNV_E_*).  mnv_last_error_string()
 *     describes the last failure of the calling thread.
 *
 * Device data layout ("SoA by field", E = number of environments, every array 16-byte aligned):
 *   d_state      f64 [4][E]        x, y, theta, speed                        (robot.py:41-45)
 *   d_velocity   f64 [2][E]        robot.velocity wrt sea floor, world frame (robot.py:98-100)
 *   d_goal       f64 [2][E]        goal position                             (marinenav_env.py:54)
 *   d_cores      f64 [3*max_c][E]  rows x_0..x_{max_c-1}, y_0.., Gs_0..  Gs = +Gamma if clockwise else -Gamma,
 *                                  Gs == 0 <=> empty slot                    (marinenav_env.py:8-15)
 *   d_obstacles  f64 [3*max_o][E]  rows x_j, y_j, r_j ; r_j <= 0 <=> empty slot; LIST ORDER IS SEMANTIC (robot.py:149,192)
 *   d_action     i32 [E]           0..8 = (a_idx*3 + w_idx)                  (robot.py:54-55)
 *   d_episode_step i32 [E]         MarineNavEnv.episode_timesteps            (marinenav_env.py:69,259)
 *   d_obs        f32 [E][4+2*n_beams] row-major: velocity_r(2), goal_r(2), per beam hit_r(2) or (0,0) (marinenav_env.py:273-326)
 *   d_reward     f32 [E] ; d_done u8 [E] ; d_info u8 [E] (MNV_INFO_*)        (marinenav_env.py:240-257)
 */
#ifndef MARINENAV_B200_H
#define MARINENAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNV_VERSION        100
#define MNV_MAX_CORES      8      /* compiled capacity of the fused step kernel */
#define MNV_MAX_OBSTACLES  32
#define MNV_MAX_BEAMS      128

/* info codes <-> info["state"] strings of marinenav_env.py:243,246,250,254,257 */
enum { MNV_INFO_NORMAL = 0, MNV_INFO_TOO_LONG = 1, MNV_INFO_COLLISION = 2, MNV_INFO_REACH_GOAL = 3, MNV_INFO_OUT_OF_BOUNDARY = 4 };

enum { MNV_E_NULL = -1, MNV_E_ALIGN = -2, MNV_E_SIZE = -3, MNV_E_CAPACITY = -4, MNV_E_PARAM = -5 };

/* Robot / sonar / reward constants.  Defaults = robot.py:7-9,28-37 and marinenav_env.py:40-64,244. */
typedef struct mnv_params {
    double dt; int32_t n_substeps;                    /* Robot.dt, Robot.N */
    double accel[3], yaw_rate[3];                     /* Robot.a, Robot.w ; action = 3*a_idx + w_idx */
    double k_drag, max_speed;                         /* Robot.k = max(a)/max_speed, Robot.max_speed */
    double robot_r, core_r, goal_dis;                 /* Robot.r, MarineNavEnv.r, MarineNavEnv.goal_dis */
    double timestep_penalty, collision_penalty, goal_reward;
    double sonar_range, sonar_angle; int32_t n_beams; /* Sonar.range, .angle, .num_beams */
    int32_t max_episode_steps;                        /* 1000, marinenav_env.py:244 */
    int32_t set_boundary; double width, height;       /* marinenav_env.py:73,40-41 */
    int32_t pdl_prefetch;                             /* launch mode of THIS call, see below (0 = plain launch; default) */
} mnv_params;

/* Map-generation constants of MarineNavEnv.reset (marinenav_env.py:40-64,86-197). */
typedef struct mnv_reset_params {
    double width, height, core_r, v_rel_max, p;
    double v_range[2], obs_r_range[2], clear_r;
    int32_t reset_start_and_goal; double start[2], goal[2];
    int32_t random_reset_state; double init_theta, init_speed, max_speed;
    int32_t num_cores, num_obs; double min_start_goal_dis;
} mnv_reset_params;

int         mnv_version(void);
const char* mnv_last_error_string(void);
void        mnv_default_params(mnv_params* p);
void        mnv_default_reset_params(mnv_reset_params* p);

/* Launch mode of mnv_step, PER CALL: mnv_params.pdl_prefetch != 0 launches the step with programmatic stream
 * serialization and lets it fetch the map tables (d_goal, d_cores, d_obstacles) BEFORE griddepcontrol.wait, i.e. while the
 * previous launch on the stream still runs (15.1 -> 12.9 us per 65 536-env step).  CONTRACT: the launch immediately before
 * that mnv_step on the stream does not write those three tables (mnv_reset must be followed by mnv_observe or any other
 * launch first; host -> device copies are fine).  Only a caller whose own launch sequences satisfy the contract sets the
 * flag (VecMarineNavEnv does, for its own launches); the default 0 has no contract.
 *
 * Process-wide tuning switches of the kernels (lab use; results are identical for every setting; parity tests run all of them).
 *   "pdl" 0|1|2  1: launch mnv_step / mnv_observe with programmatic stream serialization: the next launch on the stream
 *              is scheduled while this one drains and blocks in griddepcontrol.wait before its first global access
 *              (measured 1 % slower than 0 on back-to-back steps, profiles/README.md).  2: every mnv_step of the process
 *              behaves as if pdl_prefetch were set.  Default 0 (MNV_PDL in the environment gives the initial value).
 *   "tma" 0|1  stage the obstacle rows with the TMA bulk-copy engine (cp.async.bulk) instead of per-thread cp.async
 *              (default 0: measured slower, profiles/README.md)
 * Returns 0, or MNV_E_PARAM for an unknown key.  mnv_get_option returns the value or MNV_E_PARAM. */
int         mnv_set_option(const char* key, int32_t value);
int         mnv_get_option(const char* key);

/* MarineNavEnv.step (marinenav_env.py:199-262) for E environments: N sub-steps of get_velocity (:422-455) +
 * Robot.update_state (robot.py:102-123), then get_observation (:273-326, sonar robot.py:125-198), reward and the
 * termination priority of :240-257.  In/out: d_state, d_episode_step (+1).  Out: d_velocity (last sub-step's, Q6),
 * d_obs, d_reward, d_done, d_info; d_trajectory (optional, may be NULL) f64 [n_substeps][2][E] = the position after every
 * sub-step (Robot.trajectory, marinenav_env.py:212).  ONE kernel launch. */
int mnv_step(double* d_state, double* d_velocity, const double* d_goal, const double* d_cores, const double* d_obstacles,
             const int32_t* d_action, int32_t* d_episode_step,
             float* d_obs, float* d_reward, uint8_t* d_done, uint8_t* d_info, double* d_trajectory,
             int64_t E, int32_t max_c, int32_t max_o, const mnv_params* p, void* stream);

/* MarineNavEnv.get_observation (marinenav_env.py:273-326) of the current state, for every environment with
 * d_mask[e] != 0 (d_mask == NULL: all); rows of masked-out environments are left untouched.  If velocity_from_state != 0
 * the velocity is first set to steer + current(position) and written to d_velocity -- what reset_robot /
 * Robot.reset_state leave behind (marinenav_env.py:196-197, robot.py:79-87). */
int mnv_observe(const double* d_state, double* d_velocity, const double* d_goal, const double* d_cores,
                const double* d_obstacles, const uint8_t* d_mask, float* d_obs, int64_t E, int32_t max_c, int32_t max_o,
                const mnv_params* p, int32_t velocity_from_state, void* stream);

/* numpy.random.RandomState(seed) per environment (MarineNavEnv.seed, marinenav_env.py:75-78): legacy MT19937.
 *   d_rng_key u32 [E][624] (one contiguous MT19937 state per environment), d_rng_pos i32 [E].  d_seeds u32 [E]. */
int mnv_seed(uint32_t* d_rng_key, int32_t* d_rng_pos, const uint32_t* d_seeds, int64_t E, void* stream);

/* MarineNavEnv.reset (marinenav_env.py:86-186) for every environment with d_mask[e] != 0 (d_mask == NULL: all):
 * start/goal sampling, vortex-core and obstacle rejection sampling (check_core :344-383, check_obstacle :385-420),
 * reset_robot's draws (:188-192), taken from each environment's own MT19937 stream in the reference's draw order, so
 * that environment e reproduces MarineNavEnv(seed=seeds[e]).reset() bit-for-bit (tables, start, goal, theta0, speed0).
 * Writes d_state (start x, y, theta0, speed0), d_goal, d_cores, d_obstacles, d_start_pose f64 [4][E] (optional: start x,
 * y, init theta, init speed -- Robot.init_theta/init_speed, needed by reset_with_eval_config-style restarts),
 * d_episode_step = 0 (optional), d_n_placed u8 [2][E] (optional: cores, obstacles actually placed, Q8).
 * Does NOT write velocity / observation: follow with mnv_observe(mask, velocity_from_state = 1) -- reset_robot's
 * robot.reset_state (marinenav_env.py:196-197) + get_observation (:186).
 * rp->reset_start_and_goal == 0: start/goal = rp->start / rp->goal for all environments, unless rp->start[0] is NaN,
 * which keeps each environment's own d_start_pose[0:2] / d_goal. */
int mnv_reset(uint32_t* d_rng_key, int32_t* d_rng_pos, const uint8_t* d_mask,
              double* d_state, double* d_goal, double* d_cores, double* d_obstacles,
              double* d_start_pose, int32_t* d_episode_step, uint8_t* d_n_placed,
              int64_t E, int32_t max_c, int32_t max_o, const mnv_reset_params* rp, void* stream);

/* Host-boundary helper of the vectorised env (no counterpart in the reference, which steps ONE env and returns numpy):
 * copies the rows d_rows[e][0:row_len] (f32, row-major, e.g. the observations) of every environment with d_mask[e] != 0
 * to the same rows of h_rows_mapped, a pinned host array that is mapped into the device address space (cudaHostAlloc /
 * torch pin_memory(); under UVA its host address is its device address).  Zero-copy stores: after the launch has completed
 * (stream synchronisation) the host sees the rows.  One kernel on `stream`. */
int mnv_scatter_rows_host(const uint8_t* d_mask, const float* d_rows, float* h_rows_mapped, int64_t E, int32_t row_len,
                          void* stream);

/* Host-boundary helper: compact transport of the observation block d_obs f32 [E][obs_dim] (obs_dim = 4 + 2 * n_beams).  A beam
 * without a return is exactly (0, 0) (marinenav_env.py:318-320) and few beams carry one, so the kernel writes
 *   d_head  f32 [E][4]              the 4 head values of every row
 *   d_mask  u32 [E][W]              W = ceil(n_beams / 32); bit b of environment e: beam b has a return
 *   d_dir   u32 [ceil(E / 32)]      where the returns of environments 32 g .. 32 g + 31 start in d_vals
 *   d_count u32 [4]                 d_count[0] = number of returns (may exceed `capacity`: the tail is then not written);
 *                                   d_count[3] = launch sequence number, incremented by every call (zero it once)
 *   d_vals  f32 [capacity][2]       (x, y) of a group's returns, environment by environment, beam by beam
 * for ONE device -> host copy; libmnv_host.so (include/mnv_host.h) expands it into the dense block on the host.
 * One memset node + one kernel on `stream`. */
int mnv_pack_obs(const float* d_obs, int64_t E, int32_t obs_dim, float* d_head, uint32_t* d_mask, uint32_t* d_dir,
                 uint32_t* d_count, float* d_vals, int64_t capacity, void* stream);

/* ================================ replay buffer (thirdparty/IQN/replay_buffer.py) ===========================
 * Device-resident ReplayBuffer of the vectorised trainer.  Ring arrays (caller-owned, `capacity` transitions):
 *   d_states f32 [capacity][row_len], d_actions i64 [capacity], d_rewards f32 [capacity], d_next_states f32
 *   [capacity][row_len], d_dones f32 [capacity]  -- the tuple layout of replay_buffer.py:20,49-53.
 * A LOGICAL index i counts from the oldest stored transition (memory[i] of the reference's deque): ring slot =
 * (head + i) mod capacity.  The caller keeps head / size / pos (they advance deterministically: pos += E per append). */

/* ReplayBuffer.add (replay_buffer.py:26-41) for one vector step of E environments, ONE launch: transition e =
 * (d_obs[e], d_action[e], d_reward[e], d_next_obs[e], d_done[e]) goes through environment e's n-step window
 * (deque(maxlen=n_step), :24,29; d_win_* = [n_step][E] rows, only read / written when n_step > 1) and, once the window
 * holds n_step entries (t >= n_step - 1, t = number of vector steps appended before this one), the folded transition
 * (state_0, action_0, sum_i gamma^i reward_i in double, next_state_{n-1}, done_{n-1}) (:36-41) is stored in ring slot
 * (pos + e) mod capacity.  Like the reference, the window is NOT cleared at episode ends. */
int rpl_append(float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states, float* d_dones,
               int64_t capacity, int64_t pos, const float* d_obs, const int32_t* d_action, const float* d_reward,
               const float* d_next_obs, const uint8_t* d_done, int64_t E, int32_t row_len, int32_t n_step, double gamma,
               int64_t t, float* d_win_obs, int32_t* d_win_action, float* d_win_reward, void* stream);

/* ReplayBuffer.sample (replay_buffer.py:45-55): B picks among the `size` stored transitions, gathered into
 * d_out_states f32 [B][row_len], d_out_actions i64 [B], d_out_rewards f32 [B], d_out_next_states, d_out_dones f32 [B]
 * (what iqn_loss_grad consumes).  The picks (logical indices, also written to d_indices i64 [B]) come from the
 * counter-based Philox4x32-10 stream (seed, call): without_replacement != 0 yields an ordered tuple of B distinct
 * indices, every such tuple equally likely -- the distribution of random.sample (:47); 0 yields independent uniform picks.
 * Deterministic for given (seed, call, size, B).  B <= 8192.  Two launches (draw, gather). */
int rpl_sample(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
               const float* d_dones, int64_t capacity, int64_t head, int64_t size, uint64_t seed, uint64_t call,
               int32_t without_replacement, int64_t* d_indices, float* d_out_states, int64_t* d_out_actions,
               float* d_out_rewards, float* d_out_next_states, float* d_out_dones, int64_t B, int32_t row_len, void* stream);

/* The gather of rpl_sample for caller-provided logical indices d_indices i64 [B] (e.g. the reference's random.sample picks). */
int rpl_gather(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
               const float* d_dones, int64_t capacity, int64_t head, int64_t size, const int64_t* d_indices,
               float* d_out_states, int64_t* d_out_actions, float* d_out_rewards, float* d_out_next_states,
               float* d_out_dones, int64_t B, int32_t row_len, void* stream);

/* ======================================= IQN (thirdparty/IQN) ============================================
 * Parameters: ONE flat fp32 vector of iqn_param_count() = 35 785 floats = the 14 tensors of ObsEncoder.state_dict() in
 * order (velocity_encoder.weight [16,2], .bias, goal_encoder.*, sensor_encoder.* [176,22], cos_embedding.* [208,64],
 * hidden_layer.* [64,208], hidden_layer_2.* [64,64], output_layer.* [9,64]; model.py:125-136), torch [out][in] layout.
 * d_packed: iqn_packed_count() floats of kernel-side transposes; refresh with iqn_pack after ANY parameter change. */
int     iqn_param_count(void);
int     iqn_packed_count(void);
int     iqn_pack(const float* d_params, float* d_packed, void* stream);

/* ObsEncoder.forward / get_qvals (model.py:160-191) with caller-provided uniform samples:
 *   d_obs f32 [B][26]; d_taus f32 [B][n_tau] in [0,1) (torch.rand of model.py:149); tau <- tau * cvar (model.py:153) with
 *   cvar = d_cvar[b] if d_cvar != NULL (adaptive CVaR, agent.py:249-267) else cvar_scalar.  n_tau in {8,16,32,64}.
 * Outputs (each optional): d_quantiles f32 [B][n_tau][9]; d_qmean f32 [B][9] = mean over taus; d_greedy i32 [B] =
 * argmax_a qmean (first maximum, like np.argmax at agent.py:201). */
int     iqn_forward(const float* d_params, const float* d_packed, const float* d_obs, const float* d_taus,
                    const float* d_cvar, float cvar_scalar, float* d_quantiles, float* d_qmean, int32_t* d_greedy,
                    int64_t B, int32_t n_tau, void* stream);

/* IQNAgent.train up to loss.backward() (agent.py:276-298): target forward on next_states with d_taus_target (drawn
 * first, Q9), local forward on states with d_taus_local (both f32 [B][8]), T = r + gamma_n (1-done) max_a Q', pairwise
 * quantile Huber loss, full backward of the local network.  d_actions i64 [B]; rewards/dones f32 [B].
 * d_scratch: iqn_train_scratch_floats(B) floats.  Writes d_loss (1 float) and d_grad (35 785 floats, unclipped). */
int64_t iqn_train_scratch_floats(int64_t B);
int     iqn_loss_grad(const float* d_params_local, const float* d_packed_local, const float* d_params_target,
                      const float* d_packed_target, const float* d_states, const int64_t* d_actions,
                      const float* d_rewards, const float* d_next_states, const float* d_dones,
                      const float* d_taus_target, const float* d_taus_local, float gamma_n,
                      float* d_scratch, float* d_loss, float* d_grad, int64_t B, void* stream);

/* torch.nn.utils.clip_grad_norm_(params, max_norm) + torch.optim.Adam.step (agent.py:66,299-300) on the flat vectors;
 * the gradient is first multiplied by grad_scale (1/world_size after a summing all-reduce).  step = number of optimizer
 * steps including this one.  d_grad_norm (optional) receives the pre-clip total norm.  d_packed (iqn_pack layout) and
 * d_packed_tc (iqn_pack_tc layout, must have been initialised by iqn_pack_tc once) are optional: if given, the same
 * kernel keeps them current, so no repacking launch is needed after the update. */
int     iqn_clip_adam(float* d_params, const float* d_grad, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                      float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, int64_t step,
                      float* d_grad_norm, void* stream);

/* The data-parallel update in two launches.  iqn_loss_partials = iqn_loss_grad without the reduction: the per-tile partial
 * gradients and losses stay in d_scratch.  iqn_update_tail then does, in ONE launch, everything behind loss.backward()
 * (agent.py:298-300): fixed-order sum of the tile partials, (world > 1) a one-shot all-reduce of the gradient through peer
 * memory over NVLink -- every CTA publishes its 256-parameter slice in this rank's exchange buffer, flags the peers with
 * system-scope release stores, and sums the peers' slices in rank order (bit-identical replicas; gradient = mean over ranks)
 * -- then clip_grad_norm_(max_norm) + Adam.step + refresh of d_packed / d_packed_tc like iqn_clip_adam.  Outputs d_loss,
 * d_grad (the averaged unclipped gradient) and d_grad_norm are optional.
 *   d_sync: iqn_tail_sync_bytes() bytes, zeroed ONCE by the caller, then owned by the kernels (u64[0] completed launches,
 *   u64[1] grid-barrier arrivals, u64[2] sticky error word: 1 = a peer's slice never arrived, 2 = grid barrier timed out --
 *   both only after seconds of spinning, so that a dead rank cannot hang the surviving GPUs);
 *   peer_xchg: HOST array of `world` device pointers, entry r = rank r's exchange buffer (iqn_xchg_bytes() bytes each, zeroed
 *   once) as mapped into THIS process -- own entry from iqn_xchg_alloc, the others from iqn_xchg_open on the 64-byte handles
 *   the ranks exchange once at start-up (CUDA IPC; same node).  NULL / world == 1: single GPU.  Every rank must issue the
 *   same sequence of iqn_update_tail calls (the epochs advance in lockstep); world <= 8. */
int     iqn_loss_partials(const float* d_params_local, const float* d_packed_local, const float* d_params_target,
                          const float* d_packed_target, const float* d_states, const int64_t* d_actions,
                          const float* d_rewards, const float* d_next_states, const float* d_dones,
                          const float* d_taus_target, const float* d_taus_local, float gamma_n,
                          float* d_scratch, int64_t B, void* stream);
int64_t iqn_tail_sync_bytes(void);
int64_t iqn_xchg_bytes(void);
int32_t iqn_xchg_handle_bytes(void);
int     iqn_xchg_alloc(void** d_ptr, unsigned char* handle);
int     iqn_xchg_open(const unsigned char* handle, void** d_ptr);
int     iqn_xchg_close(void* d_ptr);
int     iqn_xchg_free(void* d_ptr);
int     iqn_update_tail(float* d_params, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                        const float* d_scratch, int64_t B, float* d_loss, float* d_grad, float* d_grad_norm,
                        void* d_sync, void* const* peer_xchg, int32_t rank, int32_t world,
                        float max_norm, float lr, float beta1, float beta2, float eps, int64_t step, void* stream);

/* Acting forward on the tensor cores (tcgen05.mma, bf16 operands, fp32 accumulation in TMEM): get_qvals + argmax for a
 * whole env batch, K = n_tau = 32 (ObsEncoder.K, model.py:118).  Same inputs as iqn_forward; outputs d_qmean f32 [B][9]
 * and/or d_greedy i32 [B].  d_packed_tc: iqn_packed_tc_bytes() bytes of bf16 weight tiles, refresh with iqn_pack_tc after
 * any parameter change.  d_scratch: iqn_act_scratch_bytes(B) bytes (the bf16 observation-encoder features written by the
 * pre-pass kernel of the same call).  d_debug (optional, NULL in production): 45 056 floats = the four raw accumulators of
 * tile 0 (128x208, 128x64, 128x64, 128x16) for diagnostics.  Q-values differ from the fp32 path by bf16 rounding (~1e-2
 * relative): use where only the argmax is consumed (agent.py:200-203); iqn_forward stays the parity path.  Two launches. */
int     iqn_packed_tc_bytes(void);
int     iqn_pack_tc(const float* d_params, void* d_packed_tc, void* stream);
int64_t iqn_act_scratch_bytes(int64_t B);
int     iqn_act_tc(const float* d_params, const void* d_packed_tc, const float* d_obs, const float* d_taus,
                   const float* d_cvar, float cvar_scalar, float* d_qmean, int32_t* d_greedy, float* d_debug,
                   void* d_scratch, int64_t B, int32_t n_tau, void* stream);

/* IQNAgent.act / act_adaptive for a whole env batch with the randomness drawn ON THE DEVICE (agent.py:186-215): the 32 taus of
 * every environment (model.py:149) and the epsilon-greedy coin + random action (agent.py:200-203: greedy iff u > eps) come
 * from the counter-based Philox4x32-10 stream keyed by `seed`, counter (environment, sub-stream, step) -- reproducible per
 * (seed, step), no tau tensor in memory, no separate random / select launches.  adaptive_cvar != 0: the CVaR level of every
 * environment is adjust_cvar(obs) (agent.py:249-267), computed by the pre-pass and left in d_cvar f32 [B]; otherwise
 * cvar_scalar.  Outputs: d_action i32 [B] (required), d_greedy i32 [B] and d_qmean f32 [B][9] (optional).  Two launches. */
int     iqn_act_tc_sample(const float* d_params, const void* d_packed_tc, const float* d_obs, int32_t adaptive_cvar,
                          float* d_cvar, float cvar_scalar, float eps, uint64_t seed, uint64_t step,
                          int32_t* d_action, int32_t* d_greedy, float* d_qmean, void* d_scratch, int64_t B, void* stream);

/* ============================ vector-step control block (CUDA-graph replays) ================================
 * A rollout + learn vector step (act -> env step -> replay append -> sample -> update, agent.py:94-173 vectorised) is a
 * chain of ~12 launches of 5 - 250 us each: launch-bound from Python.  The trainer therefore captures it ONCE as a CUDA
 * graph and replays it; what changes from step to step -- the exploration rate, the Philox counters, the ring position of
 * the replay buffer, Adam's bias corrections -- cannot be a by-value kernel argument any more (a graph freezes those), so
 * the `_ctl` entry points below read them from this 64-byte block in DEVICE memory.  The host fills a pinned copy before
 * every replay and the graph's first node copies it in.  Each `_ctl` function is its by-value counterpart with the named
 * scalars taken from the block (same kernels, same results for equal values; range checks on them are the caller's). */
typedef struct mnv_vstep_ctl {
    float    act_eps;              /* iqn_act_tc_sample: eps */
    float    adam_step_size;       /* iqn_update_tail: lr / (1 - beta1^step), computed in double like the by-value path */
    float    adam_inv_sqrt_bc2;    /* iqn_update_tail: 1 / sqrt(1 - beta2^step) */
    int32_t  reserved;
    uint64_t act_step;             /* iqn_act_tc_sample: step */
    int64_t  rpl_pos, rpl_t;       /* rpl_append: pos, t */
    int64_t  rpl_head, rpl_size;   /* rpl_sample: head, size */
    uint64_t rpl_call;             /* rpl_sample / iqn_draw_taus: call */
} mnv_vstep_ctl;

int rpl_append_ctl(float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states, float* d_dones,
                   int64_t capacity, const float* d_obs, const int32_t* d_action, const float* d_reward,
                   const float* d_next_obs, const uint8_t* d_done, int64_t E, int32_t row_len, int32_t n_step, double gamma,
                   float* d_win_obs, int32_t* d_win_action, float* d_win_reward, const mnv_vstep_ctl* d_ctl, void* stream);
int rpl_sample_ctl(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
                   const float* d_dones, int64_t capacity, uint64_t seed, int32_t without_replacement, int64_t* d_indices,
                   float* d_out_states, int64_t* d_out_actions, float* d_out_rewards, float* d_out_next_states,
                   float* d_out_dones, int64_t B, int32_t row_len, const mnv_vstep_ctl* d_ctl, void* stream);
int iqn_update_tail_ctl(float* d_params, float* d_m, float* d_v, float* d_packed, void* d_packed_tc,
                        const float* d_scratch, int64_t B, float* d_loss, float* d_grad, float* d_grad_norm,
                        void* d_sync, void* const* peer_xchg, int32_t rank, int32_t world,
                        float max_norm, float beta1, float beta2, float eps, const mnv_vstep_ctl* d_ctl, void* stream);
int iqn_act_tc_sample_ctl(const float* d_params, const void* d_packed_tc, const float* d_obs, int32_t adaptive_cvar,
                          float* d_cvar, float cvar_scalar, uint64_t seed, int32_t* d_action, int32_t* d_greedy,
                          float* d_qmean, void* d_scratch, int64_t B, const mnv_vstep_ctl* d_ctl, void* stream);

/* The quantile samples of one update drawn on the device: d_taus f32 [n] (n = 2 * B * 8: the target network's taus first,
 * then the local network's -- the draw order of agent.py:279-283, Q9) = uniform [0, 1) multiples of 2^-24 (torch.rand's
 * float32 grid) from Philox4x32-10(key = seed ^ 0x7A75, counter = (i / 4, 0x7A, call)), call = d_ctl->rpl_call if d_ctl != NULL.
 * One launch; reproducible per (seed, call). */
int iqn_draw_taus(float* d_taus, int64_t n, uint64_t seed, uint64_t call, const mnv_vstep_ctl* d_ctl, void* stream);

#ifdef __cplusplus
}
#endif
#endif

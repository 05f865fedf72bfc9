// Device-resident replay buffer of the vectorised trainer: thirdparty/IQN/replay_buffer.py:6-59 for E parallel
// environments, everything in HBM.
//
//   rpl_append   ReplayBuffer.add (:26-41) for one vector step: per environment an n-step window (deque(maxlen=n_step),
//                :24,29); once it holds n_step entries the folded transition (state_0, action_0, sum_i gamma^i r_i,
//                next_state_{n-1}, done_{n-1}) (:36-41) is appended to the ring (deque(maxlen=buffer_size), :18).  One warp
//                per environment, ONE launch for all five fields.
//   rpl_sample   ReplayBuffer.sample (:45-55): B uniform picks from the stored transitions.  without_replacement = 1 gives
//                the distribution of random.sample (:47): an ordered tuple of B DISTINCT logical indices, every such tuple
//                equally likely; the picks come from a counter-based Philox stream (reproducible per (seed, counter)).
//                The picked rows are gathered into the batch layout iqn_loss_grad consumes.
//   rpl_gather   the same gather for caller-provided indices (parity harnesses inject the reference's random.sample picks).
// Logical index i counts from the OLDEST stored transition (memory[i] of the reference's deque); the ring slot is
// (head + i) mod capacity.
#include <algorithm>

#include "mnv_common.cuh"
#include "philox.cuh"

namespace {

constexpr int kWarps = 8;

struct Ring { float* states; int64_t* actions; float* rewards; float* next_states; float* dones; long long capacity; };

__device__ __forceinline__ void copy_row(float* __restrict__ dst, const float* __restrict__ src, int row_len, int lane)
{
    if ((row_len & 1) == 0) {                       // rows are 8-byte aligned whenever row_len is even (26 floats = 104 B)
        for (int i = lane; i < (row_len >> 1); i += 32)
            reinterpret_cast<float2*>(dst)[i] = reinterpret_cast<const float2*>(src)[i];
    } else {
        for (int i = lane; i < row_len; i += 32) dst[i] = src[i];
    }
}

__global__ void __launch_bounds__(kWarps * 32)
rpl_append_kernel(Ring R, long long pos, const float* __restrict__ obs, const int32_t* __restrict__ action,
                  const float* __restrict__ reward, const float* __restrict__ next_obs, const uint8_t* __restrict__ done,
                  long long E, int row_len, int n_step, double gamma, long long t,
                  float* __restrict__ win_obs, int32_t* __restrict__ win_action, float* __restrict__ win_reward,
                  const mnv_vstep_ctl* __restrict__ ctl)
{
    if (ctl != nullptr) { pos = ctl->rpl_pos; t = ctl->rpl_t; }       // graph replays: the ring position lives in device memory
    const int lane = threadIdx.x & 31;
    const long long e = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (e >= E) return;
    const float* s0 = obs + e * row_len;
    int a0 = action[e];
    double ret = (double)reward[e];
    if (n_step > 1) {
        // the window slot of vector step t is t mod n_step; the oldest entry of a full window sits in slot (t + 1) mod n_step
        const long long cur = t % n_step;
        copy_row(win_obs + (cur * E + e) * row_len, s0, row_len, lane);
        if (lane == 0) { win_action[cur * E + e] = a0; win_reward[cur * E + e] = reward[e]; }
        if (t < n_step - 1) return;                  // len(n_step_buffer) < n_step: nothing is stored yet (:30)
        __syncwarp();
        const long long old = (t + 1) % n_step;
        s0 = win_obs + (old * E + e) * row_len;
        a0 = win_action[old * E + e];
        ret = 0.0;                                   // calc_multistep_return (:36-41): Return += gamma**idx * reward_idx, in double
        double g = 1.0;
        for (int i = 0; i < n_step; ++i) {
            const long long slot = (old + i) % n_step;
            const float r = (slot == cur) ? reward[e] : win_reward[slot * E + e];
            ret += g * (double)r;
            g *= gamma;
        }
    }
    const long long slot = (pos + e) % R.capacity;
    copy_row(R.states + slot * row_len, s0, row_len, lane);
    copy_row(R.next_states + slot * row_len, next_obs + e * row_len, row_len, lane);
    if (lane == 0) {
        R.actions[slot] = (int64_t)a0;
        R.rewards[slot] = (float)ret;
        R.dones[slot] = done[e] ? 1.0f : 0.0f;
    }
}

// n_step == 1: no window, and the ring slots of one vector step are consecutive -- (pos + e) mod capacity -- so the append is two
// flat block copies (E x row_len floats each, wrapping at the end of the ring) plus E (action, reward, done) triples: 8-byte
// vectors, every lane busy (the warp-per-row form above keeps 13 of 32 lanes busy on a 104-byte row).
__global__ void __launch_bounds__(256)
rpl_append_flat_kernel(Ring R, long long pos, const float* __restrict__ obs, const int32_t* __restrict__ action,
                       const float* __restrict__ reward, const float* __restrict__ next_obs, const uint8_t* __restrict__ done,
                       long long E, int row_len, const mnv_vstep_ctl* __restrict__ ctl)
{
    if (ctl != nullptr) pos = ctl->rpl_pos;
    const long long n2 = E * row_len / 2, ring2 = R.capacity * row_len / 2, base2 = pos * row_len / 2;   // row_len is even here
    const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float2* __restrict__ s2 = reinterpret_cast<const float2*>(obs);
    const float2* __restrict__ n2p = reinterpret_cast<const float2*>(next_obs);
    float2* d_s = reinterpret_cast<float2*>(R.states);
    float2* d_n = reinterpret_cast<float2*>(R.next_states);
    for (long long i = i0; i < n2; i += stride) {
        long long d = base2 + i;
        if (d >= ring2) d -= ring2;
        d_s[d] = s2[i];
        d_n[d] = n2p[i];
    }
    for (long long e = i0; e < E; e += stride) {
        long long slot = pos + e;
        if (slot >= R.capacity) slot -= R.capacity;
        R.actions[slot] = (int64_t)action[e];
        R.rewards[slot] = reward[e];
        R.dones[slot] = done[e] ? 1.0f : 0.0f;
    }
}

// B picks by ONE CTA.  Thread j owns picks j, j + 1024, ...  Draw r of pick j = Philox(key = seed, counter = (j, r, call)).
// Without replacement: a pick is rejected iff a LOWER-numbered pick currently holds the same value; rejected picks take
// their next draw; repeat until no pick is rejected.  The lowest holder of a value always keeps it, every rejection is
// therefore against a kept value, and the procedure commutes with any relabelling of the index range -- so the result is a
// uniformly distributed ordered tuple of distinct indices, the distribution of random.sample, and it is deterministic
// (no atomics, no races: the tie-break is the pick number).
constexpr int kDrawThreads = 1024, kMaxPicksPerThread = 8;           // B <= 8192

__device__ __forceinline__ long long draw_index(unsigned long long seed, unsigned long long call, unsigned pick, unsigned round, long long size)
{
    const philox::u4 r = philox::philox4x32_10(philox::u4{pick, round, (uint32_t)call, (uint32_t)(call >> 32)},
                                                (uint32_t)seed, (uint32_t)(seed >> 32));
    const unsigned long long r64 = ((unsigned long long)r.x << 32) | r.y;
    return (long long)__umul64hi(r64, (unsigned long long)size);      // floor(r64 * size / 2^64): bias < size / 2^64
}

__global__ void __launch_bounds__(kDrawThreads)
rpl_draw_kernel(long long* __restrict__ picks, long long B, long long size, unsigned long long seed, unsigned long long call,
                int without_replacement, const mnv_vstep_ctl* __restrict__ ctl)
{
    if (ctl != nullptr) { size = ctl->rpl_size; call = ctl->rpl_call; }
    extern __shared__ long long s_pick[];                              // [B]
    __shared__ int s_changed;
    unsigned round_of[kMaxPicksPerThread];
    for (int k = 0; k < kMaxPicksPerThread; ++k) {
        round_of[k] = 0;
        const long long j = threadIdx.x + (long long)k * kDrawThreads;
        if (j < B) s_pick[j] = draw_index(seed, call, (unsigned)j, 0u, size);
    }
    __syncthreads();
    if (without_replacement) {
        for (;;) {
            if (threadIdx.x == 0) s_changed = 0;
            __syncthreads();
            bool rejected[kMaxPicksPerThread];
            for (int k = 0; k < kMaxPicksPerThread; ++k) {
                rejected[k] = false;
                const long long j = threadIdx.x + (long long)k * kDrawThreads;
                if (j >= B) break;
                const long long v = s_pick[j];
                for (long long u = 0; u < j; ++u)                      // shared-memory broadcast reads; B^2 / 2 comparisons per round in total
                    if (s_pick[u] == v) { rejected[k] = true; break; }
            }
            __syncthreads();                                           // every comparison of this round saw the same values
            for (int k = 0; k < kMaxPicksPerThread; ++k) {
                const long long j = threadIdx.x + (long long)k * kDrawThreads;
                if (j >= B) break;
                if (rejected[k]) { s_pick[j] = draw_index(seed, call, (unsigned)j, ++round_of[k], size); s_changed = 1; }
            }
            __syncthreads();
            if (!s_changed) break;
            __syncthreads();
        }
    }
    for (long long j = threadIdx.x; j < B; j += kDrawThreads) picks[j] = s_pick[j];
}

__global__ void __launch_bounds__(kWarps * 32)
rpl_gather_kernel(Ring R, long long head, long long size, const long long* __restrict__ picks,
                  float* __restrict__ o_states, int64_t* __restrict__ o_actions, float* __restrict__ o_rewards,
                  float* __restrict__ o_next, float* __restrict__ o_dones, long long B, int row_len,
                  const mnv_vstep_ctl* __restrict__ ctl)
{
    if (ctl != nullptr) { head = ctl->rpl_head; size = ctl->rpl_size; }
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (b >= B) return;
    long long i = picks[b];
    i = i < 0 ? 0 : (i >= size ? size - 1 : i);                        // defensive: never read outside the stored range
    const long long slot = (head + i) % R.capacity;
    copy_row(o_states + b * row_len, R.states + slot * row_len, row_len, lane);
    copy_row(o_next + b * row_len, R.next_states + slot * row_len, row_len, lane);
    if (lane == 0) { o_actions[b] = R.actions[slot]; o_rewards[b] = R.rewards[slot]; o_dones[b] = R.dones[slot]; }
}

// The 2 x B x 8 quantile samples of one update (target taus first, Q9): 4 uniforms per Philox draw.
__global__ void __launch_bounds__(256)
iqn_draw_taus_kernel(float* __restrict__ taus, long long n, unsigned long long seed, unsigned long long call,
                     const mnv_vstep_ctl* __restrict__ ctl)
{
    if (ctl != nullptr) call = ctl->rpl_call;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // draw q covers taus[4 q .. 4 q + 3]
    if (4 * q >= n) return;
    const unsigned long long key = seed ^ 0x7A75ull;
    const philox::u4 r = philox::philox4x32_10(philox::u4{(uint32_t)q, 0x7Au ^ (uint32_t)((unsigned long long)q >> 32 << 8), (uint32_t)call, (uint32_t)(call >> 32)},
                                                (uint32_t)key, (uint32_t)(key >> 32));
    const float u[4] = {philox::u01(r.x), philox::u01(r.y), philox::u01(r.z), philox::u01(r.w)};
    for (int k = 0; k < 4; ++k)
        if (4 * q + k < n) taus[4 * q + k] = u[k];
}

int check_ring(const char* fn, const float* s, const int64_t* a, const float* r, const float* n, const float* d, int64_t cap, int32_t row_len)
{
    if (s == nullptr || a == nullptr || r == nullptr || n == nullptr || d == nullptr) { mnv_set_error("%s: null ring pointer", fn); return MNV_E_NULL; }
    if ((reinterpret_cast<uintptr_t>(s) & 15u) || (reinterpret_cast<uintptr_t>(n) & 15u)) { mnv_set_error("%s: ring rows not 16-byte aligned", fn); return MNV_E_ALIGN; }
    if (cap <= 0 || row_len <= 0) { mnv_set_error("%s: bad capacity / row_len (%lld / %d)", fn, (long long)cap, row_len); return MNV_E_SIZE; }
    return 0;
}

}  // namespace

static int append_impl(float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states, float* d_dones,
                       int64_t capacity, int64_t pos, const float* d_obs, const int32_t* d_action, const float* d_reward,
                       const float* d_next_obs, const uint8_t* d_done, int64_t E, int32_t row_len, int32_t n_step, double gamma,
                       int64_t t, float* d_win_obs, int32_t* d_win_action, float* d_win_reward, const mnv_vstep_ctl* d_ctl, void* stream)
{
    int rc = check_ring("rpl_append", d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, row_len);
    if (rc) return rc;
    if (d_obs == nullptr || d_action == nullptr || d_reward == nullptr || d_next_obs == nullptr || d_done == nullptr) { mnv_set_error("rpl_append: null input"); return MNV_E_NULL; }
    if (E <= 0 || E > capacity) { mnv_set_error("rpl_append: E=%lld outside (0, capacity=%lld]", (long long)E, (long long)capacity); return MNV_E_SIZE; }
    if (pos < 0 || pos >= capacity || n_step < 1 || t < 0) { mnv_set_error("rpl_append: bad pos / n_step / t"); return MNV_E_PARAM; }
    if (n_step > 1 && (d_win_obs == nullptr || d_win_action == nullptr || d_win_reward == nullptr)) { mnv_set_error("rpl_append: n_step > 1 needs the window buffers"); return MNV_E_NULL; }
    Ring R{d_states, d_actions, d_rewards, d_next_states, d_dones, capacity};
    if (n_step == 1 && (row_len & 1) == 0 && ((reinterpret_cast<uintptr_t>(d_obs) | reinterpret_cast<uintptr_t>(d_next_obs)) & 7u) == 0) {
        const long long n2 = E * row_len / 2;
        const unsigned grid = (unsigned)std::min<long long>((n2 + 255) / 256, 148ll * 8);
        rpl_append_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(R, pos, d_obs, d_action, d_reward, d_next_obs, d_done, E, row_len, d_ctl);
        return mnv_launch_status("rpl_append");
    }
    const unsigned grid = (unsigned)((E + kWarps - 1) / kWarps);
    rpl_append_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(R, pos, d_obs, d_action, d_reward, d_next_obs, d_done, E, row_len,
                                                                      n_step, gamma, t, d_win_obs, d_win_action, d_win_reward, d_ctl);
    return mnv_launch_status("rpl_append");
}

extern "C" int rpl_append(float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states, float* d_dones,
                          int64_t capacity, int64_t pos, const float* d_obs, const int32_t* d_action, const float* d_reward,
                          const float* d_next_obs, const uint8_t* d_done, int64_t E, int32_t row_len, int32_t n_step, double gamma,
                          int64_t t, float* d_win_obs, int32_t* d_win_action, float* d_win_reward, void* stream)
{
    return append_impl(d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, pos, d_obs, d_action, d_reward, d_next_obs,
                       d_done, E, row_len, n_step, gamma, t, d_win_obs, d_win_action, d_win_reward, nullptr, stream);
}

extern "C" int rpl_append_ctl(float* d_states, int64_t* d_actions, float* d_rewards, float* d_next_states, float* d_dones,
                              int64_t capacity, const float* d_obs, const int32_t* d_action, const float* d_reward,
                              const float* d_next_obs, const uint8_t* d_done, int64_t E, int32_t row_len, int32_t n_step, double gamma,
                              float* d_win_obs, int32_t* d_win_action, float* d_win_reward, const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (d_ctl == nullptr) { mnv_set_error("rpl_append_ctl: null control block"); return MNV_E_NULL; }
    return append_impl(d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, 0, d_obs, d_action, d_reward, d_next_obs,
                       d_done, E, row_len, n_step, gamma, 0, d_win_obs, d_win_action, d_win_reward, d_ctl, stream);
}

static int gather_impl(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
                       const float* d_dones, int64_t capacity, int64_t head, int64_t size, const int64_t* d_indices,
                       float* d_out_states, int64_t* d_out_actions, float* d_out_rewards, float* d_out_next_states,
                       float* d_out_dones, int64_t B, int32_t row_len, const mnv_vstep_ctl* d_ctl, void* stream)
{
    int rc = check_ring("rpl_gather", d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, row_len);
    if (rc) return rc;
    if (d_indices == nullptr || d_out_states == nullptr || d_out_actions == nullptr || d_out_rewards == nullptr || d_out_next_states == nullptr || d_out_dones == nullptr) { mnv_set_error("rpl_gather: null pointer"); return MNV_E_NULL; }
    if (B <= 0 || (d_ctl == nullptr && (size <= 0 || size > capacity || head < 0 || head >= capacity))) { mnv_set_error("rpl_gather: bad B / size / head"); return MNV_E_SIZE; }
    Ring R{const_cast<float*>(d_states), const_cast<int64_t*>(d_actions), const_cast<float*>(d_rewards), const_cast<float*>(d_next_states), const_cast<float*>(d_dones), capacity};
    const unsigned grid = (unsigned)((B + kWarps - 1) / kWarps);
    rpl_gather_kernel<<<grid, kWarps * 32, 0, (cudaStream_t)stream>>>(R, head, size, reinterpret_cast<const long long*>(d_indices), d_out_states, d_out_actions,
                                                                      d_out_rewards, d_out_next_states, d_out_dones, B, row_len, d_ctl);
    return mnv_launch_status("rpl_gather");
}

extern "C" int rpl_gather(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
                          const float* d_dones, int64_t capacity, int64_t head, int64_t size, const int64_t* d_indices,
                          float* d_out_states, int64_t* d_out_actions, float* d_out_rewards, float* d_out_next_states,
                          float* d_out_dones, int64_t B, int32_t row_len, void* stream)
{
    return gather_impl(d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, head, size, d_indices, d_out_states,
                       d_out_actions, d_out_rewards, d_out_next_states, d_out_dones, B, row_len, nullptr, stream);
}

static int sample_impl(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
                       const float* d_dones, int64_t capacity, int64_t head, int64_t size, uint64_t seed, uint64_t call,
                       int32_t without_replacement, int64_t* d_indices, float* d_out_states, int64_t* d_out_actions,
                       float* d_out_rewards, float* d_out_next_states, float* d_out_dones, int64_t B, int32_t row_len,
                       const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (d_indices == nullptr) { mnv_set_error("rpl_sample: null index buffer"); return MNV_E_NULL; }
    if (B <= 0 || B > (int64_t)kDrawThreads * kMaxPicksPerThread) { mnv_set_error("rpl_sample: B=%lld outside [1, %d]", (long long)B, kDrawThreads * kMaxPicksPerThread); return MNV_E_SIZE; }
    if (d_ctl == nullptr && (size <= 0 || (without_replacement && B > size))) { mnv_set_error("rpl_sample: %lld picks from %lld stored transitions", (long long)B, (long long)size); return MNV_E_SIZE; }
    const size_t smem = (size_t)B * sizeof(long long);
    if (smem > 48 * 1024) {
        cudaError_t a = cudaFuncSetAttribute(rpl_draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (a != cudaSuccess) { mnv_set_error("rpl_sample: cudaFuncSetAttribute: %s", cudaGetErrorString(a)); return (int)a; }
    }
    rpl_draw_kernel<<<1, kDrawThreads, smem, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(d_indices), B, size, seed, call, without_replacement, d_ctl);
    int rc = mnv_launch_status("rpl_sample(draw)");
    if (rc) return rc;
    return gather_impl(d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, head, size, d_indices, d_out_states, d_out_actions,
                       d_out_rewards, d_out_next_states, d_out_dones, B, row_len, d_ctl, stream);
}

extern "C" int rpl_sample(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
                          const float* d_dones, int64_t capacity, int64_t head, int64_t size, uint64_t seed, uint64_t call,
                          int32_t without_replacement, int64_t* d_indices, float* d_out_states, int64_t* d_out_actions,
                          float* d_out_rewards, float* d_out_next_states, float* d_out_dones, int64_t B, int32_t row_len, void* stream)
{
    return sample_impl(d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, head, size, seed, call, without_replacement,
                       d_indices, d_out_states, d_out_actions, d_out_rewards, d_out_next_states, d_out_dones, B, row_len, nullptr, stream);
}

extern "C" int rpl_sample_ctl(const float* d_states, const int64_t* d_actions, const float* d_rewards, const float* d_next_states,
                              const float* d_dones, int64_t capacity, uint64_t seed, int32_t without_replacement, int64_t* d_indices,
                              float* d_out_states, int64_t* d_out_actions, float* d_out_rewards, float* d_out_next_states,
                              float* d_out_dones, int64_t B, int32_t row_len, const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (d_ctl == nullptr) { mnv_set_error("rpl_sample_ctl: null control block"); return MNV_E_NULL; }
    return sample_impl(d_states, d_actions, d_rewards, d_next_states, d_dones, capacity, 0, 0, seed, 0, without_replacement,
                       d_indices, d_out_states, d_out_actions, d_out_rewards, d_out_next_states, d_out_dones, B, row_len, d_ctl, stream);
}

extern "C" int iqn_draw_taus(float* d_taus, int64_t n, uint64_t seed, uint64_t call, const mnv_vstep_ctl* d_ctl, void* stream)
{
    if (d_taus == nullptr) { mnv_set_error("iqn_draw_taus: null output"); return MNV_E_NULL; }
    if (n <= 0) { mnv_set_error("iqn_draw_taus: n must be > 0"); return MNV_E_SIZE; }
    const long long draws = (n + 3) / 4;
    iqn_draw_taus_kernel<<<(unsigned)((draws + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_taus, n, seed, call, d_ctl);
    return mnv_launch_status("iqn_draw_taus");
}

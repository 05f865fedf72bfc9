/*
 * mnv_host.h -- C-ABI of libmnv_host.so, the native HOST-side helper of VecMarineNavEnv.step_host (plain C, pthreads; no
 * CUDA).  It expands the compact observation packet produced on the device by mnv_pack_obs (marinenav_b200.h) into the dense
 * row-major f32 [E][obs_dim] observation block (row layout of MarineNavEnv.get_observation, marinenav_env.py:273-326).
 */
#ifndef MNV_HOST_H
#define MNV_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mnvh_pool mnvh_pool;

/* n_threads workers (the caller is one of them) for E environments with obs_dim = 4 + 2 * n_beams floats per row.
 * cpu_first >= 0 pins worker t to CPU cpu_first + t; < 0 leaves the placement to the OS.  NULL on bad arguments. */
mnvh_pool* mnvh_create(int n_threads, int64_t E, int obs_dim, int cpu_first);
void       mnvh_destroy(mnvh_pool* p);
int        mnvh_threads(const mnvh_pool* p);

/* obs holds a complete dense block (after a reset / a dense refresh): rebuild the bookkeeping of non-zero beam slots. */
void mnvh_rescan(mnvh_pool* p, float* obs);

/* One packet of mnv_pack_obs -> obs, in place: head f32 [E][4]; mask u32 [E][ceil(n_beams / 32)] (bit b: beam b has a
 * return); dir u32 [ceil(E / 32)] (where the returns of environments 32 g .. 32 g + 31 start in vals); vals f32 [.][2];
 * skip u8 [E] (may be NULL): rows the caller has already written (left alone, then re-scanned). */
void mnvh_expand(mnvh_pool* p, float* obs, const float* head, const uint8_t* skip, const uint32_t* mask, const uint32_t* dir,
                 const float* vals);

/* The same for a packet that is expanded BEFORE the rows flagged in skip have been written by the caller (the hybrid
 * transport of step_host expands the first packet while the GPU is still producing the re-observed rows): those rows are
 * neither written nor re-scanned by mnvh_expand_early; mnvh_rescan_skipped(obs, skip) re-scans them once they have landed. */
void mnvh_expand_early(mnvh_pool* p, float* obs, const float* head, const uint8_t* skip, const uint32_t* mask, const uint32_t* dir,
                       const float* vals);
void mnvh_rescan_skipped(mnvh_pool* p, float* obs, const uint8_t* skip);

#ifdef __cplusplus
}
#endif
#endif

// Shared host/device helpers of libmarinenav_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "marinenav_b200.h"

#define MNV_PI 3.14159265358979323846

void mnv_set_error(const char* fmt, ...);
enum { MNV_OPT_TMA = 0, MNV_OPT_PDL = 1, MNV_OPT_ACT_TIMING = 2, MNV_OPT_COUNT = 3 };
int mnv_option(int which);     // tuning switches, see mnv_set_option

#define MNV_CHECK_PTR(p)                                                            \
    do {                                                                            \
        if ((p) == nullptr) { mnv_set_error("%s: null pointer " #p, __func__); return MNV_E_NULL; } \
        if ((reinterpret_cast<uintptr_t>(p) & 15u) != 0) { mnv_set_error("%s: " #p " not 16-byte aligned", __func__); return MNV_E_ALIGN; } \
    } while (0)

#define MNV_CHECK_PTR_OPT(p)                                                        \
    do {                                                                            \
        if ((p) != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) != 0) { mnv_set_error("%s: " #p " not 16-byte aligned", __func__); return MNV_E_ALIGN; } \
    } while (0)

static inline int mnv_launch_status(const char* what)
{
    cudaError_t err = cudaPeekAtLastError();
    if (err != cudaSuccess) {
        mnv_set_error("%s: %s", what, cudaGetErrorString(err));
        cudaGetLastError();
        return (int)err;
    }
    return 0;
}

// Parameter layout of the IQN network (ObsEncoder, thirdparty/IQN/model.py:125-136) shared by the IQN kernels.
#pragma once
#include "mnv_common.cuh"

namespace iqn {

constexpr int kObs = 26, kFeat = 208, kCos = 64, kHid = 64, kAct = 9, kTrainTaus = 8;
#define MNV_PI_D 3.14159265358979323846

// flat fp32 parameter vector = state_dict order (velocity_encoder.weight ... output_layer.bias), 35 785 floats
constexpr int oVW = 0, oVB = 32, oGW = 48, oGB = 80, oSW = 96, oSB = 3968, oCW = 4144, oCB = 17456,
              oH1W = 17664, oH1B = 30976, oH2W = 31040, oH2B = 35136, oOW = 35200, oOB = 35776, kParams = 35785;
// per-tensor boundaries (for torch's per-tensor clip norm) in the same order
__host__ __device__ constexpr int tensor_begin(int i)
{
    constexpr int b[15] = {oVW, oVB, oGW, oGB, oSW, oSB, oCW, oCB, oH1W, oH1B, oH2W, oH2B, oOW, oOB, kParams};
    return b[i];
}

// packed transposes produced by iqn_pack (forward GEMMs want [in][out]):
//   WcT [64][208] | W1T [208][64] | W2T [64][64]
constexpr int ptWc = 0, ptW1 = 13312, ptW2 = 26624, kPacked = 30720;

}  // namespace iqn

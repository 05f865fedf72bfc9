// Parameter layout of the IQN network (ObsEncoder, thirdparty/IQN/model.py:125-136) shared by the IQN kernels.
#pragma once
#include "mnv_common.cuh"

namespace iqn {

constexpr int kObs = 26, kFeat = 208, kCos = 64, kHid = 64, kAct = 9, kTrainTaus = 8;
#define MNV_PI_D 3.14159265358979323846

// flat fp32 parameter vector = state_dict order (velocity_encoder.weight ... output_layer.bias), 35 785 floats
constexpr int oVW = 0, oVB = 32, oGW = 48, oGB = 80, oSW = 96, oSB = 3968, oCW = 4144, oCB = 17456,
              oH1W = 17664, oH1B = 30976, oH2W = 31040, oH2B = 35136, oOW = 35200, oOB = 35776, kParams = 35785;
// stride of one tile's partial gradient in the training scratch: kParams rounded up to a whole number of 128-byte lines
// (kParams is odd; the weight-gradient epilogues store 8-byte pairs)
constexpr int kPartStride = 35840;
// per-tensor boundaries (for torch's per-tensor clip norm) in the same order
__host__ __device__ constexpr int tensor_begin(int i)
{
    constexpr int b[15] = {oVW, oVB, oGW, oGB, oSW, oSB, oCW, oCB, oH1W, oH1B, oH2W, oH2B, oOW, oOB, kParams};
    return b[i];
}

// packed transposes produced by iqn_pack (forward GEMMs want [in][out]):
//   WcT [64][208] | W1T [208][64] | W2T [64][64]
constexpr int ptWc = 0, ptW1 = 13312, ptW2 = 26624, kPacked = 30720;

// ---- bf16 weight tiles of the tensor-core acting kernel (iqn_act_tc.cu) ----
// UMMA canonical K-major layout without swizzle: 8 x 8 (bf16) core matrices of 128 contiguous bytes; the core matrices of
// one 8-row group are contiguous along K (LBO = 128 B) and row groups follow each other (SBO = (K/8) * 128 B).
__host__ __device__ constexpr int tile_offset(int r, int k, int K)       // in elements
{
    return (r >> 3) * (K * 8) + (k >> 3) * 64 + (r & 7) * 8 + (k & 7);
}
constexpr int kN4 = 16;                 // output layer padded 9 -> 16 (UMMA N granularity at M = 128)
// The biases ride inside the GEMMs: every A operand carries one extra K-step whose first column is 1.0 (rest 0) and the
// weight tiles carry the bias in that column.  Reduction lengths including that step:
constexpr int kK0 = kCos + 16, kK1 = kFeat + 16, kK2 = kHid + 16, kK3 = kHid + 16;      // 80, 224, 80, 80
constexpr int kWcEl = kFeat * kK0, kW1El = kHid * kK1, kW2El = kHid * kK2, kW3El = kN4 * kK3;
constexpr int kPackedTcEl = kWcEl + kW1El + kW2El + kW3El;       // 37 376 bf16 = 74 752 bytes

// where parameter i of the flat vector lives in the packed buffers: fp32 transposes (-1: not packed) and bf16 tiles (-1: none)
__host__ __device__ inline void packed_slots(int i, int& pt, int& tc)
{
    pt = -1; tc = -1;
    if (i >= oCW && i < oCB) { const int j = i - oCW, f = j / kCos, k = j % kCos; pt = ptWc + k * kFeat + f; tc = tile_offset(f, k, kK0); }
    else if (i >= oCB && i < oH1W) { tc = tile_offset(i - oCB, kCos, kK0); }
    else if (i >= oH1W && i < oH1B) { const int j = i - oH1W, o = j / kFeat, k = j % kFeat; pt = ptW1 + k * kHid + o; tc = kWcEl + tile_offset(o, k, kK1); }
    else if (i >= oH1B && i < oH2W) { tc = kWcEl + tile_offset(i - oH1B, kFeat, kK1); }
    else if (i >= oH2W && i < oH2B) { const int j = i - oH2W, o = j / kHid, k = j % kHid; pt = ptW2 + k * kHid + o; tc = kWcEl + kW1El + tile_offset(o, k, kK2); }
    else if (i >= oH2B && i < oOW) { tc = kWcEl + kW1El + tile_offset(i - oH2B, kHid, kK2); }
    else if (i >= oOW && i < oOB) { const int j = i - oOW, a = j / kHid, k = j % kHid; tc = kWcEl + kW1El + kW2El + tile_offset(a, k, kK3); }
    else if (i >= oOB && i < kParams) { tc = kWcEl + kW1El + kW2El + tile_offset(i - oOB, kHid, kK3); }
}

}  // namespace iqn

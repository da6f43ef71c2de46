// fp32 "pointwise" GEMM used by every 1x1 convolution / linear layer on the path:
//     C[m, n] = act( (sum_k A[m,k] * ascale[m / rows_per_group, k] * W[n,k]) * cscale[n] + cbias[n] ) + R[m, n]
// A is [M, K] row-major (NHWC activations: K = input channels), W is [N, K] row-major
// (exactly the layout of a PyTorch Conv2d 1x1 / Linear weight), so both operands stream
// along K with 128-bit loads.  Tiles are staged K-major in shared memory with a
// register-prefetch double buffer.  K % 4 == 0 and N % 4 == 0 are required (true for every
// layer of EfficientNet-B2 and of the caption decoder).
#pragma once
#include "common.cuh"

namespace ac {

enum Act { ACT_NONE = 0, ACT_SWISH = 1, ACT_RELU = 2 };

struct GemmArgs {
    const float* A; const float* W; float* C;
    int M, N, K;
    const float* ascale = nullptr;  // [M / rows_per_group, K] per-group input-channel scale (SE gate)
    int rows_per_group = 1;
    const float* cscale = nullptr;  // [N] (folded BN gamma/sqrt(var+eps)); nullptr = 1
    const float* cbias = nullptr;   // [N]; nullptr = 0
    const float* R = nullptr;       // [M, N] residual added after the activation
    int act = ACT_NONE;
    int ldc = 0;                    // row stride of C / R (0 -> N)
};

int gemm_tn(const GemmArgs& g, cudaStream_t st);

}  // namespace ac

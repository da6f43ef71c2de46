// Shared between the inference bi-GRU (bigru.cu) and its training forward / backward (bigru_train.cu).
#pragma once
#include "gemm.cuh"

namespace ac {

constexpr int kGruH = 256;            // hidden size (8 CTAs x 32 units)
constexpr int kGruCluster = 8;
constexpr int kGruUnits = kGruH / kGruCluster;   // 32 hidden units per CTA
constexpr int kGruRows = 3 * kGruUnits;          // 96 rows of W_hh per CTA (r, z, n)
constexpr int kGruClips = 8;          // clips per cluster
constexpr int kGruThreads = 512;
constexpr int kGruWtStride = kGruRows + 1;       // 97: conflict-free transposed staging AND conflict-free matvec reads
constexpr size_t kGruSmem = ((size_t)kGruH * kGruWtStride + 2 * kGruH * kGruClips + 2 * kGruRows * kGruClips) * sizeof(float);

struct GruStepArgs {
    const float* G;        // [B*T_in, 2*3H] input projections (+ b_ih), both directions
    const float* whh[2];   // [3H, H] per direction
    const float* bhh[2];   // [3H]
    const int64_t* lens;   // [B]
    float* out;            // [B, T_out, 2H]
    int B, T_in, T_out;
    int ldg = 0;           // row stride of G (0 -> 6H)
    float* save = nullptr;   // SAVE: [B, T_out, 2 directions, 4 (r, z, n, hn), H]
    float* hprev = nullptr;  // SAVE: [B, T_out, 2H] hidden state each step started from (zero where inactive)
};

// one launch of the recurrence over all (direction, clip group) clusters; save = record the backward pass's inputs
int bigru_recurrence_launch(const GruStepArgs& a, bool save, cudaStream_t st);


}  // namespace ac

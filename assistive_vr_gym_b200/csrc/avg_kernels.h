// avg_kernels.h — launch interface between the C-ABI layer (avg_capi.cu) and the kernels (avg_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/avg_model.h"

#define AVG_K_WARPS_PER_BLOCK 4
#define AVG_K_MAXJ 24      /* 1-DoF joints per environment supported by the warp-per-environment kernel */
#define AVG_K_MAXMS 24     /* moving collision shapes per environment */
#define AVG_K_MAX_VARIANTS 4

struct AvgStepArgs {
    const unsigned char* models[AVG_K_MAX_VARIANTS];   // device ModelBlobs
    const int32_t* variant;                            // [n_env] model variant per environment (may be null)
    float* env;                                        // [n_env][AVG_ENV_STRIDE]
    const float* actions;                              // [n_env][n_action]
    float* obs;                                        // [n_env][n_obs]
    float* reward;                                     // [n_env]
    uint8_t* done;                                     // [n_env] (may be null)
    float* info;                                       // [n_env][2]  total_force_on_human, task_success
    float* terms;                                      // [n_env][8]  reward terms for parity tests (may be null)
    AvgContact* contacts;                              // [n_env][AVG_MAX_CONTACT] (may be null)
    int32_t* ncontacts;                                // [n_env]
    int n_env;
};

size_t avg_kernel_smem_bytes();
cudaError_t avg_launch_step(const AvgStepArgs& a, cudaStream_t stream);
cudaError_t avg_launch_reset_obs(const AvgStepArgs& a, cudaStream_t stream);

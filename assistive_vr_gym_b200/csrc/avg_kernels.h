// avg_kernels.h — launch interface between the C-ABI layer (avg_capi.cu) and the kernels (avg_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/avg_model.h"

#define AVG_K_WARPS_PER_BLOCK 4
#define AVG_K_MAXJ 24      /* 1-DoF joints per environment supported by the warp-per-environment kernels */
#define AVG_K_MAXMS 24     /* moving collision shapes per environment */
#define AVG_K_MAX_VARIANTS 32  /* model variants per handle: gender x (robot base pose | person of a `New` id); the section pointers of
                                  16 handles x 32 variants fill 56 KB of the 64 KB constant bank */

/* Scratch arena: per-environment hand-off between the kernels of one sub-step (floats; ints bit-cast). */
#define AVG_S_MAXDENSE (6 + 2 * AVG_MAX_CONTACT)
#define AVG_S_NC 0            /* contacts found by avg_collide_kernel                       */
#define AVG_S_NR 1            /* dense constraint rows (weld, contact normals, friction)    */
#define AVG_S_NS 2            /* joint-limit rows active in this sub-step                    */
#define AVG_S_NFR 3           /* first friction row (dense index)                            */
#define AVG_S_FCR 4           /* first contact normal row (dense index)                      */
#define AVG_S_NCS 5           /* contacts that received rows                                 */
#define AVG_S_OVERFLOW 6
#define AVG_S_ITERS 7         /* solver iterations accumulated over the env-step            */
#define AVG_S_NCAND 8         /* narrowphase candidates accumulated over the env-step (diagnostic) */
#define AVG_S_NSEP 9          /* entries of the separating-axis cache (persists across sub-steps and env-steps) */
#define AVG_S_NQ 10           /* candidates of this sub-step handed to the narrowphase kernel */
#define AVG_S_QD 16           /* [32] velocities after the unconstrained update             */
#define AVG_S_CONTACT 48      /* [AVG_MAX_CONTACT][14]: pa, pb, n, dist, shape a, shape b, impulse, pad */
#define AVG_S_CONTACT_STRIDE 14
#define AVG_S_ROWS_M (AVG_S_CONTACT + AVG_S_CONTACT_STRIDE * AVG_MAX_CONTACT)          /* [32][2] float4: motor row per dof */
#define AVG_S_ROWS_L (AVG_S_ROWS_M + 8 * 32)                                             /* [32][2] float4: limit row per dof */
#define AVG_S_ROWS_D (AVG_S_ROWS_L + 8 * 32)                                             /* [MAXDENSE][2] float4              */
#define AVG_S_MINV (AVG_S_ROWS_D + 8 * AVG_S_MAXDENSE)                                   /* [MAXJ][MAXJ]   */
#define AVG_S_J (AVG_S_MINV + AVG_K_MAXJ * AVG_K_MAXJ)                                  /* [MAXDENSE][32] */
#define AVG_S_W (AVG_S_J + 32 * AVG_S_MAXDENSE)
#define AVG_S_SEP (AVG_S_W + 32 * AVG_S_MAXDENSE)     /* [3][AVG_S_NSEPMAX] float4: separation certificates, see collide_warp */
#define AVG_S_NSEPMAX 32
#define AVG_S_POSE (AVG_S_SEP + 12 * AVG_S_NSEPMAX)   /* [32][8] body poses (pos, pad, quat) from the collide kernel's forward kinematics */
#define AVG_S_NPRES (AVG_S_POSE + 8 * 32)             /* [AVG_S_NQMAX][16] narrowphase results of the queued candidates: pa, pb, n, dist, shape a, shape b, hit */
#define AVG_S_NQMAX 128          /* = the broadphase candidate capacity (kMaxCand): a cup near the head queues > 64 child pairs */
#define AVG_S_TWIST (AVG_S_NPRES + 16 * AVG_S_NQMAX)  /* [32][8] start-of-step body twists about the model's reference point (ang, pad, lin, pad); particle tasks only */
#define AVG_S_FREEINV (AVG_S_TWIST + 8 * 32)          /* [2][12] inverse mass / world inverse inertia of the free bodies (tool); particle tasks only */
#define AVG_S_STRIDE (AVG_S_FREEINV + 24)

/* Particle scratch arena (Feeding / Drinking only): per-environment hand-off between the particle broadphase, the narrowphase
   and the solver of one internal step (floats; ints bit-cast). */
#define AVG_PS_NCAND 0        /* particle-vs-shape candidates queued for the narrowphase (result slots used)   */
#define AVG_PS_NPP 1          /* particle-particle contacts found by the particle broadphase                   */
#define AVG_PS_OVERFLOW 2
#define AVG_PS_CAND 16        /* [AVG_MAX_PCAND][8]: n(3), dist, particle, shape b, hit, pad                     */
#define AVG_PS_PP (AVG_PS_CAND + 8 * AVG_MAX_PCAND)           /* [AVG_PS_MAXPP][8]: n(3), dist, p, q, pad, pad */
#define AVG_PS_MAXPP 256
#define AVG_PS_REC (AVG_PS_PP + 8 * AVG_PS_MAXPP)             /* [AVG_MAX_PCONTACT][AVG_PS_REC_STRIDE] solver records  */
#define AVG_PS_REC_STRIDE 20
#define AVG_PS_SORTED (AVG_PS_REC + AVG_PS_REC_STRIDE * AVG_MAX_PCONTACT)   /* the same records in round order (what the sweeps read) */
#define AVG_PS_STRIDE (AVG_PS_SORTED + AVG_PS_REC_STRIDE * AVG_MAX_PCONTACT)
#define AVG_NP_PARTICLE 0x8000u   /* AvgNpItem.pair: shape-a field = AVG_NP_PARTICLE | particle index */

#define AVG_K_MAX_HANDLES 16   /* handles per process that can hold models at the same time (constant-memory table slots) */

/* One candidate pair that survived the culls of the collide kernel and needs the exact narrowphase. */
struct AvgNpItem {
    int32_t env;
    uint32_t pair;        /* moving shape a | other shape b << 16 */
    int32_t slot;         /* result slot in the environment's AVG_S_NPRES array */
    int32_t cert;         /* slot in the environment's certificate cache to fill in (or -1) */
};

struct AvgStepArgs {
    int slot;                                          // row of the constant-memory model table (avg_register_model)
    const unsigned char* models[AVG_K_MAX_VARIANTS];   // device ModelBlobs
    const int32_t* variant;                            // [n_env] model variant per environment (may be null)
    float* env;                                        // [n_env][AVG_ENV_STRIDE]
    float* scratch;                                    // [n_env][AVG_S_STRIDE]
    const float* actions;                              // [n_env][n_action]
    float* obs;                                        // [n_env][n_obs]
    float* reward;                                     // [n_env]
    uint8_t* done;                                     // [n_env] (may be null)
    float* info;                                       // [n_env][2]  total_force_on_human, task_success
    float* terms;                                      // [n_env][8]  reward terms for parity tests (may be null)
    AvgContact* contacts;                              // [n_env][AVG_MAX_CONTACT] (may be null)
    int32_t* ncontacts;                                // [n_env]
    int n_env;                                         // environments of the handle
    int env_begin, env_end;                            // range stepped by this launch sequence (avg_step: all; avg_step_host: one chunk)
    int maxblk;                                        // largest articulation block (dofs) over the uploaded variants
    int dbg;                                           // development switches (AVG_DBG), 0 in production
    int task;                                          // AVG_TASK_* of the uploaded models: selects the epilogue / reset-observation kernel
    int time_limit;                                    // > 0: done[e] = (env-steps of the episode >= time_limit), gym TimeLimit per environment; 0: done = 0
    AvgNpItem* np_queue;                               // [np_capacity] narrowphase work items of the current sub-step
    int* np_count;                                     // item counter: filled by the collide kernel, read by the narrowphase kernel, zeroed by the dynamics kernel
    int np_capacity;
    const uint8_t* mask;                               // reset / settle paths: environments to touch (null = all)
    float* part;                                       // [n_env][AVG_P_STRIDE] particle records (Feeding / Drinking), else null
    float* pscratch;                                   // [n_env][AVG_PS_STRIDE] particle scratch arena, else null
    int phase;                                         // avg_launch_step: 0 the whole env-step, 1 without the epilogue, 2 the epilogue only (of [env_begin, env_end))
    int post;                                          // this internal step ends a p.stepSimulation call: run the per-frame hooks (env.py:343-349)
    int n_internal;                                    // internal steps per stepSimulation (numSubSteps, feeding.py:289)
    unsigned long long* dbg_counters;                  // AVG_DBG & 32: [8] narrowphase counters (items, plane-test rejects, GJK calls, GJK iterations, SAT calls)
};

/* `launched` (optional) receives the number of kernels the call launched (22 for a ScratchItch step, 17 with the fused dynamics + solve kernel). */
cudaError_t avg_launch_step(const AvgStepArgs& a, int substeps, cudaStream_t stream, int* launched = nullptr);
/* `n` calls of p.stepSimulation without actions, hooks or epilogue for the environments of a.mask: the settle loop of
 * FeedingEnv.reset / DrinkingEnv.reset (feeding.py:318-320). */
cudaError_t avg_launch_settle(const AvgStepArgs& a, int n, cudaStream_t stream, int* launched = nullptr);
cudaError_t avg_launch_reset_obs(const AvgStepArgs& a, cudaStream_t stream);
cudaError_t avg_launch_arm_limit(const unsigned char* blob, const float* q4, float* logits, int n, cudaStream_t stream);
/* Publish the section pointers of a device ModelBlob in the constant-memory table read by the kernels (current device). */
cudaError_t avg_register_model(int slot, int variant, const unsigned char* d_blob, const AvgModelHeader* host_header);
/* Device-side episode reset (reference ScratchItchEnv.reset draws): rewrites the records of the masked environments from
 * the per-variant reset tables, bumps their episode counters, empties their scratch state. */
struct AvgResetArgs {
    const AvgResetTable* tables[AVG_K_MAX_VARIANTS];
    const unsigned char* models[AVG_K_MAX_VARIANTS];   // device ModelBlobs (the on-device IK walks the arm chain)
    int any_ik;                                        // some table has ik_enabled: launch avg_reset_ik_kernel after the draws
    int any_new;                                       // some table is a `New` id with an arm pose draw: launch avg_reset_new_kernel after the IK
    int n_variants;
    int n_per_gender;                                  // variants per gender (1 ScratchItch; BedBathing: one per robot base pose)
    float* env; float* scratch; int32_t* variant; int32_t* episode;
    float* part;                                       // particle records (Feeding / Drinking), else null
    uint8_t* retry;                                    // [n_env] start poses rejected by the self-contact test (util.py:41-46), written by avg_launch_reset_check
    int round;                                         // 0: first solve; k > 0: k-th re-solve of the rejected environments with fresh random restarts
    const uint8_t* mask;
    int n_env;
    uint32_t seed;
};
cudaError_t avg_launch_reset(const AvgResetArgs& r, cudaStream_t stream);
/* util.ik_random_restarts(step_sim=True), util.py:41-46: after the caller has run 5 stepSimulation calls from the solved start
 * pose, reject the poses whose robot touches itself (or that were pushed off their pose), re-solve those with new random
 * restarts (r.round), put every other masked environment back on its solved pose with zero velocities, and re-place the
 * particles. */
cudaError_t avg_launch_reset_check(const AvgResetArgs& r, cudaStream_t stream);
/* Policy inference for on-device rollouts: obs [n_env][n_obs] -> actions [n_env][n_act] (enjoy_vr.py:106-113). */
cudaError_t avg_launch_policy(const unsigned char* policy_blob, const float* obs, float* actions, int n_env, int n_obs, int n_act, cudaStream_t stream);

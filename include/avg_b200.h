/* avg_b200.h — C ABI of the B200 batched simulator that replaces the PyBullet call surface on the step path.
 *
 * Drop-in boundary.  The reference talks to its physics engine through the `pybullet` CPython
 * extension, one in-process server per environment (`p.connect(p.DIRECT)`, reference env.py:23).  A maintainer of
 * the reference binds THIS library instead (ctypes stub in INTEGRATION.md); each entry point below names the
 * reference interface it replaces.  Conventions: plain C types only, int return code (0 = ok, negative = error,
 * text via avg_last_error), no exceptions cross the ABI, the caller owns every I/O buffer, the library owns only
 * its state arena and model copies, all device work is enqueued on the stream passed in (a cudaStream_t cast to
 * void*, NULL = default stream) with no hidden synchronisation unless stated.  One handle per GPU; a handle is not
 * thread-safe, distinct handles are independent.
 */
#ifndef AVG_B200_H
#define AVG_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct AvgHandle AvgHandle;

/* Replaces p.connect(p.DIRECT) for n_env environments at once (env.py:23).  device = CUDA ordinal. */
int avg_create(int device, int n_env, AvgHandle** out);
/* Replaces p.disconnect (env.py:93-94). */
int avg_destroy(AvgHandle* h);
/* Last error text of this handle (or of creation when h == NULL). */
const char* avg_last_error(const AvgHandle* h);

/* Replaces loadURDF / createMultiBody / createConstraint / setCollisionFilterPair / changeDynamics /
 * setPhysicsEngineParameter / setGravity for one model variant (world_creation.py:27-93,274-365,
 * human_creation.py:275-294, scratch_itch.py:258-260): uploads a compiled ModelBlob (include/avg_model.h).
 * All variants of a handle must share task, action and observation widths.  Synchronous. */
int avg_upload_model(AvgHandle* h, int variant, const void* blob, size_t nbytes);

/* Replaces the resetJointState / resetBasePositionAndOrientation / setJointMotorControlArray calls of reset()
 * (scratch_itch.py:130-273): writes env records [env_begin, env_begin+env_count) from HOST memory
 * (env_count x AVG_ENV_STRIDE floats) and their model variant ids (may be NULL = 0).  Synchronous. */
int avg_set_state(AvgHandle* h, int env_begin, int env_count, const float* env_records, const int32_t* variants);
/* Replaces getJointStates / getBasePositionAndOrientation / getBaseVelocity bulk reads: copies env records to HOST
 * memory.  Synchronous (waits for the stream work issued so far on the default stream of the handle). */
int avg_get_state(AvgHandle* h, int env_begin, int env_count, float* env_records);
/* gym's TimeLimit wrapper (`max_episode_steps=200`, assistive_gym/__init__.py:21) evaluated on the device: with a limit > 0
 * avg_step writes done[e] = 1 when environment e has taken that many env-steps since its last reset (the environments
 * themselves never terminate, scratch_itch.py:78); that byte array is a valid `mask` for avg_reset, so staggered episodes
 * restart without a host round trip.  0 (default) switches it off: done = 0. */
int avg_set_time_limit(AvgHandle* h, int max_episode_steps);
/* Model variant of each environment (gender x robot base pose; chosen by avg_set_state or drawn by avg_reset) to HOST
 * memory: what the reference stores as `gender` in setup.pkl (scratch_itch.py:269-272).  Synchronous. */
int avg_get_variants(AvgHandle* h, int env_begin, int env_count, int32_t* variants);
/* Device pointer of the state arena ([n_env][AVG_ENV_STRIDE] floats) for zero-copy inspection. */
float* avg_state_device_ptr(AvgHandle* h);

/* Feeding / Drinking: the food / water spheres the reference creates with createMultiBody(batchPositions=...) and reads with
 * getBasePositionAndOrientation / getBaseVelocity (feeding.py:291-307,99-108; drinking.py:291-312,110-122).  One particle record
 * per environment (AVG_P_* in include/avg_model.h: positions, velocities, alive / hit / event masks), HOST buffers of
 * env_count x avg_particle_stride() floats.  Synchronous. */
int avg_set_particles(AvgHandle* h, int env_begin, int env_count, const float* records);
int avg_get_particles(AvgHandle* h, int env_begin, int env_count, float* records);
float* avg_particles_device_ptr(AvgHandle* h);
int avg_particle_stride(void);
int avg_num_particles(const AvgHandle* h);
/* Replaces the `for _ in range(100): p.stepSimulation()` settle loop of reset() (feeding.py:318-320, drinking.py:320-322,
 * bed_bathing.py:286-292): n_steps calls of stepSimulation (each numSubSteps internal steps) without actions or per-frame
 * hooks, for the environments whose byte in `mask` (DEVICE, NULL = all) is non-zero.  avg_reset runs it itself for
 * Feeding / Drinking.  Asynchronous on `stream`. */
int avg_settle(AvgHandle* h, const uint8_t* mask, int n_steps, void* stream);

/* Episode reset on the device: replaces ScratchItchEnv.reset (scratch_itch.py:130-273) for the environments whose byte
 * in `mask` (DEVICE, [n_env], NULL = all) is non-zero, without a host round trip: gender, impairment and its
 * parameters (world_creation.py:66-72), tremor amplitudes (:141), a start pose from the variant's pool of IK solutions
 * (scratch_itch.py:251-253) -- or, when the variant's table sets ik_enabled, a freshly drawn start target and an arm
 * pose solved for it on the device (util.ik_random_restarts, util.py:34-57) --, limb and the target point on it
 * (scratch_itch.py:278, util.py:118,129), with
 * counter-based random numbers keyed by (seed, environment, episode count).  Needs avg_upload_reset_table for every
 * variant (HOST pointer to an AvgResetTable, include/avg_model.h).  obs (DEVICE, may be NULL) receives the initial
 * observation of the reset environments; other rows are left alone.  Asynchronous on `stream`. */
int avg_upload_reset_table(AvgHandle* h, int variant, const void* table, size_t nbytes);
int avg_reset(AvgHandle* h, const uint8_t* mask, uint32_t seed, float* obs, void* stream);

/* Initial observation after reset, ScratchItchEnv._get_obs([0],[0,0]) (scratch_itch.py:268).  obs is a DEVICE
 * pointer [n_env][n_obs].  Asynchronous on `stream`. */
int avg_reset_obs(AvgHandle* h, float* obs, void* stream);

/* Replaces one env.step(action) for every environment: AssistiveEnv.take_step (env.py:274-351: clip/scale the action,
 * limit-masked motor targets, setJointMotorControlArray, frame_skip x {stepSimulation, enforce_hard_human_joint_limits,
 * update_targets}) followed by get_total_force / reward / _get_obs / info (scratch_itch.py:53-128).
 * All pointers are DEVICE pointers: actions [n_env][n_action] -> obs [n_env][n_obs], reward [n_env],
 * done [n_env] (uint8, may be NULL; always 0 like the reference, scratch_itch.py:78), info [n_env][2] =
 * (total_force_on_human, task_success flag).  Asynchronous on `stream`. */
int avg_step(AvgHandle* h, const float* actions, float* obs, float* reward, uint8_t* done, float* info, void* stream);

/* Same step with HOST buffers (what a NumPy caller like the reference's examples/random_actions.py holds):
 * host->device copy of the actions, the step, device->host copy of obs/reward/done/info, and returns after the
 * results are in the host buffers.  Page-locked caller buffers (avg_alloc_host, cudaHostRegister, torch pin_memory)
 * are copied in place; pageable ones are staged through pinned memory owned by the handle. */
int avg_step_host(AvgHandle* h, const float* actions, float* obs, float* reward, uint8_t* done, float* info);
/* Page-locked host memory for the buffers of avg_step_host (replaces nothing in the reference: PyBullet returns
 * fresh Python tuples; a batched caller wants stable I/O arrays). */
int avg_alloc_host(size_t nbytes, void** out);
int avg_free_host(void* p);

/* Parity / debug taps (replace getContactPoints and the per-term prints of the reference): when enabled the step
 * also records the last sub-step's contact points and the reward terms. */
int avg_enable_debug(AvgHandle* h, int enable);
/* contacts: HOST buffer [env_count][AVG_MAX_CONTACT] of AvgContact (include/avg_model.h); counts: [env_count]. */
int avg_get_contacts(AvgHandle* h, int env_begin, int env_count, void* contacts, int32_t* counts);
/* terms: HOST buffer [env_count][8]: total_force_on_human, task_success, tool_force, tool_force_at_target,
 * reward_distance, reward_action, reward_force_scratch, preferences_score. */
int avg_get_reward_terms(AvgHandle* h, int env_begin, int env_count, float* terms);

/* Policy inference for rollouts that never leave the GPU: replaces `actor_critic.act(obs, ..., deterministic=True)` on
 * VecNormalize'd observations (enjoy_vr.py:77-113) for the default a2c_ppo_acktr MLP actor.  avg_upload_policy takes a
 * HOST blob (AvgPolicyHeader + float32 arrays, include/avg_model.h); avg_policy_act maps obs (DEVICE [n_env][n_obs]) to
 * actions (DEVICE [n_env][n_action]; columns beyond the policy's outputs are zero, enjoy_vr.py:112-113).
 * Asynchronous on `stream`. */
int avg_upload_policy(AvgHandle* h, const void* blob, size_t nbytes);
int avg_policy_act(AvgHandle* h, const float* obs, float* actions, void* stream);

/* Parity tap for enforce_realistic_human_joint_limits (env.py:353-371), which the step applies after every sub-step
 * of the human-active ids: replaces human_limits_model.predict_classes (env.py:364).  q4: DEVICE [n][4] raw joint
 * angles (tz, tx, ty, qe) of human joints 7..10; logits: DEVICE [n], class 1 (valid pose) <=> logit > 0.
 * Asynchronous on `stream`. */
int avg_arm_limit_logits(AvgHandle* h, int variant, const float* q4, float* logits, int n, void* stream);

/* Shape queries. */
int avg_num_envs(const AvgHandle* h);
int avg_num_actions(const AvgHandle* h);
int avg_num_obs(const AvgHandle* h);
int avg_env_stride(void);
/* Number of kernels launched by this handle so far (bench.py reports it as gpu_launches). */
long long avg_launch_count(const AvgHandle* h);
/* Algorithmic HBM bytes one env-step moves per environment (state read + written, action, obs, reward, info),
 * computed from the record layout; bench.py's roofline uses it. */
int avg_bytes_per_env_step(const AvgHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* AVG_B200_H */

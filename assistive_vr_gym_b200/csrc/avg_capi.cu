// avg_capi.cu — C-ABI (include/avg_b200.h) over the sm_100a kernels.  Host side in C++: handle, state arena,
// model copies, pinned staging for the host-buffer step.  There is NO CPU fallback: every entry point that computes
// launches a CUDA kernel or fails with an error code.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "../../include/avg_b200.h"
#include "../../include/avg_model.h"
#include "avg_kernels.h"

struct AvgHandle {
    int slot = -1;                                     // row of the kernels' constant-memory model table
    int device = 0;
    int n_env = 0;
    int n_act = 0, n_obs = 0, task = -1;
    unsigned char* d_model[AVG_K_MAX_VARIANTS] = {nullptr, nullptr, nullptr, nullptr};
    AvgModelHeader hdr[AVG_K_MAX_VARIANTS];
    bool have[AVG_K_MAX_VARIANTS] = {false, false, false, false};
    float* d_env = nullptr;
    float* d_scratch = nullptr;
    float* d_part = nullptr;                           // particle records (Feeding / Drinking), allocated with the first such model
    float* d_pscratch = nullptr;
    int n_particle = 0, n_internal = 1;
    int substeps = 5;
    int maxblk = 0;
    int32_t* d_variant = nullptr;
    unsigned char* d_policy = nullptr;                 // uploaded policy blob (avg_upload_policy)
    int32_t* d_episode = nullptr;                      // episodes started per environment (device reset counter)
    uint8_t* d_retry = nullptr;                        // start poses rejected by the self-contact test of the device reset
    AvgResetTable* d_rtab[AVG_K_MAX_VARIANTS] = {nullptr, nullptr, nullptr, nullptr};
    // debug taps
    bool debug = false;
    AvgContact* d_contacts = nullptr;
    int32_t* d_ncontacts = nullptr;
    float* d_terms = nullptr;
    // staging for avg_step_host
    float *h_act = nullptr, *h_obs = nullptr, *h_rew = nullptr, *h_info = nullptr;
    uint8_t* h_done = nullptr;
    float *d_act = nullptr, *d_obs = nullptr, *d_rew = nullptr, *d_info = nullptr;
    uint8_t* d_done = nullptr;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    // narrowphase work queues: set 0 for avg_step and even chunks of avg_step_host, set 1 for odd chunks (second stream)
    AvgNpItem* d_npq[4] = {nullptr, nullptr, nullptr, nullptr}; int* d_npc[4] = {nullptr, nullptr, nullptr, nullptr}; int np_capacity = 0;
    cudaStream_t xstream[2] = {nullptr, nullptr};      // third / fourth stream of avg_step (AVG_STEP_CHUNKS=3|4)
    cudaEvent_t ev_xjoin[2] = {nullptr, nullptr};
    cudaStream_t stream2 = nullptr;
    bool rtab_ik[AVG_K_MAX_VARIANTS] = {}; bool any_ik = false;     // reset tables that ask for the on-device IK start pose
    bool rtab_new[AVG_K_MAX_VARIANTS] = {}; bool any_new = false;   // reset tables of `New` ids with a per-episode arm pose draw
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // avg_step on two streams: fork from / join into the caller's stream
    int step_chunks = 1;
    int time_limit = 0;                                // avg_set_time_limit
    unsigned long long* d_cnt = nullptr;               // AVG_DBG & 32 (development aid)
    std::string err;
};

static std::string g_create_error;
static bool g_slot_used[AVG_K_MAX_HANDLES];

#define AVG_CHECK(h, call)                                                                     \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

static int fail(AvgHandle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

extern "C" {

const char* avg_last_error(const AvgHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int avg_create(int device, int n_env, AvgHandle** out) {
    if (!out || n_env <= 0) return fail(nullptr, -1, "avg_create: bad arguments");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, -3, std::string("avg_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
    if (device < 0 || device >= count) return fail(nullptr, -1, "avg_create: device ordinal out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, -2, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    int slot = -1;
    for (int i = 0; i < AVG_K_MAX_HANDLES && slot < 0; ++i) if (!g_slot_used[i]) slot = i;
    if (slot < 0) return fail(nullptr, -5, "avg_create: too many live handles in this process (AVG_K_MAX_HANDLES)");
    AvgHandle* h = new AvgHandle();
    h->device = device; h->n_env = n_env; h->slot = slot; g_slot_used[slot] = true;
    if (cudaMalloc(&h->d_env, sizeof(float) * AVG_ENV_STRIDE * (size_t)n_env) != cudaSuccess ||
        cudaMalloc(&h->d_variant, sizeof(int32_t) * (size_t)n_env) != cudaSuccess ||
        cudaMalloc(&h->d_scratch, sizeof(float) * AVG_S_STRIDE * (size_t)n_env) != cudaSuccess) {
        g_slot_used[slot] = false;
        delete h;
        return fail(nullptr, -2, "avg_create: cudaMalloc of the state arena failed");
    }
    cudaMemset(h->d_env, 0, sizeof(float) * AVG_ENV_STRIDE * (size_t)n_env);
    cudaMemset(h->d_variant, 0, sizeof(int32_t) * (size_t)n_env);
    if (cudaMalloc(&h->d_episode, sizeof(int32_t) * (size_t)n_env) != cudaSuccess) {
        g_slot_used[slot] = false; cudaFree(h->d_env); cudaFree(h->d_variant); cudaFree(h->d_scratch); delete h;
        return fail(nullptr, -2, "avg_create: cudaMalloc of the episode counters failed");
    }
    cudaMemset(h->d_episode, 0, sizeof(int32_t) * (size_t)n_env);
    cudaMemset(h->d_scratch, 0, sizeof(float) * AVG_S_STRIDE * (size_t)n_env);
    h->np_capacity = n_env * 12 + 4096;                /* 1-6 candidates per environment and sub-step survive the culls (more late in
                                                          random-action episodes); overflow is flagged, never silent */
    for (int k = 0; k < 4; ++k) {
        if (cudaMalloc(&h->d_npq[k], sizeof(AvgNpItem) * (size_t)h->np_capacity) != cudaSuccess || cudaMalloc(&h->d_npc[k], 2 * sizeof(int)) != cudaSuccess) {
            g_slot_used[slot] = false; delete h;
            return fail(nullptr, -2, "avg_create: cudaMalloc of the narrowphase queue failed");
        }
        cudaMemset(h->d_npc[k], 0, 2 * sizeof(int));
    }
    cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    /* avg_step splits the batch into two halves on two streams: the kernels of one half fill the tails (and the
       sparsely populated narrowphase kernel) of the other; at small batches (one partial wave per kernel) the two dependent
       chains of short kernels overlap (+7 % at 4096 environments).  AVG_STEP_CHUNKS=1 restores the single-stream sequence. */
    h->step_chunks = (n_env >= 2048 && n_env < 131072) ? 2 : 1;      /* staggered episodes, one B200: 32768 envs +6.6 %, 65536 +1.6 %, 131072 / 196608 +-0.5 %, 393216 -1.6 % */
    { const char* c = getenv("AVG_STEP_CHUNKS"); if (c && atoi(c) > 0) h->step_chunks = atoi(c) > 4 ? 4 : atoi(c); }
    for (int k = 0; k < 2; ++k) { cudaStreamCreateWithFlags(&h->xstream[k], cudaStreamNonBlocking); cudaEventCreateWithFlags(&h->ev_xjoin[k], cudaEventDisableTiming); }
    *out = h;
    return 0;
}

int avg_destroy(AvgHandle* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->d_cnt) {
        unsigned long long c[8]; cudaMemcpy(c, h->d_cnt, 64, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[avg narrowphase counters] items %llu, rejected by the plane test %llu, GJK calls %llu, GJK iterations %llu, SAT fallbacks %llu, contacts %llu, cycles/item mean %llu max %llu\n", c[0], c[1], c[2], c[3], c[4], c[5], c[0] ? c[6] / c[0] : 0ull, c[7]);
        cudaFree(h->d_cnt);
    }
    for (int k = 0; k < 4; ++k) { cudaFree(h->d_npq[k]); cudaFree(h->d_npc[k]); }
    for (int k = 0; k < 2; ++k) { if (h->xstream[k]) cudaStreamDestroy(h->xstream[k]); if (h->ev_xjoin[k]) cudaEventDestroy(h->ev_xjoin[k]); }
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int v = 0; v < AVG_K_MAX_VARIANTS; ++v) { cudaFree(h->d_model[v]); cudaFree(h->d_rtab[v]); }
    cudaFree(h->d_episode); cudaFree(h->d_policy); cudaFree(h->d_retry); cudaFree(h->d_part); cudaFree(h->d_pscratch);
    cudaFree(h->d_env); cudaFree(h->d_scratch); cudaFree(h->d_variant); cudaFree(h->d_contacts); cudaFree(h->d_ncontacts); cudaFree(h->d_terms);
    cudaFree(h->d_act); cudaFree(h->d_obs); cudaFree(h->d_rew); cudaFree(h->d_info); cudaFree(h->d_done);
    cudaFreeHost(h->h_act); cudaFreeHost(h->h_obs); cudaFreeHost(h->h_rew); cudaFreeHost(h->h_info); cudaFreeHost(h->h_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->slot >= 0) g_slot_used[h->slot] = false;
    delete h;
    return 0;
}

int avg_upload_model(AvgHandle* h, int variant, const void* blob, size_t nbytes) {
    if (!h || !blob) return -1;
    if (variant < 0 || variant >= AVG_K_MAX_VARIANTS) return fail(h, -1, "avg_upload_model: variant out of range");
    if (nbytes < sizeof(AvgModelHeader)) return fail(h, -1, "avg_upload_model: blob too small");
    const AvgModelHeader* mh = (const AvgModelHeader*)blob;
    if (mh->magic != AVG_MAGIC || mh->version != AVG_VERSION) return fail(h, -1, "avg_upload_model: bad magic/version");
    if (mh->total_bytes != nbytes) return fail(h, -1, "avg_upload_model: size mismatch");
    if (mh->n_body + mh->n_ebody > AVG_MAX_BODY || mh->n_ebody > AVG_MAX_EBODY || mh->n_dof > AVG_MAX_DOF || mh->n_jdof > AVG_K_MAXJ || mh->n_mshape > AVG_K_MAXMS ||
        mh->n_free > 2 || mh->n_shape - mh->n_mshape > 256 || mh->n_obs_robot + mh->n_obs_human > 64 || mh->n_action_robot + mh->n_action_human > 32 ||
        mh->n_shape + mh->n_cshape + 1 >= 32768 || mh->n_particle < 0 || mh->n_particle > AVG_MAX_PARTICLE || mh->n_internal < 1 || mh->n_internal > 4)
        return fail(h, -4, "avg_upload_model: model exceeds the warp-per-environment kernel limits");
    if (mh->task < AVG_TASK_SCRATCH_ITCH || mh->task > AVG_TASK_DRINKING)
        return fail(h, -4, "avg_upload_model: unknown task");
    if ((mh->task == AVG_TASK_FEEDING || mh->task == AVG_TASK_DRINKING) && (mh->n_particle <= 0 || mh->pshape != mh->n_shape + mh->n_cshape || mh->tool_body < 0 || mh->n_frame <= AVG_F_HEAD))
        return fail(h, -4, "avg_upload_model: Feeding / Drinking models need particles, a free tool body and the head frame");
    {   /* compound shapes: children in range, at most AVG_MAX_CCHILD each */
        const AvgShape* sh = (const AvgShape*)((const char*)blob + mh->off_shape);
        int ncomp = 0;
        for (int i = 0; i < mh->n_shape; ++i) if (sh[i].type == AVG_SHAPE_COMPOUND) {
            ncomp++;
            if (i >= mh->n_mshape || sh[i].vert_off < mh->n_shape || sh[i].vert_cnt < 1 || sh[i].vert_cnt > AVG_MAX_CCHILD || sh[i].vert_off + sh[i].vert_cnt > mh->n_shape + mh->n_cshape)
                return fail(h, -4, "avg_upload_model: bad compound shape");
        }
        if (ncomp > AVG_MAX_COMPOUND) return fail(h, -4, "avg_upload_model: too many compound shapes");
    }
    if (mh->n_target < 0 || mh->n_target > AVG_MAX_TARGET) return fail(h, -4, "avg_upload_model: too many wiping targets");
    const AvgBody* bodies = (const AvgBody*)((const char*)blob + mh->off_body);
    for (int b = 0; b < mh->n_body; ++b) {
        if (bodies[b].jtype != AVG_JOINT_FREE && (bodies[b].dof != b || b >= mh->n_jdof))
            return fail(h, -4, "avg_upload_model: 1-DoF joint bodies must come first with dof == body index");
        if (bodies[b].parent >= b) return fail(h, -4, "avg_upload_model: parents must precede children");
        int cnt = 0; bool contiguous = true;
        for (int j = 0; j < mh->n_body; ++j)
            if ((bodies[j].anc_mask >> b) & 1u) { cnt++; if (j < b || j >= bodies[b].sub_end) contiguous = false; }
        if (!contiguous || cnt != bodies[b].sub_end - b) return fail(h, -4, "avg_upload_model: bodies must be in depth-first order (sub_end)");
    }
    int na = mh->n_action_robot + mh->n_action_human, no = mh->n_obs_robot + mh->n_obs_human;
    if (h->task >= 0 && (h->n_act != na || h->n_obs != no || h->task != mh->task || h->n_particle != mh->n_particle || h->n_internal != mh->n_internal))
        return fail(h, -4, "avg_upload_model: variants of one handle must share task / action / observation widths");
    cudaSetDevice(h->device);
    if (mh->n_particle > 0 && !h->d_part) {            /* particle arenas + a narrowphase queue sized for the particle candidates */
        AVG_CHECK(h, cudaMalloc(&h->d_part, sizeof(float) * AVG_P_STRIDE * (size_t)h->n_env));
        AVG_CHECK(h, cudaMalloc(&h->d_pscratch, sizeof(float) * AVG_PS_STRIDE * (size_t)h->n_env));
        AVG_CHECK(h, cudaMemset(h->d_part, 0, sizeof(float) * AVG_P_STRIDE * (size_t)h->n_env));
        AVG_CHECK(h, cudaMemset(h->d_pscratch, 0, sizeof(float) * AVG_PS_STRIDE * (size_t)h->n_env));
        const long long cap = (long long)h->n_env * (16 + 12 * mh->n_particle) + 4096;     /* worst case 16 candidates per particle; typical 3-5 */
        h->np_capacity = cap > 0x3fffffff ? 0x3fffffff : (int)cap;
        for (int k = 0; k < 4; ++k) {
            cudaFree(h->d_npq[k]); h->d_npq[k] = nullptr;
            AVG_CHECK(h, cudaMalloc(&h->d_npq[k], sizeof(AvgNpItem) * (size_t)h->np_capacity));
        }
        /* the particle solver runs ~1.4 waves at 4096 environments: halving the batch there only adds a second tail (measured:
           FeedingSawyer-v0 at 4096 envs 7.1 ms / step as one sequence, 8.1 ms as two halves) */
        if (h->n_env < 32768 && !getenv("AVG_STEP_CHUNKS")) h->step_chunks = 1;
    }
    cudaFree(h->d_model[variant]); h->d_model[variant] = nullptr;
    AVG_CHECK(h, cudaMalloc(&h->d_model[variant], nbytes));
    AVG_CHECK(h, cudaMemcpy(h->d_model[variant], blob, nbytes, cudaMemcpyHostToDevice));
    h->hdr[variant] = *mh; h->have[variant] = true;
    AVG_CHECK(h, avg_register_model(h->slot, variant, h->d_model[variant], mh));
    if (variant == 0)                                  /* variants without their own model fall back to variant 0 */
        for (int v = 1; v < AVG_K_MAX_VARIANTS; ++v) if (!h->have[v]) AVG_CHECK(h, avg_register_model(h->slot, v, h->d_model[0], mh));
    if (h->task >= 0 && h->substeps != mh->substeps && variant != 0)
        return fail(h, -4, "avg_upload_model: variants of one handle must share frame_skip");
    h->task = mh->task; h->n_act = na; h->n_obs = no; h->substeps = mh->substeps;
    h->n_particle = mh->n_particle; h->n_internal = mh->n_internal;
    for (int b = 0; b < mh->n_block; ++b) {
        int sz = mh->block_start[b + 1] - mh->block_start[b];
        if (sz > 16) return fail(h, -4, "avg_upload_model: articulation with more than 16 dofs (solver register block)");
        if (sz > h->maxblk) h->maxblk = sz;
    }
    return 0;
}

int avg_set_state(AvgHandle* h, int env_begin, int env_count, const float* env_records, const int32_t* variants) {
    if (!h || !env_records) return -1;
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_set_state: range");
    cudaSetDevice(h->device);
    /* steps run on the caller's / the handle's non-blocking streams, which do not order against the copies below */
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(h->d_env + (size_t)env_begin * AVG_ENV_STRIDE, env_records,
                            sizeof(float) * AVG_ENV_STRIDE * (size_t)env_count, cudaMemcpyHostToDevice));
    /* a new state starts with an empty separating-axis cache (and clean hand-off slots) */
    AVG_CHECK(h, cudaMemset(h->d_scratch + (size_t)env_begin * AVG_S_STRIDE, 0, sizeof(float) * AVG_S_STRIDE * (size_t)env_count));
    if (variants) {
        for (int i = 0; i < env_count; ++i)
            if (variants[i] < 0 || variants[i] >= AVG_K_MAX_VARIANTS || !h->have[variants[i]])
                return fail(h, -1, "avg_set_state: variant without an uploaded model");
        AVG_CHECK(h, cudaMemcpy(h->d_variant + env_begin, variants, sizeof(int32_t) * (size_t)env_count, cudaMemcpyHostToDevice));
    }
    AVG_CHECK(h, cudaDeviceSynchronize());              /* the memset is asynchronous: a step on another stream must not start under it */
    return 0;
}

int avg_set_particles(AvgHandle* h, int env_begin, int env_count, const float* records) {
    if (!h || !records) return -1;
    if (!h->d_part) return fail(h, -1, "avg_set_particles: this handle's task has no particles (Feeding / Drinking only)");
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_set_particles: range");
    cudaSetDevice(h->device);
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(h->d_part + (size_t)env_begin * AVG_P_STRIDE, records, sizeof(float) * AVG_P_STRIDE * (size_t)env_count, cudaMemcpyHostToDevice));
    return 0;
}

int avg_get_particles(AvgHandle* h, int env_begin, int env_count, float* records) {
    if (!h || !records) return -1;
    if (!h->d_part) return fail(h, -1, "avg_get_particles: this handle's task has no particles (Feeding / Drinking only)");
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_get_particles: range");
    cudaSetDevice(h->device);
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(records, h->d_part + (size_t)env_begin * AVG_P_STRIDE, sizeof(float) * AVG_P_STRIDE * (size_t)env_count, cudaMemcpyDeviceToHost));
    return 0;
}

float* avg_particles_device_ptr(AvgHandle* h) { return h ? h->d_part : nullptr; }
int avg_particle_stride(void) { return AVG_P_STRIDE; }
int avg_num_particles(const AvgHandle* h) { return h ? h->n_particle : 0; }

int avg_get_state(AvgHandle* h, int env_begin, int env_count, float* env_records) {
    if (!h || !env_records) return -1;
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_get_state: range");
    cudaSetDevice(h->device);
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(env_records, h->d_env + (size_t)env_begin * AVG_ENV_STRIDE,
                            sizeof(float) * AVG_ENV_STRIDE * (size_t)env_count, cudaMemcpyDeviceToHost));
    return 0;
}

float* avg_state_device_ptr(AvgHandle* h) { return h ? h->d_env : nullptr; }

static int fill_args(AvgHandle* h, AvgStepArgs& a, int qset);

int avg_settle(AvgHandle* h, const uint8_t* mask, int n_steps, void* stream) {
    if (!h || n_steps < 0) return -1;
    cudaSetDevice(h->device);
    AvgStepArgs a; memset(&a, 0, sizeof(a));
    int rc = fill_args(h, a, 0); if (rc) return rc;
    a.mask = mask;
    int nl = 0;
    AVG_CHECK(h, avg_launch_settle(a, n_steps, (cudaStream_t)stream, &nl));
    h->launches += nl;
    return 0;
}

int avg_set_time_limit(AvgHandle* h, int max_episode_steps) {
    if (!h || max_episode_steps < 0) return -1;
    h->time_limit = max_episode_steps;
    return 0;
}

int avg_get_variants(AvgHandle* h, int env_begin, int env_count, int32_t* variants) {
    if (!h || !variants) return -1;
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_get_variants: range");
    cudaSetDevice(h->device);
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(variants, h->d_variant + env_begin, sizeof(int32_t) * (size_t)env_count, cudaMemcpyDeviceToHost));
    return 0;
}

int avg_upload_policy(AvgHandle* h, const void* blob, size_t nbytes) {
    if (!h || !blob) return -1;
    if (h->task < 0) return fail(h, -1, "avg_upload_policy: upload a model first");
    if (nbytes < sizeof(AvgPolicyHeader)) return fail(h, -1, "avg_upload_policy: blob too small");
    const AvgPolicyHeader* ph = (const AvgPolicyHeader*)blob;
    if (ph->magic != AVG_POLICY_MAGIC) return fail(h, -1, "avg_upload_policy: bad magic");
    if (ph->n_in <= 0 || ph->n_in > 64 || ph->n_in > h->n_obs || ph->n_out <= 0 || ph->n_out > 32 || ph->n_out > h->n_act)
        return fail(h, -1, "avg_upload_policy: observation / action widths do not fit this environment");
    const size_t need = sizeof(AvgPolicyHeader) + sizeof(float) * ((size_t)2 * ph->n_in + (size_t)ph->n_in * 64 + 64 + 4096 + 64 + (size_t)64 * ph->n_out + ph->n_out);
    if (nbytes != need) return fail(h, -1, "avg_upload_policy: size mismatch");
    cudaSetDevice(h->device);
    cudaFree(h->d_policy); h->d_policy = nullptr;
    AVG_CHECK(h, cudaMalloc(&h->d_policy, nbytes));
    AVG_CHECK(h, cudaMemcpy(h->d_policy, blob, nbytes, cudaMemcpyHostToDevice));
    return 0;
}

int avg_policy_act(AvgHandle* h, const float* obs, float* actions, void* stream) {
    if (!h || !obs || !actions) return -1;
    if (!h->d_policy) return fail(h, -1, "avg_policy_act: no policy uploaded (avg_upload_policy)");
    cudaSetDevice(h->device);
    AVG_CHECK(h, avg_launch_policy(h->d_policy, obs, actions, h->n_env, h->n_obs, h->n_act, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

int avg_upload_reset_table(AvgHandle* h, int variant, const void* table, size_t nbytes) {
    if (!h || !table) return -1;
    if (variant < 0 || variant >= AVG_K_MAX_VARIANTS || !h->have[variant]) return fail(h, -1, "avg_upload_reset_table: upload the model of this variant first");
    if (nbytes != sizeof(AvgResetTable)) return fail(h, -1, "avg_upload_reset_table: size mismatch (AvgResetTable)");
    const AvgResetTable* t = (const AvgResetTable*)table;
    if (t->n_pool <= 0 || t->n_pool > AVG_RESET_POOL || t->n_arm > 8 || t->n_fin > 8 || t->n_hum > 8) return fail(h, -1, "avg_upload_reset_table: counts out of range");
    cudaSetDevice(h->device);
    if (!h->d_rtab[variant]) AVG_CHECK(h, cudaMalloc(&h->d_rtab[variant], sizeof(AvgResetTable)));
    AVG_CHECK(h, cudaMemcpy(h->d_rtab[variant], table, sizeof(AvgResetTable), cudaMemcpyHostToDevice));
    h->rtab_ik[variant] = t->ik_enabled != 0;
    h->any_ik = false;
    for (int v = 0; v < AVG_K_MAX_VARIANTS; ++v) h->any_ik = h->any_ik || h->rtab_ik[v];
    h->rtab_new[variant] = t->new_mode != 0 && t->hum_jitter > 0.0f;
    h->any_new = false;
    for (int v = 0; v < AVG_K_MAX_VARIANTS; ++v) h->any_new = h->any_new || h->rtab_new[v];
    return 0;
}

int avg_reset(AvgHandle* h, const uint8_t* mask, uint32_t seed, float* obs, void* stream) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    AvgResetArgs r; memset(&r, 0, sizeof(r));
    int nv = 0;
    while (nv < AVG_K_MAX_VARIANTS && h->d_rtab[nv]) { r.tables[nv] = h->d_rtab[nv]; nv++; }
    if (nv == 0) return fail(h, -1, "avg_reset: no reset table uploaded (avg_upload_reset_table)");
    r.n_variants = nv; r.n_per_gender = nv >= 2 ? nv / 2 : 1; r.env = h->d_env; r.scratch = h->d_scratch; r.variant = h->d_variant; r.episode = h->d_episode;
    r.mask = mask; r.n_env = h->n_env; r.seed = seed; r.part = h->d_part;
    for (int v = 0; v < AVG_K_MAX_VARIANTS; ++v) r.models[v] = h->d_model[v] ? h->d_model[v] : h->d_model[0];
    r.any_ik = h->any_ik ? 1 : 0; r.any_new = h->any_new ? 1 : 0;
    AVG_CHECK(h, avg_launch_reset(r, (cudaStream_t)stream));
    h->launches += 1 + r.any_ik + r.any_new + (h->d_part ? 1 : 0);
    if (r.any_ik) {
        /* util.ik_random_restarts(step_sim=True) (util.py:41-46): 5 x stepSimulation from the solved pose, poses whose robot
           touches itself (or that were pushed away) are solved again from new random restarts, up to 5 times */
        static const int rounds = getenv("AVG_IK_RETRY") ? atoi(getenv("AVG_IK_RETRY")) : 5;
        if (rounds > 0 && !h->d_retry) { AVG_CHECK(h, cudaMalloc(&h->d_retry, (size_t)h->n_env)); AVG_CHECK(h, cudaMemset(h->d_retry, 0, (size_t)h->n_env)); }
        for (int k = 1; k <= rounds; ++k) {
            int rc = avg_settle(h, mask, 5, stream); if (rc) return rc;
            r.retry = h->d_retry; r.round = k;
            AVG_CHECK(h, avg_launch_reset_check(r, (cudaStream_t)stream));
            h->launches += 2 + (h->d_part ? 1 : 0);
        }
    }
    if (h->d_part) {                                   /* "Drop food in the spoon": 100 x stepSimulation (feeding.py:318-320, drinking.py:320-322) */
        int rc = avg_settle(h, mask, 100, stream); if (rc) return rc;
    }
    if (obs) {
        AvgStepArgs a; memset(&a, 0, sizeof(a));
        int rc = fill_args(h, a, 0); if (rc) return rc;
        a.obs = obs; a.mask = mask;
        AVG_CHECK(h, avg_launch_reset_obs(a, (cudaStream_t)stream));
        h->launches++;
    }
    return 0;
}

static int fill_args(AvgHandle* h, AvgStepArgs& a, int qset) {
    if (!h->have[0]) return fail(h, -1, "no model uploaded for variant 0");
    for (int v = 0; v < AVG_K_MAX_VARIANTS; ++v) a.models[v] = h->d_model[v] ? h->d_model[v] : h->d_model[0];
    a.slot = h->slot; a.task = h->task; a.time_limit = h->time_limit;
    a.variant = h->d_variant; a.env = h->d_env; a.scratch = h->d_scratch; a.n_env = h->n_env; a.maxblk = h->maxblk;
    { const char* d = getenv("AVG_DBG"); a.dbg = d ? atoi(d) : 0; }
    a.np_queue = h->d_npq[qset]; a.np_count = h->d_npc[qset]; a.np_capacity = h->np_capacity;
    a.part = h->d_part; a.pscratch = h->d_pscratch; a.n_internal = h->n_internal; a.post = 1;
    a.env_begin = 0; a.env_end = h->n_env;
    if ((a.dbg & 32) && !h->d_cnt) { cudaMalloc(&h->d_cnt, 64); cudaMemset(h->d_cnt, 0, 64); }
    a.dbg_counters = (a.dbg & 32) ? h->d_cnt : nullptr;
    a.contacts = h->debug ? h->d_contacts : nullptr;
    a.ncontacts = h->debug ? h->d_ncontacts : nullptr;
    a.terms = h->debug ? h->d_terms : nullptr;
    return 0;
}

int avg_reset_obs(AvgHandle* h, float* obs, void* stream) {
    if (!h || !obs) return -1;
    cudaSetDevice(h->device);
    AvgStepArgs a; memset(&a, 0, sizeof(a));
    int rc = fill_args(h, a, 0); if (rc) return rc;
    a.obs = obs;
    AVG_CHECK(h, avg_launch_reset_obs(a, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

int avg_step(AvgHandle* h, const float* actions, float* obs, float* reward, uint8_t* done, float* info, void* stream) {
    if (!h || !actions || !obs || !reward || !info) return -1;
    cudaSetDevice(h->device);
    AvgStepArgs a; memset(&a, 0, sizeof(a));
    int rc = fill_args(h, a, 0); if (rc) return rc;
    a.actions = actions; a.obs = obs; a.reward = reward; a.done = done; a.info = info;
    if (h->step_chunks < 2 || h->debug) {
        int nl = 0;
        AVG_CHECK(h, avg_launch_step(a, h->substeps, (cudaStream_t)stream, &nl));
        h->launches += nl;
        return 0;
    }
    /* k equal parts on k streams (k = 2 by default; still asynchronous and capturable: event fork / join around the extra streams) */
    const int k = h->step_chunks;
    const int part = ((h->n_env + k - 1) / k + 3) & ~3;
    AVG_CHECK(h, cudaEventRecord(h->ev_fork, (cudaStream_t)stream));
    for (int c = 0; c < k; ++c) {
        AvgStepArgs b; memset(&b, 0, sizeof(b));
        rc = fill_args(h, b, c); if (rc) return rc;
        b.actions = actions; b.obs = obs; b.reward = reward; b.done = done; b.info = info;
        b.env_begin = c * part; b.env_end = (c + 1) * part < h->n_env ? (c + 1) * part : h->n_env;
        if (b.env_begin >= b.env_end) break;
        cudaStream_t st = c == 0 ? (cudaStream_t)stream : (c == 1 ? h->stream2 : h->xstream[c - 2]);
        if (c > 0) AVG_CHECK(h, cudaStreamWaitEvent(st, h->ev_fork, 0));
        int nl = 0;
        AVG_CHECK(h, avg_launch_step(b, h->substeps, st, &nl));
        h->launches += nl;
        if (c > 0) {
            cudaEvent_t ej = c == 1 ? h->ev_join : h->ev_xjoin[c - 2];
            AVG_CHECK(h, cudaEventRecord(ej, st));
            AVG_CHECK(h, cudaStreamWaitEvent((cudaStream_t)stream, ej, 0));
        }
    }
    return 0;
}

static bool is_pinned_host(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int avg_alloc_host(size_t nbytes, void** out) {
    if (!out) return -1;
    return cudaMallocHost(out, nbytes) == cudaSuccess ? 0 : -2;
}
int avg_free_host(void* p) { return cudaFreeHost(p) == cudaSuccess ? 0 : -2; }

int avg_step_host(AvgHandle* h, const float* actions, float* obs, float* reward, uint8_t* done, float* info) {
    if (!h || !actions || !obs || !reward || !info) return -1;
    cudaSetDevice(h->device);
    size_t n = (size_t)h->n_env;
    if (!h->d_act) {
        AVG_CHECK(h, cudaMalloc(&h->d_act, sizeof(float) * n * h->n_act));
        AVG_CHECK(h, cudaMalloc(&h->d_obs, sizeof(float) * n * h->n_obs));
        AVG_CHECK(h, cudaMalloc(&h->d_rew, sizeof(float) * n));
        AVG_CHECK(h, cudaMalloc(&h->d_info, sizeof(float) * n * 2));
        AVG_CHECK(h, cudaMalloc(&h->d_done, n));
    }
    /* Caller buffers that are page-locked (avg_alloc_host, cudaHostRegister, torch pin_memory) are used in place;
       pageable ones go through the handle's pinned staging buffers. */
    const bool direct = is_pinned_host(actions) && is_pinned_host(obs) && is_pinned_host(reward) && is_pinned_host(info) &&
                        (!done || is_pinned_host(done));
    if (!direct && !h->h_act) {
        AVG_CHECK(h, cudaMallocHost(&h->h_act, sizeof(float) * n * h->n_act));
        AVG_CHECK(h, cudaMallocHost(&h->h_obs, sizeof(float) * n * h->n_obs));
        AVG_CHECK(h, cudaMallocHost(&h->h_rew, sizeof(float) * n));
        AVG_CHECK(h, cudaMallocHost(&h->h_info, sizeof(float) * n * 2));
        AVG_CHECK(h, cudaMallocHost(&h->h_done, n));
    }
    const float* src_act = actions;
    float *dst_obs = obs, *dst_rew = reward, *dst_info = info; uint8_t* dst_done = done;
    if (!direct) {
        memcpy(h->h_act, actions, sizeof(float) * n * h->n_act);
        src_act = h->h_act; dst_obs = h->h_obs; dst_rew = h->h_rew; dst_info = h->h_info; dst_done = h->h_done;
    }
    /* The batch is stepped in chunks that alternate between two streams, so the device->host copy of one chunk's
       results (and the host->device copy of the next chunk's actions) overlaps the kernels of the other chunk. */
    int n_chunks = (h->n_env >= 16384 && h->n_env < 262144) ? 2 : 1;    /* measured on B200 at 196608 envs: 1 -> 22.15, 2 -> 22.08, 4 -> 22.6, 8 -> 24.0 ms;
                                                                           at 393216: 1 -> 47.19, 2 -> 47.76, 3 -> 48.43, 4 -> 48.62 ms (device-resident step: 46.93) */
    { const char* c = getenv("AVG_CHUNKS"); if (c && atoi(c) > 0) n_chunks = atoi(c); }
    /* One chunk (large batches): the sub-steps of the whole batch as one sequence, then the epilogue range by range on the same
       stream with each range's device->host copies on the second stream, so the copies overlap the epilogue of the next range
       (52 MB of results at 393216 environments: ~1 ms on PCIe 5, the epilogue ~1.1 ms). */
    int n_tail = (n_chunks == 1 && h->n_env >= 65536) ? 4 : 0;
    { const char* c = getenv("AVG_TAIL_RANGES"); if (c && n_chunks == 1) n_tail = atoi(c) > 4 ? 4 : atoi(c); }
    if (n_tail > 1) {
        cudaStream_t st = h->stream, st2 = h->stream2;
        AVG_CHECK(h, cudaMemcpyAsync(h->d_act, src_act, sizeof(float) * n * h->n_act, cudaMemcpyHostToDevice, st));
        AvgStepArgs a; memset(&a, 0, sizeof(a));
        int rc = fill_args(h, a, 0); if (rc) return rc;
        a.actions = h->d_act; a.obs = h->d_obs; a.reward = h->d_rew; a.done = h->d_done; a.info = h->d_info;
        a.phase = 1;
        int nl = 0;
        AVG_CHECK(h, avg_launch_step(a, h->substeps, st, &nl));
        h->launches += nl;
        const int per_t = ((h->n_env + n_tail - 1) / n_tail + 3) & ~3;
        for (int c = 0; c < n_tail; ++c) {
            const int b0 = c * per_t, b1 = (c + 1) * per_t < h->n_env ? (c + 1) * per_t : h->n_env;
            if (b0 >= b1) break;
            const size_t cnt = (size_t)(b1 - b0);
            a.phase = 2; a.env_begin = b0; a.env_end = b1;
            AVG_CHECK(h, avg_launch_step(a, h->substeps, st, &nl));
            h->launches += nl;
            cudaEvent_t ev = c == 0 ? h->ev_fork : (c == 1 ? h->ev_join : h->ev_xjoin[c - 2]);
            AVG_CHECK(h, cudaEventRecord(ev, st));
            AVG_CHECK(h, cudaStreamWaitEvent(st2, ev, 0));
            AVG_CHECK(h, cudaMemcpyAsync(dst_obs + (size_t)b0 * h->n_obs, h->d_obs + (size_t)b0 * h->n_obs, sizeof(float) * cnt * h->n_obs, cudaMemcpyDeviceToHost, st2));
            AVG_CHECK(h, cudaMemcpyAsync(dst_rew + b0, h->d_rew + b0, sizeof(float) * cnt, cudaMemcpyDeviceToHost, st2));
            AVG_CHECK(h, cudaMemcpyAsync(dst_info + 2 * (size_t)b0, h->d_info + 2 * (size_t)b0, sizeof(float) * cnt * 2, cudaMemcpyDeviceToHost, st2));
            if (dst_done) AVG_CHECK(h, cudaMemcpyAsync(dst_done + b0, h->d_done + b0, cnt, cudaMemcpyDeviceToHost, st2));
        }
        n_chunks = 0;                                  /* done: skip the chunk loop below */
    }
    const int per = n_chunks > 0 ? ((h->n_env + n_chunks - 1) / n_chunks + 3) & ~3 : 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int b0 = c * per, b1 = (c + 1) * per < h->n_env ? (c + 1) * per : h->n_env;
        if (b0 >= b1) break;
        const size_t cnt = (size_t)(b1 - b0);
        const int qs = c & 1;
        cudaStream_t st = qs ? h->stream2 : h->stream;
        AVG_CHECK(h, cudaMemcpyAsync(h->d_act + (size_t)b0 * h->n_act, src_act + (size_t)b0 * h->n_act, sizeof(float) * cnt * h->n_act, cudaMemcpyHostToDevice, st));
        AvgStepArgs a; memset(&a, 0, sizeof(a));
        int rc = fill_args(h, a, qs); if (rc) return rc;
        a.actions = h->d_act; a.obs = h->d_obs; a.reward = h->d_rew; a.done = h->d_done; a.info = h->d_info;
        a.env_begin = b0; a.env_end = b1;
        int nl = 0;
        AVG_CHECK(h, avg_launch_step(a, h->substeps, st, &nl));
        h->launches += nl;
        AVG_CHECK(h, cudaMemcpyAsync(dst_obs + (size_t)b0 * h->n_obs, h->d_obs + (size_t)b0 * h->n_obs, sizeof(float) * cnt * h->n_obs, cudaMemcpyDeviceToHost, st));
        AVG_CHECK(h, cudaMemcpyAsync(dst_rew + b0, h->d_rew + b0, sizeof(float) * cnt, cudaMemcpyDeviceToHost, st));
        AVG_CHECK(h, cudaMemcpyAsync(dst_info + 2 * (size_t)b0, h->d_info + 2 * (size_t)b0, sizeof(float) * cnt * 2, cudaMemcpyDeviceToHost, st));
        if (dst_done) AVG_CHECK(h, cudaMemcpyAsync(dst_done + b0, h->d_done + b0, cnt, cudaMemcpyDeviceToHost, st));
    }
    AVG_CHECK(h, cudaStreamSynchronize(h->stream));
    AVG_CHECK(h, cudaStreamSynchronize(h->stream2));
    if (!direct) {
        memcpy(obs, h->h_obs, sizeof(float) * n * h->n_obs);
        memcpy(reward, h->h_rew, sizeof(float) * n);
        memcpy(info, h->h_info, sizeof(float) * n * 2);
        if (done) memcpy(done, h->h_done, n);
    }
    return 0;
}

int avg_enable_debug(AvgHandle* h, int enable) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    if (enable && !h->d_contacts) {
        AVG_CHECK(h, cudaMalloc(&h->d_contacts, sizeof(AvgContact) * AVG_MAX_CONTACT * (size_t)h->n_env));
        AVG_CHECK(h, cudaMalloc(&h->d_ncontacts, sizeof(int32_t) * (size_t)h->n_env));
        AVG_CHECK(h, cudaMalloc(&h->d_terms, sizeof(float) * 8 * (size_t)h->n_env));
        cudaMemset(h->d_ncontacts, 0, sizeof(int32_t) * (size_t)h->n_env);
    }
    h->debug = enable != 0;
    return 0;
}

int avg_get_contacts(AvgHandle* h, int env_begin, int env_count, void* contacts, int32_t* counts) {
    if (!h || !contacts || !counts) return -1;
    if (!h->d_contacts) return fail(h, -1, "avg_get_contacts: call avg_enable_debug first");
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_get_contacts: range");
    cudaSetDevice(h->device);
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(contacts, h->d_contacts + (size_t)env_begin * AVG_MAX_CONTACT,
                            sizeof(AvgContact) * AVG_MAX_CONTACT * (size_t)env_count, cudaMemcpyDeviceToHost));
    AVG_CHECK(h, cudaMemcpy(counts, h->d_ncontacts + env_begin, sizeof(int32_t) * (size_t)env_count, cudaMemcpyDeviceToHost));
    return 0;
}

int avg_get_reward_terms(AvgHandle* h, int env_begin, int env_count, float* terms) {
    if (!h || !terms) return -1;
    if (!h->d_terms) return fail(h, -1, "avg_get_reward_terms: call avg_enable_debug first");
    if (env_begin < 0 || env_count < 0 || env_begin + env_count > h->n_env) return fail(h, -1, "avg_get_reward_terms: range");
    cudaSetDevice(h->device);
    AVG_CHECK(h, cudaDeviceSynchronize());
    AVG_CHECK(h, cudaMemcpy(terms, h->d_terms + (size_t)env_begin * 8, sizeof(float) * 8 * (size_t)env_count, cudaMemcpyDeviceToHost));
    return 0;
}

int avg_arm_limit_logits(AvgHandle* h, int variant, const float* q4, float* logits, int n, void* stream) {
    if (!h || !q4 || !logits || n < 0) return -1;
    if (variant < 0 || variant >= AVG_K_MAX_VARIANTS || !h->have[variant]) return fail(h, -1, "avg_arm_limit_logits: variant without an uploaded model");
    if (h->hdr[variant].n_mlp <= 0) return fail(h, -4, "avg_arm_limit_logits: this model carries no arm-limit classifier (human-active ids only)");
    cudaSetDevice(h->device);
    if (n == 0) return 0;
    AVG_CHECK(h, avg_launch_arm_limit(h->d_model[variant], q4, logits, n, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

int avg_num_envs(const AvgHandle* h) { return h ? h->n_env : 0; }
int avg_num_actions(const AvgHandle* h) { return h ? h->n_act : 0; }
int avg_num_obs(const AvgHandle* h) { return h ? h->n_obs : 0; }
int avg_env_stride(void) { return AVG_ENV_STRIDE; }
long long avg_launch_count(const AvgHandle* h) { return h ? h->launches : 0; }

int avg_bytes_per_env_step(const AvgHandle* h) {
    if (!h) return 0;
    /* state record read once (AVG_E_LAST floats used) + dynamic part written once + action + obs + reward + info + done;
       mirrors the load/store loops of avg_step_kernel */
    int read_state = AVG_ENV_STRIDE * 4;
    int write_state = (AVG_E_STRENGTH + (AVG_E_TARGET_ON_ARM - AVG_E_TARGET_H) + (AVG_E_LAST - AVG_E_ITERATION)) * 4;
    int io = h->n_act * 4 + h->n_obs * 4 + 4 + 8 + 1 + 4 /* variant id */;
    /* particles: position / velocity / angular velocity of each sphere read and written once, plus the mask words */
    int particles = h->n_particle > 0 ? 2 * (9 * 4 * h->n_particle) + 2 * 16 * 4 : 0;
    return read_state + write_state + io + particles;
}

}  // extern "C"

/* avg_model.h — binary layout of the compiled model ("ModelBlob") and of the per-environment state record.
 *
 * This is a DATA FORMAT definition shared by the model compiler (Python, assistive_vr_gym_b200/compiler/blob.py),
 * the CUDA product (assistive_vr_gym_b200/csrc) and the CPU oracle (oracle/avg_oracle.c, test infrastructure only).
 * It replaces what the reference keeps inside the PyBullet server after `loadURDF` / `createMultiBody` /
 * `createConstraint` / `setCollisionFilterPair` (reference world_creation.py:27-93,274-365, human_creation.py:275-294).
 *
 * Everything is little-endian, 4-byte fields, sections 16-byte aligned.  Quaternions are xyzw (PyBullet).
 */
#ifndef AVG_MODEL_H
#define AVG_MODEL_H
#include <stdint.h>

#define AVG_MAGIC   0x4D475641u  /* "AVGM" */
#define AVG_VERSION 10u

#define AVG_MAX_BODY   32   /* dynamic bodies per environment (one lane each)            */
#define AVG_MAX_DOF    32   /* velocity DoF per environment (one lane each)               */
#define AVG_MAX_EBODY   4   /* env-static bodies (pose given per environment, AVG_E_EBODY) */
#define AVG_MAX_CONTACT 32  /* articulation contact points kept per sub-step (SURVEY.md App. E budget) */
#define AVG_MAX_ROWS   160  /* oracle: constraint rows per sub-step (motors, limits, weld, 2 per contact) */
#define AVG_MAX_HULL_VERTS 48
#define AVG_MAX_PARTICLE 64 /* food / water spheres per environment (feeding.py:300, drinking.py:301) */
#define AVG_MAX_PCONTACT 320 /* particle contact points kept per internal step                */
#define AVG_MAX_PCAND   384 /* particle-vs-shape narrowphase candidates per internal step     */
#define AVG_MAX_COMPOUND 4  /* compound shapes (VHACD tool / bowl / head) per model           */
#define AVG_MAX_CCHILD  72  /* convex children of one compound                                */

enum { AVG_JOINT_REVOLUTE = 0, AVG_JOINT_PRISMATIC = 1, AVG_JOINT_FREE = 2 };
enum { AVG_SHAPE_SPHERE = 0, AVG_SHAPE_CAPSULE = 1, AVG_SHAPE_BOX = 2, AVG_SHAPE_CYLINDER = 3,
       AVG_SHAPE_HULL = 4, AVG_SHAPE_PLANE = 5,
       AVG_SHAPE_COMPOUND = 6 /* a link whose collision mesh is a VHACD set of convex hulls (spoon 64, cup 68, bowl 70, head 8-9):
                                 one broadphase entry; vert_off / vert_cnt = first child / number of children in the child section
                                 of the shape table (indices n_shape ..), which hold the hulls with body-frame AABBs */ };
enum { AVG_TASK_SCRATCH_ITCH = 0, AVG_TASK_BED_BATHING = 1, AVG_TASK_FEEDING = 2, AVG_TASK_DRINKING = 3 };
/* contact-report body ids (what the reference compares getContactPoints bodies against) */
enum { AVG_REF_ROBOT = 0, AVG_REF_HUMAN = 1, AVG_REF_TOOL = 2, AVG_REF_FURNITURE = 3, AVG_REF_PLANE = 4,
       AVG_REF_TABLE = 5, AVG_REF_BOWL = 6 /* feeding.py:111: food touching the table or the bowl is spilled */ };

/* dof flags */
#define AVG_DOF_LIMIT        1u   /* Bullet joint-limit constraint exists (revolute/prismatic with lower<=upper) */
#define AVG_DOF_MOTOR        2u   /* position motor active                                                       */
#define AVG_DOF_HUMAN        4u   /* limits scale with limit_scale, motor force with human strength, kp=human_kp */
#define AVG_DOF_HARD_LIMIT   8u   /* teleported back inside limits after every sub-step (env.py:389-410)         */

typedef struct AvgBody {          /* 32 x 4 bytes */
    int32_t  parent;              /* dynamic body index, -1 = static world                                  */
    int32_t  jtype;
    int32_t  dof;                 /* first velocity dof                                                     */
    int32_t  qidx;                /* first position coordinate in AvgEnv q[]                                */
    float    ta_pos[3];           /* parent body frame (or world) -> joint frame at q = 0                   */
    float    ta_quat[4];
    float    axis[3];             /* joint axis, joint frame                                                */
    float    tb_pos[3];           /* joint frame -> body frame (composite COM, principal axes)              */
    float    tb_quat[4];
    float    mass;
    float    inertia[3];
    float    gravity[3];          /* per-body gravity (reference setGravity(..., body=))                    */
    uint32_t anc_mask;            /* bit k: body k is this body or one of its ancestors                     */
    int32_t  ref_body;
    int32_t  ref_joint;
    int32_t  sub_end;             /* bodies are in depth-first order: the subtree of body i is [i, sub_end)      */
} AvgBody;

typedef struct AvgDof {           /* 16 x 4 bytes, one per velocity dof */
    int32_t  body;
    uint32_t flags;
    float    lower, upper;        /* enforced limits (before limit_scale)                                   */
    float    rep_lower, rep_upper;/* limits as reported by getJointInfo, used by the action mask            */
    float    kp, kd, max_force;   /* position motor (kd = PyBullet default 1.0)                             */
    int32_t  action;              /* index into the action vector, -1 = none                                */
    int32_t  human_slot;          /* index 0..9 into the reference's controllable-joint list, -1 = none     */
    float    init_target;
    float    damping;             /* joint damping torque -damping * qd (URDF <dynamics damping>, btMultibodyLink::m_jointDamping);
                                     0 in blobs written before the PR2 was compiled (the slot was padding)              */
    int32_t  pad[3];
} AvgDof;

typedef struct AvgShape {         /* 32 x 4 bytes */
    int32_t  type;
    int32_t  body;                /* dyn body, n_body+e for env-static body e, -1 = static (pose in world)   */
    int32_t  ref_body, ref_link;
    float    pos[3];              /* shape frame in the body frame (world if static)                        */
    float    quat[4];
    float    radius;              /* sphere / capsule / cylinder                                            */
    float    half[3];             /* box half extents; capsule / cylinder: half[2] = half length            */
    float    margin;              /* rounding radius added around the GJK core                              */
    int32_t  vert_off, vert_cnt;  /* hull vertices (shape frame), stored as float4 (x, y, z, 0)              */
    int32_t  plane_off, plane_cnt;/* hull face planes (n, d), n.x <= d inside                               */
    float    friction;
    float    thr;                 /* contact-breaking threshold of the owning link's compound               */
    float    aabb_c[3];           /* AABB centre / half extents: shape frame for moving shapes, world for static */
    float    aabb_h[3];
    int32_t  pad[4];
} AvgShape;

typedef struct AvgBpStatic {      /* 8 x 4 bytes: broadphase record of one static shape (same pairs as the pair table) */
    float    c[3];                /* world AABB centre                                                      */
    float    h[3];                /* world AABB half extents                                                */
    float    thr;                 /* contact-breaking threshold of the shape                                */
    uint32_t mask;                /* bit a: moving shape a may collide with this shape                      */
} AvgBpStatic;

typedef struct AvgFrame {         /* 8 x 4 bytes: a frame rigidly attached to a body */
    int32_t  body;                /* dyn body, n_body+e, or -1 (world)                                       */
    float    pos[3];
    float    quat[4];
} AvgFrame;

/* frames of interest, indices into the frame table */
enum {
    AVG_F_TOOL_TIP = 0,   /* ScratchItch: tool link 1 COM (scratch_itch.py:54,106); other tasks: tool reference frame */
    AVG_F_TOOL_BASE,      /* tool base COM frame = weld child frame                                           */
    AVG_F_WELD_PARENT,    /* weld frame on the robot end-effector link (world_creation.py:356)               */
    AVG_F_TORSO,          /* robot link 0 (15 for PR2) COM, scratch_itch.py:105                              */
    AVG_F_CHEST,          /* human link 3, scratch_itch.py:113                                               */
    AVG_F_SHOULDER,       /* human link 9  COM frame (upper arm), scratch_itch.py:118                        */
    AVG_F_ELBOW,          /* human link 11 COM frame (forearm)                                               */
    AVG_F_WRIST,          /* human link 13 COM frame (hand)                                                  */
    AVG_F_HEAD,           /* human link 27 COM frame (head), feeding.py:134,346 / drinking.py:149                     */
    AVG_F_COUNT
};

typedef struct AvgModelHeader {
    uint32_t magic, version;
    uint32_t total_bytes;
    int32_t  task;
    int32_t  n_body, n_ebody, n_dof, n_jdof, n_free;
    int32_t  n_shape, n_mshape;   /* shapes [0, n_mshape) move (dyn or env-static); the rest are static       */
    int32_t  n_vert, n_plane, n_pair, n_frame;
    int32_t  substeps;            /* frame_skip, env.py:16                                                   */
    int32_t  solver_iters;        /* numSolverIterations, scratch_itch.py:258                                */
    int32_t  n_action_robot, n_action_human, n_obs_robot, n_obs_human;
    int32_t  human_control;
    float    dt;                  /* time_step 0.02                                                          */
    float    erp;                 /* joint / contact ERP (0.2)                                               */
    float    lin_damp, ang_damp;  /* btMultiBody linear / angular damping (0.04)                             */
    float    residual_thr;        /* leastSquaresResidualThreshold (1e-7); 0 disables the early exit         */
    float    max_vel;             /* btMultiBody max coordinate velocity (100)                               */
    float    action_scale;        /* 0.05, env.py:280                                                        */
    float    weld_max_force;      /* 500, world_creation.py:364                                              */
    int32_t  weld_body_a, weld_body_b;   /* dyn body indices (robot EE composite, tool)                      */
    float    task_f[32];          /* task constants, see AVG_TF_*                                             */
    uint32_t off_body, off_dof, off_shape, off_vert, off_plane, off_pair, off_frame;
    uint32_t off_bps;             /* AvgBpStatic[n_shape - n_mshape]: broadphase records of the static shapes  */
    uint32_t off_bpm;             /* uint32[n_mshape]: bit b of entry a = moving pair (a, b), b > a, may collide */
    int32_t  n_block;             /* diagonal blocks of the joint-space mass matrix (one per articulation)     */
    int32_t  block_start[4];      /* first dof of each block; block_start[n_block] = n_jdof                    */
    uint32_t off_bcap;            /* float[n_shape][8]: bounding capsule p0(3), r, p1(3), 0 — shape frame (moving) or world (static) */
    uint32_t off_mlp;             /* arm-limit classifier (env.py:353-387): W1[4][64] b1[64] W2[64][64] b2 W3[64][64] b3 W4[64] b4   */
    int32_t  n_mlp;               /* 8705 floats when present (human-active ids), else 0                                        */
    int32_t  mlp_dof[4];          /* velocity dofs of human joints 7, 8, 9, 10 (tz, tx, ty, qe)                                  */
    uint32_t off_target;          /* BedBathing wiping targets (bed_bathing.py:360-379): float4[n_target] = position in the COM frame of
                                     human link 9 (entries [0, n_target_upper), frame AVG_F_SHOULDER) or link 11 (the rest, AVG_F_ELBOW) */
    int32_t  n_target;            /* total_target_count (129 male / 91 female), 0 for other tasks; <= AVG_MAX_TARGET                     */
    int32_t  n_target_upper;
    /* ---- version 10: compound shapes, env-static bodies, internal sub-steps, particles (Feeding / Drinking) ---- */
    int32_t  n_cshape;            /* convex children of the compound shapes: shape table entries [n_shape, n_shape + n_cshape)               */
    uint32_t off_caabb;           /* float4[n_cshape][2]: child AABB centre | half extents in the frame of the owning body                   */
    int32_t  n_internal;          /* setPhysicsEngineParameter(numSubSteps) (feeding.py:289, drinking.py:287): internal steps per
                                     p.stepSimulation, 1 when the reference passes 0; `dt` is the INTERNAL step (time_step / n_internal)     */
    int32_t  n_particle;          /* food (8, feeding.py:300) / water (64, drinking.py:301) spheres; 0 for the other tasks                   */
    int32_t  pshape;              /* shape table index of the particle template (a sphere), n_shape + n_cshape when n_particle > 0           */
    float    p_mass;              /* 0.001 (feeding.py:299)                                                                                   */
    float    p_gravity[3];        /* world gravity (0, 0, -9.81) (feeding.py:284); robot / human / tool are exempt through AvgBody.gravity    */
    int32_t  tool_body;           /* dynamic body of the spoon / cup (free body), -1 when the tool is not a single free body                  */
    int32_t  head_frozen_mask;    /* bodies of human joints 24..27: their masses are zeroed per environment (AVG_E_FROZEN) unless the
                                     episode drew a tremor or the id is human-active (feeding.py:244 + world_creation.py:157-161)             */
    float    warmstart;           /* warm-starting factor of the contact normal rows: the impulse a contact point carried in the last
                                     internal step, times this, initialises its row ([UPSTREAM-BULLET] m_warmstartingFactor: 0.85 in
                                     btContactSolverInfo, set to 0.1 by the PyBullet server -- from memory); 0 = off (blobs older than round 2) */
    uint32_t pad2[4];
} AvgModelHeader;
#define AVG_MAX_TARGET 160

/* task_f indices (config.ini + task files) */
enum {
    AVG_TF_DISTANCE_W = 0, AVG_TF_ACTION_W, AVG_TF_TOOL_FORCE_W, AVG_TF_SCRATCH_W, AVG_TF_SUCCESS_THR,
    AVG_TF_C_V, AVG_TF_C_F, AVG_TF_C_HF, AVG_TF_C_FD, AVG_TF_C_FDV,
    AVG_TF_TARGET_RADIUS,        /* 0.025, scratch_itch.py:97 */
    AVG_TF_SCRATCH_MOVE,         /* 0.01,  scratch_itch.py:66 */
    AVG_TF_FORCE_CAP,            /* 10,    scratch_itch.py:66 */
    AVG_TF_HUMAN_KP_ACTIVE,      /* human_gains passed to take_step (0.05), scratch_itch.py:45 */
    AVG_TF_HUMAN_FORCE,          /* human_forces (1.0) */
    AVG_TF_CLOSEST_RANGE,        /* BedBathing: getClosestPoints query distance (4.0), bed_bathing.py:61 */
    /* 16..18: reference point of the device spatial algebra */
    AVG_TF_MOUTH = 19,           /* [3] mouth_pos in the head frame: (0, -0.11 | -0.10, 0.03), feeding.py:253 / drinking.py:252 */
    AVG_TF_EAT_RADIUS = 22,      /* 0.02 feeding.py:102, 0.03 drinking.py:114 */
    AVG_TF_EAT_REWARD = 23,      /* +20 feeding.py:104, +10 drinking.py:117 */
    AVG_TF_SPILL_REWARD = 24,    /* -5 feeding.py:113, -1 drinking.py:126 */
    AVG_TF_Z_MIN = 25,           /* 0.5 feeding.py:111 / drinking.py:124 */
    AVG_TF_FOOD_W = 26,          /* food_reward_weight / drinking_reward_weight (config.ini:25,34) */
    AVG_TF_TILT_W = 27,          /* cup_tilt_weight 0.1 (config.ini:33); 0 for Feeding */
    AVG_TF_TILT_SIGN = 28,       /* +1: -|roll + pi/2| (Jaco), -1: -|roll - pi/2| (other robots), drinking.py:72 */
    AVG_TF_CUP_RADIUS = 29,      /* 0.05, drinking.py:112 */
    AVG_TF_CUP_TOP = 30,         /* cup_top_center_offset z = -0.055, drinking.py:278 */
    AVG_TF_CUP_BOTTOM = 31       /* cup_bottom_center_offset z = 0.07, drinking.py:279 */
};
/* Feeding / Drinking reuse AVG_TF_SUCCESS_THR for total count * task_success_threshold (feeding.py:76). */
/* BedBathing reuses AVG_TF_SCRATCH_W for wiping_reward_weight (config.ini:17) and AVG_TF_SUCCESS_THR for
 * total_target_count * task_success_threshold (bed_bathing.py:72). */

/* ---- per-environment record: AVG_ENV_STRIDE floats (int fields stored as int32 in the same slots) ---- */
#define AVG_ENV_STRIDE 192
enum {
    AVG_E_Q        = 0,     /* [32] joint positions; free body k: pos(3) + quat(4) at body.qidx                */
    AVG_E_QD       = 32,    /* [32] generalized velocities (free body: v(3) world, w(3) world)                 */
    AVG_E_MTARGET  = 64,    /* [32] motor targets per dof                                                      */
    /* episode parameters */
    AVG_E_STRENGTH = 96, AVG_E_LIMIT_SCALE = 97, AVG_E_HUMAN_KP = 98, AVG_E_TREMOR_ON = 99,
    AVG_E_TREMOR   = 100,   /* [10] world_creation.human_tremors                                               */
    AVG_E_TARGET_H = 110,   /* [10] target_human_joint_positions (mutated by the tremor path, env.py:332)      */
    AVG_E_TARGET_ON_ARM = 120, /* [3] scratch_itch.py:281 */
    AVG_E_LIMB_FRAME = 123, /* int: AVG_F_SHOULDER or AVG_F_ELBOW                                              */
    AVG_E_EBODY    = 124,   /* [4][7] env-static body poses (pos, quat)  (first 4 env-static bodies)           */
    /* task state */
    AVG_E_ITERATION = 152,  /* int */
    AVG_E_TASK_SUCCESS = 153,
    AVG_E_PREV_CONTACT = 154, /* [3] prev_target_contact_pos */
    AVG_E_VALID_POSE = 157, /* [4] right_arm_previous_valid_pose */
    AVG_E_HAS_VALID = 161,  /* int */
    AVG_E_TARGET_POS = 162, /* [3] derived each sub-step, kept for inspection */
    AVG_E_EPISODE_RETURN = 165,
    AVG_E_OVERFLOW = 166,   /* int: bit0 contact overflow, bit1 row overflow (never silently dropped)          */
    AVG_E_SOLVER_ITERS = 167, /* int: PGS iterations executed in the last env-step (diagnostic) */
    AVG_E_NCAND = 168,       /* int: narrowphase candidate pairs examined in the last env-step (diagnostic) */
    AVG_E_TARGET_MASK = 170, /* [5] uint32: BedBathing targets not yet wiped, bit t of word t/32 (bed_bathing.py:111-125 shrink the lists) */
    AVG_E_FROZEN = 175,      /* uint32: bodies whose mass / inertia count as zero in this episode (changeDynamics(mass=0), world_creation.py:157-161) */
    AVG_E_WCACHE = 176,      /* [8][2] {uint32 key = shape a | shape b << 16 (0 = empty), float normal impulse}: the contact points of the
                                last internal step that the next one warm-starts from (Bullet keeps the applied impulse in the
                                persistent manifold point, btMultiBodyConstraintSolver::setupMultiBodyContactConstraint) */
    AVG_E_LAST = 192
};
#define AVG_WCACHE_N 8

/* ---- per-environment particle record (Feeding / Drinking): AVG_P_STRIDE floats, structure of arrays so that lane p reads
 *      particle p (and p + 32) with coalesced loads.  Spheres need no orientation.  Bit p of word p / 32 in the masks. ---- */
enum {
    AVG_P_POS = 0,           /* [3][64] centre, world */
    AVG_P_VEL = 192,         /* [3][64] linear velocity */
    AVG_P_ANG = 384,         /* [3][64] angular velocity */
    AVG_P_ALIVE = 576,       /* uint32[2]: still in self.foods / self.waters (feeding.py:120, drinking.py:135); removed particles leave the simulation */
    AVG_P_HIT = 578,         /* uint32[2]: foods_hit_person (feeding.py:116-119) */
    AVG_P_TOUCH_HUMAN = 580, /* uint32[2]: has a contact point with the human in the last internal step (getContactPoints(f, human)) */
    AVG_P_TOUCH_SPILL = 582, /* uint32[2]: ... with the table or the bowl (feeding.py:111) */
    AVG_P_EV_EAT = 584,      /* uint32[2]: events of the last env-step, for parity tests: reached the mouth */
    AVG_P_EV_SPILL = 586,    /*            spilled */
    AVG_P_EV_HIT = 588,      /*            hit the person */
    AVG_P_NCONTACT = 590,    /* int: particle contact points of the last internal step; [591] int: overflow flags */
    AVG_P_STRIDE = 592
};

/* ---- episode reset on the device (reference ScratchItchEnv.reset random draws) ----
 * Per model variant, what the sampler needs as plain arrays (written by compiler/reset.py: reset_table_bytes()). */
#define AVG_RESET_POOL 64
typedef struct AvgResetTable {
    int32_t n_pool, n_arm, n_fin, n_hum;
    int32_t tool_qidx, human_control;
    int32_t task;                         /* AVG_TASK_*                                                                            */
    int32_t n_target;                     /* BedBathing: bits set in AVG_E_TARGET_MASK at reset (bed_bathing.py:369-379)           */
    float   pool_q[AVG_RESET_POOL][8];    /* robot arm joint positions of an IK start pose (scratch_itch.py:251-253, util.py:34-57) */
    float   pool_tool[AVG_RESET_POOL][8]; /* tool base pose (pos, quat, pad) that goes with it (world_creation.py:331-337)          */
    int32_t arm_qidx[8], arm_dof[8];      /* position / velocity slots of the 7 arm joints                                          */
    int32_t fin_qidx[8], fin_dof[8];      /* gripper joints, opened to 1.0 (scratch_itch.py:254)                                    */
    int32_t hum_qidx[8], hum_dof[8], hum_joint[8];   /* dynamic human joints (reference joint index 7..13)                          */
    float   hum_lower[8], hum_upper[8], hum_reset[8];
    float   limb_dims[2][2];              /* (length, radius) of upper arm and forearm (scratch_itch.py:277-280)                     */
    float   fin_open;                     /* gripper open position: 1.0 ScratchItch (scratch_itch.py:254), 1.1 BedBathing (bed_bathing.py:327) */
    /* On-device IK for the start pose (scratch_itch.py:243-253 + util.py:34-105): when ik_enabled, the pool entry is only the
     * fallback; each episode draws its own start target in the +-ik_range box around ik_target and solves the arm for it with
     * damped least squares and up to 40 random restarts (same acceptance test as util.py:51: position and quaternion
     * distance < 0.03).  ik_ee_frame = COM frame of the end-effector link (getLinkState of link 8 / 76) in the frame of
     * dynamic body ik_ee_body. */
    int32_t ik_enabled, ik_ee_body;
    float   ik_range;
    float   ik_target[8];                 /* pos(3) of the box centre, quat(4) target orientation, pad */
    float   ik_ee_frame[8];               /* pos(3), quat(4), pad */
    /* ---- version 10: Feeding / Drinking (feeding.py:171-185,242-245,276-280,291-310) ---- */
    int32_t hum_slot[8];                  /* slot of each dynamic human joint in the controllable list: joint - 4 (ScratchItch / BedBathing), joint - 24 (head) */
    float   fin_q[8];                     /* open position per finger joint (Sawyer / Baxter: +position, -position, world_creation.py:313-320) */
    int32_t n_particle, has_bowl;
    uint32_t head_mask;                   /* bodies frozen unless the episode has a tremor or the id is human-active (AVG_E_FROZEN) */
    float   ik_tol;                       /* random_restart_threshold: 0.03 ScratchItch, 0.01 Feeding / Drinking on the Jaco (feeding.py:278) */
    float   bowl_center[4], bowl_quat[4]; /* feeding.py:184-185: centre of the +-0.05 square the bowl is drawn from, its orientation */
    float   grid[AVG_MAX_PARTICLE][4];    /* particle offsets from the tool's base position, world axes (feeding.py:301-305) */
    /* ---- `New` ids (<Task><Robot>New-v0, reference __init__.py:38-50,122-134,206-218,290-302) ----
     * new_mode: human_impairment = 'none' (scratch_itch.py:159); ScratchItch / BedBathing: every dynamic arm joint starts at
     * hum_reset + U(-hum_jitter, hum_jitter) (scratch_itch.py:216-217), redrawn while the arm is closer than new_min_dist to the
     * rest of the person, the robot or the furniture (:198,219-223; bounding-capsule distances on the device, see
     * avg_reset_new_kernel).  Waist angles and hipbone_to_mouth_height are properties of the model variant. */
    int32_t new_mode;
    float   hum_jitter;                   /* 10 degrees */
    float   new_min_dist;                 /* 0.01 */
    int32_t pad_new;
} AvgResetTable;

/* Counter-based random numbers of the device reset: draw k of episode `episode` of environment `env` under `seed`.
 * A 32-bit mix (murmur3 finaliser over a running hash); the numpy mirror in compiler/reset.py computes the same bits.
 * Draw indices: 0 gender, 1 impairment, 2 limit scale, 3 strength, 4..13 tremor, 14 limb, 15 point along the limb,
 * 16 angle around the limb, 17 start pose; on-device IK: 20..22 start target, 32 + 8 r + j rest pose of joint j in restart r.
 * Feeding / Drinking: 4..7 tremor of the head joints, 8..10 head angles (joints 25..27), 11..12 bowl offset. */
#define AVG_RNG_MIX(h) do { (h) ^= (h) >> 16; (h) *= 0x85ebca6bu; (h) ^= (h) >> 13; (h) *= 0xc2b2ae35u; (h) ^= (h) >> 16; } while (0)

/* ---- policy for on-device rollouts (reference enjoy_vr.py:77-117: actor_critic.act on VecNormalize'd observations) ----
 * The default a2c_ppo_acktr actor: observations normalised with the running statistics (clip +-10), two tanh layers of
 * 64 units, a linear mean layer; deterministic action = the mean.  Blob = this header, then float32 arrays
 * ob_mean[n_in], ob_var[n_in], W1[n_in][64], b1[64], W2[64][64], b2[64], W3[64][n_out], b3[n_out]. */
#define AVG_POLICY_MAGIC 0x4c505641u  /* "AVPL" */
#define AVG_POLICY_HIDDEN 64
typedef struct AvgPolicyHeader {
    uint32_t magic;
    int32_t  n_in;                /* leading observation columns the policy reads (obs_robot_len, enjoy_vr.py:92,117) */
    int32_t  n_out;               /* leading action columns it writes (action_robot_len); the rest are zero (enjoy_vr.py:112-113) */
    float    clip_obs;            /* 10.0 (VecNormalize clipob) */
    float    eps;                 /* 1e-8 (VecNormalize epsilon) */
    int32_t  pad[3];
} AvgPolicyHeader;

/* one reported contact point (parity / debug), 16 floats */
typedef struct AvgContact {
    int32_t shape_a, shape_b;     /* shape indices, A is the moving shape listed first in the pair table        */
    float   pos_a[3], pos_b[3];   /* world */
    float   normal[3];            /* from B to A */
    float   dist;
    float   force;                /* normal impulse / dt, as getContactPoints()[9] */
    int32_t pad[3];
} AvgContact;

#endif /* AVG_MODEL_H */

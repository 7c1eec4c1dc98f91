// fm_types.h — device-side model constants and per-env state layout of the step kernel.
// Replaces (as flat fp32 tables) what dm_control/MuJoCo hold in mjModel for the scene of
// challenge_env/challenge_env/scene.py:109-169 (reference) — see factory_marl_b200/model.py:gpu_tables().
#pragma once
#include <stdint.h>

#define FM_NCUBE 10
#define FM_NARMBODY 11      // base, link1..7, gripper_base, left_plate, right_plate
#define FM_NSPH 46          // collision spheres per arm (iiwa14.xml:58-136)
#define FM_NABOX 9          // gripper boxes per arm (gripper.xml:6-41)
#define FM_MAX_SBOX 12      // table + 2 x (bucket floor + 4 fences)
#define FM_MAXCON 56        // contacts per env
#define FM_MAXARMSIDE 40    // contact sides that touch an arm (a grasped cube alone makes 16+); stored in global scratch (each stores 3x9 Jacobian + 3x9 Minv J^T)
#define FM_MAX_ARMS 6
#define FM_MAXCON_ALL 96   // largest per-scene contact capacity (debug dump arrays)
// per-scene capacities: FM_MAXCON / FM_MAXARMSIDE are the two-arm values; four arms crash into each other and the table far more
// often (and have the shared memory to spare: one block of 7 envs per SM)
template <int NA> struct FmCap {
  static constexpr int CON = NA <= 2 ? FM_MAXCON : FM_MAXCON_ALL;         // contacts per env
  static constexpr int ARMSIDE = NA <= 2 ? FM_MAXARMSIDE : 64;            // contact sides on an arm (6-bit slot)
  static constexpr int QUEUE = NA <= 2 ? 96 : (NA <= 4 ? 160 : 224);                        // cluster pairs queued for the narrow phase
  static constexpr int NCL = FM_NCUBE + NA * FM_NARMBODY;                 // sphere-bounded broad-phase clusters
  static constexpr int WS = ARMSIDE * 54 + CON * 4 + 1 + 8 * FM_MAX_ARMS + 9 * FM_MAX_ARMS + 4 * NCL; // floats of the env's global scratch row: as_J, as_B, prev_f, ct_d0, arm_qlo, bnd_ref
};
#define FM_NCLASS 4
#define FM_MAX_FRAME_SKIP 128

enum { FM_CL_PLANE = 0, FM_CL_SBOX = 1, FM_CL_BELT = 2, FM_CL_CUBE = 3, FM_CL_ARM = 4 };
enum { FM_SIDE_STATIC = 0, FM_SIDE_BELT = 1, FM_SIDE_CUBE = 2, FM_SIDE_ARM = 3 };
enum { FM_ENV_BASE_IK = 0, FM_ENV_SINGLE_FULL = 1, FM_ENV_SINGLE_DELTA = 2, FM_ENV_ALL_FULL = 3, FM_ENV_ALL_DELTA = 4,
       FM_ENV_PAUSE_TOGGLE = 5, FM_ENV_BACKUP_TOGGLE = 6 };

struct FmDevModel {
  int n_arms, n_sbox, n_cluster, n_cluster_pair;
  float h, g;                      // timestep (scene.xml:2), |gravity|
  // static boxes (world frame): pos, rot (row-major), half size, contact class, geom id
  float sbox_pos[FM_MAX_SBOX][3], sbox_R[FM_MAX_SBOX][9], sbox_half[FM_MAX_SBOX][3];
  int sbox_geom[FM_MAX_SBOX], sbox_class[FM_MAX_SBOX], sbox_aligned[FM_MAX_SBOX];
  // top face height (pos.z + half.z) of the axis-aligned big boxes as float pairs (hi = nearest float, lo = remainder of the fp64
  // sum): 1.05f + 0.04f rounds to 1.0899999, 8.6e-8 below the belt's real top — a third of a resting cube's float grid
  float sbox_top[FM_MAX_SBOX][2], belt_top[2];
  // belt (conveyor_belt.xml:4-11)
  float belt_pos[3], belt_half[3], belt_mass, belt_damp, belt_kv, belt_invw;
  int belt_geom, belt_class;
  // cubes (scene.py:9-23,121-134)
  float cube_half[FM_NCUBE], cube_mass[FM_NCUBE], cube_inertia[FM_NCUBE], cube_invw[FM_NCUBE], cube_invm[FM_NCUBE], cube_invI[FM_NCUBE];
  int cube_geom[FM_NCUBE], cube_class;
  int floor_class;
  // pair contact parameters by class: mu, solref0, solref1, solimp[5]
  float pair_param[FM_NCLASS][FM_NCLASS][8];
  // arm template
  float arm_base_pos[FM_MAX_ARMS][3], arm_base_quat[FM_MAX_ARMS][4], arm_base_R[FM_MAX_ARMS][9];
  float ab_C[FM_NARMBODY][9];      // constant rotation of each template body relative to its parent (from ab_quat)
  float ab_pos[FM_NARMBODY][3], ab_quat[FM_NARMBODY][4], ab_mass[FM_NARMBODY], ab_ipos[FM_NARMBODY][3], ab_I[FM_NARMBODY][6];
  int ab_jtype[FM_NARMBODY];       // 0 none, 1 hinge (local z), 2 slide (local x)
  float ab_range[FM_NARMBODY][2], ab_invw[FM_NARMBODY], ab_dofinvw[FM_NARMBODY];
  float sph_pos[FM_NSPH][3], sph_r[FM_NSPH];
  int sph_body[FM_NSPH], sph_geom[FM_NSPH];       // geom id offset within the arm
  int body_sph_start[FM_NARMBODY + 1];
  float abox_pos[FM_NABOX][3], abox_half[FM_NABOX][3];
  int abox_body[FM_NABOX], abox_geom[FM_NABOX];
  int body_box_start[FM_NARMBODY + 1];
  int arm_first_geom[FM_MAX_ARMS], sph_class, abox_class;
  float arm_cl[FM_NARMBODY][4];    // bounding sphere per body (local centre, radius)
  // actuators (iiwa14.xml:10, gripper.xml:61-65)
  float kp, kd, g_kp, g_kb, g_kd, g_fmin, g_fmax, ten0, ten1;
  float ctrl_lo[9], ctrl_hi[9];    // actuator_ctrlrange[0:9] (belt, 7 joints, gripper)
  float eq_solref[2], eq_solimp[5], lim_solref[2], lim_solimp[5];
  float site_pos[3]; int site_body;
  float bucket_pos[2][3], bucket_size[2][3];
  // broad-phase clusters: kind + index (static box idx / cube idx / arm*11+body)
  int cl_kind[13 + FM_NCUBE + FM_MAX_ARMS * FM_NARMBODY], cl_idx[13 + FM_NCUBE + FM_MAX_ARMS * FM_NARMBODY];
  // broad phase: conservative world AABBs of the big geoms by cluster id (static boxes; belt with its full sliding range),
  // first sphere-bounded cluster id, and the packed pair records built on the host (see fm_host_model.h)
  float big_aabb[FM_MAX_SBOX + 2][6];
  int cl_sph0;
};
enum { FM_PAIR_SPH_SPH = 0, FM_PAIR_PLANE_SPH = 1, FM_PAIR_BOX_SPH = 2, FM_PAIR_ALWAYS = 3 };

// fp64 constants of the IK base policy kernel (ik_policy.py; dm_control qpos_from_site_pose)
struct FmIkModel {
  int n_arms;
  double base_pos[FM_MAX_ARMS][3], base_quat[FM_MAX_ARMS][4];   // world pose of the arm's base body
  double player_site[FM_MAX_ARMS][3];                           // ik_policy.py:44 `_base_site`
  double body_pos[9][3], body_quat[9][4];                       // template bodies 0..8 (base .. gripper_base)
  double site_pos[3];
  double cube_group[24][4];                                     // scipy Rotation.create_group("O"), wxyz
  double bucket_pos[2][3];
  double ctrl_lo[8], ctrl_hi[8];                                // actuator_ctrlrange[1:9]
};

// task / env configuration (BaseEnv kwargs base_env.py:26-34 + reward weights environments.py:252-259)
struct FmEnvCfg {
  int env_class, n_envs, frame_skip, pgs_outer, pgs_inner, sleep_hidden;
  float pgs_tol;
  double initial_conveyor_speed, conveyor_acceleration, pt_time, force_threshold, dt, init_spawn_freq, spawn_freq_increase;
  double w_gripper_cube, w_cube_bucket, w_action_norm, base_reward;
  // (1 - alpha)^(k+1), k = sub-step of the control step, alpha = h / (h + pt_time) (base_env.py:209-211): closed form of the PT1
  // filter under a constant command (see ph_ctrl)
  float pt1_pow[FM_MAX_FRAME_SKIP];
};

/* fm_api.h — C-ABI of the B200 batched TaskEnv library (libfm_b200.so).
 *
 * The reference (nkirschi/Factory-MARL) has no FFI; its seam for this path is the Python env contract consumed
 * by stable-baselines3.  Each entry point below names the reference interface it replaces.
 * All pointers except `blob`, `cfg` and the debug outputs are DEVICE pointers owned by the caller; the library
 * owns only the handle and the device copies of the model constants.  Calls are stream-ordered on `stream`
 * (a cudaStream_t passed as void*), made from one host thread per GPU.  Return value: 0 = OK, non-zero = error,
 * message via fm_last_error().  Nothing throws across the ABI.
 */
#ifndef FM_API_H
#define FM_API_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fm_handle fm_handle;

/* env_class: the 7 env classes of src/environments.py (FM_ENV_* order = inheritance listing environments.py:13-22) */
enum { FM_FACTORY_MANIPULATION_ENV = 0, FM_SINGLE_FULL_RL = 1, FM_SINGLE_DELTA = 2, FM_ALL_FULL_RL = 3, FM_ALL_DELTA = 4,
       FM_PAUSE_IK_TOGGLE = 5, FM_BACKUP_IK_TOGGLE = 6 };

/* BaseEnv.__init__ kwargs (challenge_env/challenge_env/base_env.py:15-35) + ProgressRewardEnv weights
 * (src/environments.py:252-259) + solver knobs of this implementation. */
typedef struct fm_config {
  int32_t n_envs, n_arms, env_class, auto_reset;
  /* pgs_inner: sweep cap of the PGS islands (0 = 4), pgs_tol: stopping criterion in m/s^2 (0 = 1e-3), pgs_outer: 1 = solve arm + contact
   * islands by Gauss-Seidel sweeps only (0 = the warp-cooperative active-set Newton, exact).
   * envs_per_block: block shape of the step kernel: 0 = by batch size, 1 | 4 = small batches / tests, 14 (two arms) | 7 (four
   * arms) | 6 (six arms; also stands in for 4 there) = one block per SM (default from 2048 envs).  reserved: 1 = keep hidden cubes live (strict mode), 0 = sleep them. */
  int32_t pgs_outer, pgs_inner, envs_per_block, reserved;
  float pgs_tol, reserved_f;
  double initial_conveyor_speed, conveyor_acceleration, pt_time, force_contact_threshold, control_frequency,
      spawn_freq, spawn_freq_increase;
  double gripper_to_closest_cube_reward_factor, closest_cube_to_bucket_reward_factor, small_action_norm_reward_factor,
      base_reward;
} fm_config;

/* one env's physics state in the reference's (MuJoCo) layout; host pointers (parity tests on identical qpos/qvel/ctrl) */
typedef struct fm_state_view {
  int32_t env;
  double* qpos;        /* [nq]  = 1 + 7*10 + 9*n_arms    */
  double* qvel;        /* [nv]  = 1 + 6*10 + 9*n_arms */
  double* ctrl_target; /* [nu]  = 1 + 8*n_arms          (BaseEnv.ctrl_target, base_env.py:187) */
} fm_state_view;

/* debug view of one env after the first sub-step's position stage (tests only) */
typedef struct fm_debug_view {
  int32_t env, ncon, flags, reserved;
  int32_t pair[64][2];          /* geom ids (MuJoCo numbering of the compiled scene), sorted */
  float pos[64][3], normal[64][3], force[64][4];
  float site_xpos[6][3];
  float mass_matrix[6][45];     /* packed lower triangle of each arm's 9x9 block */
  float bias[6][9];
  float qacc_smooth[115], qacc[115];
} fm_debug_view;

/* BaseEnv.__init__ + build_scene (base_env.py:15-147, scene.py:109-169): model_blob from factory_marl_b200.model.build_blob */
int fm_create(const fm_config* cfg, const void* model_blob, size_t n, fm_handle** out);
int fm_destroy(fm_handle* h);                                       /* BaseEnv.close (base_env.py:293-295) */
int fm_obs_dim(const fm_handle* h);                                 /* observation_space.shape (environments.py:41-46,553-558) */
int fm_act_dim(const fm_handle* h);                                 /* action_space (environments.py:399-403,474-479,549) */
/* rows of the caller-allocated SoA state arrays [rows][n_envs]: float32, float64, int32 (uint64 rng is always 4 rows) */
int fm_state_rows(const fm_handle* h, int32_t* nf, int32_t* nd, int32_t* ni);
int fm_bind_state(fm_handle* h, float* f, double* d, int32_t* i, uint64_t* rng);
/* field offsets inside the SoA arrays, by name (e.g. "F_ARM_Q", "I_SCORE", "I_IGNORE"); -1 if unknown */
int fm_field_offset(const fm_handle* h, const char* name);
/* FactoryManipulationEnv.reset / BaseEnv.reset_sim (environments.py:204-248, base_env.py:177-198); mask [N] or NULL = all */
/* Optional env order for the following launches: block slot k serves env order[k] (a permutation of 0..n_envs-1 on the device,
 * owned by the caller, read at every launch; NULL = identity).  Results do not depend on it; the launch time does: a block lasts as
 * long as its most expensive env and the launch as long as the slowest SM's blocks, so ordering the envs by expected cost (state
 * rows I_ARMCON, I_NIN, I_PREV_N), most expensive first, groups like envs and starts the long blocks first
 * (VecEnv mirror: FactoryVecEnv(sort_envs=True), + 9 % at 4096 envs). */
int fm_set_env_order(fm_handle* h, const int32_t* order);
int fm_reset(fm_handle* h, const uint8_t* mask, float* obs, void* stream);
/* FactoryManipulationEnv.step (environments.py:151-202) for all envs, with SB3-VecEnv auto-reset when cfg.auto_reset:
 * actions f32[N, act_dim] (Box) or i32[N, n_arms] (MultiDiscrete); flags bit0 out_of_reach, bit1 force_terminate,
 * bits 2..5 capacity overflow diagnostics, bit 6 = the IndexError case of task_utils.py:110 (skipped); ep_info [N,4] = episode return, length, score0, score1 (Monitor + ep_score_history) */
int fm_step(fm_handle* h, const void* actions, float* obs, float* reward, uint8_t* terminated, int32_t* scores,
            uint8_t* flags, float* terminal_obs, float* ep_info, void* stream);
/* terminal_obs: rows of the envs that terminated in this step hold the observation BEFORE the auto-reset (SB3's
 * infos[i]["terminal_observation"]); for the toggling classes its IK-proposal tail (last 8 * n_arms entries) is left at zero:
 * the reference computes it with one more IK solve on the terminal state, which nothing on the path reads (no time-limit
 * truncation, so SB3 never bootstraps from it). */
/* IK base policy for all arms: IKPolicy.act + FactoryManipulationEnv._compose_control (ik_policy.py:141-282,
 * environments.py:104-127).  Writes the proposals into the state (used by the next fm_step) and, for the toggling
 * envs, into obs[:, base_dim:] (IKTogglingEnv._process_observation, environments.py:560-577). */
int fm_ik_policy(fm_handle* h, float* obs, void* stream);
int fm_get_state(fm_handle* h, fm_state_view* v);                   /* physics.data.qpos/qvel, BaseEnv.ctrl_target */
int fm_set_state(fm_handle* h, const fm_state_view* v);
/* run n raw sub-steps (mj_step2 + mj_step1 each) with ctrl = ctrl_target, no filter, no task logic; dump env `view->env` */
int fm_debug_substeps(fm_handle* h, int n_substeps, fm_debug_view* view, void* stream);
const char* fm_last_error(void);

/* ---- on-GPU PPO (replaces stable-baselines3 at src/learning.py:106-118; hyper-parameters runs/r666unuv.json) ----
 * fm_ppo_net / fm_ppo_ws are plain structs of dimensions, element offsets and caller-owned DEVICE pointers; their layout is
 * defined in factory_marl_b200/csrc/fm_ppo.cu and mirrored by ctypes in factory_marl_b200/ppo.py. */
typedef struct fm_ppo_net fm_ppo_net;
typedef struct fm_ppo_ws fm_ppo_ws;
/* ActorCriticPolicy.forward for n observations (collect_rollouts): env actions (clipped Box / int MultiDiscrete), raw actions,
 * log-probs, values.  MLP layers run as bf16 tensor-core GEMMs with fp32 accumulation. */
int fm_policy_forward(const fm_ppo_net* net, const fm_ppo_ws* ws, const float* obs, int n, uint64_t seed, int deterministic,
                      void* act_env, float* act_raw, float* logp, float* value, void* stream);
/* RolloutBuffer.compute_returns_and_advantage: arrays [T, N]; gamma learning.py:47, gae_lambda runs/r666unuv.json:42 */
int fm_gae(const float* rewards, const float* values, const uint8_t* dones, const float* last_values, int T, int N, float gamma,
           float gae_lambda, float* advantages, float* returns, void* stream);
/* one minibatch of PPO.train: forward + clipped-surrogate/value loss + backward, gradients accumulated into net->grads */
int fm_ppo_update(const fm_ppo_net* net, const fm_ppo_ws* ws, const float* obs, const float* actions, const float* old_logp,
                  const float* adv, const float* ret, const int32_t* idx, int n, float clip_range, float vf_coef, float ent_coef,
                  int normalize_advantage, float* stats, void* stream);
/* clip_grad_norm_ + Adam on the flat parameter buffer (after the NCCL all-reduce of net->grads; grad_scale = 1/world).
 * stats: 8 floats on the device (required): [5] <- gradient norm, [6] scratch, [7] Adam's step count when step <= 0 (the
 * launch sequence then has no host-side argument that changes between steps, so it can be replayed from a CUDA graph). */
int fm_adam_step(const fm_ppo_net* net, float lr, float beta1, float beta2, float eps, float max_grad_norm, float grad_scale, int step,
                 float* stats, void* stream);
int fm_ppo_sync_bf16(const fm_ppo_net* net, void* stream);
/* Monitor / ep_score_history / AdditionalMetricsCallback state (learning.py:53-72,98; environments.py:53,200-201): once per env
 * step; terminated u8[N] and ep_info f32[N,4] are fm_step's outputs; hist i16[N,100,2], first i16[N,2], count i32[N],
 * ring f32[100,2] (return, length of the last 100 episodes over all envs), ring_head i32[1] = episodes so far. */
int fm_monitor_update(const uint8_t* terminated, const float* ep_info, int n, int16_t* hist, int16_t* first, int32_t* count,
                      float* ring, int32_t* ring_head, void* stream);
/* fused 3-layer MLP forward of one net on tcgen05 (TMEM accumulators); used by fm_policy_forward and fm_ppo_update unless switched
 * off.  w1 / w2 / w3: bf16 weights in the B-operand image, 16-byte chunks [K / 8][N][8 consecutive k] (fm_ppo_net.params_pk). */
int fm_mlp_forward_tc5(const float* x, int ld_x, int d_in, const int* idx, const void* w1, const void* w2, const void* w3,
                       const float* b1, const float* b2, const float* b3, float* out, int op, void* h1, void* h2, int n, int kp,
                       int* status, void* stream);
void fm_ppo_use_tcgen05(int on);   /* bit 0: rollout forward, bit 1: the update's forward (default 3 = both; 0 = mma.sync GEMM chain) */
const char* fm_ppo_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

/* cosim_b200.h -- C ABI of the B200 batched stepping engine (libcosim_b200.so).
 *
 * Drop-in boundary for the reference's per-env MuJoCo loop.  The reference has no FFI of its own
 * (it is pure Python on top of the `mujoco` wheel); the operator boundary it exposes is the BaseEnv
 * API of /root/reference/envs/wrappers.py:8-85 as assembled by build_env
 * (/root/reference/envs/build.py:8-24).  Each entry point below names the reference interface it
 * replaces.  All array arguments are DEVICE pointers unless the name ends in `_host`; rows are
 * per-env and contiguous ([N][dim], float32).  Functions return 0 or a negative error code and never
 * throw; cosim_last_error() gives the message.  A handle is not thread-safe; work is enqueued on
 * the caller's CUDA stream (cudaStream_t passed as void*).
 */
#ifndef COSIM_B200_H
#define COSIM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cosim_handle cosim_handle;

#define COSIM_OK 0
#define COSIM_ERR_ARG (-1)
#define COSIM_ERR_CUDA (-2)
#define COSIM_ERR_MODEL (-3)
#define COSIM_ERR_FIELD (-4)

/* build_env(config) -> one engine for num_envs instances (envs/build.py:8-24; <Robot>.__init__ incl.
 * XMLManager.get_model_path draws, e.g. envs/flamingo_p_v3/flamingo_p_v3.py:17-113, and
 * MjModel.from_xml_path / mj_setConst).  `blob` is the flat model of cosim_b200/model.py
 * (include/cosim_blob.h).  Global env ids are env_offset .. env_offset+num_envs-1 (RNG substreams). */
int cosim_create(const void* blob, size_t nbytes, int num_envs, int device, uint64_t seed,
                 uint32_t env_offset, cosim_handle** out);
void cosim_destroy(cosim_handle* h);                       /* env.close(), wrappers.py:416-417 */
const char* cosim_last_error(const cosim_handle* h);

/* CommandWrapper.reset() (wrappers.py:385-389 -> 303-307 -> 245-256 -> <Robot>.reset_model,
 * flamingo_p_v3.py:235-255).  mask: uint8[N] (NULL = all).  command: applied command [N][command_dim]
 * (NULL = zeros).  state_out: [N][state_dim]. */
int cosim_reset(cosim_handle* h, const uint8_t* mask, const float* command, float* state_out, void* stream);

/* CommandWrapper.step(action) (wrappers.py:391-405 -> 309-320 -> 258-269 -> <Robot>.step,
 * flamingo_p_v3.py:150-199: delay, PD, clip, frame_skip x mj_step, mj_rnePostConstraint, obs, done).
 * action [N][action_dim] in [-1,1]; command = applied (scaled) command; user_command = raw command
 * used for the reporter statistics (NULL = command).  terminated/truncated: uint8[N]. */
int cosim_step(cosim_handle* h, const float* action, const float* command, const float* user_command,
               float* state_out, uint8_t* terminated, uint8_t* truncated, void* stream);

/* Same call with HOST buffers (pageable or pinned): H2D of action/command, step, D2H of
 * state/terminated/truncated, then a stream sync.  This is the end-to-end path bench.py times. */
int cosim_step_host(cosim_handle* h, const float* action_host, const float* command_host,
                    float* state_out_host, uint8_t* terminated_host, uint8_t* truncated_host);

/* env.event("push", v) (flamingo_p_v3.py:257-266).  mask uint8[N] (NULL = all), vel [N][3] world frame. */
int cosim_push(cosim_handle* h, const uint8_t* mask, const float* vel_world, void* stream);

/* env.get_data() replacement (wrappers.py:410-411): copy a named per-env field into dst ([N][dim] f32,
 * device).  Fields: qpos qvel qacc_warmstart body_mass invweight_dof invweight_body frictionloss
 * geom_mu scal kp kd torque info last_action counters(int32) stats, and with debug enabled:
 * contacts heightmap hm_cell(int32) cfrc_ext sens qacc iters(int32).  cosim_field_dim gives dim. */
int cosim_field_dim(const cosim_handle* h, const char* field);
int cosim_get(cosim_handle* h, const char* field, void* dst, void* stream);
int cosim_set(cosim_handle* h, const char* field, const void* src, void* stream);   /* qpos qvel qacc_warmstart torque */
/* One raw physics sub-step (= one mujoco.mj_step of gymnasium's do_simulation loop, called from
 * flamingo_p_v3.py:189) from the stored qpos/qvel/qacc_warmstart with ctrl = the last applied torque; no env layer.
 * Parity aid: lets tests compare single sub-steps with the oracle. */
int cosim_substep(cosim_handle* h, void* stream);
int cosim_set_debug(cosim_handle* h, int enable);   /* allocate + fill the debug dumps */
int cosim_field_is_int(const cosim_handle* h, const char* field);   /* 1 = int32 rows, 0 = float32 rows */
/* Counter-based RNG probe (replaces the reference's unseeded random / numpy / scipy draws,
 * manager/control_manager.py:18, manager/xml_manager.py:50, utils/noise_generator_utils.py:13,25-27):
 * out[N][nidx] (uint32, device) = Philox4x32-10 draws of (global env id, rng_stream, step, idx). */
int cosim_rng_probe(cosim_handle* h, uint32_t rng_stream, uint32_t step, int nidx, uint32_t* out, void* stream);

/* Reporter statistics (core/reporter.py:210-218 accumulates info; SURVEY.md C-17 defines the
 * reductions): sums the per-env accumulators into out[COSIM_NSTAT] (device, float64). */
#define COSIM_NSTAT 16
int cosim_stats_reduce(cosim_handle* h, double* out, void* stream);
int cosim_stats_clear(cosim_handle* h, void* stream);

/* introspection */
int cosim_num_envs(const cosim_handle* h);
int cosim_dim(const cosim_handle* h, const char* name);      /* "state_dim", "action_dim", "command_dim", "nq", "nv", ... */
int cosim_launch_count(const cosim_handle* h);               /* kernels launched so far */
int cosim_smem_bytes_per_env(const cosim_handle* h);
int cosim_warps_per_block(const cosim_handle* h);
int cosim_pool_size(const cosim_handle* h);                  /* environments per CTA of the pooled step kernel; 0 = lock-step kernel */
int cosim_general_path(const cosim_handle* h);               /* 1 = the model runs on the general constraint path (condim 1/4/6, elliptic cone, PGS, impratio) */

/* ---- policy MLP (core/policy.py:11-21 MLPPolicy.get_action: clip(MLP(state), -1, 1)) ---- */
typedef struct cosim_policy cosim_policy;
/* layer sizes dims[0..nlayers] (dims[0] = state_dim, dims[nlayers] = action_dim); weights are
 * row-major [out][in] float32 on the HOST, biases [out]; hidden activation: 0 = ELU, 1 = tanh, 2 = ReLU */
int cosim_policy_create(int device, int nlayers, const int* dims, const float* const* weights_host,
                        const float* const* biases_host, int activation, cosim_policy** out);
void cosim_policy_destroy(cosim_policy* p);
int cosim_policy_forward(cosim_policy* p, const float* state, int num_envs, float* action_out, void* stream);
int cosim_policy_launch_count(const cosim_policy* p);
/* `activation | COSIM_POLICY_RAW_OUTPUT` in cosim_policy_create: the last layer writes its plain linear output (no clip) and
 * may be up to 2048 wide -- used for the gate pre-activations of LSTMPolicy (core/policy.py:24-47). */
#define COSIM_POLICY_RAW_OUTPUT 0x100
#define COSIM_POLICY_ACTIVATED_OUTPUT 0x200   /* the output gets the hidden activation instead of the clip (encoders in front of an LSTM) */
/* ONNX LSTM cell update, gate order i, o, f, c: gates [N][4H] -> c [N][H], h [N][H] updated in place (device pointers). */
int cosim_lstm_cell(const float* gates, float* c, float* h, int num_envs, int hidden, void* stream);
/* profiling builds only (-DCOSIM_PHASE_TIMING): out_host[144] = 16 per-phase cycle / event counters + two 64-bin histograms
   (8 K-cycle bins) of the per-warp collision and Newton phase times; zeros in the product build (tools/phase_profile.py) */
int cosim_phase_cycles(cosim_handle* h, unsigned long long* out_host, int reset);

#ifdef __cplusplus
}
#endif
#endif

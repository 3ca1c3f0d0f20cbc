/*
 * prl_b200 - C ABI of the B200-native data-parallel PPO hot path.
 *
 * Drop-in boundary for Raven4567/Parallel-Reinforcement-Learning.  The reference has no FFI of its own:
 * its boundary is the Python API (PPO/, AsyncTools/), so these are the entry points a ctypes binding in
 * the reference's Python would call; each one names the reference code it replaces (file:line under the
 * reference repo).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (cudaMalloc'd / torch CUDA tensor .data_ptr()) unless its name ends
 *     in `_host`; sizes are element counts; `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = ok, negative = error (PRL_ERR_*); prl_last_error() has the message (thread-local).
 *   - no global state, no allocation inside the library; scratch memory is passed in as `ws`
 *     (prl_scan_ws_bytes / prl_update_ws_floats tell how much).  Calls are asynchronous on `stream`.
 *   - layouts: env state is SoA fp64 [S][E]; the rollout buffer is time-major [T][C][E] float32;
 *     PPO.memory is env-major flat: states [N][O], actions [N] or [N][A], rewards [N], dones [N] float32;
 *     network parameters are ONE flat float32 buffer in torch `.parameters()` order (prl_policy_param_count).
 */
#ifndef PRL_B200_H
#define PRL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRL_VERSION 100

enum { PRL_OK = 0, PRL_ERR_INVALID = -1, PRL_ERR_CUDA = -2, PRL_ERR_CAPACITY = -3 };
enum { PRL_ENV_CARTPOLE = 0, PRL_ENV_PENDULUM = 1, PRL_ENV_ACROBOT = 2, PRL_ENV_MOUNTAINCAR = 3, PRL_ENV_MOUNTAINCARCONT = 4 };
enum { PRL_ACT_I32 = 0, PRL_ACT_I64 = 1, PRL_ACT_F32 = 2 };

const char *prl_last_error(void);
int prl_version(void);

/* gym.make(id) facts the reference reads from env.observation_space / action_space (train.py:12-13). */
int prl_env_info(int env_id, int *state_dim, int *obs_dim, int *action_dim, int *is_continuous, int *max_episode_steps);
/* number of float32 in ActorCritic(is_continuous, O, A).parameters()  (PPO/ActorCritic.py:14-64) */
int64_t prl_policy_param_count(int is_continuous, int obs_dim, int action_dim);
/* number of float32 in ONE of RND's two nets (PPO/RND.py:25-31) */
int64_t prl_rnd_param_count(int in_features, int out_features);
size_t prl_scan_ws_bytes(int64_t n);

/* ---------------------------------------------------------------- EnvVectorizer (AsyncTools/AsyncPPO.py:35-102) */
/* reset(): AsyncPPO.py:48-62.  Draws every env's start state from Philox(seed, episode) in the env's reset box,
 * zeroes the TimeLimit counters and the terminal mask, writes obs [E][O] float32. */
int prl_env_reset(int env_id, int E, uint64_t seed, uint64_t episode, double *state, int32_t *elapsed,
                  uint8_t *terminal, float *obs, void *stream);
/* reset() from numpy's seeded stream - what E gymnasium envs seeded with env.reset(seed=seeds[e]) draw (the per-env
 * call at AsyncPPO.py:53): rng is [4][E] uint64 {state_hi, state_lo, inc_hi, inc_lo} of one PCG64 per env.  seeds != NULL:
 * generator e is first created as np.random.PCG64(np.random.SeedSequence(seeds[e])); seeds == NULL: the stored
 * generators continue (a later reset() of the same envs).  States = Generator.uniform(low, high), bit for bit. */
int prl_env_reset_numpy(int env_id, int E, const uint64_t *seeds, uint64_t *rng, double *state, int32_t *elapsed,
                        uint8_t *terminal, float *obs, void *stream);
/* reset() with injected start states (teacher forcing): state_aos is [E][S] fp64. */
int prl_env_set_state(int env_id, int E, const double *state_aos, double *state, int32_t *elapsed,
                      uint8_t *terminal, float *obs, void *stream);
int prl_env_get_state(int env_id, int E, const double *state, double *state_aos, void *stream);
/* step(actions): AsyncPPO.py:64-102.  Steps the n non-terminal envs active_idx[0..n) (ascending env index) with
 * the compact actions[rank]; returns compact next obs [n][O] f32, rewards [n] f64, dones/truncs [n] u8.
 * TimeLimit: elapsed += 1; trunc = elapsed >= max_episode_steps. */
int prl_env_step(int env_id, int E, int n, const int32_t *active_idx, const void *actions, int action_dtype,
                 double *state, int32_t *elapsed, int max_episode_steps, float *obs, double *rewards,
                 uint8_t *dones, uint8_t *truncs, void *stream);

/* ---------------------------------------------------------------- AsyncTools/utils.py */
/* indexes_of_active_environments (utils.py:3-4) + number_of_active_environments (:6-7): idx = arange(E)[~terminal],
 * *count = len(idx).  Generic form: select where flags[i] == want. */
int prl_compact_indices(const uint8_t *flags, int64_t n, int want, int32_t *idx, int32_t *count, void *ws,
                        size_t ws_bytes, void *stream);
/* inactive_states_dropout (utils.py:14-15): out = rows[idx] for a compacted idx; rows are `width` float32 wide. */
int prl_gather_rows(const float *rows, const int32_t *idx, const int32_t *count, int64_t max_rows, int width,
                    float *out, void *stream);
/* update_active_environments_list (utils.py:38-43): terminal[active_idx[i]] = dones[i]. */
int prl_mask_update(uint8_t *terminal, const int32_t *active_idx, const uint8_t *dones, int n, void *stream);
/* buffer_append (utils.py:17-36) into the device VecMemory (AsyncPPO.py:11-33): for rank i, env e = active_idx[i]:
 * slot t = lengths[e]; buf_*[t][.][e] = float32(item[i]); lengths[e] = t + 1.  Returns PRL_ERR_CAPACITY via
 * *overflow != 0 (device flag) when t >= T_cap. */
int prl_buffer_append(int E, int T_cap, int n, const int32_t *active_idx, const float *states, int obs_dim,
                      const float *actions, int act_width, const float *rewards, const float *dones,
                      float *buf_states, float *buf_actions, float *buf_rewards, float *buf_dones,
                      int32_t *lengths, int32_t *overflow, void *stream);
/* buffer_to_target_buffer_transfer (utils.py:45-50): env-major, time-minor concatenation of the per-env
 * episodes behind the `base` items already in PPO.memory; zeroes lengths (buffer.clear()).  *total (device int64)
 * = base + sum(lengths). */
int prl_buffer_transfer(int E, int T_cap, int obs_dim, int act_width, const float *buf_states,
                        const float *buf_actions, const float *buf_rewards, const float *buf_dones,
                        int32_t *lengths, int64_t base, int64_t capacity, float *mem_states, float *mem_actions,
                        float *mem_rewards, float *mem_dones, int64_t *total, void *ws, size_t ws_bytes,
                        void *stream);

/* The same with up to 4 extra per-transition planes [T][E] -> flat [N] rows (host arrays of device pointers): what the fused
 * worker produced besides the reference's four fields - the acting policy's log-prob and state value of every transition
 * (prl_rollout_eval) and the GAE returns computed on the time-major buffer (prl_gae_columns) - travels with them. */
int prl_buffer_transfer_ex(int E, int T_cap, int obs_dim, int act_width, const float *buf_states,
                           const float *buf_actions, const float *buf_rewards, const float *buf_dones, int n_extra,
                           const float *const *extra_planes, float *const *extra_rows, int32_t *lengths, int64_t base,
                           int64_t capacity, float *mem_states, float *mem_actions, float *mem_rewards,
                           float *mem_dones, int64_t *total, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------- PPO.get_action (PPO/PPO.py:82-96) */
/* states [n][O] f32 -> sampled actions: discrete int64 [n] (Categorical(probs).sample()), continuous f32 [n][A]
 * (tanh(mu + std*eps) * action_scaling).  Randomness: Philox(seed; row id, call_index) where row id = row_ids[i] (the
 * env index of compact row i; NULL -> i), so a step-by-step loop with call_index = (episode << 32) | t draws the
 * same numbers as the fused prl_rollout.  Optional output (may be NULL): dist [n][A] probs or [n][2A] (mu, std). */
int prl_policy_act(const float *params, int is_continuous, int obs_dim, int action_dim, float action_scaling,
                   const float *states, const int32_t *row_ids, int64_t n, uint64_t seed, uint64_t call_index,
                   void *actions, float *dist, void *stream);
/* ActorCritic.get_evaluate (PPO/ActorCritic.py:118-146) forward: logp [n], value [n], entropy_sum (device double,
 * ACCUMULATED: caller zeroes it; mean = sum / n). */
int prl_policy_evaluate(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                        const float *actions, int64_t n, float *logp, float *value, double *entropy_sum,
                        void *stream);

/* ---------------------------------------------------------------- AsyncPPO.worker (AsyncPPO.py:117-146), fused */
/* One episode per env in ONE launch: obs -> policy forward -> sample -> physics -> TimeLimit -> buffer write, looped
 * until done|trunc, at most T_cap (= max_episode_steps) steps.  tape != NULL replaces the sampler with taped actions
 * tape[t][e] (int32) / tape[t][e][A] (f32) - teacher forcing.  Outputs: time-major buffer, lengths, final state,
 * terminal mask (all 1), scores[0] = sum of rewards (f64), scores[1] = number of env steps (as f64). */
int prl_rollout(int env_id, int E, int T_cap, const float *params, float action_scaling, uint64_t seed,
                uint64_t episode, const void *tape, double *state, int32_t *elapsed, uint8_t *terminal,
                float *buf_states, float *buf_actions, float *buf_rewards, float *buf_dones, int32_t *lengths,
                double *scores, void *stream);

/* prl_rollout + the old-policy evaluation of PPO.learn (PPO/PPO.py:134-154) taken where the network outputs already exist:
 * buf_logp[t][e] = log-prob of the stored action, buf_values[t][e] = V(s_t), both under `params` - bit-identical to
 * prl_policy_evaluate on the same rows (same forward, same epilogue arithmetic).  Both NULL = prl_rollout.  Discrete
 * policies and continuous ones with action_dim == 1.
 * auto_reset_horizon > 0 (opt-in; 0 = the reference's worker, where a finished env drops out - AsyncPPO.py:118,143-146):
 * an env whose episode ends - terminated, or auto_reset_horizon (= the TimeLimit) steps into the episode - is reset in
 * place (its k-th reset draws what prl_env_reset draws in episode `episode | k << 40`) and keeps stepping, so every env
 * fills all T_cap slots; the last slot closes the running episode with done = 1.  lengths[e] = T_cap.
 * score_ws != NULL (prl_rollout_score_ws_doubles(E) doubles, zeroed once): the reward sum is accumulated in a fixed order
 * (per-CTA partials, added in CTA order by the last CTA to finish) and is bit-reproducible; NULL: double atomicAdd. */
int prl_rollout_eval(int env_id, int E, int T_cap, const float *params, float action_scaling, uint64_t seed,
                     uint64_t episode, const void *tape, double *state, int32_t *elapsed, uint8_t *terminal,
                     float *buf_states, float *buf_actions, float *buf_rewards, float *buf_dones, float *buf_logp,
                     float *buf_values, int32_t *lengths, double *scores, int auto_reset_horizon, double *score_ws,
                     void *stream);
/* doubles of score_ws for E envs (prl_rollout_eval; zeroed once by the caller) */
size_t prl_rollout_score_ws_doubles(int E);

/* ---------------------------------------------------------------- PPO.compute_gae (PPO/PPO.py:107-120) */
/* Flat reverse scan over the env-major buffer, float32, same operation order as the reference;
 * next_value = values[N-1] as in PPO.py:188 when next_value_ptr == NULL. */
int prl_gae(const float *rewards, const float *dones, const float *values, const float *next_value_ptr,
            double gamma, double gae_lambda, int64_t N, float *returns, void *ws, size_t ws_bytes, void *stream);
size_t prl_gae_ws_bytes(int64_t N);
/* Column form on the time-major rollout buffer [T][E] (each env's episode scanned backwards from lengths[e]-1). */
int prl_gae_columns(const float *rewards, const float *dones, const float *values, const int32_t *lengths, int E,
                    int T_cap, double gamma, double gae_lambda, float *returns, void *stream);
/* PPO.py:198-199: adv = returns - values; (adv - mean) / (std_unbiased + 1e-8).  stats (device double[4]) receives
 * {sum, sum of squared deviations, count, unused}; pass stats_in to normalise with externally reduced statistics
 * (multi-GPU: allreduce of stats between the two phases). phase 1 = statistics, 2 = normalise, 3 = both. */
int prl_adv_normalize(const float *returns, const float *values, int64_t N, float *adv, double *stats, int phase,
                      void *stream);

/* ---------------------------------------------------------------- PPO.learn minibatch step (PPO/PPO.py:219-252) */
size_t prl_update_ws_floats(int is_continuous, int obs_dim, int action_dim, int64_t batch);
/* forward + clipped-surrogate/SmoothL1/entropy loss + backward for ONE minibatch of b rows; writes the flat
 * gradient of loss.mean() (same layout as params) and loss_out = {sum of per-row policy terms, sum of SmoothL1
 * terms, sum of entropies, b} as device doubles.  inv_count = 1 / (global minibatch rows) so that sharded
 * minibatches sum to the global mean gradient. */
int prl_ppo_grad(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                 const float *actions, const float *old_logp, const float *adv, const float *returns, int64_t b,
                 float policy_clip, float inv_count, float *grad, double *loss_out, float *ws, size_t ws_floats,
                 void *stream);
/* Tensor-core form of prl_ppo_grad (csrc/update_tc.cu): same contract and gradient layout; tcgen05.mma (bf16x3 split
 * operands, fp32 accumulation in tensor memory) for every contraction over features or rows.  Policies with
 * observ_dim <= 16 and action_dim <= 8; prl_ppo_grad_tc_supported returns 1 for discrete policies (every form below), 2 for
 * continuous ones (prl_ppo_grad_tc only: a float32 forward pre-pass evaluates the tanh-Gaussian loss and its gradient with
 * respect to the mu / log_std head outputs, then the two-head kernel runs twice - {mu, critic} and {log_std, critic weighted 0} -
 * and the two reduced gradients are added; ActorCritic.py:28-42), 0 otherwise.  The workspace must be zeroed once before its first
 * use: ws[0] is a sticky status word that prl_ppo_grad_tc_status (host-synchronising) reads - 1 if a tensor-core phase
 * of any call never completed. */
int prl_ppo_grad_tc_supported(int is_continuous, int obs_dim, int action_dim);
size_t prl_update_tc_ws_floats(int is_continuous, int obs_dim, int action_dim, int64_t batch);
int prl_ppo_grad_tc(const float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                    const float *actions, const float *old_logp, const float *adv, const float *returns, int64_t b,
                    float policy_clip, float inv_count, float *grad, double *loss_out, float *ws, size_t ws_floats,
                    void *stream);
int prl_ppo_grad_tc_status(const float *ws, int *status_host, void *stream);
/* prl_ppo_grad_tc + clip_grad_norm_ + AdamW in ONE cooperative launch (single-GPU path), with no grid barrier: every CTA
 * writes its partial-gradient row and loss sums and adds one to an arrival counter (release); every CTA waits for the
 * counter to reach the grid size, reduces its 64-parameter slices over the CTAs' rows in a fixed order (bit-identical to
 * the separate reduction of prl_ppo_grad_tc) and publishes the squared norm of its slices as a tagged 64-bit word
 * {payload, launch number}; the leader CTA collects those and publishes the clip coefficient the same way; then every
 * CTA applies clip_grad_norm_ + AdamW to its slices of `params` in place.  `grad` receives the reduced gradient,
 * step_counter is the 24-byte optimiser clock of prl_adamw_step_dev.  Every wait is bounded in wall-clock time
 * (PRL_TC_TIMEOUT_MS, default generous): status word 2 = a wait on another CTA timed out. */
int prl_ppo_step_tc(float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                    const float *actions, const float *old_logp, const float *adv, const float *returns, int64_t b,
                    float policy_clip, float inv_count, float *grad, double *loss_out, float *exp_avg,
                    float *exp_avg_sq, int64_t *step_counter, float lr, float weight_decay, float max_norm,
                    double *grad_norm_out, float *ws, size_t ws_floats, void *stream);
/* Sharded form of prl_ppo_step_tc: the gradient allreduce happens INSIDE the kernel over NVLink peer memory.  Every rank
 * owns an exchange buffer (prl_p2p_alloc of prl_p2p_exchange_bytes; shared with the other ranks of the node through CUDA
 * IPC handles: prl_p2p_get_handle / prl_p2p_open_handle) = inbox[2 (step parity)][world (sender)][P] of 8-byte words
 * {float bits, step number}; peer_bufs is a DEVICE array of `world` pointers to the ranks' buffers (own one at [rank]).
 * After its slice reduction a CTA pushes every value of its slices, tagged with the optimiser step number, into all
 * ranks' inboxes (one naturally aligned 64-bit store per value and rank: the tag cannot arrive without its value - no
 * fence, no flag), polls the `world` words of each of its parameters in its own inbox until all carry this step's number
 * and sums them in rank order (bit-identical on every rank); then squared norm, clip and AdamW as on one GPU.  b may be
 * 0 (a rank without rows in this minibatch still takes part).  Status word 3 = a peer never signalled (bounded wait). */
int prl_ppo_step_tc_p2p(float *params, int is_continuous, int obs_dim, int action_dim, const float *states,
                        const float *actions, const float *old_logp, const float *adv, const float *returns, int64_t b,
                        float policy_clip, float inv_count, float *grad, double *loss_out, float *exp_avg,
                        float *exp_avg_sq, int64_t *step_counter, float lr, float weight_decay, float max_norm,
                        double *grad_norm_out, void *const *peer_bufs, int rank, int world, float *ws, size_t ws_floats,
                        void *stream);
size_t prl_p2p_exchange_bytes(int is_continuous, int obs_dim, int action_dim, int world);
int prl_p2p_alloc(size_t bytes, void **ptr);
int prl_p2p_free(void *ptr);
int prl_p2p_get_handle(void *ptr, unsigned char *handle64);
int prl_p2p_open_handle(const unsigned char *handle64, void **ptr);
int prl_p2p_close_handle(void *ptr);
/* nn.utils.clip_grad_norm_(params, max_norm) + AdamW.step (PPO.py:250-252; torch defaults betas (0.9,0.999),
 * eps 1e-8, weight_decay 0.01).  step = 1-based optimiser step count. max_norm <= 0 disables clipping. */
int prl_adamw_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, int64_t step,
                   float lr, float weight_decay, float max_norm, double *grad_norm_out, void *stream);
/* same with the optimiser clock in device memory: step_counter points to 24 bytes {int64 step, double beta1^step, double
 * beta2^step} (zero-initialised = step 0); the kernel advances and stores them, so the identical launch can be replayed
 * (CUDA graphs over the k_epochs x minibatch loop). */
int prl_adamw_step_dev(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                       int64_t *step_counter, float lr, float weight_decay, float max_norm, double *grad_norm_out,
                       void *stream);

/* ---------------------------------------------------------------- RND (PPO/RND.py:71-115) */
/* compute_intrinsic_reward: out[i] = beta * || pred(s_i) - target(s_i) ||_2 ; add_to != NULL: out = add_to + that */
int prl_rnd_intrinsic(const float *target_params, const float *pred_params, int in_features, int out_features,
                      const float *states, int64_t n, float beta, const float *add_to, float *out, void *stream);
/* update_pred for one chunk: gradient of MSELoss(mean) wrt pred_params (flat), loss_out[0] += sum of squared err */
int prl_rnd_grad(const float *target_params, const float *pred_params, int in_features, int out_features,
                 const float *states, int64_t n, float *grad, double *loss_out, float *ws, size_t ws_floats,
                 void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PRL_B200_H */

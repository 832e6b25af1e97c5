// env.cu -- device-resident batched AHRAGGymEnv: B environments advance with ONE launch per step.
//
// Reference (sequential Python, one graph copy per environment): src/ah_rag/agent/gym_env.py:41-125 over
// src/ah_rag/agent/environment.py:155-241 (reset/semantic_anchor), :340-355 (commit_selection), :301-335
// (query_node_details), :393-480 (expand_*), src/ah_rag/agent/featurizer.py:7-56, src/ah_rag/agent/reward.py:6-26.
//
// Per environment the device keeps what the reference keeps in Python sets/dicts, reduced to what the gym step
// consumes: selection_set and frontier_set as unique id lists, and the CURRENT observation: the nodes shown in
// obs["selection"] (ids + rounded score/semantic), len(obs["seeds"]), obs["step"], and the sizes of
// obs["state"].selection_ids / frontier_ids[:50] (or "no state", as query_node_details returns, :325-333).
// One warp per environment; the expansion is the same ordered walk as ahr_expand (graph_shared.cuh).
#include "common.cuh"
#include "graph_shared.cuh"
#include "kernels.h"

namespace ahr {

constexpr int ENV_SHOWN = 10;   // featurizer looks at the first 10 selection entries (featurizer.py:21)
constexpr int ENV_LIMIT = 10;   // expand_* default limit used by the gym step (gym_env.py:94-98)
constexpr int ENV_FEATS = 84;   // 4 globals + 10 x 8
constexpr int ENV_WARPS = 4;

struct EnvView {
  int32_t n_env, cap_sel, cap_fr;
  int32_t *sel_ids, *sel_cnt;      // selection_set
  int32_t *fr_ids, *fr_cnt;        // frontier_set
  int32_t *shown_ids;              // [n_env, 10] obs["selection"] node ids
  double *shown_score, *shown_sem; // [n_env, 10]
  int32_t *shown_cnt;              // len(obs["selection"])
  int32_t *n_seeds;                // len(obs["seeds"])
  int32_t *obs_step;               // obs["step"]
  int32_t *obs_sel, *obs_fr;       // sizes in obs["state"]; -1 = the observation carries no "state"
  int32_t *env_steps;              // GraphEnvironment.step_count
  int32_t *gym_steps;              // AHRAGGymEnv._cur_step
  int32_t *last_action;            // -1 = none
  uint8_t *done;
  int32_t *overflow;               // [1] set when a selection / frontier list outgrew its capacity
};

__device__ __forceinline__ bool list_contains(const int32_t *list, int cnt, int32_t v, int lane) {
  bool hit = false;
  for (int i = lane; i < cnt; i += 32) hit |= (list[i] == v);
  return __any_sync(0xffffffffu, hit);
}

// featurize_observation (featurizer.py:21-56): float64 feature values cast to float32
__device__ __forceinline__ void env_featurize(const EnvView &e, const GraphView &g, int env, int lane, float *feat,
                                              uint8_t *mask) {
  float *f = feat + (int64_t)env * ENV_FEATS;
  const int shown = min(e.shown_cnt[env], ENV_SHOWN);
  if (lane == 0) {
    const bool has_state = e.obs_sel[env] >= 0;
    f[0] = (float)e.obs_step[env];
    f[1] = has_state ? (float)e.obs_sel[env] : 0.f;
    f[2] = has_state ? (float)e.obs_fr[env] : 0.f;
    f[3] = (float)e.n_seeds[env];
  }
  for (int j = lane; j < ENV_SHOWN; j += 32) {
    float *o = f + 4 + j * 8;
    if (j < shown) {
      const int32_t v = e.shown_ids[env * ENV_SHOWN + j];
      const NodeHead rec = g.line[v].head;  // one sector per shown node
      const uint8_t t = rec.type;
      const uint8_t lv = rec.level;
      const double judge = rec.judge, conf = rec.conf;
      o[0] = t == AHR_NT_ENTITY ? 1.f : 0.f;
      o[1] = t == AHR_NT_SUMMARY ? 1.f : 0.f;
      o[2] = (t != AHR_NT_ENTITY && t != AHR_NT_SUMMARY) ? 1.f : 0.f;
      // layer = 0 if entity else (level or (1 if summary else 0))          environment.py:98
      o[3] = t == AHR_NT_ENTITY ? 0.f : (lv != 0 ? (float)lv : (t == AHR_NT_SUMMARY ? 1.f : 0.f));
      o[4] = __double2float_rn(e.shown_score[env * ENV_SHOWN + j]);
      o[5] = __double2float_rn(e.shown_sem[env * ENV_SHOWN + j]);
      o[6] = judge == judge ? __double2float_rn(judge) : 0.f;  // float(n.get("judge_overall") or 0.0)
      o[7] = conf == conf ? __double2float_rn(conf) : 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) o[c] = 0.f;
    }
  }
  // get_action_mask (gym_env.py:54-75): without a top node only end_episode is valid
  if (lane < 6) mask[env * 6 + lane] = (e.shown_cnt[env] > 0 || lane == 5) ? 1 : 0;
}

// reset: state from the anchor result (environment.py:155-241 via gym_env.py:41-52, top_k = 5)
__global__ void __launch_bounds__(ENV_WARPS * 32)
env_reset_kernel(EnvView e, GraphView g, const int32_t *__restrict__ nodes, const double *__restrict__ sem,
                 const double *__restrict__ score, const int32_t *__restrict__ count, int32_t top_k, int32_t n_seeds,
                 float *__restrict__ feat, uint8_t *__restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * ENV_WARPS + (threadIdx.x >> 5);
  if (env >= e.n_env) return;
  const int cnt = count[env];
  int32_t *fr = e.fr_ids + (int64_t)env * e.cap_fr;
  // frontier_set = set(reranked ids) (:225): reranked ids are unique by construction
  for (int j = lane; j < cnt; j += 32) fr[j] = nodes[(int64_t)env * top_k + j];
  for (int j = lane; j < ENV_SHOWN; j += 32) {
    const bool in = j < cnt;
    e.shown_ids[env * ENV_SHOWN + j] = in ? nodes[(int64_t)env * top_k + j] : -1;
    e.shown_score[env * ENV_SHOWN + j] = in ? score[(int64_t)env * top_k + j] : 0.0;
    e.shown_sem[env * ENV_SHOWN + j] = in ? sem[(int64_t)env * top_k + j] : 0.0;
  }
  __syncwarp();
  if (lane == 0) {
    e.sel_cnt[env] = 0;
    e.fr_cnt[env] = cnt;
    e.shown_cnt[env] = cnt;
    e.n_seeds[env] = n_seeds;
    e.obs_step[env] = 1;  // reset() zeroes step_count, semantic_anchor() increments it (:161, :197)
    e.env_steps[env] = 1;
    e.gym_steps[env] = 0;
    e.obs_sel[env] = 0;
    e.obs_fr[env] = min(cnt, 50);
    e.last_action[env] = -1;
    e.done[env] = 0;
  }
  __syncwarp();
  env_featurize(e, g, env, lane, feat, mask);
}

// one gym step for every environment (gym_env.py:85-125)
__global__ void __launch_bounds__(ENV_WARPS * 32)
env_step_kernel(EnvView e, GraphView g, const int32_t *__restrict__ actions, int32_t max_steps, double repeat_penalty,
                float *__restrict__ feat, double *__restrict__ rewards, uint8_t *__restrict__ done_out,
                uint8_t *__restrict__ mask) {
  __shared__ int32_t sm_out[ENV_WARPS][ENV_LIMIT];
  __shared__ int32_t sm_in[ENV_WARPS][4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * ENV_WARPS + w;
  if (env >= e.n_env) return;
  const int a = actions[env];
  int32_t *sel = e.sel_ids + (int64_t)env * e.cap_sel;
  int32_t *fr = e.fr_ids + (int64_t)env * e.cap_fr;
  const int shown = e.shown_cnt[env];
  // previous observation's state as step_reward sees it (no "state" key -> empty sets)
  const int prev_sel = max(e.obs_sel[env], 0), prev_fr = max(e.obs_fr[env], 0);
  bool done = false;
  bool obs_changed = true;
  int gym_steps = e.gym_steps[env] + 1;

  if (a >= 0 && a <= 2) {
    // expand_parents / expand_children on the top 2, expand_related on the top 1 (gym_env.py:93-98)
    const int take = min(a == 2 ? 1 : 2, min(shown, ENV_SHOWN));
    if (lane < take) sm_in[w][lane] = e.shown_ids[env * ENV_SHOWN + lane];
    __syncwarp();
    ExpandState st;
    st.out = sm_out[w];
    st.count = 0; st.raw = 0; st.limit = ENV_LIMIT;
    expand_walk(st, g, a, sm_in[w], take, lane);
    __syncwarp();
    const int n_new = st.count;
    // the observation is built BEFORE the frontier is updated (environment.py:409-411)
    const int sel_n = e.sel_cnt[env], fr_n = e.fr_cnt[env];
    for (int j = lane; j < ENV_SHOWN; j += 32) {
      e.shown_ids[env * ENV_SHOWN + j] = j < n_new ? sm_out[w][j] : -1;
      e.shown_score[env * ENV_SHOWN + j] = 0.0;
      e.shown_sem[env * ENV_SHOWN + j] = 0.0;
    }
    // frontier_set.update(expanded)
    int fr_cnt = fr_n;
    for (int j = 0; j < n_new; ++j) {
      const int32_t v = sm_out[w][j];
      if (!list_contains(fr, fr_cnt, v, lane)) {
        if (fr_cnt < e.cap_fr) {
          if (lane == 0) fr[fr_cnt] = v;
          ++fr_cnt;
        } else if (lane == 0) {
          atomicExch(e.overflow, 1);
        }
        __syncwarp();
      }
    }
    if (lane == 0) {
      e.shown_cnt[env] = n_new;
      e.n_seeds[env] = n_new;
      e.env_steps[env] += 1;
      e.obs_step[env] = e.env_steps[env];
      e.obs_sel[env] = sel_n;
      e.obs_fr[env] = min(fr_n, 50);
      e.fr_cnt[env] = fr_cnt;
    }
  } else if (a == 3) {
    // commit_selection(top 3) (environment.py:340-355): add, drop from the frontier, show what was added
    const int take = min(3, min(shown, ENV_SHOWN));
    if (lane < take) sm_in[w][lane] = e.shown_ids[env * ENV_SHOWN + lane];
    __syncwarp();
    int sel_cnt = e.sel_cnt[env], fr_cnt = e.fr_cnt[env], n_added = 0;
    for (int j = 0; j < take; ++j) {
      const int32_t v = sm_in[w][j];
      if (v < 0 || v >= g.n_nodes) continue;
      if (list_contains(sel, sel_cnt, v, lane)) continue;
      if (sel_cnt < e.cap_sel) {
        if (lane == 0) sel[sel_cnt] = v;
        ++sel_cnt;
      } else if (lane == 0) {
        atomicExch(e.overflow, 1);
      }
      if (lane == 0) sm_out[w][n_added] = v;
      ++n_added;
      // frontier_set.discard(v): swap-remove (the set is unordered; only membership and size matter)
      int pos = -1;
      for (int i = lane; i < fr_cnt; i += 32)
        if (fr[i] == v) pos = i;
      pos = __reduce_max_sync(0xffffffffu, pos);
      if (pos >= 0) {
        if (lane == 0) fr[pos] = fr[fr_cnt - 1];
        --fr_cnt;
      }
      __syncwarp();
    }
    for (int j = lane; j < ENV_SHOWN; j += 32) {
      e.shown_ids[env * ENV_SHOWN + j] = j < n_added ? sm_out[w][j] : -1;
      e.shown_score[env * ENV_SHOWN + j] = 0.0;
      e.shown_sem[env * ENV_SHOWN + j] = 0.0;
    }
    if (lane == 0) {
      e.shown_cnt[env] = n_added;
      e.n_seeds[env] = 0;
      e.env_steps[env] += 1;
      e.obs_step[env] = e.env_steps[env];
      e.sel_cnt[env] = sel_cnt;
      e.fr_cnt[env] = fr_cnt;
      e.obs_sel[env] = sel_cnt;
      e.obs_fr[env] = min(fr_cnt, 50);
    }
  } else if (a == 4) {
    // query_node_details(top 1): a single detailed node, no seeds and NO "state" in the observation (:325-333)
    if (shown > 0) {
      const int32_t v = e.shown_ids[env * ENV_SHOWN + 0];
      __syncwarp();
      for (int j = lane; j < ENV_SHOWN; j += 32) {
        e.shown_ids[env * ENV_SHOWN + j] = j == 0 ? v : -1;
        e.shown_score[env * ENV_SHOWN + j] = 0.0;  // the brief has no score / semantic keys -> `or 0.0`
        e.shown_sem[env * ENV_SHOWN + j] = 0.0;
      }
      if (lane == 0) {
        e.shown_cnt[env] = 1;
        e.n_seeds[env] = 0;
        e.env_steps[env] += 1;
        e.obs_step[env] = e.env_steps[env];
        e.obs_sel[env] = -1;
        e.obs_fr[env] = -1;
      }
    } else {
      obs_changed = false;  // obs = prev_raw
    }
  } else {
    done = true;  // end_episode: obs = prev_raw
    obs_changed = false;
  }
  __syncwarp();

  if (lane == 0) {
    double r = 0.0;
    if (!done) {
      // step_reward(prev, cur) (reward.py:6-26) on set sizes: the selection only grows within an episode
      const int cur_sel = obs_changed ? max(e.obs_sel[env], 0) : prev_sel;
      const int cur_fr = obs_changed ? max(e.obs_fr[env], 0) : prev_fr;
      const int add_sel = max(cur_sel - prev_sel, 0);
      const int add_fr = max(0, cur_fr - prev_fr);
      // Python evaluates (1.0*a + 0.05*b) - 0.05 with one rounding per operation: no FMA contraction
      r = __dadd_rn(__dadd_rn((double)add_sel, __dmul_rn(0.05, (double)min(add_fr, 10))), -0.05);
      if (e.last_action[env] >= 0 && a == e.last_action[env] && repeat_penalty > 0) r = __dadd_rn(r, -repeat_penalty);
    }
    done = done || gym_steps >= max_steps;
    rewards[env] = r;
    done_out[env] = done ? 1 : 0;
    e.done[env] = done ? 1 : 0;
    e.gym_steps[env] = gym_steps;
    e.last_action[env] = a;
  }
  __syncwarp();
  env_featurize(e, g, env, lane, feat, mask);
}

static EnvView make_view(const ahr_env_batch &b) {
  EnvView e;
  e.n_env = b.n_env; e.cap_sel = b.cap_sel; e.cap_fr = b.cap_fr;
  e.sel_ids = b.sel_ids; e.sel_cnt = b.sel_cnt; e.fr_ids = b.fr_ids; e.fr_cnt = b.fr_cnt;
  e.shown_ids = b.shown_ids; e.shown_score = b.shown_score; e.shown_sem = b.shown_sem; e.shown_cnt = b.shown_cnt;
  e.n_seeds = b.n_seeds; e.obs_step = b.obs_step; e.obs_sel = b.obs_sel; e.obs_fr = b.obs_fr;
  e.env_steps = b.env_steps; e.gym_steps = b.gym_steps; e.last_action = b.last_action; e.done = b.done;
  e.overflow = b.overflow;
  return e;
}

int env_reset(const GraphView &g, const ahr_env_batch &b, const int32_t *nodes, const double *sem, const double *score,
              const int32_t *count, int32_t top_k, int32_t n_seeds, float *feat, uint8_t *mask, cudaStream_t stream) {
  AHR_REQUIRE(b.n_env >= 0 && b.cap_fr >= top_k && b.cap_sel >= 3, "env_reset: list capacities too small");
  if (b.n_env == 0) return AHR_OK;
  env_reset_kernel<<<(b.n_env + ENV_WARPS - 1) / ENV_WARPS, ENV_WARPS * 32, 0, stream>>>(make_view(b), g, nodes, sem, score,
                                                                                      count, top_k, n_seeds, feat, mask);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int env_step(const GraphView &g, const ahr_env_batch &b, const int32_t *actions, int32_t max_steps, double repeat_penalty,
             float *feat, double *rewards, uint8_t *done, uint8_t *mask, cudaStream_t stream) {
  if (b.n_env == 0) return AHR_OK;
  env_step_kernel<<<(b.n_env + ENV_WARPS - 1) / ENV_WARPS, ENV_WARPS * 32, 0, stream>>>(make_view(b), g, actions, max_steps,
                                                                                     repeat_penalty, feat, rewards, done,
                                                                                     mask);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

}  // namespace ahr

/* mgb200 -- B200-native batched MiniGrid simulator: the C-ABI drop-in boundary.
 *
 * One shared library (libmgb200.so, built for sm_100a only) replaces the hot path
 * of rohitrango/gym-minigrid:
 *     MiniGridEnv.step          gym_minigrid/minigrid.py:1227-1325
 *     gen_obs / gen_obs_grid    gym_minigrid/minigrid.py:1327-1381
 *     Grid.slice/rotate_left/process_vis/encode/decode
 *                               gym_minigrid/minigrid.py:439-473,571-648
 *     reset + _gen_grid         gym_minigrid/minigrid.py:831-858 and
 *                               envs/{empty,doorkey,fourrooms,dynamicobstacles,keycorridor}.py, roomgrid.py
 * for a *batch* of independent environments resident in HBM.
 *
 * Conventions
 *   - every entry point returns 0 on success, <0 on error; mgb_last_error() gives the
 *     thread-local message.  Nothing throws across the boundary.
 *   - all data pointers are DEVICE pointers unless the name ends in _host; the caller
 *     owns them and keeps them alive until `stream` reaches the call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls
 *     enqueue and return immediately unless stated otherwise.
 *   - there is no CPU fallback: without a CUDA device mgb_create fails.
 *   - a handle is bound to one device and is not thread-safe; its calls all read and write the
 *     handle's env state, so issue them one at a time, in stream order (different handles are
 *     independent).  Nothing about a launch is kept on the host: mgb_reset / mgb_step /
 *     mgb_rollout may be captured in a CUDA graph and replayed.
 *   - device memory per env: the state block (80-1500 bytes) and, for the
 *     KeyCorridor / crossing / MultiRoom generators, a spare-layout block of about twice that.
 */
#ifndef MGB200_H
#define MGB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_OBS_BYTES 147      /* 7*7*3 uint8, layout [vx][vy][c] = Grid.encode (minigrid.py:571-594); 3*V*V for view size V */
#define MGB_MAX_OBSTACLES 8

/* layout generators, one per reference env file on the path */
enum {
    MGB_GEN_EMPTY = 0,        /* envs/empty.py:30-57 */
    MGB_GEN_DOORKEY = 1,      /* envs/doorkey.py:15-44 */
    MGB_GEN_FOURROOMS = 2,    /* envs/fourrooms.py:19-69 */
    MGB_GEN_DYNOBS = 3,       /* envs/dynamicobstacles.py:35-89 */
    MGB_GEN_KEYCORRIDOR = 4,  /* roomgrid.py:118-359 + envs/keycorridor.py:26-59 */
    MGB_GEN_POOL = 5,         /* no on-device generator: reset draws one of the uploaded levels (mgb_set_level_pool).
                                 For envs whose step() is the base MiniGridEnv.step or one of the MGB_HOOK_* rules
                                 -- SURVEY §8(f) rank 2 */
    MGB_GEN_CROSSING = 6,     /* envs/crossing.py:24-99: gen_param0 = num_crossings, gen_param1 = ori (0 h, 1 v, 2 both) | 4 if the
                                 obstacles are walls (SimpleCrossing) instead of lava */
    MGB_GEN_LAVAGAP = 7,      /* envs/lavagap.py:21-60: gen_param0 = const (gap column fixed at width/2), gen_param1 = 1 if walls */
    MGB_GEN_MULTIROOM = 8,    /* envs/multiroom.py:41-241: gen_param0 = number of rooms (min == max in every registered id),
                                 gen_param1 = maxRoomSize */
    MGB_GEN_DISTSHIFT = 9     /* envs/distshift.py:30-52: a fixed layout (no draws): goal at (width-2, 1), lava strips in rows 1 and
                                 gen_param0 = strip2_row, agent at (1,1) facing right.  Runs on the Empty kernels. */
};

/* static per-env-id configuration: what the reference bakes into constructor kwargs
 * (SURVEY.md Appendix B).  Replaces MiniGridEnv.__init__ kwargs, minigrid.py:767-778. */
typedef struct {
    int32_t gen;            /* MGB_GEN_* */
    int32_t width, height;  /* grid size */
    int32_t max_steps;
    int32_t see_through;    /* see_through_walls: skip process_vis (minigrid.py:1344-1347) */
    int32_t n_actions;      /* action_space.n: 7, or 3 for Dynamic-Obstacles (dynamicobstacles.py:32) */
    int32_t n_obstacles;    /* Dynamic-Obstacles */
    int32_t room_size;      /* RoomGrid (KeyCorridor) */
    int32_t num_rows;       /* RoomGrid (KeyCorridor); num_cols is 3 */
    int32_t random_start;   /* agent_start_pos=None variants (Empty-Random-*, Dynamic-Obstacles-Random-*) */
    int32_t lava_v1;        /* 'v1' in class name => lava gives reward -1, not done (minigrid.py:1263-1266) */
    int32_t agent_view_size;/* minigrid.py:776,795 / ViewSizeWrapper (wrappers.py:579-608): 0 or 7 = default; 3, 5, 9, 11
                               also built.  Every obs buffer is [..][V][V][3], i.e. 3*V*V bytes per env-step */
    int32_t hook;           /* MGB_HOOK_*: subclass step() post-hook for MGB_GEN_POOL handles (0 = base step only) */
    int32_t gen_param0;     /* generator parameters of MGB_GEN_CROSSING / LAVAGAP / MULTIROOM (see the enum), else 0 */
    int32_t gen_param1;
} mgb_config;

/* step() post-hooks of the stock env files that only add a success/failure rule on top of MiniGridEnv.step.
 * Per-level parameters come with the level pool (mgb_set_level_pool, `hook_params`). */
enum {
    MGB_HOOK_NONE = 0,
    MGB_HOOK_PICKUP_TARGET = 1, /* unlockpickup.py:34-42, blockedunlockpickup.py:38-46 (= keycorridor.py:51-59): picking up `obj` */
    MGB_HOOK_UNLOCK = 2,        /* unlock.py:33-41: toggle while the door (pos A) is open */
    MGB_HOOK_FETCH = 3,         /* fetch.py:74-86: carrying anything ends the episode; reward iff it is the target */
    MGB_HOOK_GOTODOOR = 4,      /* gotodoor.py:72-93: `done` next to the target door; done next to any door (A..D) */
    MGB_HOOK_GOTOOBJECT = 5,    /* gotoobject.py:68-84: toggle ends; `done` within 1 cell of the target */
    MGB_HOOK_PUTNEAR = 6,       /* putnear.py:91-112: wrong pickup ends; dropping the move object ends, reward iff near the target */
    MGB_HOOK_REDBLUEDOORS = 7,  /* redbluedoors.py:44-66: red door = pos A, blue door = pos B */
    MGB_HOOK_MEMORY = 8         /* memory.py:88-100: pickup acts as toggle; success pos A, failure pos B */
};
#define MGB_HOOK_PARAMS 16      /* int32 per level: target_type, target_color, move_type, move_color, target_x, target_y,
                                   A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y, 0, 0 */

typedef struct mgb_handle mgb_handle;

const char *mgb_version(void);
const char *mgb_last_error(void);

/* Replaces gym.make(id) x num_envs (register.py:5-21 -> MiniGridEnv.__init__).
 * env_id_base: global id of env 0 of this handle; the Philox stream of an env is keyed by
 * (seed, global env id, episode), so results do not depend on how envs are sharded over GPUs.
 * The envs are NOT reset: call mgb_reset. */
int mgb_create(const mgb_config *cfg, int64_t num_envs, int device, uint64_t seed,
               int64_t env_id_base, mgb_handle **out);
int mgb_destroy(mgb_handle *h);
int64_t mgb_num_envs(const mgb_handle *h);

/* auto-reset on done (default 1).  With 0 a done env simply keeps stepping, exactly like a
 * reference env whose caller ignores `done`. */
int mgb_set_autoreset(mgb_handle *h, int on);

/* Replaces env.seed(s) (minigrid.py:860-863): re-keys the Philox streams and rewinds the
 * episode counters; takes effect at the next reset. */
int mgb_seed(mgb_handle *h, uint64_t seed);

/* Replaces MiniGridEnv.reset (minigrid.py:831-858) for envs with mask[i]!=0 (all when NULL).
 * obs [N][147] (may be NULL), dir [N] (may be NULL). */
int mgb_reset(mgb_handle *h, const uint8_t *mask, uint8_t *obs, uint8_t *dir, void *stream);

/* Replaces MiniGridEnv.step + subclass hooks + gen_obs (minigrid.py:1227-1381) for all N envs.
 * actions [N]; obs [N][147]; reward [N] fp64 (0.0, -1.0 or 1-0.9*(step_count/max_steps) with
 * the reference's three roundings); done [N]; dir [N].  With auto-reset, obs/dir of a done env
 * are those of the first observation of its next episode. */
int mgb_step(mgb_handle *h, const uint8_t *actions, uint8_t *obs, double *reward,
             uint8_t *done, uint8_t *dir, void *stream);

/* T consecutive steps in ONE persistent kernel; env state stays in shared memory between steps.
 * actions [T][N]; obs [T][N][147]; reward/done/dir [T][N].  Any output may be NULL (not written).
 * Alignment is a matter of speed only, never of results: obs blocks leave with bulk copies when obs is 16-byte aligned
 * (else byte stores), and action rows are fetched 16 bytes at a time when actions and N are multiples of 16 (else bytes). */
int mgb_rollout(mgb_handle *h, int32_t T, const uint8_t *actions, uint8_t *obs, double *reward,
                uint8_t *done, uint8_t *dir, void *stream);

/* T steps under the uniform random policy of the reference's own drivers (`env.action_space.sample()`, run_tests.py:43,
 * benchmark.py:27-33), drawn on the device from a counter-based stream -- no action input from the host: the action of
 * step t of the e-th call of this function on this handle, for global env id g, is
 *   mulhi32(Philox4x32-10(counter (t>>2, e, g lo, g hi), key (seed lo, seed hi ^ 0x41435431))[t&3], n_actions).
 * actions_out [T][N] (or NULL) receives the actions taken; everything else as mgb_rollout. */
int mgb_rollout_random(mgb_handle *h, int32_t T, uint8_t *actions_out, uint8_t *obs, double *reward, uint8_t *done,
                       uint8_t *dir, void *stream);

/* One step with HOST buffers (pinned memory recommended): H2D of actions, the step kernel and
 * D2H of obs/reward/done/dir are pipelined over internal streams in env chunks.  Synchronous:
 * returns when the outputs are in host memory.  This is the end-to-end path a CPU-side caller
 * of env.step() sees. */
int mgb_step_host(mgb_handle *h, const uint8_t *actions_host, uint8_t *obs_host,
                  double *reward_host, uint8_t *done_host, uint8_t *dir_host);

/* State exchange in the reference's own encoding (Grid.encode, minigrid.py:571-594):
 *   grid      [count][W][H][3]   (type, colour, state)
 *   aux       [count][W][H]      bit0 = Goal.overlap (minigrid.py:160) -- hidden attribute
 *   agent     [count][4] int32   x, y, dir, step_count
 *   carrying  [count][3]         (0,0,0) = nothing
 *   obstacles [count][8][2] int16 obstacle list in reference order (dynamicobstacles.py:53-56)
 *   target    [count][2]         (type, colour) of KeyCorridor.obj (keycorridor.py:48)
 *   rng       [count][2] uint32  (#resets so far, draws consumed in the current episode)
 * Any pointer may be NULL (left unchanged / not written).  This is also the checkpoint format. */
int mgb_set_state(mgb_handle *h, int64_t first, int64_t count, const uint8_t *grid,
                  const uint8_t *aux, const int32_t *agent, const uint8_t *carrying,
                  const int16_t *obstacles, const uint8_t *target, const uint32_t *rng, void *stream);
int mgb_get_state(mgb_handle *h, int64_t first, int64_t count, uint8_t *grid, uint8_t *aux,
                  int32_t *agent, uint8_t *carrying, int16_t *obstacles, uint8_t *target,
                  uint32_t *rng, void *stream);

/* Level pool for MGB_GEN_POOL handles: n_levels layouts in the reference's encoding (e.g. snapshots of
 * reference envs after reset()).  grid [K][W][H][3], aux [K][W][H] (may be NULL), agent [K][3] int32 = x,y,dir.
 * Device pointers; the pool is copied, the buffers may be freed once `stream` has passed the call.
 * reset / auto-reset of env e in episode k uses level  mulhi32(philox(seed, e, k).word0, K).
 * agent [K][3]; hook_params: per-level parameters of mgb_config.hook (positions in grid coordinates). */
int mgb_set_level_pool(mgb_handle *h, int32_t n_levels, const uint8_t *grid, const uint8_t *aux,
                       const int32_t *agent, const int32_t *hook_params /* [K][MGB_HOOK_PARAMS] or NULL */, void *stream);
/* which level each env is playing: levels int32 [N] (device).  set: after mgb_set_state, to restore a checkpoint. */
int mgb_get_levels(mgb_handle *h, int32_t *levels, void *stream);
int mgb_set_levels(mgb_handle *h, const int32_t *levels, void *stream);

/* RNG-tape parity mode: env i consumes draws[offsets[i] ...) in order instead of Philox
 * (values are final randint results).  NULL switches back to Philox.  Device pointers,
 * offsets has N+1 entries. */
int mgb_set_rng_tape(mgb_handle *h, const int32_t *draws, const int64_t *offsets);

/* FullyObsWrapper.observation (wrappers.py:311-338): full grid encode with the agent cell set
 * to (10, 0, dir).  out [N][W][H][3]. */
int mgb_full_obs(mgb_handle *h, uint8_t *out, void *stream);

/* ---- observation wrappers: stateless batched kernels on the current device ---- */

/* OneHotPartialObsWrapper.observation (wrappers.py:203-243) / FullyObsOneHotWrapper.observation
 * (wrappers.py:340-415): cells [n_cells][3] (type, colour, state) -> out [n_cells][n_classes+n_colors+n_states]
 * with out[type'] = out[n_classes+colour] = out[n_classes+n_colors+state] = 1.  class_map[11] maps type ->
 * type' (NULL = identity; FullyObsOneHotWrapper's keep_classes); n_colors may be 0 (drop_color). */
int mgb_onehot(const uint8_t *cells, uint8_t *out, int64_t n_cells, const uint8_t *class_map,
               int32_t n_classes, int32_t n_colors, int32_t n_states, void *stream);

/* FlatObsWrapper.observation (wrappers.py:528-577): out[n] = concat(float32(img[n]), mission_table[mission_idx[n]])
 * img [N][img_bytes]; mission_table [M][mission_len] float32 (27 x maxStrLen one-hot, built on the host);
 * mission_idx [N] (NULL = row 0 for every env); out [N][img_bytes + mission_len] float32. */
int mgb_flat_obs(const uint8_t *img, int32_t img_bytes, const float *mission_table, int32_t mission_len,
                 const uint8_t *mission_idx, float *out, int64_t N, void *stream);

/* ---- RGB observation wrappers: tile-atlas gathers, pixel-exact with the reference rasteriser ----
 * atlas uint8 [231][10][tile][tile][3] (device): tile for cell encoding code = type*21 + colour*3 + state and variant
 * 0 plain, 1 highlighted, 2..5 agent facing dir 0..3, 6 agent facing up + highlight, 7..9 agent facing dir 0..2 + highlight; built from the reference's
 * Grid.render_tile by oracle/gen_atlas.py and shipped as gym_minigrid_b200/data/tile_atlas_t8.npz.  tile % 8 == 0. */

/* RGBImgPartialObsWrapper.observation = MiniGridEnv.get_obs_render (wrappers.py:283-309, minigrid.py:1383-1398):
 * obs [N][V][V][3] -> out [N][V*tile][V*tile][3]; the agent is drawn at (V/2, V-1) facing up, every cell whose
 * type != unseen is highlighted. */
int mgb_render_partial(const uint8_t *obs, int32_t view, const uint8_t *atlas, int32_t tile, uint8_t *out,
                       int64_t N, void *stream);
/* MiniGridEnv.render('rgb_array', highlight=...) (minigrid.py:1400-1466) and RGBImgObsWrapper.observation, which is
 * render(highlight=False) (wrappers.py:245-281): out [N][H*tile][W*tile][3] from the handle's current state.
 * obs == NULL: no highlight.  obs = the current partial observation [N][V][V][3] (as written by mgb_reset/mgb_step):
 * the cells the agent sees are highlighted. */
int mgb_render_full(mgb_handle *h, const uint8_t *obs, const uint8_t *atlas, int32_t tile, uint8_t *out, void *stream);

/* ---- bookkeeping wrappers: element-wise kernels on the outputs of a step ---- */

/* ActionBonus.step (wrappers.py:87-119; by_action = 1, key (agent_pos, agent_dir, action), table = W*H*4*n_actions)
 * and StateBonus.step (wrappers.py:121-154; by_action = 0, key agent_pos, table = W*H):
 * counts uint32 [N][table] (caller-owned, zero-initialised, never cleared by reset -- the reference keeps its dict
 * across episodes); c = ++counts[n][key(state of env n)]; reward[n] += 1 / sqrt(c) in fp64.  Reads the handle's
 * CURRENT state, so call it before the env is reset (the facade switches auto-reset off and resets after it). */
int mgb_visit_bonus(mgb_handle *h, int32_t by_action, const uint8_t *actions, uint32_t *counts, int64_t table,
                    double *reward, void *stream);

/* DACWrapper.step (wrappers.py:56-77): an env whose episode ended keeps returning the blank observation
 * (image*0+1, reset-time direction), reward 0, and done only once count >= max_steps.
 * done_in [N] = done of the wrapped step; envdone_in/envdone_out [N] = the wrapper's env_done flag before/after
 * (two different buffers); obs [N][obs_bytes], reward [N], dir [N] are modified in place; done_out [N] != done_in. */
int mgb_dac(mgb_handle *h, int32_t count_ge_max, const uint8_t *done_in, const uint8_t *envdone_in, uint8_t *envdone_out,
            const uint8_t *reset_dir, uint8_t *obs, double *reward, uint8_t *done_out, uint8_t *dir, void *stream);

/* AppendActionWrapper (wrappers.py:418-458): hist uint8 [N][K] = the last K action indices, 255 = none.
 * actions == NULL (reset) clears every history; otherwise hist is shifted and actions[n] appended, and envs with
 * done[n] != 0 (auto-reset) are cleared.  Then out [N][D + A*K] = obs[n] (D bytes) ++ K one-hot vectors of A bytes. */
int mgb_append_action(int64_t N, int32_t D, int32_t A, int32_t K, const uint8_t *obs, const uint8_t *actions,
                      const uint8_t *done, uint8_t *hist, uint8_t *out, void *stream);

/* Episode statistics (not in the reference, which only prints): run_ret / run_len [N] are the caller-owned
 * running return and length (zero-initialised); after a step, out_ret / out_len [N] hold the totals of the episodes that
 * ended at this step (0 elsewhere) and the running values of those envs restart.  totals (or NULL): uint64 [2] device
 * counters, += number of finished episodes and += their summed length. */
int mgb_episode_stats(int64_t N, const double *reward, const uint8_t *done, double *run_ret, int32_t *run_len,
                      double *out_ret, int32_t *out_len, uint64_t *totals, void *stream);

/* GoalPolicyWrapper._get_goals (wrappers.py:476-497): obs [n_cells][planes] one-hot rows of FullyObsOneHotWrapper ->
 * achieved (goal plane cleared) and desired (agent cell -> empty, goal cell -> agent). */
int mgb_goal_policy(int64_t n_cells, int32_t planes, int32_t agent_idx, int32_t empty_idx, int32_t goal_idx,
                    const uint8_t *obs, uint8_t *achieved, uint8_t *desired, void *stream);

/* Synchronises `stream` and returns the sticky device error flags (then clears them):
 *   1 unknown action (reference: assert False, minigrid.py:1316-1318)   2 RNG tape exhausted
 *   4 tape value outside [low,high)    8 rejection sampling gave up (RecursionError in reset)
 *  16 agent/cell index out of bounds   32 unsupported cell code in set_state   64 reset without a level pool */
int mgb_error_flags(mgb_handle *h, void *stream, uint32_t *flags_host);

/* number of kernels this handle has launched so far */
int64_t mgb_kernel_launches(const mgb_handle *h);

/* duration in ms of the most recent mgb_step/mgb_rollout kernel, measured with CUDA events on
 * the launching stream (blocks until that kernel has finished); <0 if timing is disabled. */
int mgb_set_kernel_timing(mgb_handle *h, int on);
double mgb_last_kernel_ms(mgb_handle *h);

#ifdef __cplusplus
}
#endif
#endif

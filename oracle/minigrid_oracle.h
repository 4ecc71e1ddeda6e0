/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference hot path
 * of rohitrango/gym-minigrid: MiniGridEnv.step / gen_obs / gen_obs_grid /
 * Grid.slice / rotate_left / process_vis / encode and reset/_gen_grid of the
 * five BASELINE configs (+ their registered size variants).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product path
 * (gym_minigrid_b200 -> libmgb200.so) never does.
 *
 * Parity pin: tests/golden/ holds traces produced by the *live* reference
 * (oracle/gen_golden.py, run in the build container where /root/reference is
 * mounted); tests/test_oracle_golden.py replays them through this library and
 * requires bit-exact obs / direction / reward(fp64 bits) / done / full grid.
 */
#ifndef MINIGRID_ORACLE_H
#define MINIGRID_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* generator kinds (one per reference env file on the path) */
enum {
    ORC_GEN_EMPTY = 0,       /* envs/empty.py:30-57          */
    ORC_GEN_DOORKEY = 1,     /* envs/doorkey.py:15-44        */
    ORC_GEN_FOURROOMS = 2,   /* envs/fourrooms.py:19-69      */
    ORC_GEN_DYNOBS = 3,      /* envs/dynamicobstacles.py:35-58 */
    ORC_GEN_KEYCORRIDOR = 4, /* roomgrid.py:118-169 + envs/keycorridor.py:26-49 */
    ORC_GEN_POOL = 5,        /* reset = one of the uploaded reference layouts (base MiniGridEnv.step only) */
    ORC_GEN_CROSSING = 6,    /* envs/crossing.py:24-99; gen_param0 = num_crossings, gen_param1 = ori | 4 * (obstacle_type == Wall) */
    ORC_GEN_LAVAGAP = 7,     /* envs/lavagap.py:21-60; gen_param0 = const, gen_param1 = (obstacle_type == Wall) */
    ORC_GEN_MULTIROOM = 8,   /* envs/multiroom.py:41-241; gen_param0 = minNumRooms == maxNumRooms, gen_param1 = maxRoomSize */
    ORC_GEN_DISTSHIFT = 9    /* envs/distshift.py:30-52; gen_param0 = strip2_row */
};

typedef struct {
    int32_t gen;            /* ORC_GEN_* */
    int32_t width, height;
    int32_t max_steps;
    int32_t see_through;    /* see_through_walls */
    int32_t n_actions;      /* action_space.n (7; 3 for DynObs) */
    int32_t n_obstacles;    /* DynObs */
    int32_t room_size;      /* KeyCorridor / RoomGrid */
    int32_t num_rows;       /* KeyCorridor / RoomGrid (num_cols == 3) */
    int32_t random_start;   /* Empty-Random / DynObs-Random: agent_start_pos=None */
    int32_t lava_v1;        /* 'v1' in type(env).__name__ (minigrid.py:1263-1266): true for e.g. DoorKeyEn*v1*6x16 */
    int32_t view_size;      /* agent_view_size (minigrid.py:776,795; ViewSizeWrapper wrappers.py:579-608); 0 = 7 */
    int32_t hook;           /* ORC_GEN_POOL only: 1 pickup target (unlockpickup.py:34-42) 2 unlock (unlock.py:33-41)
                               3 fetch (fetch.py:74-86) 4 gotodoor (gotodoor.py:72-93) 5 gotoobject (gotoobject.py:68-84)
                               6 putnear (putnear.py:91-112) 7 redbluedoors (redbluedoors.py:44-66) 8 memory (memory.py:88-100) */
    int32_t gen_param0, gen_param1;   /* see ORC_GEN_CROSSING / LAVAGAP / MULTIROOM */
} orc_config;

#define ORC_OBS_BYTES 147
#define ORC_MAX_OBST 8

typedef struct orc_vec orc_vec;

/* n independent envs with global ids env0 .. env0+n-1; Philox key = seed */
orc_vec *orc_vec_create(const orc_config *cfg, uint64_t seed, int64_t env0, int32_t n);
void orc_vec_destroy(orc_vec *v);
void orc_set_threads(int nthreads);

/* level pool for ORC_GEN_POOL: grid [K][W][H][3], aux [K][W][H] or NULL, agent [K][3] = x,y,dir.
 * reset picks level rand_int(0, K) (first draw of the episode's stream). */
/* hook_params [K][16] (NULL if hook == 0): target_type, target_color, move_type, move_color, target_x, target_y,
 * A.x, A.y, B.x, B.y, C.x, C.y, D.x, D.y, 0, 0 */
int orc_vec_set_level_pool(orc_vec *v, int32_t n_levels, const uint8_t *grid, const uint8_t *aux, const int32_t *agent,
                           const int32_t *hook_params);
int orc_vec_get_levels(orc_vec *v, int32_t *levels);
int orc_vec_set_levels(orc_vec *v, const int32_t *levels);

/* RNG tape (parity mode 2): draws[offsets[i] .. offsets[i+1]) are the raw randint
 * results env i will consume, in order.  NULL disables. */
int orc_vec_set_tape(orc_vec *v, const int32_t *draws, const int64_t *offsets);

/* obs records are view_size*view_size*3 bytes (147 for the default 7).
 * reset envs where mask[i]!=0 (all if mask NULL): obs [n][147], dir [n] */
int orc_vec_reset(orc_vec *v, const uint8_t *mask, uint8_t *obs, uint8_t *dir);

/* one step; if autoreset, a done env is reset and obs/dir are those of the new episode */
int orc_vec_step(orc_vec *v, const uint8_t *actions, int autoreset,
                 uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir);

/* T steps; actions [T][n]; outputs [T][n]...; any output may be NULL */
#define ORC_ACTION_KEY 0x41435431u   /* "ACT1": key tweak of the random-policy stream */
int orc_policy_action(uint64_t seed, int64_t env_id, uint32_t epoch, uint32_t t, int n_actions);
/* T steps under the uniform random policy (run_tests.py:43); the actions taken are written to actions_out [T][n] */
int orc_vec_rollout_random(orc_vec *v, int32_t T, int autoreset, uint8_t *actions_out,
                           uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir);
int orc_vec_rollout(orc_vec *v, int32_t T, const uint8_t *actions, int autoreset,
                    uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir);

/* state exchange, layout = Grid.encode (minigrid.py:571-594): grid [n][W][H][3];
 * aux [n][W][H] bit0 = Goal.overlap; agent [n][4] = x,y,dir,step_count;
 * carrying [n][3] ((0,0,0)=none); obstacles [n][8][2]; target [n][2]=(type,color);
 * rng [n][2] = (episode, ndraws) */
int orc_vec_get_state(orc_vec *v, uint8_t *grid, uint8_t *aux, int32_t *agent,
                      uint8_t *carrying, int16_t *obstacles, uint8_t *target, uint32_t *rng);
int orc_vec_set_state(orc_vec *v, const uint8_t *grid, const uint8_t *aux, const int32_t *agent,
                      const uint8_t *carrying, const int16_t *obstacles, const uint8_t *target,
                      const uint32_t *rng);

/* Philox4x32-10 known-answer hook */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

const char *orc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif

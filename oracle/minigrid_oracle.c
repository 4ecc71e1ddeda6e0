/* TEST INFRASTRUCTURE ONLY -- see minigrid_oracle.h.
 *
 * A deliberately literal CPU restatement of the reference algorithm.  Every
 * function cites the reference lines it follows (paths relative to
 * /root/reference/gym_minigrid/).  It keeps the reference's *structure*
 * (object grid, slice, rotate_left x (dir+1), process_vis double sweep,
 * encode with vis_mask) so that it is an independent check of the closed
 * forms used by the CUDA kernel.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared).
 * -ffp-contract=off matters: _reward() is three separately rounded fp64 ops
 * in Python (minigrid.py:933-937); an FMA would change the last bit.
 */
#include "minigrid_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- minigrid.py:27-61 ------------------------------------------------ */
enum { T_UNSEEN = 0, T_EMPTY = 1, T_WALL = 2, T_FLOOR = 3, T_DOOR = 4, T_KEY = 5,
       T_BALL = 6, T_BOX = 7, T_GOAL = 8, T_LAVA = 9, T_AGENT = 10 };
enum { C_RED = 0, C_GREEN = 1, C_BLUE = 2, C_PURPLE = 3, C_YELLOW = 4, C_GREY = 5, C_WHITE = 6 };
/* COLOR_NAMES = sorted(COLORS.keys()) (minigrid.py:24):
 * blue, green, grey, purple, red, white, yellow */
static const uint8_t COLOR_NAMES_IDX[7] = { C_BLUE, C_GREEN, C_GREY, C_PURPLE, C_RED, C_WHITE, C_YELLOW };
/* DIR_TO_VEC minigrid.py:64-73 */
static const int DIRX[4] = { 1, 0, -1, 0 };
static const int DIRY[4] = { 0, 1, 0, -1 };
enum { A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6 };

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int g_threads = 0;
void orc_set_threads(int n) { g_threads = n; }

/* ---- WorldObj and subclasses, by value (minigrid.py:75-364) ------------ */
typedef struct {
    uint8_t has;          /* 0 == Python None */
    uint8_t type, color;
    uint8_t is_open, is_locked;   /* Door  :240-243 */
    int8_t toggletimes;           /* Goal :158, Box :336 */
    uint8_t overlap;              /* Goal :160 */
    uint8_t is_target;            /* identity of KeyCorridor.obj (keycorridor.py:48) */
    uint8_t contains_key;         /* Box.contains = Key(colour k): k+1, 0 = None (minigrid.py:335; obstructedmaze.py:68-73) */
} Obj;

static const Obj NONE = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };

static Obj mk(uint8_t type, uint8_t color) {
    Obj o = NONE; o.has = 1; o.type = type; o.color = color; return o;
}
static Obj mk_wall(void) { return mk(T_WALL, C_GREY); }               /* :229-231 */
static Obj mk_goal(int toggletimes) {                                 /* :156-162 */
    Obj o = mk(T_GOAL, C_GREEN); o.toggletimes = (int8_t)toggletimes; o.overlap = toggletimes <= 0; return o;
}
static Obj mk_door(uint8_t color, int is_open, int is_locked) {       /* :239-243 */
    Obj o = mk(T_DOOR, color); o.is_open = (uint8_t)is_open; o.is_locked = (uint8_t)is_locked; return o;
}
static Obj mk_box(uint8_t color) { Obj o = mk(T_BOX, color); o.toggletimes = 1; return o; } /* :332-337 */

static int can_overlap(const Obj *o) {   /* :93-95,164-166,192-193,211-212,245-247,342-343 */
    switch (o->type) {
    case T_GOAL: case T_FLOOR: case T_LAVA: return 1;
    case T_DOOR: return o->is_open;
    case T_BOX: return 0;               /* color == triage_color(None) */
    default: return 0;
    }
}
static int can_pickup(const Obj *o) {    /* :97-99,305-306,326-327,339-340 */
    return o->type == T_KEY || o->type == T_BALL || o->type == T_BOX;
}
static int see_behind(const Obj *o) {    /* :105-107,233-234,249-250 */
    if (o->type == T_WALL) return 0;
    if (o->type == T_DOOR) return o->is_open;
    return 1;
}
static void obj_encode(const Obj *o, uint8_t out[3]) {   /* :113-115, Door :264-275 */
    out[0] = o->type; out[1] = o->color; out[2] = 0;
    if (o->type == T_DOOR) {
        if (o->is_open) out[2] = 0;
        else if (o->is_locked) out[2] = 2;
        else out[2] = 1;
    }
}
static int obj_decode(int t, int c, int s, Obj *out) {   /* :117-150 */
    *out = NONE;
    if (t < 0 || t > T_LAVA || c < 0 || c > 6) return -1;
    if (t == T_EMPTY || t == T_UNSEEN) return 0;
    switch (t) {
    case T_WALL: case T_FLOOR: case T_BALL: case T_KEY: *out = mk((uint8_t)t, (uint8_t)c); break;
    case T_BOX: *out = mk_box((uint8_t)c); break;
    case T_DOOR: *out = mk_door((uint8_t)c, s == 0, s == 2); break;
    case T_GOAL: *out = mk_goal(1); break;       /* Goal() ignores the colour byte :143-144 */
    case T_LAVA: *out = mk(T_LAVA, C_RED); break;
    default: return -1;
    }
    return 0;
}

/* ---- Grid (minigrid.py:366-718) ---------------------------------------- */
#define VIEW 11          /* maximum agent_view_size; the actual size is cfg.view_size */
typedef struct { int w, h; Obj *c; } Grid;       /* c[j*w + i]  (:414) */
typedef struct { int w, h; Obj c[VIEW * VIEW]; } VGrid;

static __thread int g_oob; /* set when an assert in Grid.get/set (:412-419) would fire */

static Obj grid_get(const Grid *g, int i, int j) {
    if (i < 0 || i >= g->w || j < 0 || j >= g->h) { g_oob = 1; return NONE; }
    return g->c[j * g->w + i];
}
static void grid_set(Grid *g, int i, int j, Obj v) {
    if (i < 0 || i >= g->w || j < 0 || j >= g->h) { g_oob = 1; return; }
    g->c[j * g->w + i] = v;
}
static void horz_wall(Grid *g, int x, int y, int length) {   /* :421-425 */
    if (length < 0) length = g->w - x;
    for (int i = 0; i < length; i++) grid_set(g, x + i, y, mk_wall());
}
static void vert_wall(Grid *g, int x, int y, int length) {   /* :427-431 */
    if (length < 0) length = g->h - y;
    for (int j = 0; j < length; j++) grid_set(g, x, y + j, mk_wall());
}
static void wall_rect(Grid *g, int x, int y, int w, int h) { /* :433-437 */
    horz_wall(g, x, y, w); horz_wall(g, x, y + h - 1, w);
    vert_wall(g, x, y, h); vert_wall(g, x + w - 1, y, h);
}
static Obj vget(const VGrid *g, int i, int j) { return g->c[j * g->w + i]; }
static void vset(VGrid *g, int i, int j, Obj v) { g->c[j * g->w + i] = v; }

static void grid_slice(const Grid *g, int topX, int topY, int width, int height, VGrid *out) { /* :453-473 */
    out->w = width; out->h = height;
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++) {
            int x = topX + i, y = topY + j;
            Obj v;
            if (x >= 0 && x < g->w && y >= 0 && y < g->h) v = grid_get(g, x, y);
            else v = mk_wall();
            vset(out, i, j, v);
        }
}
static void rotate_left(const VGrid *g, VGrid *out) {        /* :439-451 */
    out->w = g->h; out->h = g->w;
    for (int i = 0; i < g->w; i++)
        for (int j = 0; j < g->h; j++)
            vset(out, j, out->h - 1 - i, vget(g, i, j));
}
/* default branch only (:617-648, 712-718); mask[i][j] */
static void process_vis(VGrid *g, int ax, int ay, uint8_t mask[VIEW][VIEW]) {
    memset(mask, 0, VIEW * VIEW);
    mask[ax][ay] = 1;
    for (int j = g->h - 1; j >= 0; j--) {
        for (int i = 0; i < g->w - 1; i++) {
            if (!mask[i][j]) continue;
            Obj cell = vget(g, i, j);
            if (cell.has && !see_behind(&cell)) continue;
            mask[i + 1][j] = 1;
            if (j > 0) { mask[i + 1][j - 1] = 1; mask[i][j - 1] = 1; }
        }
        for (int i = g->w - 1; i >= 1; i--) {
            if (!mask[i][j]) continue;
            Obj cell = vget(g, i, j);
            if (cell.has && !see_behind(&cell)) continue;
            mask[i - 1][j] = 1;
            if (j > 0) { mask[i - 1][j - 1] = 1; mask[i][j - 1] = 1; }
        }
    }
    for (int j = 0; j < g->h; j++)
        for (int i = 0; i < g->w; i++)
            if (!mask[i][j]) vset(g, i, j, NONE);
}
/* Grid.encode(vis_mask) (:571-594): array[i][j][:]  -> offset (i*h + j)*3 */
static void vgrid_encode(const VGrid *g, uint8_t mask[VIEW][VIEW], uint8_t *out) {
    memset(out, 0, (size_t)g->w * g->h * 3);
    for (int i = 0; i < g->w; i++)
        for (int j = 0; j < g->h; j++)
            if (mask[i][j]) {
                Obj v = vget(g, i, j);
                uint8_t *p = out + (i * g->h + j) * 3;
                if (!v.has) { p[0] = T_EMPTY; p[1] = 0; p[2] = 0; }
                else obj_encode(&v, p);
            }
}

/* ---- Philox4x32-10 ------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- RoomGrid bookkeeping (roomgrid.py:14-37) --------------------------- */
#define MAXR 3
typedef struct {
    int top[2], size[2];
    uint8_t doors[4];          /* None / Door / True  -> truthiness only */
    uint8_t has_door_pos[4];
    int door_pos[4][2];
    int nb[4][2];              /* neighbour (i,j) or -1 */
    uint8_t locked;
} Room;

/* ---- MiniGridEnv (minigrid.py:720-1381) --------------------------------- */
struct orc_vec;
typedef struct {
    orc_config cfg;
    Grid grid;
    int has_agent; int ax, ay, adir;
    Obj carrying;
    int step_count;
    int obst[ORC_MAX_OBST][2];
    uint8_t target_type, target_color;
    uint32_t episode, ndraws;
    int64_t env_id;
    uint64_t seed;
    const int32_t *tape; int64_t tape_len;
    int err;
    Room rooms[MAXR][MAXR];    /* room_grid[j][i] */
    const struct orc_vec *owner;
    int level;                 /* ORC_GEN_POOL: level being played */
} Env;

struct orc_vec {
    orc_config cfg;
    int n;
    uint32_t policy_epoch;   /* number of random-policy rollouts so far */
    Env *envs;
    Obj *cells;
    int pool_n;          /* ORC_GEN_POOL */
    Obj *pool_cells;     /* [pool_n][W*H], index j*W+i like Grid */
    int *pool_agent;     /* [pool_n][3] */
    int *pool_hook;      /* [pool_n][16] or NULL */
};

#define ORC_DRAW_ODD_MULT 0x9E3779B1u
static int rand_int(Env *e, int low, int high) {   /* minigrid.py:939-944 */
    if (e->tape) {
        if ((int64_t)e->ndraws >= e->tape_len) { e->err |= 2; return low; }
        int v = e->tape[e->ndraws++];   /* in tape mode ndraws is the tape cursor and never rewinds */
        if (v < low || v >= high) e->err |= 4;
        return v;
    }
    /* stream id = index of the current episode = (#resets so far) - 1.  Every Philox word serves two consecutive
     * draws: draw n = word (n>>1)&3 of block n>>3, multiplied by ORC_DRAW_ODD_MULT when n is odd (DESIGN.md "RNG") */
    uint32_t ctr[4] = { e->ndraws >> 3, e->episode - 1u, (uint32_t)e->env_id, (uint32_t)((uint64_t)e->env_id >> 32) };
    uint32_t key[2] = { (uint32_t)e->seed, (uint32_t)(e->seed >> 32) };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    uint32_t u = out[(e->ndraws >> 1) & 3];
    if (e->ndraws & 1) u *= ORC_DRAW_ODD_MULT;
    e->ndraws++;
    return low + (int)(((uint64_t)u * (uint32_t)(high - low)) >> 32);
}

/* place_obj (minigrid.py:1003-1061). max_tries<0 == math.inf. reject_next_to = roomgrid.py:3-12.
 * returns 0, or -1 for the RecursionError. */
static int place_obj(Env *e, Obj obj, int has_top, int topx, int topy, int has_size, int sx, int sy,
                     int reject_next_to, int max_tries, int *px, int *py) {
    Grid *g = &e->grid;
    if (!has_top) { topx = 0; topy = 0; }
    else { if (topx < 0) topx = 0; if (topy < 0) topy = 0; }
    if (!has_size) { sx = g->w; sy = g->h; }
    int num_tries = 0, x, y;
    for (;;) {
        if (max_tries >= 0 && num_tries > max_tries) return -1;
        num_tries++;
        int hx = topx + sx < g->w ? topx + sx : g->w;
        int hy = topy + sy < g->h ? topy + sy : g->h;
        x = rand_int(e, topx, hx);
        y = rand_int(e, topy, hy);
        if (e->err & 6) return -1;
        if (grid_get(g, x, y).has) continue;
        if (e->has_agent && x == e->ax && y == e->ay) continue;
        if (reject_next_to) {
            int d = abs(e->ax - x) + abs(e->ay - y);
            if (d < 2) continue;
        }
        break;
    }
    grid_set(g, x, y, obj);
    if (px) *px = x;
    if (py) *py = y;
    return 0;
}
/* MiniGridEnv.place_agent (minigrid.py:1072-1090) */
static int place_agent(Env *e, int has_top, int topx, int topy, int has_size, int sx, int sy, int max_tries) {
    e->has_agent = 0;
    int x, y;
    if (place_obj(e, NONE, has_top, topx, topy, has_size, sx, sy, 0, max_tries, &x, &y)) return -1;
    e->ax = x; e->ay = y; e->has_agent = 1;
    e->adir = rand_int(e, 0, 4);
    return 0;
}

/* envs/empty.py:30-57 (extra == 0 only) */
static void gen_empty(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    wall_rect(g, 0, 0, W, H);
    grid_set(g, W - 2, H - 2, mk_goal(1));
    if (!e->cfg.random_start) { e->ax = 1; e->ay = 1; e->adir = 0; e->has_agent = 1; }
    else place_agent(e, 0, 0, 0, 0, 0, 0, -1);
}
/* envs/doorkey.py:15-44 */
static void gen_doorkey(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    wall_rect(g, 0, 0, W, H);
    grid_set(g, W - 2, H - 2, mk_goal(1));
    int split = rand_int(e, 2, W - 2);
    vert_wall(g, split, 0, -1);
    place_agent(e, 0, 0, 0, 1, split, H, -1);
    int door = rand_int(e, 1, W - 2);
    grid_set(g, split, door, mk_door(C_YELLOW, 0, 1));
    place_obj(e, mk(T_KEY, C_YELLOW), 1, 0, 0, 1, split, H, 0, -1, NULL, NULL);
}
/* envs/fourrooms.py:19-69 (agent_pos=None, goal_pos=None) */
static void gen_fourrooms(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    horz_wall(g, 0, 0, -1); horz_wall(g, 0, H - 1, -1);
    vert_wall(g, 0, 0, -1); vert_wall(g, W - 1, 0, -1);
    int room_w = W / 2, room_h = H / 2;
    for (int j = 0; j < 2; j++)
        for (int i = 0; i < 2; i++) {
            int xL = i * room_w, yT = j * room_h, xR = xL + room_w, yB = yT + room_h;
            if (i + 1 < 2) {
                vert_wall(g, xR, yT, room_h);
                int y = rand_int(e, yT + 1, yB);
                grid_set(g, xR, y, NONE);
            }
            if (j + 1 < 2) {
                horz_wall(g, xL, yB, room_w);
                int x = rand_int(e, xL + 1, xR);
                grid_set(g, x, yB, NONE);
            }
        }
    place_agent(e, 0, 0, 0, 0, 0, 0, -1);
    place_obj(e, mk_goal(1), 0, 0, 0, 0, 0, 0, 0, -1, NULL, NULL);
}
/* envs/dynamicobstacles.py:35-58 */
static void gen_dynobs(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    wall_rect(g, 0, 0, W, H);
    grid_set(g, W - 2, H - 2, mk_goal(1));
    if (!e->cfg.random_start) { e->ax = 1; e->ay = 1; e->adir = 0; e->has_agent = 1; }
    else place_agent(e, 0, 0, 0, 0, 0, 0, -1);
    for (int k = 0; k < e->cfg.n_obstacles; k++) {
        int x = 0, y = 0;
        if (place_obj(e, mk(T_BALL, C_BLUE), 0, 0, 0, 0, 0, 0, 0, 100, &x, &y)) e->err |= 8; /* RecursionError escapes reset */
        e->obst[k][0] = x; e->obst[k][1] = y;
    }
}

/* ---- RoomGrid + KeyCorridor (roomgrid.py, envs/keycorridor.py) ---------- */
static Room *get_room(Env *e, int i, int j) { return &e->rooms[j][i]; }
static int rand_color(Env *e) { return COLOR_NAMES_IDX[rand_int(e, 0, 7)]; } /* minigrid.py:960-991 */

static void roomgrid_gen(Env *e) {                 /* roomgrid.py:118-169 */
    Grid *g = &e->grid;
    int rs = e->cfg.room_size, rows = e->cfg.num_rows, cols = 3;
    for (int j = 0; j < rows; j++)
        for (int i = 0; i < cols; i++) {
            Room *r = get_room(e, i, j);
            memset(r, 0, sizeof(*r));
            r->top[0] = i * (rs - 1); r->top[1] = j * (rs - 1);
            r->size[0] = rs; r->size[1] = rs;
            for (int k = 0; k < 4; k++) r->nb[k][0] = r->nb[k][1] = -1;
            wall_rect(g, r->top[0], r->top[1], rs, rs);
        }
    for (int j = 0; j < rows; j++)
        for (int i = 0; i < cols; i++) {
            Room *r = get_room(e, i, j);
            int x_l = r->top[0] + 1, y_l = r->top[1] + 1;
            int x_m = r->top[0] + r->size[0] - 1, y_m = r->top[1] + r->size[1] - 1;
            if (i < cols - 1) {
                r->nb[0][0] = i + 1; r->nb[0][1] = j;
                r->door_pos[0][0] = x_m; r->door_pos[0][1] = rand_int(e, y_l, y_m); r->has_door_pos[0] = 1;
            }
            if (j < rows - 1) {
                r->nb[1][0] = i; r->nb[1][1] = j + 1;
                r->door_pos[1][0] = rand_int(e, x_l, x_m); r->door_pos[1][1] = y_m; r->has_door_pos[1] = 1;
            }
            if (i > 0) {
                r->nb[2][0] = i - 1; r->nb[2][1] = j;
                Room *n = get_room(e, i - 1, j);
                r->door_pos[2][0] = n->door_pos[0][0]; r->door_pos[2][1] = n->door_pos[0][1]; r->has_door_pos[2] = n->has_door_pos[0];
            }
            if (j > 0) {
                r->nb[3][0] = i; r->nb[3][1] = j - 1;
                Room *n = get_room(e, i, j - 1);
                r->door_pos[3][0] = n->door_pos[1][0]; r->door_pos[3][1] = n->door_pos[1][1]; r->has_door_pos[3] = n->has_door_pos[1];
            }
        }
    e->ax = (cols / 2) * (rs - 1) + rs / 2;
    e->ay = (rows / 2) * (rs - 1) + rs / 2;
    e->adir = 0; e->has_agent = 1;
}
/* roomgrid.py:212-246 with door_idx, color and locked all given */
static void add_door(Env *e, int i, int j, int door_idx, int color, int locked) {
    Room *r = get_room(e, i, j);
    r->locked = (uint8_t)locked;
    grid_set(&e->grid, r->door_pos[door_idx][0], r->door_pos[door_idx][1], mk_door((uint8_t)color, 0, locked));
    Room *n = get_room(e, r->nb[door_idx][0], r->nb[door_idx][1]);
    r->doors[door_idx] = 1;
    n->doors[(door_idx + 2) % 4] = 1;
}
/* roomgrid.py:248-282 */
static void remove_wall(Env *e, int i, int j, int wall_idx) {
    Room *r = get_room(e, i, j);
    Grid *g = &e->grid;
    int tx = r->top[0], ty = r->top[1], w = r->size[0], h = r->size[1];
    if (wall_idx == 0) for (int k = 1; k < h - 1; k++) grid_set(g, tx + w - 1, ty + k, NONE);
    else if (wall_idx == 1) for (int k = 1; k < w - 1; k++) grid_set(g, tx + k, ty + h - 1, NONE);
    else if (wall_idx == 2) for (int k = 1; k < h - 1; k++) grid_set(g, tx, ty + k, NONE);
    else for (int k = 1; k < w - 1; k++) grid_set(g, tx + k, ty, NONE);
    Room *n = get_room(e, r->nb[wall_idx][0], r->nb[wall_idx][1]);
    r->doors[wall_idx] = 1;
    n->doors[(wall_idx + 2) % 4] = 1;
}
/* roomgrid.py:171-210: add_object(i, j, kind, color) -> place_in_room */
static void add_object(Env *e, int i, int j, Obj obj) {
    Room *r = get_room(e, i, j);
    if (place_obj(e, obj, 1, r->top[0], r->top[1], 1, r->size[0], r->size[1], 1, 1000, NULL, NULL)) e->err |= 8;
}
/* RoomGrid.place_agent(i, j) (roomgrid.py:284-303) */
static void room_place_agent(Env *e, int i, int j) {
    Room *r = get_room(e, i, j);
    for (;;) {
        if (place_agent(e, 1, r->top[0], r->top[1], 1, r->size[0], r->size[1], 1000)) { e->err |= 8; return; }
        Obj f = grid_get(&e->grid, e->ax + DIRX[e->adir], e->ay + DIRY[e->adir]);
        if (!f.has || f.type == T_WALL) break;
    }
}
/* roomgrid.py:305-359 */
static void connect_all(Env *e) {
    int rs = e->cfg.room_size, rows = e->cfg.num_rows, cols = 3;
    int si = e->ax / (rs - 1), sj = e->ay / (rs - 1);   /* room_from_pos :99-111 */
    int num_itrs = 0;
    for (;;) {
        if (num_itrs > 5000) { e->err |= 8; return; }
        num_itrs++;
        /* find_reach */
        uint8_t reach[MAXR][MAXR]; memset(reach, 0, sizeof(reach));
        int stack[64][2], sp = 0, count = 0;
        stack[sp][0] = si; stack[sp][1] = sj; sp++;
        while (sp > 0) {
            sp--; int ci = stack[sp][0], cj = stack[sp][1];
            if (reach[cj][ci]) continue;
            reach[cj][ci] = 1; count++;
            Room *r = get_room(e, ci, cj);
            for (int k = 0; k < 4; k++)
                if (r->doors[k] && sp < 63) { stack[sp][0] = r->nb[k][0]; stack[sp][1] = r->nb[k][1]; sp++; }
        }
        if (count == rows * cols) break;
        int i = rand_int(e, 0, cols);
        int j = rand_int(e, 0, rows);
        int k = rand_int(e, 0, 4);
        if (e->err & 6) return;
        Room *r = get_room(e, i, j);
        if (!r->has_door_pos[k] || r->doors[k]) continue;
        if (r->locked || get_room(e, r->nb[k][0], r->nb[k][1])->locked) continue;
        int color = rand_color(e);
        add_door(e, i, j, k, color, 0);
    }
}
/* envs/keycorridor.py:26-49 (obj_type == "ball") */
static void gen_keycorridor(Env *e) {
    int rows = e->cfg.num_rows;
    roomgrid_gen(e);
    for (int j = 1; j < rows; j++) remove_wall(e, 1, j, 3);
    int room_idx = rand_int(e, 0, rows);
    int door_color = rand_color(e);                       /* add_door(2, room_idx, 2, locked=True) */
    add_door(e, 2, room_idx, 2, door_color, 1);
    int obj_color = rand_color(e);                        /* add_object(2, room_idx, kind="ball") */
    Obj ball = mk(T_BALL, (uint8_t)obj_color); ball.is_target = 1;
    add_object(e, 2, room_idx, ball);
    int key_room = rand_int(e, 0, rows);
    add_object(e, 0, key_room, mk(T_KEY, (uint8_t)door_color));
    room_place_agent(e, 1, rows / 2);
    connect_all(e);
    e->target_type = T_BALL; e->target_color = (uint8_t)obj_color;
}

/* ---- the RandomState methods crossing.py calls besides randint, as oracle/ref_shim.py defines them on the stream:
 * shuffle = Fisher-Yates from the back, one randint(0, i+1) per position; choice(range(a, b)) = a + randint(0, b-a) ---- */
static void rand_shuffle(Env *e, int *x, int n) {
    for (int i = n - 1; i >= 1; i--) {
        int j = rand_int(e, 0, i + 1);
        int t = x[i]; x[i] = x[j]; x[j] = t;
    }
}
static int rand_choice_range(Env *e, int lo, int hi) { return lo + rand_int(e, 0, hi - lo); }
static void sort_ints(int *x, int n) {
    for (int i = 1; i < n; i++) { int v = x[i], j = i; while (j > 0 && x[j - 1] > v) { x[j] = x[j - 1]; j--; } x[j] = v; }
}

/* envs/crossing.py:24-99.  A river is (direction, position): direction v (vertical, a column) = 0, h (a row) = 1. */
static void gen_crossing(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    const int num_crossings = e->cfg.gen_param0, ori = e->cfg.gen_param1 & 3;
    const Obj obstacle = (e->cfg.gen_param1 & 4) ? mk_wall() : mk(T_LAVA, C_RED);
    wall_rect(g, 0, 0, W, H);                                    /* :31 */
    e->ax = 1; e->ay = 1; e->adir = 0; e->has_agent = 1;         /* :34-35 */
    grid_set(g, W - 2, H - 2, mk_goal(1));                       /* :38 */
    int rivers[64], nr = 0;                                      /* :44-52, encoded dir*64 + pos */
    if (ori == 0) { for (int j = 2; j < W - 2; j += 2) rivers[nr++] = 64 + j; }
    else if (ori == 1) { for (int i = 2; i < H - 2; i += 2) rivers[nr++] = i; }
    else { for (int i = 2; i < H - 2; i += 2) rivers[nr++] = i; for (int j = 2; j < W - 2; j += 2) rivers[nr++] = 64 + j; }
    rand_shuffle(e, rivers, nr);                                 /* :53 */
    if (nr > num_crossings) nr = num_crossings;                  /* :54 */
    int rv[32], rh[32], nv = 0, nh = 0;
    for (int k = 0; k < nr; k++) { if (rivers[k] < 64) rv[nv++] = rivers[k]; else rh[nh++] = rivers[k] - 64; }
    sort_ints(rv, nv); sort_ints(rh, nh);                        /* :55-56 */
    for (int i = 1; i < W - 1; i++) for (int k = 0; k < nh; k++) grid_set(g, i, rh[k], obstacle);   /* :57-62: product(range(1, width-1), rivers_h) */
    for (int k = 0; k < nv; k++) for (int j = 1; j < H - 1; j++) grid_set(g, rv[k], j, obstacle);   /*          product(rivers_v, range(1, height-1)) */
    int path[64], np = 0;                                        /* :65: [h] * len(rivers_v) + [v] * len(rivers_h) */
    for (int k = 0; k < nv; k++) path[np++] = 1;
    for (int k = 0; k < nh; k++) path[np++] = 0;
    rand_shuffle(e, path, np);                                   /* :66 */
    int limits_v[34], limits_h[34];                              /* :69-70 */
    limits_v[0] = 0; for (int k = 0; k < nv; k++) limits_v[k + 1] = rv[k]; limits_v[nv + 1] = H - 1;
    limits_h[0] = 0; for (int k = 0; k < nh; k++) limits_h[k + 1] = rh[k]; limits_h[nh + 1] = W - 1;
    int room_i = 0, room_j = 0;
    for (int k = 0; k < np; k++) {                               /* :72-85 */
        int i, j;
        if (path[k] == 1) {                                      /* direction is h */
            i = limits_v[room_i + 1];
            j = rand_choice_range(e, limits_h[room_j] + 1, limits_h[room_j + 1]);
            room_i++;
        } else {
            i = rand_choice_range(e, limits_v[room_i] + 1, limits_v[room_i + 1]);
            j = limits_h[room_j + 1];
            room_j++;
        }
        grid_set(g, i, j, NONE);
    }
}

/* envs/lavagap.py:21-60 */
static void gen_lavagap(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    const Obj obstacle = e->cfg.gen_param1 ? mk_wall() : mk(T_LAVA, C_RED);
    wall_rect(g, 0, 0, W, H);
    e->ax = 1; e->ay = 1; e->adir = 0; e->has_agent = 1;
    grid_set(g, W - 2, H - 2, mk_goal(1));
    int gx, gy;
    if (!e->cfg.gen_param0) { gx = rand_int(e, 2, W - 2); gy = rand_int(e, 1, H - 1); }     /* :40-44 */
    else { gx = W / 2; gy = rand_int(e, 1, H - 1); }                                         /* :45-49 */
    for (int j = 0; j < H - 2; j++) grid_set(g, gx, 1 + j, obstacle);                        /* :52 vert_wall(x, 1, height-2, obstacle_type) */
    grid_set(g, gx, gy, NONE);                                                               /* :55 */
}

/* envs/multiroom.py:41-241 */
typedef struct { int topX, topY, sizeX, sizeY, entryX, entryY; } MRoom;
static int mr_place_room(Env *e, int numLeft, MRoom *list, int *n, int minSz, int maxSz, int entryDoorWall, int ex, int ey) {  /* :123-241 */
    Grid *g = &e->grid;
    int sizeX = rand_int(e, minSz, maxSz + 1);
    int sizeY = rand_int(e, minSz, maxSz + 1);
    int topX, topY;
    if (*n == 0) { topX = ex; topY = ey; }
    else if (entryDoorWall == 0) { topX = ex - sizeX + 1; topY = rand_int(e, ey - sizeY + 2, ey); }
    else if (entryDoorWall == 1) { topX = rand_int(e, ex - sizeX + 2, ex); topY = ey - sizeY + 1; }
    else if (entryDoorWall == 2) { topX = ex; topY = rand_int(e, ey - sizeY + 2, ey); }
    else { topX = rand_int(e, ex - sizeX + 2, ex); topY = ey; }
    if (e->err & 6) return 0;
    if (topX < 0 || topY < 0) return 0;                                            /* :164-167 */
    if (topX + sizeX > g->w || topY + sizeY >= g->h) return 0;
    for (int k = 0; k + 1 < *n; k++) {                                             /* :170-178: roomList[:-1] */
        const MRoom *r = &list[k];
        int nonOverlap = topX + sizeX < r->topX || r->topX + r->sizeX <= topX || topY + sizeY < r->topY || r->topY + r->sizeY <= topY;
        if (!nonOverlap) return 0;
    }
    MRoom m = { topX, topY, sizeX, sizeY, ex, ey };
    list[(*n)++] = m;                                                              /* :181-186 */
    if (numLeft == 1) return 1;                                                    /* :189-190 */
    for (int i = 0; i < 8; i++) {                                                  /* :193-239 */
        int walls[3], nw = 0;
        for (int w = 0; w < 4; w++) if (w != entryDoorWall) walls[nw++] = w;       /* sorted(wallSet) */
        int exitDoorWall = walls[rand_int(e, 0, 3)];
        int nextEntryWall = (exitDoorWall + 2) % 4;
        int dx, dy;
        if (exitDoorWall == 0) { dx = topX + sizeX - 1; dy = topY + rand_int(e, 1, sizeY - 1); }
        else if (exitDoorWall == 1) { dx = topX + rand_int(e, 1, sizeX - 1); dy = topY + sizeY - 1; }
        else if (exitDoorWall == 2) { dx = topX; dy = topY + rand_int(e, 1, sizeY - 1); }
        else { dx = topX + rand_int(e, 1, sizeX - 1); dy = topY; }
        if (e->err & 6) return 1;
        if (mr_place_room(e, numLeft - 1, list, n, minSz, maxSz, nextEntryWall, dx, dy)) break;
    }
    return 1;
}
static void gen_multiroom(Env *e) {
    Grid *g = &e->grid; int W = g->w;
    const int numRooms = e->cfg.gen_param0, maxRoomSize = e->cfg.gen_param1;       /* :44: _rand_int(min, max+1) */
    MRoom best[8], cur[8];
    int nbest = 0;
    const int nr = rand_int(e, numRooms, numRooms + 1);
    for (int guard = 0; nbest < nr; guard++) {                                     /* :46-64 */
        if (guard > 100000 || (e->err & 6)) { e->err |= 8; return; }
        int ncur = 0;
        int ex = rand_int(e, 0, W - 2), ey = rand_int(e, 0, W - 2);
        mr_place_room(e, nr, cur, &ncur, 4, maxRoomSize, 2, ex, ey);
        if (ncur > nbest) { memcpy(best, cur, sizeof(MRoom) * (size_t)ncur); nbest = ncur; }
    }
    int prev = -1;                                                                 /* prevDoorColor */
    for (int idx = 0; idx < nbest; idx++) {                                        /* :77-108 */
        const MRoom *r = &best[idx];
        for (int i = 0; i < r->sizeX; i++) { grid_set(g, r->topX + i, r->topY, mk_wall()); grid_set(g, r->topX + i, r->topY + r->sizeY - 1, mk_wall()); }
        for (int j = 0; j < r->sizeY; j++) { grid_set(g, r->topX, r->topY + j, mk_wall()); grid_set(g, r->topX + r->sizeX - 1, r->topY + j, mk_wall()); }
        if (idx > 0) {
            int cols[7], nc = 0;                                                   /* sorted(doorColors): COLOR_NAMES order minus the previous one */
            for (int k = 0; k < 7; k++) if (COLOR_NAMES_IDX[k] != prev) cols[nc++] = COLOR_NAMES_IDX[k];
            int color = cols[rand_int(e, 0, nc)];
            grid_set(g, r->entryX, r->entryY, mk_door((uint8_t)color, 0, 0));
            prev = color;
        }
    }
    if (place_agent(e, 1, best[0].topX, best[0].topY, 1, best[0].sizeX, best[0].sizeY, -1)) e->err |= 8;          /* :111 */
    if (place_obj(e, mk_goal(1), 1, best[nbest - 1].topX, best[nbest - 1].topY, 1, best[nbest - 1].sizeX, best[nbest - 1].sizeY, 0, -1, NULL, NULL)) e->err |= 8;   /* :114 */
}

/* envs/distshift.py:30-52 (agent_start_pos = (1,1), agent_start_dir = 0 in every registered id) */
static void gen_distshift(Env *e) {
    Grid *g = &e->grid; int W = g->w, H = g->h;
    wall_rect(g, 0, 0, W, H);                                           /* :35 */
    grid_set(g, W - 2, 1, mk_goal(1));                                  /* :38, goal_pos = (width-2, 1) (:19) */
    for (int i = 0; i < W - 6; i++) {                                   /* :41-43 */
        grid_set(g, 3 + i, 1, mk(T_LAVA, C_RED));
        grid_set(g, 3 + i, e->cfg.gen_param0, mk(T_LAVA, C_RED));
    }
    e->ax = 1; e->ay = 1; e->adir = 0; e->has_agent = 1;                /* :46-48 */
}

static void gen_pool(Env *e);
/* reset (minigrid.py:831-858) */
static void gen_obs(Env *e, uint8_t *obs, uint8_t *dir);
static void env_reset(Env *e, uint8_t *obs, uint8_t *dir) {
    Grid *g = &e->grid;
    for (int k = 0; k < g->w * g->h; k++) g->c[k] = NONE;
    e->has_agent = 0;
    if (!e->tape) e->ndraws = 0;
    e->episode++;
    switch (e->cfg.gen) {
    case ORC_GEN_EMPTY: gen_empty(e); break;
    case ORC_GEN_DOORKEY: gen_doorkey(e); break;
    case ORC_GEN_FOURROOMS: gen_fourrooms(e); break;
    case ORC_GEN_DYNOBS: gen_dynobs(e); break;
    case ORC_GEN_KEYCORRIDOR: gen_keycorridor(e); break;
    case ORC_GEN_POOL: gen_pool(e); break;
    case ORC_GEN_CROSSING: gen_crossing(e); break;
    case ORC_GEN_LAVAGAP: gen_lavagap(e); break;
    case ORC_GEN_MULTIROOM: gen_multiroom(e); break;
    case ORC_GEN_DISTSHIFT: gen_distshift(e); break;
    }
    e->carrying = NONE;
    e->step_count = 0;
    if (obs) gen_obs(e, obs, dir);
}

/* get_view_exts (minigrid.py:1162-1189) */
static void get_view_exts(const Env *e, int *topX, int *topY) {
    const int sz = e->cfg.view_size;
    if (e->adir == 0) { *topX = e->ax; *topY = e->ay - sz / 2; }
    else if (e->adir == 1) { *topX = e->ax - sz / 2; *topY = e->ay; }
    else if (e->adir == 2) { *topX = e->ax - sz + 1; *topY = e->ay - sz / 2; }
    else { *topX = e->ax - sz / 2; *topY = e->ay - sz + 1; }
}
/* gen_obs_grid + gen_obs (minigrid.py:1327-1381) */
static void gen_obs(Env *e, uint8_t *obs, uint8_t *dir) {
    int topX, topY;
    get_view_exts(e, &topX, &topY);
    VGrid a, b, *cur = &a, *nxt = &b;
    const int V = e->cfg.view_size;
    grid_slice(&e->grid, topX, topY, V, V, cur);
    for (int i = 0; i < e->adir + 1; i++) { rotate_left(cur, nxt); VGrid *t = cur; cur = nxt; nxt = t; }
    uint8_t mask[VIEW][VIEW];
    if (!e->cfg.see_through) process_vis(cur, V / 2, V - 1, mask);
    else memset(mask, 1, sizeof(mask));
    vset(cur, cur->w / 2, cur->h - 1, e->carrying.has ? e->carrying : NONE);
    vgrid_encode(cur, mask, obs);
    if (dir) *dir = (uint8_t)e->adir;
}

static double reward_fn(const Env *e) {   /* _reward minigrid.py:933-937; 3 rounded fp64 ops */
    volatile double q = (double)e->step_count / (double)e->cfg.max_steps;
    volatile double m = 0.9 * q;
    volatile double r = 1.0 - m;
    return r;
}

/* toggle (Door :252-262, Box :355-364, Goal :171-181, base :109-111) */
static void obj_toggle(Env *e, int x, int y) {
    Grid *g = &e->grid;
    Obj o = grid_get(g, x, y);
    if (o.type == T_DOOR) {
        if (o.is_locked) {
            if (e->carrying.has && e->carrying.type == T_KEY && e->carrying.color == o.color) {
                o.is_locked = 0; o.is_open = 1; grid_set(g, x, y, o);
            }
            return;
        }
        o.is_open = !o.is_open; grid_set(g, x, y, o);
    } else if (o.type == T_BOX) {
        o.toggletimes -= 1;
        if (o.toggletimes <= 0) grid_set(g, x, y, o.contains_key ? mk(T_KEY, (uint8_t)(o.contains_key - 1)) : NONE);   /* cell := contents */
        else grid_set(g, x, y, o);
    } else if (o.type == T_GOAL) {
        if (o.toggletimes > 0) {
            o.toggletimes -= 1;
            if (o.toggletimes <= 0) grid_set(g, x, y, NONE);
            else grid_set(g, x, y, o);
        }
    }
}

/* MiniGridEnv.step (minigrid.py:1227-1325); returns -1 for the `assert False, "unknown action"` */
static int base_step(Env *e, int action, double *reward, int *done) {
    Grid *g = &e->grid;
    int bad = 0;
    e->step_count += 1;
    *reward = 0; *done = 0;
    int fx = e->ax + DIRX[e->adir], fy = e->ay + DIRY[e->adir];
    Obj fwd = grid_get(g, fx, fy);
    /* left/right cells are fetched (and bounds-asserted) too :1242-1243 */
    (void)grid_get(g, e->ax + DIRX[(e->adir + 3) % 4], e->ay + DIRY[(e->adir + 3) % 4]);
    (void)grid_get(g, e->ax + DIRX[(e->adir + 1) % 4], e->ay + DIRY[(e->adir + 1) % 4]);
    if (action == A_LEFT) {
        e->adir -= 1; if (e->adir < 0) e->adir += 4;
    } else if (action == A_RIGHT) {
        e->adir = (e->adir + 1) % 4;
    } else if (action == A_FORWARD) {
        if (!fwd.has || can_overlap(&fwd)) { e->ax = fx; e->ay = fy; }
        if (fwd.has && fwd.type == T_GOAL && fwd.overlap) { *done = 1; *reward = 1 * reward_fn(e); }
        if (fwd.has && fwd.type == T_LAVA) {                /* :1262-1268 */
            if (e->cfg.lava_v1) { *done = 0; *reward = -1; }
            else *done = 1;
        }
    } else if (action == A_PICKUP) {
        if (fwd.has && can_pickup(&fwd)) {
            if (!e->carrying.has) { e->carrying = fwd; grid_set(g, fx, fy, NONE); }
        }
    } else if (action == A_DROP) {
        if (!fwd.has && e->carrying.has) { grid_set(g, fx, fy, e->carrying); e->carrying = NONE; }
    } else if (action == A_TOGGLE) {
        if (fwd.has) obj_toggle(e, fx, fy);
    } else if (action == A_DONE) {
        /* pass */
    } else {
        bad = 1;
    }
    if (e->step_count >= e->cfg.max_steps) *done = 1;
    return bad ? -1 : 0;
}

static int door_is_open(Env *e, int x, int y) { Obj o = grid_get(&e->grid, x, y); return o.has && o.type == T_DOOR && o.is_open; }
static int adj4(const Env *e, int x, int y) { return (e->ax == x && abs(e->ay - y) == 1) || (e->ay == y && abs(e->ax - x) == 1); }

/* step() of the stock env files that add a success/failure rule around MiniGridEnv.step; the per-level
 * attributes (self.obj, self.door, target_pos, ...) come with the level pool */
static int pool_hook_step(Env *e, int action, double *reward, int *done) {
    const struct orc_vec *v = e->owner;
    const int *hp = v->pool_hook + (size_t)e->level * 16;
    const int ttype = hp[0], tcol = hp[1], mtype = hp[2], mcol = hp[3], tx = hp[4], ty = hp[5];
    const int Ax = hp[6], Ay = hp[7], Bx = hp[8], By = hp[9], Cx = hp[10], Cy = hp[11], Dx = hp[12], Dy = hp[13];
    const Obj pre = e->carrying;
    int red_before = 0, blue_before = 0;
    if (e->cfg.hook == 8 && action == A_PICKUP) action = A_TOGGLE;                    /* memory.py:89-90 */
    if (e->cfg.hook == 7) { red_before = door_is_open(e, Ax, Ay); blue_before = door_is_open(e, Bx, By); }
    int rc = base_step(e, action, reward, done);
    switch (e->cfg.hook) {
    case 1:   /* unlockpickup.py:34-42, blockedunlockpickup.py:38-46: carrying == self.obj */
        if (action == A_PICKUP && e->carrying.has && e->carrying.type == ttype && e->carrying.color == tcol) { *reward = reward_fn(e); *done = 1; }
        break;
    case 2:   /* unlock.py:33-41 */
        if (action == A_TOGGLE && door_is_open(e, Ax, Ay)) { *reward = reward_fn(e); *done = 1; }
        break;
    case 3:   /* fetch.py:74-86 */
        if (e->carrying.has) {
            if (e->carrying.color == tcol && e->carrying.type == ttype) { *reward = reward_fn(e); *done = 1; }
            else { *reward = 0; *done = 1; }
        }
        break;
    case 4:   /* gotodoor.py:72-93 */
        if (action == A_DONE) {
            if (adj4(e, tx, ty)) *reward = reward_fn(e);
            if (adj4(e, Ax, Ay) || adj4(e, Bx, By) || adj4(e, Cx, Cy) || adj4(e, Dx, Dy)) *done = 1;
        }
        break;
    case 5:   /* gotoobject.py:68-84 */
        if (action == A_TOGGLE) *done = 1;
        if (action == A_DONE) { if (abs(e->ax - tx) <= 1 && abs(e->ay - ty) <= 1) *reward = reward_fn(e); *done = 1; }
        break;
    case 6: { /* putnear.py:91-112 */
        const int ox = e->ax + DIRX[e->adir], oy = e->ay + DIRY[e->adir];
        if (action == A_PICKUP && e->carrying.has)
            if (e->carrying.type != mtype || e->carrying.color != mcol) *done = 1;
        if (action == A_DROP && pre.has) {
            Obj f = grid_get(&e->grid, ox, oy);
            /* `self.grid.get(ox, oy) is preCarrying`: the drop happened this step */
            if (!e->carrying.has && f.has && f.type == pre.type && f.color == pre.color)
                if (abs(ox - tx) <= 1 && abs(oy - ty) <= 1) *reward = reward_fn(e);
            *done = 1;
        }
        break;
    }
    case 7: { /* redbluedoors.py:44-66 */
        const int red_after = door_is_open(e, Ax, Ay), blue_after = door_is_open(e, Bx, By);
        if (blue_after) { if (red_before) { *reward = reward_fn(e); *done = 1; } else { *reward = 0; *done = 1; } }
        else if (red_after) { if (blue_before) { *reward = 0; *done = 1; } }
        break;
    }
    case 8:   /* memory.py:88-100 */
        if (e->ax == Ax && e->ay == Ay) { *reward = reward_fn(e); *done = 1; }
        if (e->ax == Bx && e->ay == By) { *reward = 0; *done = 1; }
        break;
    }
    return rc;
}

/* env.step including the subclass hooks */
static int env_step(Env *e, int action, double *reward, int *done) {
    int rc;
    if (e->cfg.gen == ORC_GEN_DYNOBS) {             /* envs/dynamicobstacles.py:60-89 */
        Grid *g = &e->grid;
        if (action >= e->cfg.n_actions) action = 0;
        Obj front = grid_get(g, e->ax + DIRX[e->adir], e->ay + DIRY[e->adir]);
        int not_clear = front.has && front.type != T_GOAL;
        for (int k = 0; k < e->cfg.n_obstacles; k++) {
            int ox = e->obst[k][0], oy = e->obst[k][1];
            Obj ball = grid_get(g, ox, oy);
            int nx, ny;
            if (place_obj(e, ball, 1, ox - 1, oy - 1, 1, 3, 3, 0, 100, &nx, &ny) == 0) {
                e->obst[k][0] = nx; e->obst[k][1] = ny;
                grid_set(g, ox, oy, NONE);
            }
        }
        rc = base_step(e, action, reward, done);
        if (action == A_FORWARD && not_clear) { *reward = -1; *done = 1; }
        return rc;
    }
    if (e->cfg.gen == ORC_GEN_POOL && e->cfg.hook) return pool_hook_step(e, action, reward, done);
    rc = base_step(e, action, reward, done);
    if (e->cfg.gen == ORC_GEN_KEYCORRIDOR) {        /* envs/keycorridor.py:51-59 */
        if (action == A_PICKUP)
            if (e->carrying.has && e->carrying.is_target) { *reward = reward_fn(e); *done = 1; }
    }
    return rc;
}

/* ---- vector front-end --------------------------------------------------- */

static void gen_pool(Env *e) {
    const struct orc_vec *v = e->owner;
    if (!v || v->pool_n < 1) { e->err |= 64; return; }
    const int lvl = rand_int(e, 0, v->pool_n);
    const size_t cells = (size_t)e->grid.w * e->grid.h;
    memcpy(e->grid.c, v->pool_cells + (size_t)lvl * cells, cells * sizeof(Obj));
    e->ax = v->pool_agent[lvl * 3]; e->ay = v->pool_agent[lvl * 3 + 1]; e->adir = v->pool_agent[lvl * 3 + 2];
    e->has_agent = 1;
    e->level = lvl;
}

int orc_vec_set_level_pool(orc_vec *v, int32_t n_levels, const uint8_t *grid, const uint8_t *aux, const int32_t *agent,
                           const int32_t *hook_params) {
    const int W = v->cfg.width, H = v->cfg.height;
    const size_t cells = (size_t)W * H;
    free(v->pool_cells); free(v->pool_agent); free(v->pool_hook);
    v->pool_hook = NULL;
    if (hook_params) { v->pool_hook = (int *)calloc((size_t)n_levels * 16, sizeof(int)); for (int i = 0; i < n_levels * 16; i++) v->pool_hook[i] = hook_params[i]; }
    v->pool_cells = (Obj *)calloc((size_t)n_levels * cells, sizeof(Obj));
    v->pool_agent = (int *)calloc((size_t)n_levels * 3, sizeof(int));
    v->pool_n = n_levels;
    for (int l = 0; l < n_levels; l++) {
        for (int i = 0; i < W; i++)
            for (int j = 0; j < H; j++) {
                const size_t ci = ((size_t)l * W + i) * H + j;
                Obj o;
                if (obj_decode(grid[ci * 3], grid[ci * 3 + 1], grid[ci * 3 + 2], &o)) { snprintf(g_err, sizeof g_err, "orc_vec_set_level_pool: bad cell code"); return -1; }
                if (o.has && o.type == T_GOAL) { o.color = grid[ci * 3 + 1]; if (aux && (aux[ci] & 1)) { o.toggletimes = 0; o.overlap = 1; } }
                if (o.has && o.type == T_BOX && aux) o.contains_key = (aux[ci] >> 1) & 7;
                v->pool_cells[(size_t)l * cells + (size_t)j * W + i] = o;
            }
        for (int k = 0; k < 3; k++) v->pool_agent[l * 3 + k] = agent[l * 3 + k];
    }
    return 0;
}

orc_vec *orc_vec_create(const orc_config *cfg, uint64_t seed, int64_t env0, int32_t n) {
    if (cfg->view_size != 0 && (cfg->view_size < 3 || cfg->view_size > VIEW || cfg->view_size % 2 == 0)) {
        snprintf(g_err, sizeof g_err, "orc_vec_create: view_size must be odd, 3..11"); return NULL;
    }
    if (cfg->width < 3 || cfg->height < 3 || cfg->width > 64 || cfg->height > 64 || n < 0 ||
        cfg->n_obstacles > ORC_MAX_OBST || cfg->num_rows > MAXR) {
        snprintf(g_err, sizeof g_err, "orc_vec_create: bad config"); return NULL;
    }
    orc_vec *v = (orc_vec *)calloc(1, sizeof(*v));
    v->cfg = *cfg; v->n = n;
    if (v->cfg.view_size == 0) v->cfg.view_size = 7;
    v->envs = (Env *)calloc((size_t)(n > 0 ? n : 1), sizeof(Env));
    size_t cells = (size_t)cfg->width * cfg->height;
    v->cells = (Obj *)calloc((size_t)(n > 0 ? n : 1) * cells, sizeof(Obj));
    for (int i = 0; i < n; i++) {
        Env *e = &v->envs[i];
        e->cfg = v->cfg; e->grid.w = cfg->width; e->grid.h = cfg->height; e->grid.c = v->cells + (size_t)i * cells;
        e->env_id = env0 + i; e->seed = seed; e->owner = v;
    }
    return v;
}
void orc_vec_destroy(orc_vec *v) { if (!v) return; free(v->envs); free(v->cells); free(v->pool_cells); free(v->pool_agent); free(v->pool_hook); free(v); }
int orc_vec_get_levels(orc_vec *v, int32_t *levels) { for (int i = 0; i < v->n; i++) levels[i] = v->envs[i].level; return 0; }
int orc_vec_set_levels(orc_vec *v, const int32_t *levels) { for (int i = 0; i < v->n; i++) v->envs[i].level = levels[i]; return 0; }

int orc_vec_set_tape(orc_vec *v, const int32_t *draws, const int64_t *offsets) {
    for (int i = 0; i < v->n; i++) {
        Env *e = &v->envs[i];
        if (!draws) { e->tape = NULL; continue; }
        e->tape = draws + offsets[i]; e->tape_len = offsets[i + 1] - offsets[i]; e->ndraws = 0;
    }
    return 0;
}

/* tiny static-chunk parallel-for over envs (pthreads; no OpenMP runtime needed) */
typedef int (*range_fn)(void *ctx, int lo, int hi);
typedef struct { range_fn fn; void *ctx; int lo, hi, rc; } Job;
static void *job_main(void *p) { Job *j = (Job *)p; j->rc = j->fn(j->ctx, j->lo, j->hi); return NULL; }
static int nthreads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static int parallel_for(int n, range_fn fn, void *ctx) {
    int nt = nthreads();
    if (nt > n / 16) nt = n / 16;
    if (nt <= 1) return fn(ctx, 0, n);
    if (nt > 256) nt = 256;
    pthread_t th[256]; Job jobs[256];
    int rc = 0;
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].rc = 0;
        jobs[t].lo = (int)((int64_t)n * t / nt); jobs[t].hi = (int)((int64_t)n * (t + 1) / nt);
        if (t + 1 < nt && pthread_create(&th[t], NULL, job_main, &jobs[t]) != 0) {
            jobs[t].rc = fn(ctx, jobs[t].lo, jobs[t].hi); th[t] = 0;
        }
    }
    jobs[nt - 1].rc = fn(ctx, jobs[nt - 1].lo, jobs[nt - 1].hi);
    for (int t = 0; t < nt; t++) { if (t + 1 < nt && th[t]) pthread_join(th[t], NULL); rc |= jobs[t].rc; }
    return rc;
}

typedef struct { orc_vec *v; const uint8_t *mask; uint8_t *obs; uint8_t *dir; } ResetCtx;
static int reset_range(void *p, int lo, int hi) {
    ResetCtx *c = (ResetCtx *)p; orc_vec *v = c->v;
    int bad = 0;
    for (int i = lo; i < hi; i++) {
        if (c->mask && !c->mask[i]) continue;
        Env *e = &v->envs[i];
        g_oob = 0;
        env_reset(e, c->obs ? c->obs + (size_t)i * (3 * v->cfg.view_size * v->cfg.view_size) : NULL, c->dir ? c->dir + i : NULL);
        bad |= e->err | (g_oob ? 16 : 0);
    }
    return bad;
}
int orc_vec_reset(orc_vec *v, const uint8_t *mask, uint8_t *obs, uint8_t *dir) {
    ResetCtx c = { v, mask, obs, dir };
    int bad = parallel_for(v->n, reset_range, &c);
    if (bad) { snprintf(g_err, sizeof g_err, "orc_vec_reset: error flags 0x%x", bad); return -1; }
    return 0;
}

/* The uniform random policy of run_tests.py:43 / benchmark.py (`env.action_space.sample()`), as a counter-based stream
 * so that the device can draw the same actions without a host round trip: action of step t of the `epoch`-th random
 * rollout of a vector, for env id g = mulhi32(Philox4x32-10(ctr = (t>>2, epoch, g), key = (seed lo, seed hi ^ ORC_ACTION_KEY))[t&3],
 * n_actions).  The key differs from the env's own stream, so the policy never correlates with the layout draws. */
int orc_policy_action(uint64_t seed, int64_t env_id, uint32_t epoch, uint32_t t, int n_actions) {
    uint32_t ctr[4] = { t >> 2, epoch, (uint32_t)env_id, (uint32_t)((uint64_t)env_id >> 32) };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) ^ ORC_ACTION_KEY };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return (int)(((uint64_t)out[t & 3] * (uint32_t)n_actions) >> 32);
}

typedef struct { orc_vec *v; int T; const uint8_t *actions; int autoreset;
                 uint8_t *obs; double *reward; uint8_t *done; uint8_t *dir; } RollCtx;
static int rollout_range(void *p, int lo, int hi) {
    RollCtx *c = (RollCtx *)p; orc_vec *v = c->v;
    const int T = c->T, autoreset = c->autoreset;
    const uint8_t *actions = c->actions; uint8_t *obs = c->obs; double *reward = c->reward;
    uint8_t *done = c->done, *dir = c->dir;
    int bad = 0;
    const size_t n = (size_t)v->n;
    for (int i = lo; i < hi; i++) {
        Env *e = &v->envs[i];
        g_oob = 0;
        uint8_t scratch[3 * VIEW * VIEW];
        const size_t ob = (size_t)3 * v->cfg.view_size * v->cfg.view_size;
        for (int t = 0; t < T; t++) {
            size_t o = (size_t)t * n + (size_t)i;
            double r; int d; uint8_t dd;
            if (env_step(e, actions[o], &r, &d)) bad |= 1;
            uint8_t *op = obs ? obs + o * ob : scratch;
            if (d && autoreset) env_reset(e, op, &dd);
            else gen_obs(e, op, &dd);
            if (reward) reward[o] = r;
            if (done) done[o] = (uint8_t)d;
            if (dir) dir[o] = dd;
        }
        bad |= e->err | (g_oob ? 16 : 0);
    }
    return bad;
}
int orc_vec_rollout(orc_vec *v, int32_t T, const uint8_t *actions, int autoreset,
                    uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir) {
    if (!actions) { snprintf(g_err, sizeof g_err, "orc_vec_rollout: actions is NULL"); return -1; }
    RollCtx c = { v, T, actions, autoreset, obs, reward, done, dir };
    int bad = parallel_for(v->n, rollout_range, &c);
    if (bad) { snprintf(g_err, sizeof g_err, "orc_vec_rollout: error flags 0x%x (1=unknown action, 2=tape exhausted, 4=tape value out of range, 8=rejection sampling failed, 16=grid index out of bounds)", bad); return -1; }
    return 0;
}

int orc_vec_rollout_random(orc_vec *v, int32_t T, int autoreset, uint8_t *actions_out,
                           uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir) {
    if (!actions_out) { snprintf(g_err, sizeof g_err, "orc_vec_rollout_random: actions_out is NULL"); return -1; }
    for (int t = 0; t < T; t++)
        for (int i = 0; i < v->n; i++)
            actions_out[(size_t)t * v->n + i] = (uint8_t)orc_policy_action(v->envs[i].seed, v->envs[i].env_id, v->policy_epoch, (uint32_t)t, v->cfg.n_actions);
    v->policy_epoch++;
    return orc_vec_rollout(v, T, actions_out, autoreset, obs, reward, done, dir);
}

int orc_vec_step(orc_vec *v, const uint8_t *actions, int autoreset,
                 uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir) {
    return orc_vec_rollout(v, 1, actions, autoreset, obs, reward, done, dir);
}

int orc_vec_get_state(orc_vec *v, uint8_t *grid, uint8_t *aux, int32_t *agent,
                      uint8_t *carrying, int16_t *obstacles, uint8_t *target, uint32_t *rng) {
    const int W = v->cfg.width, H = v->cfg.height;
    for (int n = 0; n < v->n; n++) {
        Env *e = &v->envs[n];
        for (int i = 0; i < W; i++)
            for (int j = 0; j < H; j++) {
                Obj o = e->grid.c[j * W + i];
                size_t ci = ((size_t)n * W + i) * H + j;
                if (grid) {
                    uint8_t *p = grid + ci * 3;
                    if (!o.has) { p[0] = T_EMPTY; p[1] = 0; p[2] = 0; } else obj_encode(&o, p);
                }
                if (aux) aux[ci] = (uint8_t)(((o.has && o.type == T_GOAL && o.overlap) ? 1 : 0) | ((o.has && o.type == T_BOX) ? (o.contains_key << 1) : 0));
            }
        if (agent) { agent[n * 4] = e->ax; agent[n * 4 + 1] = e->ay; agent[n * 4 + 2] = e->adir; agent[n * 4 + 3] = e->step_count; }
        if (carrying) {
            uint8_t *p = carrying + n * 3; p[0] = p[1] = p[2] = 0;
            if (e->carrying.has) { obj_encode(&e->carrying, p); if (e->carrying.type == T_BOX) p[2] = (uint8_t)(e->carrying.contains_key << 1); }
        }
        if (obstacles) for (int k = 0; k < ORC_MAX_OBST; k++) {
            obstacles[(n * ORC_MAX_OBST + k) * 2] = (int16_t)(k < v->cfg.n_obstacles ? e->obst[k][0] : 0);
            obstacles[(n * ORC_MAX_OBST + k) * 2 + 1] = (int16_t)(k < v->cfg.n_obstacles ? e->obst[k][1] : 0);
        }
        if (target) { target[n * 2] = e->target_type; target[n * 2 + 1] = e->target_color; }
        if (rng) { rng[n * 2] = e->episode; rng[n * 2 + 1] = e->ndraws; }
    }
    return 0;
}

int orc_vec_set_state(orc_vec *v, const uint8_t *grid, const uint8_t *aux, const int32_t *agent,
                      const uint8_t *carrying, const int16_t *obstacles, const uint8_t *target,
                      const uint32_t *rng) {
    const int W = v->cfg.width, H = v->cfg.height;
    for (int n = 0; n < v->n; n++) {
        Env *e = &v->envs[n];
        if (target) { e->target_type = target[n * 2]; e->target_color = target[n * 2 + 1]; }
        for (int i = 0; i < W; i++)
            for (int j = 0; j < H; j++) {
                size_t ci = ((size_t)n * W + i) * H + j;
                Obj o;
                if (obj_decode(grid[ci * 3], grid[ci * 3 + 1], grid[ci * 3 + 2], &o)) {
                    snprintf(g_err, sizeof g_err, "orc_vec_set_state: bad cell code"); return -1;
                }
                if (o.has && o.type == T_GOAL) { o.color = grid[ci * 3 + 1]; if (aux && (aux[ci] & 1)) { o.toggletimes = 0; o.overlap = 1; } }
                if (o.has && o.type == T_BOX && aux) o.contains_key = (aux[ci] >> 1) & 7;
                if (o.has && e->target_type && o.type == e->target_type && o.color == e->target_color) o.is_target = 1;
                e->grid.c[j * W + i] = o;
            }
        e->ax = agent[n * 4]; e->ay = agent[n * 4 + 1]; e->adir = agent[n * 4 + 2]; e->step_count = agent[n * 4 + 3];
        e->has_agent = 1;
        e->carrying = NONE;
        if (carrying && carrying[n * 3]) {
            if (obj_decode(carrying[n * 3], carrying[n * 3 + 1], carrying[n * 3] == T_BOX ? 0 : carrying[n * 3 + 2], &e->carrying)) return -1;
            if (e->carrying.type == T_BOX) e->carrying.contains_key = (carrying[n * 3 + 2] >> 1) & 7;
            if (e->target_type && e->carrying.type == e->target_type && e->carrying.color == e->target_color) e->carrying.is_target = 1;
        }
        if (obstacles) for (int k = 0; k < v->cfg.n_obstacles; k++) {
            e->obst[k][0] = obstacles[(n * ORC_MAX_OBST + k) * 2];
            e->obst[k][1] = obstacles[(n * ORC_MAX_OBST + k) * 2 + 1];
        }
        if (rng) { e->episode = rng[n * 2]; e->ndraws = rng[n * 2 + 1]; }
    }
    return 0;
}

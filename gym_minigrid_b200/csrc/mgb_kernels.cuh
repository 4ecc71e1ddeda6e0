// mgb200 device code: the batched MiniGrid hot path for sm_100a.
//
// Mapping: one THREAD owns one environment, one WARP owns a "group" of 32
// environments and never talks to another warp.  A group's state lives in HBM as a
// lane-interleaved block  state[group][word k][lane]  so that
//   * loading/storing it is a run of perfectly coalesced 128-byte rows, and
//   * once in shared memory every lane only ever touches bank == lane: the per-env random
//     grid accesses of the view gather are bank-conflict free by construction.
// Grid cells are 1-byte codes (type*21 + colour*3 + state, 231 combinations; 231..237 =
// terminal goal of colour c), four per word.  A 256-entry shared-memory LUT expands a code to
// (type | colour<<8 | state<<16 | flags<<24).
//
// Per step a thread applies the transition (minigrid.py:1227-1325 + subclass hooks), gathers
// its 7x7 view with the closed form of slice+rotate_left, runs process_vis as
// carry-propagation floods on 7-bit row masks, packs the 147 output bytes into 37
// registers, realigns them with funnel shifts + one warp shuffle into a contiguous 4704-byte
// shared-memory block per warp, and ONE lane hands that block to the TMA engine
// (cp.async.bulk.global.shared::cta) -- the observation stream, which is ~93% of the
// algorithmic bytes, leaves the SM as full-line bulk writes without occupying the LSU.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mgb {

constexpr int VIEW = 7;                           // default agent_view_size
constexpr int OBS_BYTES = 147;
constexpr int GROUP = 32;
constexpr int STAGE_BYTES = GROUP * OBS_BYTES;   // 4704 = 294 * 16
// the kernel is a template on the (odd) view size V: record = 3*V*V bytes, staging block = 32 records
__host__ __device__ constexpr int obs_bytes(int V) { return 3 * V * V; }
// Staging block of a warp: the 32 records of a step, contiguous here and in global memory, stored with one bulk copy.
// Multiple of 16 bytes for every V; at least 1024: the block also parks 32 rows x 32 bytes of actions between two observations.
__host__ __device__ constexpr int stage_bytes(int V) { return GROUP * 3 * V * V < 1024 ? 1024 : GROUP * 3 * V * V; }
constexpr int MAX_WARPS_PER_BLOCK = 8;           // the host picks 2..8 warps per CTA to maximise resident warps/SM
constexpr int MAX_THREADS = MAX_WARPS_PER_BLOCK * 32;
constexpr int MAX_OBST = 8;
constexpr int XWORDS = 4;                        // agent, steps/target, episode, ndraws
constexpr int OBST_WORDS = 4;                    // 8 x (x,y) bytes
// CTA-shared tables at the start of dynamic shared memory:
//   LUT: 256 entries of one word, code -> type | colour<<8 | state<<16 | flags<<24.  The three low bytes are the cell's
//   output bytes (no PRMT of the pack ever selects byte 3); "opaque" is bit 7 of the flags, i.e. bit 31 of the word, which
//   is what the occluded path shifts into its row masks; the transition reads the flags.  Entry pitch: 12 bytes -- the
//   address code*12+base is then an immediate-form IMAD on the otherwise idle FMA pipe, where a 4-byte pitch becomes an
//   LEA on the ALU pipe that bounds these kernels (and IMAD with the pitch in a uniform register measured 1.3 % slower
//   than the immediate form) -- except for the Dynamic-Obstacles kernels, whose resident warps are limited by shared
//   memory: there the 2 KB saved by a 4-byte pitch buy a 16th warp per SM in two evenly filled 8-warp CTAs.  Any 32
//   consecutive codes map to 32 different banks at either pitch.
//   AXIS tables: shared-memory offset of grid coordinate v (index v + AXIS_BIAS) along x and along y, out-of-grid
//   entries = offset of the wall pad word (see observe()).
__host__ __device__ constexpr int lut_pitch_words(int gen) { return gen == 3 /* GEN_DYNOBS */ ? 1 : 3; }
__host__ __device__ constexpr int lut_bytes(int gen) { return 256 * 4 * lut_pitch_words(gen); }
constexpr int AXIS_ENTRIES = 88;                                 // v in [-(V-1), 64+V-2] for V <= 11: index v + AXIS_BIAS
constexpr int AXIS_BIAS = 10;
constexpr int MBAR_BYTES = MAX_WARPS_PER_BLOCK * 8;               // one mbarrier per warp (bulk load of the state block)
__host__ __device__ constexpr int table_bytes(int gen) { return (lut_bytes(gen) + 2 * AXIS_ENTRIES * 4 + MBAR_BYTES + 127) / 128 * 128; }   // 3840 / 1792
// Every kernel with a device generator keeps a copy of the static layout behind the tables: a reset copies it into the
// env's column from shared memory.  (From global memory that copy was GW dependent L2 round trips -- with episode ends
// spread over time, i.e. one resetting lane per warp, 10-30 k cycles per reset: Empty-8x8 -24 %, FourRooms -28 %.)
__host__ __device__ inline int tmpl_smem_bytes(int gen, int GW) { return gen == 5 /* GEN_POOL: no template */ ? 0 : (GW * 4 + 127) / 128 * 128; }

// minigrid.py:40-52 / 27-35 / 57-61
enum : int { T_UNSEEN = 0, T_EMPTY = 1, T_WALL = 2, T_FLOOR = 3, T_DOOR = 4, T_KEY = 5, T_BALL = 6,
             T_BOX = 7, T_GOAL = 8, T_LAVA = 9, T_AGENT = 10 };
enum : int { C_RED = 0, C_GREEN = 1, C_BLUE = 2, C_PURPLE = 3, C_YELLOW = 4, C_GREY = 5, C_WHITE = 6 };
enum : int { A_LEFT = 0, A_RIGHT = 1, A_FORWARD = 2, A_PICKUP = 3, A_DROP = 4, A_TOGGLE = 5, A_DONE = 6 };
enum : int { GEN_EMPTY = 0, GEN_DOORKEY = 1, GEN_FOURROOMS = 2, GEN_DYNOBS = 3, GEN_KEYCORRIDOR = 4, GEN_POOL = 5,
             // kernel template value for the three procedural generators with the base step (the hot loop is the same;
             // generate<GEN_PROC> switches on DevCfg::gen, which keeps the precise id below)
             GEN_PROC = 6,
             GEN_CROSSING = 6, GEN_LAVAGAP = 7, GEN_MULTIROOM = 8 };
enum : uint32_t { ERR_ACTION = 1, ERR_TAPE_END = 2, ERR_TAPE_RANGE = 4, ERR_SAMPLING = 8, ERR_BOUNDS = 16, ERR_CODE = 32, ERR_NO_POOL = 64 };

__host__ __device__ constexpr int code_of(int t, int c, int s) { return t * 21 + c * 3 + s; }
constexpr int CODE_EMPTY = code_of(T_EMPTY, 0, 0);            // 21
constexpr int CODE_WALL = code_of(T_WALL, C_GREY, 0);          // out-of-grid cells (minigrid.py:469)
constexpr int CODE_GOAL = code_of(T_GOAL, C_GREEN, 0);
constexpr int CODE_TGOAL0 = 231;                               // + colour: Goal(toggletimes=0), overlap=True
constexpr int CODE_KEYBOX0 = 238;                              // + key colour: grey Box(contains=Key(colour)) (obstructedmaze.py:68-73)
constexpr uint32_t EMPTY_WORD = 0x15151515u;                   // 4 x CODE_EMPTY
static_assert(CODE_EMPTY == 0x15, "EMPTY_WORD");

enum : int { HOOK_NONE = 0, HOOK_PICKUP_TARGET = 1, HOOK_UNLOCK = 2, HOOK_FETCH = 3, HOOK_GOTODOOR = 4, HOOK_GOTOOBJECT = 5,
             HOOK_PUTNEAR = 6, HOOK_REDBLUEDOORS = 7, HOOK_MEMORY = 8 };
constexpr int POOL_XW = 5;     // pool record = GW grid words + agent word + 4 hook-parameter words
enum : uint32_t { F_OPAQUE = 0x80, F_OVERLAP = 2, F_PICKUP = 4, F_TGOAL = 8, F_LAVA = 16 };   // F_OPAQUE = bit 31 of the LUT word

// code -> type | colour<<8 | state<<16 | flags<<24   (WorldObj.encode + the predicates
// can_overlap / can_pickup / see_behind, minigrid.py:93-115,164-166,192-193,211-212,233-250,305-343)
__host__ __device__ inline uint32_t lut_entry(int code) {
    int t, c, s;
    uint32_t f = 0;
    if (code >= CODE_KEYBOX0) {
        if (code >= CODE_KEYBOX0 + 7) return 0;
        t = T_BOX; c = C_GREY; s = 0;                    // looks like any grey box; the key appears when it is toggled
    } else if (code >= CODE_TGOAL0) {
        t = T_GOAL; c = code - CODE_TGOAL0; s = 0; f = F_TGOAL;
    } else {
        t = code / 21; c = (code % 21) / 3; s = code % 3;
    }
    if (t == T_WALL) f |= F_OPAQUE;
    if (t == T_DOOR) f |= (s != 0) ? F_OPAQUE : F_OVERLAP;
    if (t == T_EMPTY || t == T_UNSEEN || t == T_FLOOR || t == T_GOAL || t == T_LAVA) f |= F_OVERLAP;
    if (t == T_LAVA) f |= F_LAVA;
    if (t == T_KEY || t == T_BALL || t == T_BOX) f |= F_PICKUP;
    return (uint32_t)t | ((uint32_t)c << 8) | ((uint32_t)s << 16) | (f << 24);
}

struct DevCfg {
    int32_t gen, W, H, max_steps, see_through, n_actions, n_obst, room_size, num_rows, random_start, lava_v1, hook;
    int32_t gp0, gp1;   // generator parameters (mgb_config.gen_param0/1)
    int32_t goal_idx;   // Empty / DistShift: cell index x*HP + y of the template's goal (FLAG_GOAL_GONE), else -1
    int32_t HP;   // grid column pitch in cells: H rounded up to a multiple of 4 (one column = HP/4 words)
    int32_t GW;   // grid words per env = W*HP/4; cell (x,y) = byte (y&3) of word x*HP/4 + (y>>2)
    int32_t S;    // state words per env
};

struct RolloutParams {
    DevCfg cfg;
    uint32_t *state;            // [n_groups][S][32]
    uint32_t *spare;            // [n_groups][spare_words(GW)][32]: pre-generated next layouts (spare_gen kernels), or NULL
    const uint32_t *tmpl;       // [GW] static part of the layout (walls, fixed goal)
    int64_t n_envs;
    int32_t group0, n_groups;   // group sub-range handled by this launch
    int32_t T;                  // steps (0 = reset/observe only)
    int32_t do_reset;
    int32_t autoreset;
    const uint8_t *reset_mask;
    const uint8_t *actions;     // [T][stride]
    uint8_t *obs;               // [T][stride][147]
    double *reward;
    uint8_t *done;
    uint8_t *dir;
    int64_t stride;             // elements between step t and t+1 in the outputs
    uint64_t seed;
    int64_t env_id_base;
    const int32_t *tape;
    const int64_t *tape_off;
    const uint32_t *pool;       // GEN_POOL: [pool_n][GW + POOL_XW] words: grid, x|y<<8|dir<<16, then 4 hook words:
                                //   tcode | mcode<<8 ; tx|ty<<8|A.x<<16|A.y<<24 ; B.x|B.y<<8|C.x<<16|C.y<<24 ; D.x|D.y<<8
    int32_t pool_n;
    uint32_t *err;
    uint32_t *ticket;           // [2] {next ticket, warps done}: groups beyond a warp's first are handed out by the counter; the last
                                // warp to leave zeroes both words for the next launch (NULL: fixed stride)
};

// ------------------------------------------------------------------------------------------
// per-thread environment context (registers) + shared-memory column
// ------------------------------------------------------------------------------------------
struct Env {                    // hot: stays in registers (its address never escapes)
    int ax, ay, dir, carry, steps, target;
    int flags;                  // bit0 "pristine": the grid equals the static template plus a blue ball at every recorded
                                // obstacle position.  Set by the Empty / Dynamic-Obstacles generators, cleared by a grid
                                // edit or an upload.  The grid rows of a pristine env in HBM are DON'T-CARE: loads rebuild
                                // the grid from the template (L2-resident) and the ball list, write-backs skip it, and the
                                // state readers (k_get_state, k_render_full) reconstruct it the same way.
    bool dirty;                 // grid words modified since load
};
struct PoolCtx {                // GEN_POOL kernels only: the level being played and its hook parameters
    int level;
    uint32_t hp0, hp1, hp2, hp3;
};
struct Rng {                    // cold: passed by reference to the out-of-line draw routine
    uint32_t episode, ndraws;
    uint32_t rb0, rb1, rb2, rb3, rblk;   // cached Philox block
    uint32_t err;
    int64_t gid;                // global env id (Philox counter)
    int64_t lid;                // env index inside this handle (tape offsets)
};

__device__ __forceinline__ uint32_t cell_rd(const uint32_t *st, int idx) {
    return __byte_perm(st[(idx >> 2) * 32], 0, 0x4440 | (idx & 3));
}
__device__ __forceinline__ void cell_wr(uint32_t *st, int idx, uint32_t code) {
    uint32_t *p = &st[(idx >> 2) * 32];
    const int sh = (idx & 3) * 8;
    *p = (*p & ~(0xFFu << sh)) | (code << sh);
}

// Philox4x32-10, Salmon et al. SC'11
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t &o0, uint32_t &o1,
                                              uint32_t &o2, uint32_t &o3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

// The env's random stream ("RNG"): Philox4x32-10, key = seed, counter = (block, episode, global env id).
// Every 32-bit output word serves TWO consecutive draws: draw n takes word (n>>1)&3 of block n>>3, as it is for even n
// and multiplied by DRAW_ODD_MULT (2^32 / golden ratio, odd: a bijection of the 32-bit words) for odd n.  The pair
// (w, w*M) is the 2-D lattice of a multiplicative congruential generator with a good multiplier, so two consecutive
// bounded draws (the x and the y of a placement try) are jointly uniform far below the resolution any span here needs
// (span_x * span_y <= 2^12 against 2^32 lattice points; tests/test_host_logic.py checks the joint histogram).  This
// halves the Philox work of Dynamic-Obstacles, whose rejection sampling draws ~21 numbers per env-step.
constexpr uint32_t DRAW_ODD_MULT = 0x9E3779B1u;
__device__ __forceinline__ uint32_t draw_word(uint32_t w, uint32_t n) { return (n & 1u) ? w * DRAW_ODD_MULT : w; }

// One Philox block of a stream, out of line and PURE (arguments and result in registers, no memory): the layout
// generators call it through rand_int_inl below.
__device__ __noinline__ uint4 philox_block(uint32_t blk, uint32_t stream, uint32_t gid_lo, uint32_t gid_hi, uint32_t seed_lo, uint32_t seed_hi) {
    uint4 o;
    philox4x32_10(blk, stream, gid_lo, gid_hi, seed_lo, seed_hi, o.x, o.y, o.z, o.w);
    return o;
}
// the tape entry of RNG-tape mode (parity against the reference's own MT19937 draws), out of line
__device__ __noinline__ int tape_draw(const RolloutParams &p, int64_t lid, uint32_t &ndraws, uint32_t &err, int low, int high) {
    const int64_t off = p.tape_off[lid], len = p.tape_off[lid + 1] - off;
    if ((int64_t)ndraws >= len) { err |= ERR_TAPE_END; return low; }
    const int v = p.tape[off + ndraws++];
    if (v < low || v >= high) err |= ERR_TAPE_RANGE;
    return v;
}

// MiniGridEnv._rand_int (minigrid.py:939-944): low + mulhi32(u32, high-low) on the stream
// (seed, global env id, episode); or the next tape entry in RNG-tape mode.
// The Rng it works on must be a LOCAL of the caller (registers): the generators used to reach their Rng through a
// reference across an out-of-line rand_int -- half a dozen local-memory accesses per draw, and with ~220 KB of each SM
// carved out as shared memory there is next to no L1 left, so each was an L2 round trip (KeyCorridor: 11 ms to
// regenerate 2^20 layouts, 9 launches' worth of stepping).
__device__ __forceinline__ int rand_int_inl(Rng &e, const RolloutParams &p, int low, int high) {
    if (p.tape) {
        uint32_t nd = e.ndraws, er = e.err;
        const int v = tape_draw(p, e.lid, nd, er, low, high);
        e.ndraws = nd; e.err = er;
        return v;
    }
    const uint32_t blk = e.ndraws >> 3;
    if (blk != e.rblk) {
        const uint4 o = philox_block(blk, e.episode - 1u, (uint32_t)e.gid, (uint32_t)((uint64_t)e.gid >> 32), (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
        e.rb0 = o.x; e.rb1 = o.y; e.rb2 = o.z; e.rb3 = o.w;
        e.rblk = blk;
    }
    const uint32_t sel = (e.ndraws >> 1) & 3;
    const uint32_t u = draw_word(sel == 0 ? e.rb0 : sel == 1 ? e.rb1 : sel == 2 ? e.rb2 : e.rb3, e.ndraws);
    e.ndraws++;
    return low + (int)__umulhi(u, (uint32_t)(high - low));
}
__device__ __forceinline__ int rand_int(Rng &e, const RolloutParams &p, int low, int high) { return rand_int_inl(e, p, low, high); }
// out-of-line copy working through the reference: only for the Empty kernels' random-start generator, whose (headline)
// hot loop is sensitive to how much code and register pressure the cold reset path carries (measured: -1.1 % with the
// inline form there); its three draws per reset do not matter
__device__ __noinline__ int rand_int_ool(Rng &e, const RolloutParams &p, int low, int high) { return rand_int_inl(e, p, low, high); }

// Dynamic-Obstacles consumes 2 draws (one stream word) per ball try, ~11 tries per step.  Instead of computing a Philox
// block inside the (divergent) try loop, a lane pre-computes DYN_BLOCKS consecutive blocks of its stream in
// straight-line code (all lanes active, independent chains -> ILP) into its column of the warp's staging buffer,
// which is idle between two observations.  Word i of the window at draws[i*32]; 5 blocks = 20 words = 20 tries
// (P(a step needs more) ~ 1e-3 per env; those continue on demand).  Small views have a shorter staging column.
#ifndef MGB_DYN_BLOCKS
#define MGB_DYN_BLOCKS 5
#endif
constexpr int DYN_BLOCKS = MGB_DYN_BLOCKS;
__host__ __device__ constexpr int draw_blocks(int V) { return (3 * V * V / 4) / 4 < DYN_BLOCKS ? (3 * V * V / 4) / 4 : DYN_BLOCKS; }
template <int NB>
__device__ __noinline__ void prefetch_draws(uint32_t *draws, uint32_t first_block, uint32_t stream, int64_t gid, uint64_t seed) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        uint32_t o0, o1, o2, o3;
        philox4x32_10(first_block + j, stream, (uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1, o2, o3);
        draws[(4 * j + 0) * 32] = o0; draws[(4 * j + 1) * 32] = o1; draws[(4 * j + 2) * 32] = o2; draws[(4 * j + 3) * 32] = o3;
    }
}

__device__ __forceinline__ int ldg_u8(const uint8_t *g) {        // volatile asm: stays where it is written
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(g) : "memory");
    return (int)v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %1; st.shared.u16 [%0], t; }" ::"r"(a), "r"(v) : "memory");
}

constexpr int HARD_TRY_CAP = 1 << 16;   // the reference would spin forever; we flag ERR_SAMPLING

// MiniGridEnv.place_obj (minigrid.py:1003-1061).  max_tries < 0 == math.inf.
// check_agent: "don't place the object where the agent is" (agent_pos may be None -> false).
template <bool INL = false, bool OOL = false>
__device__ __forceinline__ bool place_obj(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p, int code, int topx, int topy,
                                          int sx, int sy, bool reject_next_to, int max_tries,
                                          bool check_agent, int &ox, int &oy) {
    const int W = p.cfg.W, H = p.cfg.H, HP = p.cfg.HP;
    topx = max(topx, 0); topy = max(topy, 0);
    const int hx = min(topx + sx, W), hy = min(topy + sy, H);
    int tries = 0, x, y;
    for (;;) {
        if ((max_tries >= 0 && tries > max_tries) || tries > HARD_TRY_CAP) return false;
        tries++;
        x = OOL ? rand_int_ool(rg, p, topx, hx) : rand_int(rg, p, topx, hx);
        y = OOL ? rand_int_ool(rg, p, topy, hy) : rand_int(rg, p, topy, hy);
        if (rg.err & (ERR_TAPE_END | ERR_TAPE_RANGE)) return false;
        if (cell_rd(st, x * HP + y) != CODE_EMPTY) continue;
        if (check_agent && x == e.ax && y == e.ay) continue;
        if (reject_next_to && (abs(e.ax - x) + abs(e.ay - y) < 2)) continue;   // roomgrid.py:3-12
        break;
    }
    if (code != CODE_EMPTY) { cell_wr(st, x * HP + y, (uint32_t)code); e.dirty = true; }
    ox = x; oy = y;
    return true;
}

// MiniGridEnv.place_agent (minigrid.py:1072-1090)
template <bool OOL = false>
__device__ __forceinline__ bool place_agent(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p, int topx, int topy, int sx,
                                            int sy, int max_tries) {
    int x, y;
    if (!place_obj<false, OOL>(st, e, rg, p, CODE_EMPTY, topx, topy, sx, sy, false, max_tries, false, x, y)) return false;
    e.ax = x; e.ay = y;
    e.dir = OOL ? rand_int_ool(rg, p, 0, 4) : rand_int(rg, p, 0, 4);
    return true;
}

// COLOR_NAMES = sorted(COLORS) = blue, green, grey, purple, red, white, yellow (minigrid.py:24)
__device__ __forceinline__ int rand_color(Rng &rg, const RolloutParams &p) {
    const int k = rand_int(rg, p, 0, 7);
    return (int)((0x4603512u >> (4 * k)) & 0xF);   // [2,1,5,3,0,6,4]
}

// obstacle k = byte pair (k&1)*2 of word GW+XWORDS+(k>>1)
__device__ __forceinline__ void obst_get(const uint32_t *st, const DevCfg &c, int k, int &x, int &y) {
    const uint32_t w = st[(c.GW + XWORDS + (k >> 1)) * 32] >> ((k & 1) * 16);
    x = w & 0xFF; y = (w >> 8) & 0xFF;
}
__device__ __forceinline__ void obst_set(uint32_t *st, const DevCfg &c, int k, int x, int y) {
    uint32_t *q = &st[(c.GW + XWORDS + (k >> 1)) * 32];
    const int sh = (k & 1) * 16;
    *q = (*q & ~(0xFFFFu << sh)) | ((uint32_t)(x | (y << 8)) << sh);
}

constexpr int FLAG_PRISTINE = 1;
constexpr int FLAG_GOAL_GONE = 16;       // Empty kernels, with FLAG_PRISTINE: ... except that the goal was toggled away (the one edit an
                                         // Empty grid can suffer, minigrid.py:171-177): still implied, still never crosses HBM
constexpr int FLAGS_IMPLIED = FLAG_PRISTINE | FLAG_GOAL_GONE;
constexpr int FLAG_SPARE = 2;            // the env's spare block holds the layout of its NEXT episode (see reset_lanes)
constexpr int LONE_SHIFT = 26;           // two bits of the flags byte: the warp's count of lone resets (reset_with_spares)
constexpr int SPARE_XW = 5;              // spare block = GW grid words + agent word, target, draws consumed, error bits, episode,
__host__ __device__ constexpr int spare_words(int GW) { return 2 * GW + SPARE_XW; }      // + GW words where a live grid is parked
// kernels whose generator is worth more than a trip to HBM (measured: for DoorKey and FourRooms it is not -- fetching a
// spare layout with one lane costs more than generating it)
#ifndef MGB_SPARES
#define MGB_SPARES 1
#endif
#ifndef MGB_DYNAMIC_GROUPS
#define MGB_DYNAMIC_GROUPS 1        // 0: every launch hands out groups by a fixed stride (the A/B of profiles/r2_ab_group_tickets.txt)
#endif
#ifndef MGB_PDL
#define MGB_PDL 1      // 0: plain launches (the A/B of profiles/r2_ab_pdl.txt)
#endif
#ifndef MGB_TICKET_LEAD
#define MGB_TICKET_LEAD 4           // Empty rollouts: steps before a group's last at which the next ticket is taken
#endif
__host__ __device__ constexpr bool spare_gen(int gen) { return MGB_SPARES && (gen == GEN_KEYCORRIDOR || gen == GEN_PROC); }
__host__ __device__ constexpr bool template_gen(int gen) { return gen == GEN_EMPTY || gen == GEN_DYNOBS; }
constexpr uint32_t CODE_BLUE_BALL = (uint32_t)code_of(T_BALL, C_BLUE, 0);

// grid of a pristine env into the lane's shared-memory column: template words + the balls of the obstacle list
// (which is already in shared memory: words GW+XWORDS..)
// `tmpl_s`: the CTA's shared-memory copy of the template (all lanes read the same word: a broadcast, no global latency)
// template -> the lane's column, four words per shared-memory load (16-byte aligned: tmpl_s sits behind the 128-byte-padded tables)
__device__ __forceinline__ void copy_template(uint32_t *st, const uint32_t *tmpl_s, int GW) {
    const uint4 *t4 = reinterpret_cast<const uint4 *>(tmpl_s);
    int k = 0;
#pragma unroll 4
    for (; k + 4 <= GW; k += 4) {
        const uint4 v = t4[k >> 2];
        st[k * 32] = v.x; st[(k + 1) * 32] = v.y; st[(k + 2) * 32] = v.z; st[(k + 3) * 32] = v.w;
    }
    for (; k < GW; ++k) st[k * 32] = tmpl_s[k];
}
__device__ __forceinline__ void rebuild_pristine_grid(uint32_t *st, const RolloutParams &p, const uint32_t *tmpl_s, int flags) {
    const DevCfg &c = p.cfg;
    copy_template(st, tmpl_s, c.GW);
    for (int j = 0; j < c.n_obst; ++j) {
        int ox, oy;
        obst_get(st, c, j, ox, oy);
        cell_wr(st, ox * c.HP + oy, CODE_BLUE_BALL);
    }
    if (flags & FLAG_GOAL_GONE) cell_wr(st, c.goal_idx, CODE_EMPTY);
}

// grid word k of env `base` (column of the state block in HBM) as the state readers must see it
__device__ __forceinline__ uint32_t grid_word(const DevCfg &c, const uint32_t *base, const uint32_t *tmpl, int k) {
    const uint32_t fl = base[(c.GW + 1) * 32] >> 24;
    const bool pristine = template_gen(c.gen) && (fl & FLAG_PRISTINE);
    if (!pristine) return base[k * 32];
    uint32_t w = tmpl[k];
    if ((fl & FLAG_GOAL_GONE) && (c.goal_idx >> 2) == k) { const int sh = (c.goal_idx & 3) * 8; w = (w & ~(0xFFu << sh)) | ((uint32_t)CODE_EMPTY << sh); }
    for (int j = 0; j < c.n_obst; ++j) {
        const uint32_t o = base[(c.GW + XWORDS + (j >> 1)) * 32] >> ((j & 1) * 16);
        const int idx = (int)(o & 0xFF) * c.HP + (int)((o >> 8) & 0xFF);
        if ((idx >> 2) == k) { const int sh = (idx & 3) * 8; w = (w & ~(0xFFu << sh)) | (CODE_BLUE_BALL << sh); }
    }
    return w;
}

// ------------------------------------------------------------------------------------------
// layout generators (reset): the static part comes from the template, the random part mirrors
// the reference draw for draw.
// ------------------------------------------------------------------------------------------
// RoomGrid bookkeeping (roomgrid.py:14-37) for the <= 3 x 3 rooms of KeyCorridor, room r = j*3 + i, sides k = right, down,
// left, up -- all of it in registers (a struct of byte arrays indexed at run time lives in local memory, and with the
// shared-memory carve-out of these kernels local memory is an L2 round trip per access):
//   dpr / dpd: 6 bits per room, the random coordinate of the room's right-side / down-side door position (the other
//              coordinate is the wall's; the left / up positions are the neighbour's right / down ones, roomgrid.py:151-168)
//   doors:     bit k*16 + r: side k of room r has a door or an opening;  locked: bit r
// `door_pos[k] is not None` (has) is geometry: right i < 2, down j < rows-1, left i > 0, up j > 0.
struct Rooms { uint64_t dpr, dpd, doors; uint32_t locked; };
__device__ __forceinline__ int room_nb(int r, int k, int rows) {   // right, down, left, up
    const int i = r % 3, j = r / 3;
    if (k == 0) return i < 2 ? r + 1 : -1;
    if (k == 1) return j < rows - 1 ? r + 3 : -1;
    if (k == 2) return i > 0 ? r - 1 : -1;
    return j > 0 ? r - 3 : -1;
}
__device__ __forceinline__ bool room_has(int r, int k, int rows) { return room_nb(r, k, rows) >= 0; }
__device__ __forceinline__ bool room_door(const Rooms &R, int r, int k) { return (R.doors >> (k * 16 + r)) & 1u; }
__device__ __forceinline__ void room_door_pos(const Rooms &R, int r, int k, int rs, int &x, int &y) {
    if (k == 2) { r -= 1; k = 0; }
    if (k == 3) { r -= 3; k = 1; }
    const int i = r % 3, j = r / 3;
    if (k == 0) { x = i * (rs - 1) + rs - 1; y = (int)((R.dpr >> (6 * r)) & 63u); }
    else { x = (int)((R.dpd >> (6 * r)) & 63u); y = j * (rs - 1) + rs - 1; }
}
__device__ __forceinline__ void add_door(uint32_t *st, const DevCfg &c, Rooms &R, int r, int k, int color, bool locked) {
    // roomgrid.py:212-246 with door_idx, colour and locked given
    R.locked = (R.locked & ~(1u << r)) | ((uint32_t)locked << r);
    int x, y;
    room_door_pos(R, r, k, c.room_size, x, y);
    cell_wr(st, x * c.HP + y, (uint32_t)code_of(T_DOOR, color, locked ? 2 : 1));
    R.doors |= (1ull << (k * 16 + r)) | (1ull << (((k + 2) & 3) * 16 + room_nb(r, k, c.num_rows)));
}

// ---- procedural generators of the stock env files with the base step: crossing.py, lavagap.py,
// multiroom.py.  np_random.shuffle / choice are defined on the stream like numpy defines them (DESIGN.md "RNG"): Fisher-Yates
// from the back with one randint(0, i+1) per position; choice(range(a, b)) = a + randint(0, b - a).
__device__ __forceinline__ void rand_shuffle(Rng &rg, const RolloutParams &p, uint8_t *x, int n) {
    for (int i = n - 1; i >= 1; --i) {
        const int j = rand_int(rg, p, 0, i + 1);
        const uint8_t t = x[i]; x[i] = x[j]; x[j] = t;
    }
}
__device__ __forceinline__ void sort_small(uint8_t *x, int n) {
    for (int i = 1; i < n; ++i) { const uint8_t v = x[i]; int j = i; while (j > 0 && x[j - 1] > v) { x[j] = x[j - 1]; --j; } x[j] = v; }
}

// envs/crossing.py:24-99 (walls and goal come from the template).  A river is direction*64 + position, direction 0 = a column.
__device__ __forceinline__ void gen_crossing(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p) {
    const DevCfg &c = p.cfg;
    const int W = c.W, H = c.H, HP = c.HP, ori = c.gp1 & 3;
    const uint32_t obstacle = (c.gp1 & 4) ? (uint32_t)CODE_WALL : (uint32_t)code_of(T_LAVA, C_RED, 0);
    e.ax = 1; e.ay = 1; e.dir = 0;
    uint8_t rivers[64], rv[32], rh[32], path[64], lim_v[34], lim_h[34];
    int nr = 0, nv = 0, nh = 0, np = 0;
    if (ori != 0) for (int i = 2; i < H - 2 && nr < 31; i += 2) rivers[nr++] = (uint8_t)i;            // (v, i), :46-49
    if (ori != 1) for (int j = 2; j < W - 2 && nr < 62; j += 2) rivers[nr++] = (uint8_t)(64 + j);     // (h, j), :44,50
    rand_shuffle(rg, p, rivers, nr);                                                                  // :53
    nr = min(nr, c.gp0);                                                                              // :54
    for (int k = 0; k < nr; ++k) { if (rivers[k] < 64) rv[nv++] = rivers[k]; else rh[nh++] = rivers[k] - 64; }
    sort_small(rv, nv); sort_small(rh, nh);                                                           // :55-56
    for (int k = 0; k < nh; ++k) for (int i = 1; i < W - 1; ++i) cell_wr(st, i * HP + rh[k], obstacle);   // :57-62 (rows first, then
    for (int k = 0; k < nv; ++k) for (int j = 1; j < H - 1; ++j) cell_wr(st, rv[k] * HP + j, obstacle);   //  columns: same cells, same object)
    for (int k = 0; k < nv; ++k) path[np++] = 1;                                                      // :65: h = 1, v = 0
    for (int k = 0; k < nh; ++k) path[np++] = 0;
    rand_shuffle(rg, p, path, np);                                                                    // :66
    lim_v[0] = 0; for (int k = 0; k < nv; ++k) lim_v[k + 1] = rv[k]; lim_v[nv + 1] = (uint8_t)(H - 1);   // :69-70
    lim_h[0] = 0; for (int k = 0; k < nh; ++k) lim_h[k + 1] = rh[k]; lim_h[nh + 1] = (uint8_t)(W - 1);
    int room_i = 0, room_j = 0;
    for (int k = 0; k < np; ++k) {                                                                    // :72-85
        int i, j;
        if (path[k]) { i = lim_v[room_i + 1]; j = lim_h[room_j] + 1 + rand_int(rg, p, 0, lim_h[room_j + 1] - lim_h[room_j] - 1); ++room_i; }
        else { i = lim_v[room_i] + 1 + rand_int(rg, p, 0, lim_v[room_i + 1] - lim_v[room_i] - 1); j = lim_h[room_j + 1]; ++room_j; }
        cell_wr(st, i * HP + j, CODE_EMPTY);
    }
}

// envs/lavagap.py:21-60
__device__ __forceinline__ void gen_lavagap(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p) {
    const DevCfg &c = p.cfg;
    const int W = c.W, H = c.H, HP = c.HP;
    const uint32_t obstacle = c.gp1 ? (uint32_t)CODE_WALL : (uint32_t)code_of(T_LAVA, C_RED, 0);
    e.ax = 1; e.ay = 1; e.dir = 0;
    const int gx = c.gp0 ? W / 2 : rand_int(rg, p, 2, W - 2);                                         // :40-49
    const int gy = rand_int(rg, p, 1, H - 1);
    for (int j = 1; j <= H - 2; ++j) cell_wr(st, gx * HP + j, obstacle);                              // :52
    cell_wr(st, gx * HP + gy, CODE_EMPTY);                                                            // :55
}

// envs/multiroom.py:41-241.  _placeRoom recurses, but never backtracks: a call either fails before appending its room, or
// appends it and tries up to 8 times to place the next one -- the first success ends every loop above it.  So a chain of
// rooms grows one room at a time; it is grown here in a loop.
// The two room lists (the chain being grown, the longest so far: 2 x 8 rooms x 8 bytes) live in the lane's column of the
// warp's staging block -- shared memory, bank == lane -- which is idle between two observations (k_rollout waits for the
// previous block's bulk store before a reset of this kernel).  Arrays indexed at run time would otherwise sit in local
// memory, an L2 round trip per access under these kernels' shared-memory carve-out.
struct MRoom { int topX, topY, sizeX, sizeY, entryX, entryY, entryWall; };
__device__ __forceinline__ void mr_put(uint32_t *scr, int slot, const MRoom &m) {
    scr[(2 * slot) * 32] = (uint32_t)(m.topX & 0xFF) | ((uint32_t)(m.topY & 0xFF) << 8) | ((uint32_t)m.sizeX << 16) | ((uint32_t)m.sizeY << 24);
    scr[(2 * slot + 1) * 32] = (uint32_t)(m.entryX & 0xFF) | ((uint32_t)(m.entryY & 0xFF) << 8) | ((uint32_t)m.entryWall << 16);
}
__device__ __forceinline__ MRoom mr_get(const uint32_t *scr, int slot) {
    const uint32_t a = scr[(2 * slot) * 32], b = scr[(2 * slot + 1) * 32];
    MRoom m;
    m.topX = (int)(a & 0xFF); m.topY = (int)((a >> 8) & 0xFF); m.sizeX = (int)((a >> 16) & 0xFF); m.sizeY = (int)(a >> 24);
    m.entryX = (int)(b & 0xFF); m.entryY = (int)((b >> 8) & 0xFF); m.entryWall = (int)((b >> 16) & 0xFF);
    return m;
}
// one _placeRoom call up to the point where the room is appended (multiroom.py:123-186); list = slots base .. base+n-1
__device__ __forceinline__ bool mr_try_room(Rng &rg, const RolloutParams &p, uint32_t *scr, int base, int &n, int maxSz, int wall, int ex, int ey) {
    const DevCfg &c = p.cfg;
    const int sizeX = rand_int(rg, p, 4, maxSz + 1), sizeY = rand_int(rg, p, 4, maxSz + 1);
    int topX, topY;
    if (n == 0) { topX = ex; topY = ey; }
    else if (wall == 0) { topX = ex - sizeX + 1; topY = rand_int(rg, p, ey - sizeY + 2, ey); }
    else if (wall == 1) { topX = rand_int(rg, p, ex - sizeX + 2, ex); topY = ey - sizeY + 1; }
    else if (wall == 2) { topX = ex; topY = rand_int(rg, p, ey - sizeY + 2, ey); }
    else { topX = rand_int(rg, p, ex - sizeX + 2, ex); topY = ey; }
    if (topX < 0 || topY < 0) return false;                                                           // :164-167
    if (topX + sizeX > c.W || topY + sizeY >= c.H) return false;
    for (int k = 0; k + 1 < n; ++k) {                                                                 // :170-178: roomList[:-1]
        const MRoom r = mr_get(scr, base + k);
        const bool nonOverlap = topX + sizeX < r.topX || r.topX + r.sizeX <= topX || topY + sizeY < r.topY || r.topY + r.sizeY <= topY;
        if (!nonOverlap) return false;
    }
    MRoom m;
    m.topX = topX; m.topY = topY; m.sizeX = sizeX; m.sizeY = sizeY; m.entryX = ex; m.entryY = ey; m.entryWall = wall;
    mr_put(scr, base + n, m);
    ++n;
    return true;
}
__device__ __noinline__ bool gen_multiroom(uint32_t *st, Env &e_, Rng &rg_, const RolloutParams &p, uint32_t *scr) {
    Env e = e_; Rng rg = rg_;                                     // register copies: see rand_int_inl
    const DevCfg &c = p.cfg;
    const int W = c.W, HP = c.HP, maxSz = c.gp1;
    constexpr int BEST = 0, CUR = 8;                              // slot ranges of the two lists
    int nbest = 0;
    bool ok = true;
    const int numRooms = min(rand_int(rg, p, c.gp0, c.gp0 + 1), 8);                                   // :44
    for (int guard = 0; nbest < numRooms; ++guard) {                                                  // :46-64
        if (guard > 100000 || (rg.err & (ERR_TAPE_END | ERR_TAPE_RANGE))) { ok = false; break; }
        int ncur = 0;
        const int ex0 = rand_int(rg, p, 0, W - 2), ey0 = rand_int(rg, p, 0, W - 2);
        bool grow = mr_try_room(rg, p, scr, CUR, ncur, maxSz, 2, ex0, ey0);
        while (grow && ncur < numRooms) {                                                             // a placed room with numLeft > 1: :193-239
            const MRoom r = mr_get(scr, CUR + ncur - 1);
            grow = false;
            for (int i = 0; i < 8 && !grow; ++i) {
                const int pick = rand_int(rg, p, 0, 3);                                               // _rand_elem(sorted(wallSet - {entryDoorWall}))
                const int exitWall = pick + (pick >= r.entryWall ? 1 : 0);
                const int nextEntryWall = (exitWall + 2) & 3;
                int dx, dy;
                if (exitWall == 0) { dx = r.topX + r.sizeX - 1; dy = r.topY + rand_int(rg, p, 1, r.sizeY - 1); }
                else if (exitWall == 1) { dx = r.topX + rand_int(rg, p, 1, r.sizeX - 1); dy = r.topY + r.sizeY - 1; }
                else if (exitWall == 2) { dx = r.topX; dy = r.topY + rand_int(rg, p, 1, r.sizeY - 1); }
                else { dx = r.topX + rand_int(rg, p, 1, r.sizeX - 1); dy = r.topY; }
                if (rg.err & (ERR_TAPE_END | ERR_TAPE_RANGE)) { ok = false; break; }
                grow = mr_try_room(rg, p, scr, CUR, ncur, maxSz, nextEntryWall, dx, dy);
            }
            if (!ok) break;
        }
        if (!ok) break;
        if (ncur > nbest) {
            for (int k = 0; k < ncur; ++k) { scr[(2 * (BEST + k)) * 32] = scr[(2 * (CUR + k)) * 32]; scr[(2 * (BEST + k) + 1) * 32] = scr[(2 * (CUR + k) + 1) * 32]; }
            nbest = ncur;
        }
    }
    if (ok) {
        int prev = -1;                                                                                // prevDoorColor, :77-108
        for (int idx = 0; idx < nbest; ++idx) {
            const MRoom r = mr_get(scr, BEST + idx);
            for (int i = 0; i < r.sizeX; ++i) { cell_wr(st, (r.topX + i) * HP + r.topY, CODE_WALL); cell_wr(st, (r.topX + i) * HP + r.topY + r.sizeY - 1, CODE_WALL); }
            for (int j = 0; j < r.sizeY; ++j) { cell_wr(st, r.topX * HP + r.topY + j, CODE_WALL); cell_wr(st, (r.topX + r.sizeX - 1) * HP + r.topY + j, CODE_WALL); }
            if (idx > 0) {
                // sorted(doorColors): COLOR_NAMES (sorted) without the previous door's colour
                const int k = rand_int(rg, p, 0, prev < 0 ? 7 : 6);
                int color = -1;
                for (int q = 0, seen = 0; q < 7; ++q) {
                    const int cq = (int)((0x4603512u >> (4 * q)) & 0xF);                              // COLOR_NAMES order -> colour index
                    if (cq == prev) continue;
                    if (seen++ == k) { color = cq; break; }
                }
                cell_wr(st, r.entryX * HP + r.entryY, (uint32_t)code_of(T_DOOR, color, 1));           // Door(color): closed, unlocked
                prev = color;
            }
        }
        const MRoom f = mr_get(scr, BEST), l = mr_get(scr, BEST + nbest - 1);
        ok = place_agent(st, e, rg, p, f.topX, f.topY, f.sizeX, f.sizeY, -1);                         // :111
        int x, y;
        ok = ok && place_obj(st, e, rg, p, CODE_GOAL, l.topX, l.topY, l.sizeX, l.sizeY, false, -1, true, x, y);   // :114
    }
    e_ = e; rg_ = rg;
    return ok;
}

template <int GEN>
__device__ __forceinline__ void generate_body(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p, PoolCtx *pc, uint32_t *scr, const uint32_t *tmpl_s);
// out of line; works on register copies of the caller's Env / Rng (see rand_int_inl)
template <int GEN>
__device__ __noinline__ void generate(uint32_t *st, Env &e_, Rng &rg_, const RolloutParams &p, PoolCtx *pc, uint32_t *scr, const uint32_t *tmpl_s) {
    if (GEN == GEN_EMPTY) { generate_body<GEN>(st, e_, rg_, p, pc, scr, tmpl_s); return; }     // see rand_int_ool
    Env e = e_;
    Rng rg = rg_;
    generate_body<GEN>(st, e, rg, p, pc, scr, tmpl_s);
    e_ = e;
    rg_ = rg;
}
template <int GEN>
__device__ __forceinline__ void generate_body(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p, PoolCtx *pc, uint32_t *scr, const uint32_t *tmpl_s) {
    const DevCfg &c = p.cfg;
    const int W = c.W, H = c.H, HP = c.HP;
    if (GEN == GEN_POOL) {
        // no generator on the device: draw one of the uploaded reference layouts
        rg.episode++;
        if (!p.tape) rg.ndraws = 0;
        rg.rblk = 0xFFFFFFFFu;
        e.carry = 0; e.steps = 0; e.target = 0; e.dirty = true;
        if (p.pool_n <= 0) { rg.err |= ERR_NO_POOL; return; }
        const int lvl = rand_int(rg, p, 0, p.pool_n);
        const uint32_t *src = p.pool + (size_t)lvl * (c.GW + POOL_XW);
#pragma unroll 8
        for (int k = 0; k < c.GW; ++k) st[k * 32] = __ldg(&src[k]);        // independent loads, eight in flight
        const uint32_t a = __ldg(&src[c.GW]);
        e.ax = a & 0xFF; e.ay = (a >> 8) & 0xFF; e.dir = (a >> 16) & 3;
        pc->level = lvl;
        pc->hp0 = __ldg(&src[c.GW + 1]); pc->hp1 = __ldg(&src[c.GW + 2]); pc->hp2 = __ldg(&src[c.GW + 3]); pc->hp3 = __ldg(&src[c.GW + 4]);
        return;
    }
    // Grid(width,height) + static walls/goal
    if (GEN == GEN_DYNOBS && (e.flags & FLAG_PRISTINE)) {
        // nothing but the balls ever changes in this env (actions >= 3 are clamped, dynamicobstacles.py:62-63):
        // removing the old balls restores the template
        for (int k = 0; k < c.n_obst; ++k) { int ox, oy; obst_get(st, c, k, ox, oy); cell_wr(st, ox * HP + oy, CODE_EMPTY); }
    } else {
        copy_template(st, tmpl_s, c.GW);                                // from the CTA's shared-memory copy
    }
    e.dirty = true;
    rg.episode++;
    if (!p.tape) rg.ndraws = 0;
    rg.rblk = 0xFFFFFFFFu;
    e.carry = 0; e.steps = 0; e.target = 0;
    bool ok = true;
    int x, y;
    if (GEN == GEN_EMPTY) {                              // envs/empty.py:30-57 (extra == 0)
        if (!c.random_start) { e.ax = 1; e.ay = 1; e.dir = 0; }
        else ok = place_agent<true>(st, e, rg, p, 0, 0, W, H, -1);
        e.flags = (e.flags & ~FLAG_GOAL_GONE) | FLAG_PRISTINE;
    } else if (GEN == GEN_DOORKEY) {                     // envs/doorkey.py:15-44
        const int split = rand_int(rg, p, 2, W - 2);
        for (int j = 0; j < H; ++j) cell_wr(st, split * HP + j, CODE_WALL);
        ok = place_agent(st, e, rg, p, 0, 0, split, H, -1);
        const int door = rand_int(rg, p, 1, W - 2);
        cell_wr(st, split * HP + door, code_of(T_DOOR, C_YELLOW, 2));
        ok = ok && place_obj(st, e, rg, p, code_of(T_KEY, C_YELLOW, 0), 0, 0, split, H, false, -1, true, x, y);
    } else if (GEN == GEN_FOURROOMS) {                   // envs/fourrooms.py:19-69
        const int rw = W / 2, rh = H / 2;
        // walls are in the template; gaps in reference draw order (j,i) = (0,0),(0,1),(1,0)
        const int g1 = rand_int(rg, p, 1, rh);            cell_wr(st, rw * HP + g1, CODE_EMPTY);
        const int g2 = rand_int(rg, p, 1, rw);            cell_wr(st, g2 * HP + rh, CODE_EMPTY);
        const int g3 = rand_int(rg, p, rw + 1, 2 * rw);   cell_wr(st, g3 * HP + rh, CODE_EMPTY);
        const int g4 = rand_int(rg, p, rh + 1, 2 * rh);   cell_wr(st, rw * HP + g4, CODE_EMPTY);
        ok = place_agent(st, e, rg, p, 0, 0, W, H, -1);
        ok = ok && place_obj(st, e, rg, p, CODE_GOAL, 0, 0, W, H, false, -1, true, x, y);
    } else if (GEN == GEN_DYNOBS) {                      // envs/dynamicobstacles.py:35-58
        if (!c.random_start) { e.ax = 1; e.ay = 1; e.dir = 0; }
        else ok = place_agent(st, e, rg, p, 0, 0, W, H, -1);
        for (int k = 0; k < c.n_obst; ++k) {
            const bool placed = place_obj<true>(st, e, rg, p, code_of(T_BALL, C_BLUE, 0), 0, 0, W, H, false, 100, true, x, y);   // resets are frequent here: inline Philox
            // 101 rejected tries (the reference raises RecursionError): no position to record -- the ball is parked on the
            // agent's start cell record-wise (a move from there finds no free neighbour or steps off normally) and the
            // failure is flagged
            obst_set(st, c, k, placed ? x : e.ax, placed ? y : e.ay);
            ok = ok && placed;
        }
        e.flags |= FLAG_PRISTINE;
    } else if (GEN == GEN_PROC) {                        // crossing.py / lavagap.py / multiroom.py
        if (c.gen == GEN_CROSSING) gen_crossing(st, e, rg, p);
        else if (c.gen == GEN_LAVAGAP) gen_lavagap(st, e, rg, p);
        else ok = gen_multiroom(st, e, rg, p, scr);
    } else if (GEN == GEN_KEYCORRIDOR) {                 // roomgrid.py:118-169 + envs/keycorridor.py:26-49
        Rooms R;
        R.dpr = R.dpd = R.doors = 0; R.locked = 0;
        const int rs = c.room_size, rows = c.num_rows;
        for (int j = 0; j < rows; ++j)
            for (int i = 0; i < 3; ++i) {                              // roomgrid.py:139-168: door positions, row-major draw order
                const int r = j * 3 + i, tx = i * (rs - 1), ty = j * (rs - 1);
                const int x_l = tx + 1, y_l = ty + 1, x_m = tx + rs - 1, y_m = ty + rs - 1;
                if (i < 2) R.dpr |= (uint64_t)rand_int(rg, p, y_l, y_m) << (6 * r);
                if (j < rows - 1) R.dpd |= (uint64_t)rand_int(rg, p, x_l, x_m) << (6 * r);
            }
        e.ax = 1 * (rs - 1) + rs / 2; e.ay = (rows / 2) * (rs - 1) + rs / 2; e.dir = 0;
        // remove_wall(1, j, 3) (roomgrid.py:248-282)
        for (int j = 1; j < rows; ++j) {
            const int r = j * 3 + 1, tx = rs - 1, ty = j * (rs - 1);
            for (int m = 1; m < rs - 1; ++m) cell_wr(st, (tx + m) * HP + ty, CODE_EMPTY);
            R.doors |= (1ull << (3 * 16 + r)) | (1ull << (1 * 16 + r - 3));
        }
        const int room_idx = rand_int(rg, p, 0, rows);
        const int door_color = rand_color(rg, p);                     // add_door(2, room_idx, 2, locked=True)
        add_door(st, c, R, room_idx * 3 + 2, 2, door_color, true);
        const int obj_color = rand_color(rg, p);                      // add_object(2, room_idx, "ball")
        ok = place_obj(st, e, rg, p, code_of(T_BALL, obj_color, 0), 2 * (rs - 1), room_idx * (rs - 1), rs, rs, true, 1000, true, x, y);
        const int key_room = rand_int(rg, p, 0, rows);                // add_object(0, ri, "key", door.color)
        ok = place_obj(st, e, rg, p, code_of(T_KEY, door_color, 0), 0, key_room * (rs - 1), rs, rs, true, 1000, true, x, y) && ok;
        // RoomGrid.place_agent(1, rows//2) (roomgrid.py:284-303)
        for (int guard = 0; ok && guard < HARD_TRY_CAP; ++guard) {
            ok = place_agent(st, e, rg, p, rs - 1, (rows / 2) * (rs - 1), rs, rs, 1000);
            if (!ok) break;
            const int dx = (e.dir & 1) ? 0 : 1 - e.dir, dy = (e.dir & 1) ? 2 - e.dir : 0;
            const uint32_t f = cell_rd(st, (e.ax + dx) * HP + (e.ay + dy));
            if (f == CODE_EMPTY || f / 21 == T_WALL) break;
        }
        // connect_all (roomgrid.py:305-359).  Reachability (:312-327) as a fixpoint on 9-bit room masks: a reached room with
        // a door on its right / down / left / up side reaches room r+1 / r+3 / r-1 / r-3.
        const int start = (e.ay / (rs - 1)) * 3 + e.ax / (rs - 1);
        const uint32_t all_rooms = (1u << (rows * 3)) - 1u;
        for (int it = 0; ok; ++it) {
            if (it > 5000) { ok = false; break; }
            uint32_t reach = 1u << start;
            for (;;) {
                const uint32_t d0 = (uint32_t)R.doors & 0x1FFu, d1 = (uint32_t)(R.doors >> 16) & 0x1FFu;
                const uint32_t d2 = (uint32_t)(R.doors >> 32) & 0x1FFu, d3 = (uint32_t)(R.doors >> 48) & 0x1FFu;
                const uint32_t nr = (reach | ((reach & d0) << 1) | ((reach & d1) << 3) | ((reach & d2) >> 1) | ((reach & d3) >> 3)) & all_rooms;
                if (nr == reach) break;
                reach = nr;
            }
            if (reach == all_rooms) break;
            const int i = rand_int(rg, p, 0, 3);
            const int j = rand_int(rg, p, 0, rows);
            const int k = rand_int(rg, p, 0, 4);
            if (rg.err & (ERR_TAPE_END | ERR_TAPE_RANGE)) { ok = false; break; }
            const int r = j * 3 + i;
            if (!room_has(r, k, rows) || room_door(R, r, k)) continue;
            if (((R.locked >> r) & 1u) || ((R.locked >> room_nb(r, k, rows)) & 1u)) continue;
            const int color = rand_color(rg, p);
            add_door(st, c, R, r, k, color, false);
        }
        e.target = code_of(T_BALL, obj_color, 0);
    }
    if (!ok) rg.err |= ERR_SAMPLING;
}

// ------------------------------------------------------------------------------------------
// transition: MiniGridEnv.step (minigrid.py:1227-1325) + subclass hooks
// ------------------------------------------------------------------------------------------
// DynamicObstaclesEnv.step, obstacle update (envs/dynamicobstacles.py:70-78) on the Philox stream.  Balls are taken
// in list order by all lanes in lock step.  The tries of one ball see an unchanging grid (it only changes when a try
// succeeds), so the first DYN_SPEC tries are evaluated speculatively in straight-line code -- independent loads,
// no branches -- and the first valid one wins; a lane whose DYN_SPEC tries all failed (p ~ 0.25^4) or whose draw
// window is nearly used up continues one try at a time in a (divergent, rare) loop, computing Philox blocks on demand
// once it runs past the prefetched window.  A try is one stream word when the draw counter is even (x from the word,
// y from the word times DRAW_ODD_MULT); after an odd number of draws (the -Random- ids: place_agent's direction draw)
// it straddles two words.
constexpr int DYN_SPEC = 4;
// byte offset of grid cell (x,y) inside a lane's column: word x*HP/4 + (y>>2) at pitch 128, byte y&3
__device__ __forceinline__ uint32_t cell_off(int x, int y, int HP) { return (uint32_t)(x * (HP * 32) + y + (y >> 2) * 124); }

// word `widx` of the stream (seed, gid, stream): word widx&3 of Philox block widx>>2 (cold path: one block per call)
__device__ __forceinline__ uint32_t stream_word(uint32_t widx, uint32_t stream, int64_t gid, uint64_t seed) {
    uint32_t o0, o1, o2, o3;
    philox4x32_10(widx >> 2, stream, (uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), (uint32_t)seed, (uint32_t)(seed >> 32), o0, o1, o2, o3);
    return (widx & 2) ? ((widx & 1) ? o3 : o2) : ((widx & 1) ? o1 : o0);
}

// Rare continuation of one ball's rejection sampling after the speculative tries (divergent: ~1 lane of a warp in a
// third of the ball steps), one try at a time, out of line so that the hot loop stays small.  Words come from the
// prefetched window while they last, from a Philox block computed on the spot after that.  Returns the shared-memory
// address of the accepted cell, or 0 when place_obj's 101 tries are used up (minigrid.py:1028-1031); nd advances by
// two draws per try.
struct MoreTries { uint32_t nsa, npos, nd; };
__device__ __noinline__ MoreTries dynobs_more_tries(uint32_t st_sa, uint32_t dr_sa, uint32_t wbase, uint32_t winw, uint32_t nd, int tries,
                                                    int tx, int ty, uint32_t sx, uint32_t sy, int HP, uint32_t stream, int64_t gid, uint64_t seed) {
    MoreTries r;
    r.nsa = 0; r.npos = 0;
    const uint32_t par = nd & 1u;
    for (; tries <= 100; ++tries) {
        const uint32_t wi = (nd - wbase) >> 1;                    // window word of draw nd
        uint32_t w0, w1 = 0;
        if (wi + par < winw) {
            w0 = lds_u32(dr_sa + wi * 128u);
            if (par) w1 = lds_u32(dr_sa + wi * 128u + 128u);
        } else {
            w0 = stream_word((wbase >> 1) + wi, stream, gid, seed);
            if (par) w1 = stream_word((wbase >> 1) + wi + 1u, stream, gid, seed);
        }
        nd += 2;
        const uint32_t wm = w0 * DRAW_ODD_MULT;
        const int x = tx + (int)__umulhi(par ? wm : w0, sx);
        const int y = ty + (int)__umulhi(par ? w1 : wm, sy);
        const uint32_t sa = st_sa + cell_off(x, y, HP);
        if (lds_u8(sa) == CODE_EMPTY) { r.nsa = sa; r.npos = (uint32_t)(x | (y << 8)); break; }
    }
    r.nd = nd;
    return r;
}

template <int V>
__device__ __forceinline__ void dynobs_move(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p, uint32_t *draws) {
    const DevCfg &c = p.cfg;
    const int W = c.W, H = c.H, HP = c.HP, nob = c.n_obst;
    constexpr uint32_t WINW = 4 * draw_blocks(V);                 // words (= aligned tries) in the window
    const uint32_t wbase = rg.ndraws & ~7u;                       // draw index of draws[0]
    prefetch_draws<draw_blocks(V)>(draws, wbase >> 3, rg.episode - 1u, rg.gid, p.seed);
    rg.rblk = 0xFFFFFFFFu;
    uint32_t nd = rg.ndraws;
    const uint32_t par = nd & 1u;                                 // every try takes two draws: the parity holds for the step
    const uint32_t st_sa = (uint32_t)__cvta_generic_to_shared(st);
    const uint32_t dr_sa = (uint32_t)__cvta_generic_to_shared(draws);
    uint32_t ob_p = st_sa + (uint32_t)(c.GW + XWORDS) * 128u;    // obstacle k: 16 bits at +(k>>1)*128 + (k&1)*2
    const int XP = HP * 32;                                       // bytes between two grid columns of a lane
    // "not where the agent is" (minigrid.py:1044): an empty cell under the agent is made non-empty for the duration
    // of the moves, so that a try is valid iff its cell is empty
    const uint32_t ag_sa = st_sa + cell_off(e.ax, e.ay, HP);
    const bool ag_mark = lds_u8(ag_sa) == CODE_EMPTY;
    if (ag_mark) sts_u8(ag_sa, CODE_WALL);
    // The next ball's record and address arithmetic are computed one ball ahead (a move only rewrites the mover's own
    // record), next to this ball's dependent chain instead of behind its stores.
    struct Ball { int ox, oy, tx, ty; uint32_t sx, sy, colb, old_sa; };
    auto ball_of = [&](uint32_t opos) {
        Ball b;
        b.ox = (int)(opos & 0xFF); b.oy = (int)(opos >> 8);
        b.tx = max(b.ox - 1, 0); b.ty = max(b.oy - 1, 0);
        b.sx = (uint32_t)(min(b.tx + 3, W) - b.tx); b.sy = (uint32_t)(min(b.ty + 3, H) - b.ty);
        b.colb = st_sa + (uint32_t)(b.tx * XP);                   // column tx of the lane's grid
        b.old_sa = st_sa + cell_off(b.ox, b.oy, HP);
        return b;
    };
    Ball nb = ball_of(lds_u16(ob_p));
    for (int k = 0; k < nob; ++k) {
        const Ball b = nb;
        const uint32_t ob_n = ob_p + ((k & 1) ? 126u : 2u);
        nb = ball_of(lds_u16(ob_n));                              // k = nob-1: the row behind the records, unused
        const uint32_t oldcode = lds_u8(b.old_sa);                // the ball's own code, moved with it
        uint32_t npos = 0, nsa = 0;
        const uint32_t wi = (nd - wbase) >> 1;                    // window word of draw nd
        const bool spec = wi + DYN_SPEC + par <= WINW;            // all speculative draws are inside the window
        const uint32_t wa = dr_sa + (spec ? wi : 0u) * 128u;
        uint32_t w[DYN_SPEC + 1];
#pragma unroll
        for (int j = 0; j <= DYN_SPEC; ++j) w[j] = lds_u32(wa + j * 128);     // word DYN_SPEC is only used when par
        int sel = DYN_SPEC;
#pragma unroll
        for (int j = DYN_SPEC - 1; j >= 0; --j) {                 // descending: the lowest valid try overwrites
            const uint32_t wm = w[j] * DRAW_ODD_MULT;
            const uint32_t ux = par ? wm : w[j], uy = par ? w[j + 1] : wm;
            const int dx = (int)__umulhi(ux, b.sx);
            const int y = b.ty + (int)__umulhi(uy, b.sy);
            const uint32_t sa = b.colb + (uint32_t)(dx * XP) + (uint32_t)(y + (y >> 2) * 124);
            // the ball's own cell counts as occupied: it must move (minigrid.py:1040-1041)
            if (lds_u8(sa) == CODE_EMPTY) { sel = j; npos = (uint32_t)(dx | (y << 8)); nsa = sa; }
        }
        npos += (uint32_t)b.tx;
        if (spec) nd += 2u * (uint32_t)min(sel + 1, DYN_SPEC);
        if (!spec || sel == DYN_SPEC) {
            // a tight loop over the rest of the window, one try at a time (a second speculative round measured 4 % slower: it
            // costs the whole warp a round's instructions for the one or two lanes that need it); the rest out of line
            int tries = spec ? DYN_SPEC : 0;
            nsa = 0;                                              // !spec: the speculative block looked at the wrong words
            uint32_t wp = dr_sa + ((nd - wbase) >> 1) * 128u;
            const uint32_t wend = dr_sa + (WINW - par) * 128u;
            while (wp < wend) {                                   // tries <= 100 holds: the window has at most 20 words
                const uint32_t w0 = lds_u32(wp), w1 = lds_u32(wp + 128u);
                wp += 128u; nd += 2u; ++tries;
                const uint32_t wm = w0 * DRAW_ODD_MULT;
                const int dx = (int)__umulhi(par ? wm : w0, b.sx);
                const int y = b.ty + (int)__umulhi(par ? w1 : wm, b.sy);
                const uint32_t sa = b.colb + (uint32_t)(dx * XP) + (uint32_t)(y + (y >> 2) * 124);
                if (lds_u8(sa) == CODE_EMPTY) { nsa = sa; npos = (uint32_t)((b.tx + dx) | (y << 8)); break; }
            }
            if (!nsa) {
                const MoreTries r = dynobs_more_tries(st_sa, dr_sa, wbase, WINW, nd, tries, b.tx, b.ty, b.sx, b.sy, HP, rg.episode - 1u, rg.gid, p.seed);
                nsa = r.nsa; npos = r.npos; nd = r.nd;
            }
        }
        if (nsa) {                                                // a failed placement (RecursionError, swallowed) leaves the ball
            sts_u8(nsa, oldcode);
            sts_u8(b.old_sa, CODE_EMPTY);
            sts_u16(ob_p, npos);
        }
        ob_p = ob_n;
    }
    if (ag_mark) sts_u8(ag_sa, CODE_EMPTY);
    rg.ndraws = nd;
}

// The same update driven by the RNG tape (parity against the reference's own MT19937 draws): for each ball in list
// order, place_obj(top=old-(1,1), size=(3,3), max_tries=100), then clear the old cell; a failed placement
// (RecursionError, swallowed) leaves the ball where it is.  Cold path, out of line.
__device__ __noinline__ void dynobs_move_tape(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p) {
    const DevCfg &c = p.cfg;
    const int W = c.W, H = c.H, HP = c.HP;
    for (int k = 0; k < c.n_obst; ++k) {
        int ox, oy;
        obst_get(st, c, k, ox, oy);
        const uint32_t ball = cell_rd(st, ox * HP + oy);
        const int tx = max(ox - 1, 0), ty = max(oy - 1, 0), hx = min(tx + 3, W), hy = min(ty + 3, H);
        for (int tries = 0; tries <= 100; ++tries) {              // 101 tries (minigrid.py:1028-1031)
            const int x = rand_int_inl(rg, p, tx, hx);
            const int y = rand_int_inl(rg, p, ty, hy);
            if (rg.err & (ERR_TAPE_END | ERR_TAPE_RANGE)) return;
            if (cell_rd(st, x * HP + y) != CODE_EMPTY) continue;  // the ball's own cell counts: it must move
            if (x == e.ax && y == e.ay) continue;
            cell_wr(st, x * HP + y, ball);
            obst_set(st, c, k, x, y);
            cell_wr(st, ox * HP + oy, CODE_EMPTY);
            break;
        }
    }
}

// Warp-cooperative DynamicObstaclesEnv._gen_grid (envs/dynamicobstacles.py:35-58) for the env in column `src` of
// the warp's state block: agent at (1,1) facing right (or place_agent() for the -Random- ids), then n_obstacles times
// place_obj(Ball(), max_tries=100) over the whole grid.  With a uniform random policy an episode lasts ~12 steps, so
// nearly every warp-step has a lane or two that must reset; doing that inside one lane stalls the other 30.  Here
// lane t evaluates try t of the new episode's stream (draws 2t, 2t+1: word t&3 of Philox block t>>2).  place_obj accepts
// a try iff its cell is free in the static layout, is not the agent's, and does not hold an earlier ball -- i.e. iff
// it is free and the first try with that position (an earlier try with the same position was either not free, and
// then neither is this one, or was itself accepted).  So: ballot(free), match_any(position) for "first with that
// position", and the n_obstacles lowest accepted lanes each write their own ball.  No loop.
// Returns false (nothing but the removal of the old balls done) if 32 tries were not enough; the caller then runs the
// scalar generator, which replays the same stream from its start.
__device__ __forceinline__ bool dynobs_coop_reset(uint32_t *st_warp, int src, int lane, const RolloutParams &p, const uint32_t *tmpl_s, int64_t gid,
                                                  uint32_t stream, uint32_t &consumed, uint32_t &agent) {
    const DevCfg &c = p.cfg;
    const int HP = c.HP, nob = c.n_obst;
    const uint32_t col_sa = (uint32_t)__cvta_generic_to_shared(st_warp + src);
    const uint32_t ob_sa = col_sa + (uint32_t)(c.GW + XWORDS) * 128u;
    if (lane < nob) {                                             // grid := static layout
        const uint32_t o = lds_u16(ob_sa + (uint32_t)((lane >> 1) * 128 + (lane & 1) * 2));
        sts_u8(col_sa + cell_off((int)(o & 0xFF), (int)(o >> 8), HP), CODE_EMPTY);
    }
    __syncwarp();
    uint32_t o0, o1, o2, o3;
    philox4x32_10((uint32_t)(lane >> 2), stream, (uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), (uint32_t)p.seed, (uint32_t)(p.seed >> 32),
                  o0, o1, o2, o3);
    const uint32_t wl = (lane & 2) ? ((lane & 1) ? o3 : o2) : ((lane & 1) ? o1 : o0);      // stream word `lane`
    const uint32_t ux = wl, uy = wl * DRAW_ODD_MULT;                                     // draws 2*lane, 2*lane+1
    // the static layout is read from the CTA's shared-memory copy of the template (no global-memory latency on this path)
    auto static_free = [&](int x, int y) { const int i = x * HP + y; return ((tmpl_s[i >> 2] >> ((i & 3) * 8)) & 0xFFu) == CODE_EMPTY; };
    const uint32_t lt = (1u << lane) - 1u;
    int x = (int)__umulhi(ux, (uint32_t)c.W), y = (int)__umulhi(uy, (uint32_t)c.H);
    uint32_t eligible = 0xFFFFFFFFu;                              // lanes whose try is a ball try
    uint32_t extra = 0;                                           // draws before ball try 0
    agent = 1u | (1u << 8);                                       // fixed start: (1,1) facing right (dynamicobstacles.py:44-47)
    if (c.random_start) {
        // place_agent() (minigrid.py:1072-1090): tries on the pairs (2t, 2t+1) until the cell is empty in the static
        // layout (agent_pos is None meanwhile), then one draw for the direction; the ball tries that follow are the
        // odd-aligned pairs (2m+1, 2m+2), m > t_agent: lane m's second draw and lane m+1's first.
        const uint32_t afree = __ballot_sync(0xFFFFFFFFu, static_free(x, y));
        if (afree == 0 || (afree & 0x3FFFFFFFu) == 0) return false;           // needs the direction draw and room for balls
        const int ta = __ffs((int)afree) - 1;
        const uint32_t apos = __shfl_sync(0xFFFFFFFFu, (uint32_t)(x | (y << 8)), ta);
        const uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, ux, 1);             // draw 2*lane+2
        const uint32_t adir = __umulhi(__shfl_sync(0xFFFFFFFFu, nxt, ta), 4u);  // draw 2*ta+2
        agent = apos | (adir << 16);
        x = (int)__umulhi(uy, (uint32_t)c.W); y = (int)__umulhi(nxt, (uint32_t)c.H);
        eligible = (ta >= 30) ? 0u : ((0xFFFFFFFFu << (ta + 1)) & 0x7FFFFFFFu);  // lane 31 has no draw 64
        extra = 1;
    }
    const uint32_t pos = (uint32_t)(x | (y << 8));
    const bool is_free = static_free(x, y) && pos != (agent & 0xFFFFu) && ((eligible >> lane) & 1u);
    const bool first = (__match_any_sync(0xFFFFFFFFu, pos) & lt & eligible) == 0;
    const uint32_t acc = __ballot_sync(0xFFFFFFFFu, is_free && first);
    if (__popc(acc) < nob) return false;
    const int rank = __popc(acc & lt);
    const bool mine = ((acc >> lane) & 1u) && rank < nob;
    // draws up to and including the last ball's try: pairs (2m, 2m+1), or (2m+1, 2m+2) after a random start
    consumed = 2u * (uint32_t)__ffs((int)__ballot_sync(0xFFFFFFFFu, mine && rank == nob - 1)) + extra;
    if (mine) {
        sts_u8(col_sa + cell_off(x, y, HP), (uint32_t)code_of(T_BALL, C_BLUE, 0));
        sts_u16(ob_sa + (uint32_t)((rank >> 1) * 128 + (rank & 1) * 2), pos);
    }
    __syncwarp();
    return true;
}

__device__ __forceinline__ double reward_formula(int steps, int max_steps) {
    // _reward (minigrid.py:933-937): 1 - 0.9 * (step_count / max_steps), three separately
    // rounded fp64 operations -- explicit _rn intrinsics so that nvcc cannot contract to an FMA.
    return __dsub_rn(1.0, __dmul_rn(0.9, __ddiv_rn((double)steps, (double)max_steps)));
}

template <int GEN, bool SEE, int V>
__device__ __forceinline__ void transition(uint32_t *st, Env &e, Rng &rg, const RolloutParams &p, const uint32_t *lut, int action,
                                           double &reward, bool &done, uint32_t *draws, const PoolCtx &pc) {
    const DevCfg &c = p.cfg;
    const int W = c.W, H = c.H, HP = c.HP;
    reward = 0.0; done = false;
    bool not_clear = false;
    if (GEN == GEN_DYNOBS) {                             // envs/dynamicobstacles.py:60-78
        if ((unsigned)action >= 3u) action = 0;                    // n_actions == 3 (mgb_create checks): only left / right / forward exist here,
                                                          // so the pickup / drop / toggle logic below folds away
        const int dx0 = (e.dir & 1) ? 0 : 1 - e.dir, dy0 = (e.dir & 1) ? 2 - e.dir : 0;
        const int fx0 = e.ax + dx0, fy0 = e.ay + dy0;
        uint32_t front = CODE_WALL;
        if ((unsigned)fx0 < (unsigned)W && (unsigned)fy0 < (unsigned)H) front = cell_rd(st, fx0 * HP + fy0);
        not_clear = front != CODE_EMPTY && (lut[front * lut_pitch_words(GEN)] & 0xFF) != T_GOAL;
        // Update obstacle positions (dynamicobstacles.py:70-78)
        if (!p.tape) dynobs_move<V>(st, e, rg, p, draws);
        else { Env te = e; Rng tr = rg; dynobs_move_tape(st, te, tr, p); e = te; rg = tr; }     // parity-only mode, out of line
    } else if (action >= c.n_actions) {
        rg.err |= ERR_ACTION;                             // reference: assert False, "unknown action"
        action = A_DONE;
    }
    // ---- subclass step() pre-hooks of the level-pool envs
    const int pre_carry = e.carry;
    bool red_before = false, blue_before = false;
    if (GEN == GEN_POOL) {
        if (c.hook == HOOK_MEMORY && action == A_PICKUP) action = A_TOGGLE;                  // memory.py:89-90
        if (c.hook == HOOK_REDBLUEDOORS) {                                                     // redbluedoors.py:45-46
            const uint32_t ra = lut[cell_rd(st, (int)((pc.hp1 >> 16) & 0xFF) * HP + (int)(pc.hp1 >> 24)) * lut_pitch_words(GEN)];
            const uint32_t rb = lut[cell_rd(st, (int)(pc.hp2 & 0xFF) * HP + (int)((pc.hp2 >> 8) & 0xFF)) * lut_pitch_words(GEN)];
            red_before = (ra & 0xFF) == T_DOOR && ((ra >> 16) & 0xFF) == 0;
            blue_before = (rb & 0xFF) == T_DOOR && ((rb >> 16) & 0xFF) == 0;
        }
    }
    e.steps++;
    const int dx = (e.dir & 1) ? 0 : 1 - e.dir, dy = (e.dir & 1) ? 2 - e.dir : 0;
    const int fx = e.ax + dx, fy = e.ay + dy;
    uint32_t fc = CODE_WALL;
    const int fidx = fx * HP + fy;
    const bool f_in = (unsigned)fx < (unsigned)W && (unsigned)fy < (unsigned)H;
    if (f_in) fc = cell_rd(st, fidx); else rg.err |= ERR_BOUNDS;
    const uint32_t fw = lut[fc * lut_pitch_words(GEN)];
    const uint32_t ff = fw >> 24;
    const int ftype = fw & 0xFF;
    // select form of the action switch (minigrid.py:1245-1318): one rarely-taken branch for grid edits
    const int turn = action == A_LEFT ? 3 : (action == A_RIGHT ? 1 : 0);
    e.dir = (e.dir + turn) & 3;
    const bool fwd = action == A_FORWARD;
    const bool move = fwd && (ff & F_OVERLAP);
    e.ax = move ? fx : e.ax;
    e.ay = move ? fy : e.ay;
    const bool goal = fwd && (ff & F_TGOAL);                  // only Goal(toggletimes<=0) terminates (:157-160,:1259)
    const bool lava = fwd && (ff & F_LAVA);
    done = goal || (lava && !c.lava_v1);
    if (lava && c.lava_v1) reward = -1.0;                     // 'v1' in class name (:1263-1266)
    if (goal) reward = reward_formula(e.steps, c.max_steps);
    const int ds = (fw >> 16) & 0xFF, dcol = (fw >> 8) & 0xFF;
    const int ns = (ds == 2) ? ((e.carry == code_of(T_KEY, dcol, 0)) ? 0 : 2) : (ds ^ 1);   // Door.toggle :252-262
    constexpr bool MANIP = GEN != GEN_DYNOBS;               // Dynamic-Obstacles has no pickup / drop / toggle (dynamicobstacles.py:32,62-63)
    const bool tog = MANIP && action == A_TOGGLE && f_in;
    const bool tog_door = tog && ftype == T_DOOR && ns != ds;
    // default Box (contains None) and default Goal (toggletimes 1) vanish when toggled (:171-177, :355-360)
    const bool tog_vanish = tog && (ftype == T_BOX || (ftype == T_GOAL && !(ff & F_TGOAL)));
    const bool pick = MANIP && action == A_PICKUP && (ff & F_PICKUP) && e.carry == 0 && f_in;
    const bool drop = MANIP && action == A_DROP && fc == CODE_EMPTY && e.carry != 0 && f_in;
    if (pick || drop || tog_door || tog_vanish) {
        uint32_t nv = CODE_EMPTY;
        if (drop) nv = (uint32_t)e.carry;
        if (tog_door) nv = fc - ds + ns;
        if (tog_vanish && fc >= CODE_KEYBOX0) nv = (uint32_t)code_of(T_KEY, (int)fc - CODE_KEYBOX0, 0);   // Box.toggle: cell := contents (:355-360)
        cell_wr(st, fidx, nv);
        e.dirty = true;
        if (GEN == GEN_EMPTY && tog_vanish && fidx == c.goal_idx) e.flags |= FLAG_GOAL_GONE;      // still implied (harmless if not pristine)
        else e.flags &= ~FLAGS_IMPLIED;
        e.carry = pick ? (int)fc : (drop ? 0 : e.carry);
    }
    if (e.steps >= c.max_steps) done = true;
    if (GEN == GEN_KEYCORRIDOR) {                        // envs/keycorridor.py:51-59
        if (action == A_PICKUP && e.carry != 0 && e.carry == e.target) { reward = reward_formula(e.steps, c.max_steps); done = true; }
    }
    if (GEN == GEN_POOL && c.hook != HOOK_NONE) {        // step() post-hooks of the level-pool envs
        const int tcode = pc.hp0 & 0xFF, mcode = (pc.hp0 >> 8) & 0xFF;
        const int tx = pc.hp1 & 0xFF, ty = (pc.hp1 >> 8) & 0xFF;
        const int Ax = (pc.hp1 >> 16) & 0xFF, Ay = pc.hp1 >> 24, Bx = pc.hp2 & 0xFF, By = (pc.hp2 >> 8) & 0xFF;
        const int Cx = (pc.hp2 >> 16) & 0xFF, Cy = pc.hp2 >> 24, Dx = pc.hp3 & 0xFF, Dy = (pc.hp3 >> 8) & 0xFF;
        auto adj4 = [&](int x, int y) { return (e.ax == x && abs(e.ay - y) == 1) || (e.ay == y && abs(e.ax - x) == 1); };
        auto door_open = [&](int x, int y) {
            const uint32_t w = lut[cell_rd(st, x * HP + y) * lut_pitch_words(GEN)];
            return (w & 0xFF) == T_DOOR && ((w >> 16) & 0xFF) == 0;
        };
        bool win = false;
        switch (c.hook) {
        case HOOK_PICKUP_TARGET:                          // unlockpickup.py:34-42, blockedunlockpickup.py:38-46
            if (action == A_PICKUP && e.carry != 0 && e.carry == tcode) { win = true; done = true; }
            break;
        case HOOK_UNLOCK:                                 // unlock.py:33-41
            if (action == A_TOGGLE && door_open(Ax, Ay)) { win = true; done = true; }
            break;
        case HOOK_FETCH:                                  // fetch.py:74-86
            if (e.carry != 0) { if (e.carry == tcode) win = true; else reward = 0.0; done = true; }
            break;
        case HOOK_GOTODOOR:                               // gotodoor.py:72-93
            if (action == A_DONE) {
                if (adj4(tx, ty)) win = true;
                if (adj4(Ax, Ay) || adj4(Bx, By) || adj4(Cx, Cy) || adj4(Dx, Dy)) done = true;
            }
            break;
        case HOOK_GOTOOBJECT:                             // gotoobject.py:68-84
            if (action == A_TOGGLE) done = true;
            if (action == A_DONE) { if (abs(e.ax - tx) <= 1 && abs(e.ay - ty) <= 1) win = true; done = true; }
            break;
        case HOOK_PUTNEAR: {                              // putnear.py:91-112
            if (action == A_PICKUP && e.carry != 0 && e.carry != mcode) done = true;
            if (action == A_DROP && pre_carry != 0) {
                const int ox = e.ax + dx, oy = e.ay + dy;                 // dir is unchanged by drop
                if (e.carry == 0 && abs(ox - tx) <= 1 && abs(oy - ty) <= 1) win = true;   // dropped: the cell in front now holds it
                done = true;
            }
            break;
        }
        case HOOK_REDBLUEDOORS: {                         // redbluedoors.py:44-66
            const bool red_after = door_open(Ax, Ay), blue_after = door_open(Bx, By);
            if (blue_after) { if (red_before) win = true; else reward = 0.0; done = true; }
            else if (red_after) { if (blue_before) { reward = 0.0; done = true; } }
            break;
        }
        case HOOK_MEMORY:                                 // memory.py:88-100
            if (e.ax == Ax && e.ay == Ay) { win = true; done = true; }
            if (e.ax == Bx && e.ay == By) { win = false; reward = 0.0; done = true; }
            break;
        }
        if (win) reward = reward_formula(e.steps, c.max_steps);
    }
    if (GEN == GEN_DYNOBS) {                             // envs/dynamicobstacles.py:84-87
        e.dirty = true;
        if (action == A_FORWARD && not_clear) { reward = -1.0; done = true; }
    }
}

// ------------------------------------------------------------------------------------------
// observation: gen_obs_grid + Grid.encode (minigrid.py:1327-1381, 571-594)
// ------------------------------------------------------------------------------------------
// (v & bit) ? x : 0.  Spelled as and/setp/selp so that ptxas turns the V tests of one row mask into a single R2P
// (register bits -> predicates) plus one SEL per cell, instead of shift-left / arithmetic-shift-right / and per cell.
__device__ __forceinline__ uint32_t sel_bit(uint32_t v, uint32_t bit, uint32_t x) {
    uint32_t r;
    asm("{ .reg .pred p; .reg .b32 t; and.b32 t, %1, %3; setp.ne.u32 p, t, 0; selp.b32 %0, %2, 0, p; }" : "=r"(r) : "r"(v), "r"(x), "r"(bit));
    return r;
}

// code -> LUT word; the address is formed with a multiply-add (see the LUT comment at the top)
template <int GEN>
__device__ __forceinline__ uint32_t lut_ld(uint32_t lut_sa, uint32_t code) {
    uint32_t a;
    asm("mad.lo.u32 %0, %1, %3, %2;" : "=r"(a) : "r"(code), "r"(lut_sa), "n"(lut_pitch_words(GEN) * 4));
    return lds_u32(a);
}
// Addressing of the view gather: the shared-memory offset of grid cell (x,y) inside a lane's column is
// offx(x) + offy(y) with offx(x) = x*HP*32 and offy(y) = (y>>2)*128 + (y&3).  Out-of-grid coordinates map to
// the offset of a pad word that holds CODE_WALL; any sum with an out-of-grid term is >= that offset, so one
// min() clamps it onto the pad -- no per-cell bounds test (minigrid.py:465-469).  offx/offy are tabulated
// once per CTA (axis tables).

template <int GEN, bool SEE, int V>
__device__ __forceinline__ void observe(const uint32_t *st, const Env &e, const RolloutParams &p, const uint32_t *lut,
                                        uint32_t *stage_w, int lane) {
    const DevCfg &c = p.cfg;
    constexpr int R = 3 * V * V;            // record bytes (147)
    constexpr int FW = R >> 2;              // full words of a record (36); word FW holds the last R&3 bytes
    constexpr int NG = (V * V + 3) / 4;     // groups of 4 cells (13)
    constexpr int AGENT_CI = (V / 2) * V + (V - 1);   // the agent's own cell (3,6) in output order
    constexpr uint32_t VMASK = (1u << V) - 1u;
    const uint32_t st_sa = (uint32_t)__cvta_generic_to_shared(st);     // 32-bit shared address of the lane's column
    const uint32_t lut_sa = (uint32_t)__cvta_generic_to_shared(lut);
    const int odd = e.dir & 1;
    const int sgn = 1 - (e.dir & 2);                 // +1 for dir 0/1, -1 for dir 2/3
    const int wall_sa = c.S * 128 + (int)st_sa;       // address of the pad word (index S of the lane's column)
    // world(vx,vy) = agent + d*(6-vy) + r*(vx-3), d = DIR_TO_VEC[dir], r = (-d.y, d.x)   
    //   even dir: x = ax + sgn*(V-1-vy) (rows)    y = ay + sgn*(vx-V/2) (columns)
    //   odd  dir: x = ax - sgn*(vx-V/2) (columns) y = ay + sgn*(V-1-vy) (rows)
    // P[vx] / Q[vy] = shared-memory offsets of those coordinates, read from the CTA's axis tables
    const uint32_t ax_sa = (uint32_t)__cvta_generic_to_shared(lut) + lut_bytes(GEN);      // table of x offsets
    const uint32_t ay_sa = ax_sa + AXIS_ENTRIES * 4;                                 // table of y offsets
    const int p0 = odd ? e.ax + (V / 2) * sgn : e.ay - (V / 2) * sgn, pstep4 = (odd ? -sgn : sgn) * 4;
    const int q6 = odd ? e.ay : e.ax, qstep4 = sgn * 4;               // row vy = V-1 is the agent's own row
    const uint32_t pa = (odd ? ax_sa : ay_sa) + (uint32_t)(p0 + AXIS_BIAS) * 4;
    const uint32_t qa = (odd ? ay_sa : ax_sa) + (uint32_t)(q6 + AXIS_BIAS) * 4;
    int P[V], Q[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        P[k] = (int)lds_u32(pa + k * pstep4) + (int)st_sa;            // P carries the column base address
        Q[k] = (int)lds_u32(qa + (V - 1 - k) * qstep4);
    }
    const uint32_t own = e.carry ? lds_u32(lut_sa + (uint32_t)e.carry * (lut_pitch_words(GEN) * 4)) : (uint32_t)T_EMPTY;   // minigrid.py:1349-1356

    // Realignment of the record to byte offset lane*147 of the warp's 4704-byte block: block word q+j takes the
    // high bytes of record word j-1 and the low bytes of record word j -- one funnel shift per word.
    const int boff = lane * R;
    const int q = boff >> 2;
    const uint32_t s8 = (boff & 3) * 8;
    uint32_t first = 0, w36 = 0, w37 = 0, spill = 0;
    auto emit = [&](int j, uint32_t a) {
        const uint32_t o = __funnelshift_l(spill, a, s8);      // spill holds the previous un-shifted word
        spill = a;
        if (j == 0) first = o;
        else if (j < FW) stage_w[q + j] = o;
        else if (j == FW) w36 = o;
        else w37 = o;
    };

    if (SEE) {
        // no occlusion: stream cells in output order (vx-major); 4 cells (3 bytes each) -> 3 words.  The four cell
        // loads of a group are issued before the first dependent LUT load, and all LUT loads before the pack.
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            uint32_t code[4], x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ci = g * 4 + i;
                code[i] = 0;
                if (ci < V * V && ci != AGENT_CI) {
                    const int vx = ci / V, vy = ci % V;
                    code[i] = lds_u8((uint32_t)min(P[vx] + Q[vy], wall_sa));
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ci = g * 4 + i;
                x[i] = 0;
                if (ci < V * V) x[i] = (ci == AGENT_CI) ? own : lut_ld<GEN>(lut_sa, code[i]);
            }
            emit(g * 3, __byte_perm(x[0], x[1], 0x4210));                              // x0.b0 x0.b1 x0.b2 x1.b0
            if (g * 3 + 1 <= FW + 1) emit(g * 3 + 1, __byte_perm(x[1], x[2], 0x5421));     // x1.b1 x1.b2 x2.b0 x2.b1
            if (g * 3 + 2 <= FW + 1) emit(g * 3 + 2, __byte_perm(x[2], x[3], 0x6542));     // x2.b2 x3.b0 x3.b1 x3.b2
        }
    } else {
        // (A) all V*V cell + LUT loads, independent of each other, into registers in output order; (B) the flood on
        // the V row masks, invisible cells zeroed in place; (C) the streaming pack of the see-through path.
        uint32_t xs[V * V + 3], opq[V];
        xs[V * V] = xs[V * V + 1] = xs[V * V + 2] = 0;
#pragma unroll
        for (int vy = V - 1; vy >= 0; --vy) {
            uint32_t code[V];
#pragma unroll
            for (int vx = V - 1; vx >= 0; --vx) code[vx] = lds_u8((uint32_t)min(P[vx] + Q[vy], wall_sa));
            uint32_t o = 0;
#pragma unroll
            for (int vx = V - 1; vx >= 0; --vx) {
                // one 32-bit word per cell: 24-bit encoding + flags, "opaque" in bit 31 (byte 3 is never selected by the PRMTs below)
                const uint32_t x = lut_ld<GEN>(lut_sa, code[vx]);
                xs[vx * V + vy] = x;
                o = __funnelshift_l(x, o, 1);                             // (o << 1) | (x >> 31)
            }
            opq[vy] = o;
        }
        xs[AGENT_CI] = own;
        // Both sweeps of a row (process_vis, minigrid.py:617-648) with one carry chain: the word holds the row in bits
        // 0..V-1 and, bit-reversed, in bits 31..32-V, so the "towards higher bits" fill ((v&t)+t)^t runs up the row in
        // the low field and down the row in the high field at once; or-ing the word with its own reversal merges the two.
        // The sweeps compute the closure of "a visible see-through cell shows both neighbours", which is the union of the
        // two fills (tests/test_kernel_algebra.py).  Bits V and 31-V collect carries / shifted-out seeds; no mask has
        // them and no test reads them.
        uint32_t vw = (1u << (V / 2)) | (0x80000000u >> (V / 2));            // mask[(3,6)] = True (minigrid.py:619)
#pragma unroll
        for (int vy = V - 1; vy >= 0; --vy) {
            const uint32_t t = ~opq[vy] & VMASK;
            const uint32_t tw = t | __brev(t);
            const uint32_t fw = (((vw & tw) + tw) ^ tw) | vw;
            const uint32_t vis = fw | __brev(fw);
            const uint32_t sw = vis & tw;
            vw = sw | (sw << 1) | (sw >> 1);                               // seeds of row vy-1
#pragma unroll
            for (int vx = 0; vx < V; ++vx) xs[vx * V + vy] = sel_bit(vis, 1u << vx, xs[vx * V + vy]);
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const uint32_t *y = &xs[g * 4];
            emit(g * 3, __byte_perm(y[0], y[1], 0x4210));
            if (g * 3 + 1 <= FW + 1) emit(g * 3 + 1, __byte_perm(y[1], y[2], 0x5421));
            if (g * 3 + 2 <= FW + 1) emit(g * 3 + 2, __byte_perm(y[2], y[3], 0x6542));
        }
    }
    // the partial last word of lane t-1 shares a 32-bit word with the head of lane t
    const uint32_t tail = (s8 >= 16) ? w37 : w36;
    const uint32_t ptail = __shfl_up_sync(0xFFFFFFFFu, tail, 1);
    if (s8 != 0 && lane > 0) first |= ptail;
    stage_w[q] = first;
    if (s8 != 0) stage_w[q + FW] = w36;
}

// ------------------------------------------------------------------------------------------
// TMA bulk store of the staged observation block
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_copy(void *gptr, const void *sptr, uint32_t bytes) {
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sptr);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gptr), "r"(saddr), "r"(bytes) : "memory");
}
// state block HBM -> shared memory with one bulk copy that signals the warp's mbarrier
__device__ __forceinline__ void mbar_init(uint32_t mbar_sa, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_sa), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst_sa, const void *gptr, uint32_t bytes, uint32_t mbar_sa) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_sa), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_sa), "l"(gptr), "r"(bytes), "r"(mbar_sa) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar_sa, uint32_t parity) {
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @!p bra W; }" ::"r"(mbar_sa), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 32 steps of a group's actions.  Row t of the [T][N] action array holds the 32 bytes of the group at
// actions + t*stride + group*32: lane t fetches row t0+t (two 16-byte loads when rows are aligned, bytes otherwise),
// parks it in the warp's staging block (idle between two observations), and every lane then gathers its own column
// into 32 x 4 bits (values >= 15 stay invalid: n_actions <= 9).  Issue and store/pack are separate so that the
// state block's load can wait in between.
struct ActionRow { uint4 lo, hi; };
__device__ __forceinline__ void actions_issue(ActionRow &r, const uint8_t *row, bool mine, bool fast) {
    r.lo = r.hi = make_uint4(0, 0, 0, 0);
    if (mine && fast) {
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.lo.x), "=r"(r.lo.y), "=r"(r.lo.z), "=r"(r.lo.w) : "l"(row) : "memory");
        asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.hi.x), "=r"(r.hi.y), "=r"(r.hi.z), "=r"(r.hi.w) : "l"(row + 16) : "memory");
    }
}
__device__ __forceinline__ void actions_pack(const ActionRow &r, const uint8_t *row, bool mine, bool fast, int nvalid, uint32_t *stage_w, int lane,
                                             uint32_t &q0, uint32_t &q1, uint32_t &q2, uint32_t &q3) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the staging block is free
    __syncwarp();
    if (mine) {
        if (fast) {
            reinterpret_cast<uint4 *>(stage_w)[lane * 2] = r.lo;
            reinterpret_cast<uint4 *>(stage_w)[lane * 2 + 1] = r.hi;
        } else {
            uint8_t *sb = reinterpret_cast<uint8_t *>(stage_w) + lane * 32;
            for (int k = 0; k < nvalid; ++k) sb[k] = row[k];
        }
    }
    __syncwarp();
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(stage_w) + (uint32_t)lane;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 32; ++k) w[k >> 3] |= min(lds_u8(sa + k * 32), 15u) << ((k & 7) * 4);
    q0 = w[0]; q1 = w[1]; q2 = w[2]; q3 = w[3];
    __syncwarp();
}
// One bulk copy (one HBM latency) instead of ceil(S/16) dependent batches of LDG -> STS.  The previous group's
// generic-proxy accesses to the block are ordered before the async-proxy write by the fence.  Once per
// group.
__device__ __forceinline__ void load_state_block(uint32_t dst_sa, const uint32_t *src, uint32_t bytes, uint32_t mbar_sa, uint32_t phase, int lane) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the previous group's write-back has been read
        bulk_load(dst_sa, src, bytes, mbar_sa);
    }
    mbar_wait(mbar_sa, phase);
}
__device__ __forceinline__ void bulk_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Reset of the lanes with `need` set (auto-reset after a done step, or mgb_reset).  A generator is a long scalar routine:
// run for ONE finished env it costs the warp as much as run for all 32, and with episode ends spread over time (a
// KeyCorridor policy that picks up its target) that is what happens -- 2.9e10 -> 1.7e10 env-steps/s.  So whenever some
// lane has to generate, every lane of the warp that holds no spare layout generates the layout of its own NEXT episode in
// the same pass (it depends only on seed, env id and episode number) and parks it in the env's spare block in HBM; a later
// reset of such a lane is a copy.  A lane that pre-generates parks its live grid in HBM, runs the generator in place in
// shared memory, stores the result and fetches its grid back.  Results are bit-identical to generating at the reset.
#ifdef MGB_DEBUG_SPARES
__device__ unsigned long long g_spare_dbg[8];     // calls, consumed, stale, passes, lanes generating live, lanes generating ahead
#define SPARE_DBG(i, n) atomicAdd(&g_spare_dbg[i], (unsigned long long)(n))
#else
#define SPARE_DBG(i, n)
#endif
// 4-byte asynchronous copies global -> shared (LDGSTS): a lane fetches a whole grid column with every word in flight at
// once and no registers.  It matters: under the kernel's write stream a dependent HBM read takes several microseconds.
__device__ __forceinline__ void cp_async_word(uint32_t dst_sa, const uint32_t *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_sa), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
template <int GEN>
__device__ __noinline__ void reset_with_spares(bool need, bool valid, uint32_t *st, int group, Env &e, Rng &rg,
                                               const RolloutParams &p, uint32_t *scr, const uint32_t *tmpl_s) {
    const int GW = p.cfg.GW;
    const bool spares = p.spare != nullptr && p.tape == nullptr;
    uint32_t *const spc = p.spare + (size_t)group * spare_words(GW) * 32 + (threadIdx.x & 31);      // words 0..GW+4: the spare
    uint32_t *const park = spc + (size_t)(GW + SPARE_XW) * 32;                                      // words GW+5..: parking rows
    const uint32_t st_sa = (uint32_t)__cvta_generic_to_shared(st);
    uint32_t &fw = st[(GW + 1) * 32];                                 // FLAG_SPARE stays in the state word (k_rollout carries it over)
    if ((threadIdx.x & 31) == 0) SPARE_DBG(0, 1);
    const int n_need = __popc(__ballot_sync(0xFFFFFFFFu, need));    // lanes that reset in this call (with or without a spare)
    if (spares && need && (fw & (FLAG_SPARE << 24))) {              // 1. a reset with a spare at hand is a copy: ONE round trip
        fw &= ~((uint32_t)FLAG_SPARE << 24);
        for (int k = 0; k < GW; ++k) cp_async_word(st_sa + (uint32_t)k * 128u, spc + k * 32);
        const uint32_t a = __ldcg(&spc[GW * 32]), tg = __ldcg(&spc[(GW + 1) * 32]), nd = __ldcg(&spc[(GW + 2) * 32]);
        const uint32_t er = __ldcg(&spc[(GW + 3) * 32]), ep = __ldcg(&spc[(GW + 4) * 32]);
        cp_async_wait_all();
        SPARE_DBG(ep == rg.episode + 1u ? 1 : 2, 1);
        if (ep == rg.episode + 1u) {                                  // the layout of exactly the episode that starts now (else: the
            e.ax = a & 0xFF; e.ay = (a >> 8) & 0xFF; e.dir = (a >> 16) & 3; e.carry = 0; e.steps = 0;     // generator below overwrites it)
            e.target = (int)tg; e.dirty = true;
            rg.episode++; rg.ndraws = nd; rg.rblk = 0xFFFFFFFFu; rg.err |= er;
            need = false;
        }
    }
    if (!__any_sync(0xFFFFFFFFu, need)) return;
    // Generating ahead pays when the warp's episodes really end at scattered steps.  A warp in lock step with one straggler
    // (an env that once ended early) would run a 32-lane pass -- ~100 us on one warp of a persistent launch, i.e. on its
    // tail -- for every lone reset of that env: sustained KeyCorridor lost 5 % to it.  So the warp counts its lone resets
    // (two bits of the flags byte, zeroed by a reset of eight or more lanes) and generates ahead from the second on.
    uint32_t lone = (fw >> LONE_SHIFT) & 3u;
    lone = n_need >= 8 ? 0u : min(lone + 1u, 3u);
    fw = (fw & ~(3u << LONE_SHIFT)) | (lone << LONE_SHIFT);
    if (GEN == GEN_PROC && p.cfg.gen == GEN_MULTIROOM) {              // the MultiRoom generator scribbles in the staging block
        if ((threadIdx.x & 31) == 0) bulk_store_wait_read();
        __syncwarp();
    }
    const bool pre = spares && lone >= 2u && valid && !need && !(fw & (FLAG_SPARE << 24));   // 2. somebody generates: so does everybody without a spare
    if ((threadIdx.x & 31) == 0) SPARE_DBG(3, 1);
    if (need) SPARE_DBG(4, 1);
    if (pre) SPARE_DBG(5, 1);
    if (pre) {
#pragma unroll 8
        for (int k = 0; k < GW; ++k) __stcg(&park[k * 32], st[k * 32]);         // the live grid waits in HBM meanwhile
    }
    if (need || pre) {
        Env te = e; Rng tr = rg;
        generate<GEN>(st, te, tr, p, nullptr, scr, tmpl_s);
        if (!pre) { e = te; rg = tr; }
        else {
#pragma unroll 8
            for (int k = 0; k < GW; ++k) __stcg(&spc[k * 32], st[k * 32]);
            __stcg(&spc[GW * 32], (uint32_t)te.ax | ((uint32_t)te.ay << 8) | ((uint32_t)te.dir << 16));
            __stcg(&spc[(GW + 1) * 32], (uint32_t)te.target);
            __stcg(&spc[(GW + 2) * 32], tr.ndraws);
            __stcg(&spc[(GW + 3) * 32], tr.err & ~rg.err);
            __stcg(&spc[(GW + 4) * 32], tr.episode);
            fw |= (uint32_t)FLAG_SPARE << 24;
            __threadfence();                                             // the parked rows were written by this thread, long ago
            for (int k = 0; k < GW; ++k) cp_async_word(st_sa + (uint32_t)k * 128u, park + k * 32);
            cp_async_wait_all();
        }
    }
}
template <int GEN>
__device__ __forceinline__ void reset_lanes(bool need, bool valid, uint32_t *st, int group, Env &e, Rng &rg, PoolCtx &pc,
                                            const RolloutParams &p, uint32_t *scr, const uint32_t *tmpl_s) {
    if (spare_gen(GEN)) {
        if (__any_sync(0xFFFFFFFFu, need)) {
            Env te = e; Rng tr = rg;                                  // copy-in/out keeps e, rg in registers
            reset_with_spares<GEN>(need, valid, st, group, te, tr, p, scr, tmpl_s);
            e = te; rg = tr;
        }
    } else {
        // Empty with a fixed start, grid untouched (it always is): the reset is a handful of register moves.  Out of line it
        // costs a call with its spills and fills -- local memory is an L2 round trip here -- which is what a policy that
        // reaches the goal every dozen steps would pay on nearly every warp-step.
        if (GEN == GEN_EMPTY && need && !p.cfg.random_start && (e.flags & FLAG_PRISTINE)) {
            if (e.flags & FLAG_GOAL_GONE) { cell_wr(st, p.cfg.goal_idx, CODE_GOAL); e.flags &= ~FLAG_GOAL_GONE; }
            e.ax = 1; e.ay = 1; e.dir = 0; e.carry = 0; e.steps = 0; e.target = 0;
            rg.episode++; if (!p.tape) rg.ndraws = 0; rg.rblk = 0xFFFFFFFFu;
            need = false;
        }
        if (!need) return;
        Env te = e; Rng tr = rg; PoolCtx tp = pc;
        generate<GEN>(st, te, tr, p, GEN == GEN_POOL ? &tp : nullptr, scr, tmpl_s);
        e = te; rg = tr; if (GEN == GEN_POOL) pc = tp;
    }
}

// ------------------------------------------------------------------------------------------
// the persistent rollout kernel (also serves reset and single step)
// ------------------------------------------------------------------------------------------
#ifndef MGB_DYN_MIN_BLOCKS
#define MGB_DYN_MIN_BLOCKS 2
#endif
#define MGB_MIN_BLOCKS(GEN) ((GEN) == GEN_DYNOBS ? MGB_DYN_MIN_BLOCKS : 0)     // 0 = no hint
// NOTE: no minBlocksPerSM argument on purpose -- with it ptxas spends up to 157 registers/thread and the
// occupancy loss costs more than it gains (measured: profiles/README.md, A/B table)
template <int GEN, bool SEE, int V>
__global__ void __launch_bounds__(MAX_THREADS, MGB_MIN_BLOCKS(GEN)) k_rollout(
    const __grid_constant__ RolloutParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
#if MGB_PDL
    // programmatic dependent launch: the next launch on the stream may place its CTAs on SMs that this grid has left and run
    // its prologue (tables, template) there; it touches nothing a predecessor wrote before griddepcontrol.wait below
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    constexpr bool PACKED = !SEE;     // rollouts of the occluded kernels hold 32 steps of actions in 4 registers (see below)
    const DevCfg &c = p.cfg;
    // the warp index goes through a lane-0 broadcast so that ptxas knows it is warp-uniform: everything derived from
    // it (group, staging block, output addresses) then lives in uniform registers and the bulk copies below need no
    // per-lane "waterfall" loop around their uniform-register operands
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xFFFFFFFFu, (int)(threadIdx.x >> 5), 0), wpb = blockDim.x >> 5;
    uint32_t *lut = reinterpret_cast<uint32_t *>(smem_raw);                       // 256 words
    uint32_t *axis = lut + lut_bytes(GEN) / 4;                                           // [2][AXIS_ENTRIES]
    uint32_t *tmpl_s = reinterpret_cast<uint32_t *>(smem_raw + table_bytes(GEN));
    uint8_t *stage_base = smem_raw + table_bytes(GEN) + tmpl_smem_bytes(GEN, c.GW);
    if (GEN != GEN_POOL)
        for (int i = threadIdx.x; i < c.GW; i += blockDim.x) tmpl_s[i] = __ldg(&p.tmpl[i]);
    constexpr int SB = stage_bytes(V), OB = obs_bytes(V), SB_OBS = GROUP * OB;
    uint32_t *stage_w = reinterpret_cast<uint32_t *>(stage_base + warp * SB);
    uint32_t *st_warp = reinterpret_cast<uint32_t *>(stage_base + wpb * SB) + warp * ((c.S + 1) * 32);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        lut[i * lut_pitch_words(GEN)] = lut_entry(i);
    }
    for (int i = threadIdx.x; i < AXIS_ENTRIES; i += blockDim.x) {
        const int v = i - AXIS_BIAS, wall = c.S * 128;
        axis[i] = ((unsigned)v < (unsigned)c.W) ? (uint32_t)(v * c.HP * 32) : (uint32_t)wall;                       // x: column pitch
        axis[AXIS_ENTRIES + i] = ((unsigned)v < (unsigned)c.H) ? (uint32_t)(((v >> 2) << 7) + (v & 3)) : (uint32_t)wall;   // y: word + byte
    }
    const uint32_t mbar_sa = (uint32_t)__cvta_generic_to_shared(axis + 2 * AXIS_ENTRIES) + (uint32_t)warp * 8u;
    if (lane == 0) mbar_init(mbar_sa, 1);
    __syncthreads();
#if MGB_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");       // everything before us on the stream has completed and is visible
#endif

    const int S = c.S, GW = c.GW;
    const int64_t stride = p.stride;
    uint32_t phase = 0;
    bool cols_hold_template = false;       // warp-uniform: every column of the warp's state block holds exactly the template grid
    // Groups: a warp's first is fixed; the following ones come from a ticket counter -- the time a group takes varies
    // (rejection sampling, generator passes), and with a fixed stride the launch ends with the unluckiest warp's sum of ~14
    // groups.  The ticket is taken at the top of the group it follows: its latency is hidden.  Every warp that had a first
    // group ends on a ticket beyond the last group; the last such warp to leave zeroes the counter for the next launch (all
    // other warps have taken their last ticket by then), so nothing about it lives on the host: a launch replayed from a CUDA
    // graph finds the counter as a fresh one does.
    const int g_first = blockIdx.x * wpb + warp;
    for (int g = g_first; (unsigned)g < (unsigned)p.n_groups;) {
        int g_next = g + gridDim.x * wpb;
        uint32_t tk = 0;        // Empty rollouts take it as late as its latency allows (MGB_TICKET_LEAD steps before the group's last):
                                // a warp that claims its next group early still holds it when the tickets run out -- a longer tail
        constexpr bool LATE_TICKET = GEN == GEN_EMPTY && SEE;      // measured: Empty +1.8 %; DoorKey -3.4 % (the test in its step loop costs
                                                                    // ten registers), FourRooms / KeyCorridor / Dynamic-Obstacles +-1 %
        if (p.T <= 1 && p.ticket != nullptr && lane == 0) tk = atomicAdd(p.ticket, 1u);     // single steps, resets: at the top
        const int group = p.group0 + g;
        uint32_t *gst = p.state + (size_t)group * S * 32 + lane;
        // The step loop holds no global load: a lane fetches the actions of its env for 32 steps at once (one HBM latency
        // per 32 steps, overlapped with the state block's) and keeps them as 32 x 4 bits in four registers.  With a load
        // per step, ptxas gave it a scoreboard that the first instruction after the transition also waited on: 8-15 % of
        // all stall samples of the occluded kernels (+8-13 % there).  The see-through kernels are bound by the shared-memory
        // pipe, not by that stall, and measured 6 % slower with this scheme: they keep the per-step load.
        ActionRow arow;
        const uint8_t *arow_p = p.actions + (int64_t)lane * p.stride + (int64_t)group * 32;
        const bool afast = ((p.stride & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.actions) & 15) == 0) && ((int64_t)group * 32 + 32 <= p.n_envs);
        if (PACKED && p.T > 1) actions_issue(arow, arow_p, lane < p.T, afast);
        // ---- load the group's state block: S coalesced 128-byte rows -> bank == lane ----
        // Single-step launches of the template-grid kernels (Empty, Dynamic-Obstacles) first fetch only the non-grid rows;
        // when every env of the group is pristine -- the normal case -- the grid rows never cross HBM at all (they are
        // rebuilt below), which cuts the state round trip of a step from S to S-GW rows.
        const uint32_t st_warp_sa = (uint32_t)__cvta_generic_to_shared(st_warp);
        const uint32_t *gblock = p.state + (size_t)group * S * 32;
        if (template_gen(GEN) && p.T <= 1) {
            load_state_block(st_warp_sa + (uint32_t)GW * 128u, gblock + (size_t)GW * 32, (uint32_t)(S - GW) * 128u, mbar_sa, phase, lane);
            phase ^= 1u;
            if (!__all_sync(0xFFFFFFFFu, (st_warp[(GW + 1) * 32 + lane] >> 24) & FLAG_PRISTINE)) {      // rare: an uploaded / edited grid
                load_state_block(st_warp_sa, gblock, (uint32_t)GW * 128u, mbar_sa, phase, lane);
                phase ^= 1u;
                cols_hold_template = false;
            }
        } else {
            load_state_block(st_warp_sa, gblock, (uint32_t)S * 128u, mbar_sa, phase, lane);
            phase ^= 1u;
            cols_hold_template = false;
        }
        st_warp[S * 32 + lane] = (uint32_t)CODE_WALL * 0x01010101u;        // out-of-grid pad (minigrid.py:469)
        // Single-step launches are bound by the latency of the state round trip.  For the Empty kernels (4 rows of state per
        // group once the grid is implied) the rows and actions of the warp's NEXT group are pulled into L2 now, so that its
        // bulk load and action load are L2 hits: +3 %.  Measured and rejected for the kernels with large state blocks, which
        // are HBM-bound in this mode (DoorKey-16x16 -16 %, FourRooms -10 %, Dynamic-Obstacles -1.5 %).
        // (After this group's own load has arrived: by then the ticket naming the next group has too.)
        if (GEN == GEN_EMPTY && p.T <= 1) {
            int gn = g_next;
            if (p.ticket != nullptr) gn = gridDim.x * wpb + (int)__shfl_sync(0xFFFFFFFFu, tk, 0);
            if ((unsigned)gn < (unsigned)p.n_groups) {
                const uint32_t *nb = p.state + (size_t)(p.group0 + gn) * S * 32;
                for (int r = GW + lane; r < S; r += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + (size_t)r * 32));
                if (lane == 0 && p.T > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.actions + (int64_t)(p.group0 + gn) * 32));
            }
        }
        Env e;
        Rng rg;
        PoolCtx pc;
        pc.level = 0; pc.hp0 = pc.hp1 = pc.hp2 = pc.hp3 = 0;
        uint32_t *const st = st_warp + lane;   // this lane's column: word k at st[k*32]; bank == lane
        rg.lid = (int64_t)group * 32 + lane;
        rg.gid = p.env_id_base + rg.lid;
        const int64_t lid = rg.lid;
        const bool valid = lid < p.n_envs;
        {
            const uint32_t w0 = st[(GW + 0) * 32], w1 = st[(GW + 1) * 32];
            e.ax = w0 & 0xFF; e.ay = (w0 >> 8) & 0xFF; e.dir = (w0 >> 16) & 3; e.carry = w0 >> 24;
            e.steps = w1 & 0xFFFF; e.target = (w1 >> 16) & 0xFF; e.flags = (w1 >> 24) & FLAGS_IMPLIED;
            rg.episode = st[(GW + 2) * 32]; rg.ndraws = st[(GW + 3) * 32];
            if (GEN == GEN_POOL) {
                pc.level = (int)st[(GW + XWORDS) * 32];
                if (p.pool_n > 0 && (unsigned)pc.level < (unsigned)p.pool_n) {
                    const uint32_t *src = p.pool + (size_t)pc.level * (GW + POOL_XW) + GW;
                    pc.hp0 = __ldg(&src[1]); pc.hp1 = __ldg(&src[2]); pc.hp2 = __ldg(&src[3]); pc.hp3 = __ldg(&src[4]);
                }
            }
        }
        rg.rblk = 0xFFFFFFFFu; rg.err = 0; e.dirty = false;
        rg.rb0 = rg.rb1 = rg.rb2 = rg.rb3 = 0;
        if (template_gen(GEN)) {
            // the grid rows of a pristine env in HBM are don't-care: rebuild.  An Empty grid IS the template, so once a warp's
            // columns hold it (left there by the previous all-pristine group of a single-step launch) there is nothing to do.
            if (e.flags & FLAG_PRISTINE) {
                if (!(GEN == GEN_EMPTY && cols_hold_template)) rebuild_pristine_grid(st, p, tmpl_s, e.flags);
                else if (e.flags & FLAG_GOAL_GONE) cell_wr(st, c.goal_idx, CODE_EMPTY);
            }
        }

        const bool full = ((int64_t)group * 32 + 32) <= p.n_envs;
        const int nvalid = full ? 32 : (int)max((int64_t)0, p.n_envs - (int64_t)group * 32);

        if (p.do_reset) {
            const bool m = valid && (!p.reset_mask || p.reset_mask[lid]);
            reset_lanes<GEN>(m, valid, st, group, e, rg, pc, p, stage_w + lane, tmpl_s);
        }
        const int nsteps = p.T > 0 ? p.T : 1;
        int a_next = 0;
        uint32_t aq0 = 0, aq1 = 0, aq2 = 0, aq3 = 0;
        if (PACKED) {
            if (p.T > 1) actions_pack(arow, arow_p, lane < p.T, afast, nvalid, stage_w, lane, aq0, aq1, aq2, aq3);
            if (p.T == 1 && valid) a_next = min((int)p.actions[lid], 15);
        } else if (p.T > 0 && valid) a_next = p.actions[lid];
        // loop invariants spelled out: ptxas otherwise re-derives them from the parameter bank on every step
        // (69 of the occluded kernel's 880 instructions per step were this bookkeeping).  HOIST is off for the
        // see-through kernels: at their 64 registers the extra live values cost more than the bookkeeping (measured -4 %).
        constexpr bool HOIST = !SEE;
        const bool stepping = p.T > 0, multi = PACKED && p.T > 1;
        const bool w_rew = valid && stepping && p.reward != nullptr, w_done = valid && stepping && p.done != nullptr;
        const bool w_dir = valid && p.dir != nullptr;
        int64_t o = lid;                                                  // this env's slot in the [t][N] outputs
        uint8_t *gobs = p.obs ? p.obs + (int64_t)group * 32 * OB : nullptr;   // the group's block of step t (warp-uniform)
        const int64_t obs_pitch = stride * OB;
        const int t_ticket = max(nsteps - MGB_TICKET_LEAD, 0);
        for (int t = 0; t < nsteps; ++t) {
            if (LATE_TICKET && p.T > 1 && t == t_ticket && p.ticket != nullptr && lane == 0) tk = atomicAdd(p.ticket, 1u);
            double reward = 0.0; bool done = false;
            if (!HOIST) {
                o = (int64_t)t * stride + lid;
                gobs = p.obs ? p.obs + ((int64_t)t * stride + (int64_t)group * 32) * OB : nullptr;
            }
            if (HOIST ? stepping : p.T > 0) {
                if (HOIST ? multi : (PACKED && p.T > 1)) {
                    if ((t & 31) == 0 && t > 0) {             // rollouts longer than 32 steps: next chunk
                        ActionRow r;
                        const uint8_t *rp = arow_p + (int64_t)t * stride;
                        actions_issue(r, rp, t + lane < p.T, afast);
                        actions_pack(r, rp, t + lane < p.T, afast, nvalid, stage_w, lane, aq0, aq1, aq2, aq3);
                    }
                    a_next = (int)(aq0 & 15u);
                    aq0 = __funnelshift_r(aq0, aq1, 4); aq1 = __funnelshift_r(aq1, aq2, 4); aq2 = __funnelshift_r(aq2, aq3, 4); aq3 >>= 4;
                }
                const int action = a_next;
                if (!PACKED && t + 1 < p.T && valid) a_next = p.actions[(int64_t)(t + 1) * stride + lid];
                if (GEN == GEN_DYNOBS) {                        // the staging block doubles as the draw window
                    if (lane == 0) bulk_store_wait_read();
                    __syncwarp();
                }
                bool need_reset = false;
                if (valid) {
                    transition<GEN, SEE, V>(st, e, rg, p, lut, action, reward, done, stage_w + lane, pc);
                    need_reset = done && p.autoreset;
                }
                if (GEN == GEN_DYNOBS) {                        // frequent resets: regenerate finished envs with the whole warp
                    uint32_t rm = __ballot_sync(0xFFFFFFFFu, need_reset && !p.tape && (e.flags & FLAG_PRISTINE));
                    while (rm) {
                        const int src = __ffs((int)rm) - 1;
                        rm &= rm - 1;
                        uint32_t consumed = 0, agent = 0;
                        const uint32_t stream = __shfl_sync(0xFFFFFFFFu, rg.episode, src);
                        if (dynobs_coop_reset(st_warp, src, lane, p, tmpl_s, p.env_id_base + (int64_t)group * 32 + src, stream, consumed, agent) && lane == src) {
                            e.ax = (int)(agent & 0xFF); e.ay = (int)((agent >> 8) & 0xFF); e.dir = (int)(agent >> 16);
                            e.carry = 0; e.steps = 0; e.target = 0; e.dirty = true;
                            rg.episode++; rg.ndraws = consumed; rg.rblk = 0xFFFFFFFFu;
                            need_reset = false;
                        }
                    }
                }
                reset_lanes<GEN>(need_reset, valid, st, group, e, rg, pc, p, stage_w + lane, tmpl_s);
            }
            if (gobs) {
                if (lane == 0) bulk_store_wait_read();          // previous block has left shared memory
                __syncwarp();
                observe<GEN, SEE, V>(st, e, p, lut, stage_w, lane);
                if (full && ((reinterpret_cast<uintptr_t>(gobs) & 15) == 0)) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { bulk_copy(gobs, stage_w, (uint32_t)SB_OBS); bulk_commit(); }
                } else {                                         // ragged tail group / unaligned base
                    __syncwarp();
                    const uint8_t *sb = reinterpret_cast<const uint8_t *>(stage_w);
                    for (int b = lane; b < nvalid * OB; b += 32) gobs[b] = sb[b];
                    __syncwarp();
                }
                if (HOIST) gobs += obs_pitch;
            }
            if (HOIST) {
                if (w_rew) p.reward[o] = reward;
                if (w_done) p.done[o] = done ? 1 : 0;
                if (w_dir) p.dir[o] = (uint8_t)e.dir;
                o += stride;
            } else if (valid) {
                if (p.T > 0) {
                    if (p.reward) p.reward[o] = reward;
                    if (p.done) p.done[o] = done ? 1 : 0;
                }
                if (p.dir) p.dir[o] = (uint8_t)e.dir;
            }
        }
        if (!LATE_TICKET && p.T > 1 && p.ticket != nullptr && lane == 0) tk = atomicAdd(p.ticket, 1u);     // the other rollouts: after the step loop
        // ---- write the state back ----
        st[(GW + 0) * 32] = (uint32_t)e.ax | ((uint32_t)e.ay << 8) | ((uint32_t)e.dir << 16) | ((uint32_t)e.carry << 24);
        st[(GW + 1) * 32] = (uint32_t)(e.steps & 0xFFFF) | ((uint32_t)e.target << 16) | ((uint32_t)e.flags << 24)
                            | (spare_gen(GEN) ? st[(GW + 1) * 32] & (((uint32_t)FLAG_SPARE << 24) | (3u << LONE_SHIFT)) : 0u);
        st[(GW + 2) * 32] = rg.episode;
        st[(GW + 3) * 32] = rg.ndraws;
        if (GEN == GEN_POOL) st[(GW + XWORDS) * 32] = (uint32_t)pc.level;
        // grid rows go back only for envs whose grid is not implied by the template and the ball list
        const bool any_dirty = __any_sync(0xFFFFFFFFu, e.dirty && !(template_gen(GEN) && (e.flags & FLAG_PRISTINE)));
        const int k0 = any_dirty ? 0 : GW;                            // an untouched grid stays where it is
        if (any_dirty || p.T > 1) {      // a clean single step writes 4-5 rows: the loop is cheaper
            // one bulk copy (rows k0..S-1 are contiguous here and in HBM) instead of a 7-instruction loop per word;
            // load_state_block waits for it to have left shared memory before the block is reused
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_copy(p.state + (size_t)group * S * 32 + (size_t)k0 * 32, st_warp + k0 * 32, (uint32_t)(S - k0) * 128u);
                bulk_commit();
            }
        } else {
            __syncwarp();
            for (int k = k0; k < S; ++k) gst[k * 32] = st_warp[k * 32 + lane];
        }
        if (GEN == GEN_EMPTY) cols_hold_template = __all_sync(0xFFFFFFFFu, (e.flags & FLAGS_IMPLIED) == FLAG_PRISTINE);
        if (rg.err) atomicOr(p.err, rg.err);
        __syncwarp();
        if (p.ticket != nullptr) g_next = gridDim.x * wpb + (int)__shfl_sync(0xFFFFFFFFu, tk, 0);
        g = g_next;
    }
    if (p.ticket != nullptr && lane == 0 && g_first < p.n_groups) {
        const uint32_t takers = (uint32_t)min((int64_t)gridDim.x * wpb, (int64_t)p.n_groups);
        if (atomicAdd(p.ticket + 1, 1u) == takers - 1u) { p.ticket[0] = 0u; p.ticket[1] = 0u; }
    }
    if (lane == 0) bulk_store_wait_all();
}

// ------------------------------------------------------------------------------------------
// state exchange kernels (K3): reference encoding <-> device layout
// ------------------------------------------------------------------------------------------
struct StateIO {
    DevCfg cfg;
    uint32_t *state;
    const uint32_t *tmpl;
    int64_t first, count;
    uint8_t *grid;        // [count][W][H][3]
    uint8_t *aux;         // [count][W][H]
    int32_t *agent;       // [count][4]
    uint8_t *carrying;    // [count][3]
    int16_t *obstacles;   // [count][8][2]
    uint8_t *target;      // [count][2]
    uint32_t *rng;        // [count][2]
    uint32_t *err;
};

__device__ __forceinline__ int encode_cell(int t, int c, int s, int auxbits, uint32_t &err) {
    if (t > T_LAVA || c > 6 || s > 2) { err |= ERR_CODE; return CODE_EMPTY; }
    // 'unseen' and 'empty' both decode to None (minigrid.py:124-125).  Written as one range test on
    // purpose: the `if (t==0) t=1; ... if (t==1)` form made ptxas 12.9 emit VIMNMX.U16x2 with a
    // predicate output whose sense came out inverted on sm_100a (every cell decoded as empty).
    if (t <= T_EMPTY) return CODE_EMPTY;
    if (t != T_DOOR) s = 0;                               // WorldObj.decode ignores state for non-doors
    if (t == T_GOAL && (auxbits & 1)) return CODE_TGOAL0 + c;
    if (t == T_BOX && ((auxbits >> 1) & 7)) {           // Box.contains = Key(colour k): aux bits 1-3 = k+1; only grey boxes
        if (c != C_GREY || ((auxbits >> 1) & 7) > 7) { err |= ERR_CODE; return CODE_EMPTY; }
        return CODE_KEYBOX0 + ((auxbits >> 1) & 7) - 1;
    }
    return code_of(t, c, s);
}

// Before a partial upload: the grid rows of pristine envs in [first, first+count) become real (template + balls) and the
// envs stop being pristine, so that whatever k_set_state leaves untouched is what a reader saw before.  One thread per env.
__global__ void k_materialize(DevCfg c, uint32_t *state, const uint32_t *__restrict__ tmpl, int64_t first, int64_t count) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= count || !template_gen(c.gen)) return;
    const int64_t env = first + n;
    uint32_t *base = state + (env >> 5) * c.S * 32 + (env & 31);
    const uint32_t fw = base[(c.GW + 1) * 32];
    if (!((fw >> 24) & FLAG_PRISTINE)) return;
    for (int k = 0; k < c.GW; ++k) base[k * 32] = grid_word(c, base, tmpl, k);
    base[(c.GW + 1) * 32] = fw & ~((uint32_t)FLAGS_IMPLIED << 24);
}

// clears flag bits of every env (mgb_seed: another seed invalidates the pre-generated layouts)
__global__ void k_clear_flags(uint32_t *state, int S, int GW, int64_t n, uint32_t bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    state[((i >> 5) * S + GW + 1) * 32 + (i & 31)] &= ~(bits << 24);
}

// one thread per (env, state word)
__global__ void k_set_state(const StateIO io) {
    const DevCfg &c = io.cfg;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / c.S;
    const int k = (int)(tid % c.S);
    if (n >= io.count) return;
    const int64_t env = io.first + n;
    uint32_t *dst = io.state + ((env >> 5) * c.S + k) * 32 + (env & 31);
    uint32_t err = 0;
    const int cells = c.W * c.H;
    const int HW = c.HP >> 2;
    if (k < c.GW) {
        if (!io.grid) return;
        uint32_t w = 0;
        const int x = k / HW, y0 = (k % HW) * 4;
        for (int b = 0; b < 4; ++b) {
            int code = CODE_EMPTY;
            if (y0 + b < c.H) {
                const size_t idx = (size_t)n * cells + x * c.H + y0 + b;     // Grid.encode: [x][y]
                const uint8_t *g = io.grid + idx * 3;
                code = encode_cell(g[0], g[1], g[2], io.aux ? io.aux[idx] : 0, err);
            }
            w |= (uint32_t)code << (8 * b);
        }
        *dst = w;
    } else if (k == c.GW) {
        uint32_t w = *dst;
        if (io.agent) {
            const int32_t *a = io.agent + n * 4;
            if (a[0] < 0 || a[0] >= c.W || a[1] < 0 || a[1] >= c.H || a[2] < 0 || a[2] > 3) err |= ERR_BOUNDS;
            w = (w & 0xFF000000u) | (uint32_t)(a[0] & 0xFF) | ((uint32_t)(a[1] & 0xFF) << 8) | ((uint32_t)(a[2] & 3) << 16);
        }
        if (io.carrying) {
            const uint8_t *q = io.carrying + n * 3;
            int code = 0;
            if (q[0] != 0) { code = encode_cell(q[0], q[1], q[0] == T_BOX ? 0 : q[2], q[0] == T_BOX ? q[2] : 0, err); if (code == CODE_EMPTY) code = 0; }
            w = (w & 0x00FFFFFFu) | ((uint32_t)code << 24);
        }
        *dst = w;
    } else if (k == c.GW + 1) {
        uint32_t w = *dst;
        if (io.grid || io.obstacles) w &= ~((uint32_t)FLAGS_IMPLIED << 24);      // an uploaded grid / ball list ends 'template + balls' (state_io materialises the grid first)
        if (io.rng) w &= ~((uint32_t)FLAG_SPARE << 24);        // another episode counter: the pre-generated next layout is not this env's any more
        if (io.agent) w = (w & 0xFFFF0000u) | (uint32_t)(io.agent[n * 4 + 3] & 0xFFFF);
        if (io.target) {
            const uint8_t *q = io.target + n * 2;
            const int code = q[0] ? code_of(q[0], q[1], 0) : 0;
            w = (w & 0xFF00FFFFu) | ((uint32_t)code << 16);
        }
        *dst = w;
    } else if (k == c.GW + 2) {
        if (io.rng) *dst = io.rng[n * 2];
    } else if (k == c.GW + 3) {
        if (io.rng) *dst = io.rng[n * 2 + 1];
    } else {
        if (!io.obstacles) return;
        const int o = (k - c.GW - XWORDS) * 2;
        const int16_t *q = io.obstacles + (n * MAX_OBST + o) * 2;
        *dst = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
    }
    if (err) atomicOr(io.err, err);
}

// one thread per (env, state word); full_obs != 0: FullyObsWrapper (wrappers.py:311-338)
__global__ void k_get_state(const StateIO io, int full_obs) {
    const DevCfg &c = io.cfg;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = tid / c.S;
    const int k = (int)(tid % c.S);
    if (n >= io.count) return;
    const int64_t env = io.first + n;
    const uint32_t *base = io.state + (env >> 5) * c.S * 32 + (env & 31);
    const uint32_t w = k < c.GW ? grid_word(c, base, io.tmpl, k) : base[k * 32];
    const int cells = c.W * c.H;
    const int HW = c.HP >> 2;
    if (k < c.GW) {
        int aidx = -1, adir = 0;
        if (full_obs) {
            const uint32_t w0 = base[c.GW * 32];
            aidx = (int)(w0 & 0xFF) * c.H + (int)((w0 >> 8) & 0xFF);
            adir = (w0 >> 16) & 3;
        }
        const int cx = k / HW, y0 = (k % HW) * 4;
        for (int b = 0; b < 4; ++b) {
            if (y0 + b >= c.H) break;
            const int idx = cx * c.H + y0 + b;                                // Grid.encode: [x][y]
            const uint32_t x = lut_entry((w >> (8 * b)) & 0xFF);
            if (io.grid) {
                uint8_t *g = io.grid + ((size_t)n * cells + idx) * 3;
                if (idx == aidx) { g[0] = T_AGENT; g[1] = 0; g[2] = (uint8_t)adir; }
                else { g[0] = x & 0xFF; g[1] = (x >> 8) & 0xFF; g[2] = (x >> 16) & 0xFF; }
            }
            const uint32_t cc = (w >> (8 * b)) & 0xFF;
            if (io.aux) io.aux[(size_t)n * cells + idx] = (((x >> 24) & F_TGOAL) ? 1 : 0) | (cc >= CODE_KEYBOX0 && cc < CODE_KEYBOX0 + 7 ? ((cc - CODE_KEYBOX0 + 1) << 1) : 0);
        }
    } else if (k == c.GW) {
        if (io.agent) { int32_t *a = io.agent + n * 4; a[0] = w & 0xFF; a[1] = (w >> 8) & 0xFF; a[2] = (w >> 16) & 3; }
        if (io.carrying) {
            uint8_t *q = io.carrying + n * 3;
            const uint32_t code = w >> 24;
            const uint32_t x = code ? lut_entry(code) : 0;
            q[0] = x & 0xFF; q[1] = (x >> 8) & 0xFF; q[2] = (x >> 16) & 0xFF;
            // a carried box keeps its contents: the (otherwise always 0) state byte carries the aux bits
            if (code >= CODE_KEYBOX0 && code < CODE_KEYBOX0 + 7) q[2] = (uint8_t)((code - CODE_KEYBOX0 + 1) << 1);
        }
    } else if (k == c.GW + 1) {
        if (io.agent) io.agent[n * 4 + 3] = w & 0xFFFF;
        if (io.target) {
            const uint32_t code = (w >> 16) & 0xFF;
            const uint32_t x = code ? lut_entry(code) : 0;
            io.target[n * 2] = x & 0xFF; io.target[n * 2 + 1] = (x >> 8) & 0xFF;
        }
    } else if (k == c.GW + 2) {
        if (io.rng) io.rng[n * 2] = w;
    } else if (k == c.GW + 3) {
        if (io.rng) io.rng[n * 2 + 1] = w;
    } else {
        if (!io.obstacles) return;
        const int o = (k - c.GW - XWORDS) * 2;
        int16_t *q = io.obstacles + (n * MAX_OBST + o) * 2;
        q[0] = w & 0xFF; q[1] = (w >> 8) & 0xFF; q[2] = (w >> 16) & 0xFF; q[3] = (w >> 24) & 0xFF;
    }
}

// level pool upload: one thread per (level, pool word); same cell packing as k_set_state
__global__ void k_pack_levels(DevCfg c, int n_levels, const uint8_t *__restrict__ grid, const uint8_t *__restrict__ aux,
                              const int32_t *__restrict__ agent, const int32_t *__restrict__ hookp,
                              uint32_t *__restrict__ pool, uint32_t *err_out) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int PW = c.GW + POOL_XW;
    const int64_t lvl = tid / PW;
    const int k = (int)(tid % PW);
    if (lvl >= n_levels) return;
    uint32_t err = 0, w = 0;
    const int cells = c.W * c.H, HW = c.HP >> 2;
    if (k < c.GW) {
        const int x = k / HW, y0 = (k % HW) * 4;
        for (int b = 0; b < 4; ++b) {
            int code = CODE_EMPTY;
            if (y0 + b < c.H) {
                const size_t idx = (size_t)lvl * cells + x * c.H + y0 + b;
                const uint8_t *g = grid + idx * 3;
                code = encode_cell(g[0], g[1], g[2], aux ? aux[idx] : 0, err);
            }
            w |= (uint32_t)code << (8 * b);
        }
    } else if (k == c.GW) {
        const int32_t *a = agent + lvl * 3;
        if (a[0] < 0 || a[0] >= c.W || a[1] < 0 || a[1] >= c.H || a[2] < 0 || a[2] > 3) err |= ERR_BOUNDS;
        w = (uint32_t)(a[0] & 0xFF) | ((uint32_t)(a[1] & 0xFF) << 8) | ((uint32_t)(a[2] & 3) << 16);
    } else if (hookp) {
        const int32_t *q = hookp + lvl * 16;
        auto b = [&](int i) { return (uint32_t)(q[i] & 0xFF); };
        const int j = k - c.GW;
        if (j == 1) w = (q[0] ? (uint32_t)code_of(q[0], q[1], 0) : 0u) | ((q[2] ? (uint32_t)code_of(q[2], q[3], 0) : 0u) << 8);
        else if (j == 2) w = b(4) | (b(5) << 8) | (b(6) << 16) | (b(7) << 24);
        else if (j == 3) w = b(8) | (b(9) << 8) | (b(10) << 16) | (b(11) << 24);
        else w = b(12) | (b(13) << 8);
    }
    pool[lvl * PW + k] = w;
    if (err) atomicOr(err_out, err);
}

__global__ void k_levels(uint32_t *state, int S, int word, int64_t n, int32_t *out, const int32_t *in) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t *w = state + ((i >> 5) * S + word) * 32 + (i & 31);
    if (in) *w = (uint32_t)in[i];
    if (out) out[i] = (int32_t)*w;
}

// ------------------------------------------------------------------------------------------
// K4: observation-wrapper kernels (stateless, HBM-bound elementwise/gather)
// ------------------------------------------------------------------------------------------
struct ClassMap { uint8_t m[16]; };

// one thread per output byte; out[cell][bit]
__global__ void k_onehot(const uint8_t *__restrict__ cells, uint8_t *__restrict__ out, int64_t n_cells,
                         ClassMap cm, int n_classes, int n_colors, int n_states) {
    const int nbits = n_classes + n_colors + n_states;
    const int64_t total = n_cells * nbits;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = i / nbits;
        const int bit = (int)(i - cell * nbits);
        const uint8_t *q = cells + cell * 3;
        const int t = cm.m[q[0] & 15], c = q[1], st = q[2];
        const bool on = bit == t || (n_colors > 0 && bit == n_classes + c) || bit == n_classes + n_colors + st;
        out[i] = on ? 1 : 0;
    }
}

// one thread per output float; out[n][img_bytes + mission_len]
__global__ void k_flat_obs(const uint8_t *__restrict__ img, int img_bytes, const float *__restrict__ table, int mlen,
                           const uint8_t *__restrict__ midx, float *__restrict__ out, int64_t N) {
    const int row = img_bytes + mlen;
    const int64_t total = N * row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / row;
        const int k = (int)(i - n * row);
        float v;
        if (k < img_bytes) v = (float)img[n * img_bytes + k];
        else v = __ldg(&table[(size_t)(midx ? midx[n] : 0) * mlen + (k - img_bytes)]);
        out[i] = v;
    }
}

// RGB wrappers: one thread copies one tile row (tile*3 bytes, a multiple of 8) from the atlas to the image.
// Consecutive threads write consecutive segments of an image row, so the output stream is fully coalesced;
// the atlas (<= 1.3 MB) stays in L2/L1.
constexpr int ATLAS_VARIANTS = 10;   // 0 plain, 1 highlight, 2..5 agent dir 0..3, 6 agent dir 3 + highlight, 7..9 agent dir 0..2 + highlight
__device__ __forceinline__ void copy_tile_row(const uint8_t *__restrict__ atlas, int tile, int tile_id, int py, uint8_t *dst) {
    const uint2 *src = reinterpret_cast<const uint2 *>(atlas + ((size_t)tile_id * tile + py) * tile * 3);
    uint2 *d = reinterpret_cast<uint2 *>(dst);
    for (int k = 0; k < tile * 3 / 8; ++k) d[k] = __ldg(&src[k]);
}

// IDX = uint32_t whenever the launch has < 2^31 tile rows (the common case): the three div/mods by run-time
// divisors are several times cheaper in 32 bits, and they are what bounds these kernels, not HBM.
template <typename IDX>
__global__ void k_render_partial(const uint8_t *__restrict__ obs, int V, const uint8_t *__restrict__ atlas, int tile,
                                 uint8_t *__restrict__ out, int64_t N) {
    const IDX rows = (IDX)(V * tile), uV = (IDX)V, utile = (IDX)tile;
    const IDX total = (IDX)N * rows * uV;                       // (env, pixel row, cell column)
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const IDX r = i / uV;
        const int cx = (int)(i - r * uV);
        const IDX n = r / rows;
        const int row = (int)(r - n * rows);
        const int cy = row / (int)utile, py = row - cy * (int)utile;
        const uint8_t *q = obs + (((size_t)n * V + cx) * V + cy) * 3;   // obs[n][vx][vy][c]
        const int t = q[0] > 9 ? 0 : q[0], c = q[1] > 6 ? 0 : q[1], st = q[2] > 2 ? 0 : q[2];
        const int code = t * 21 + c * 3 + st;
        int variant = t != 0 ? 1 : 0;                           // vis_mask = (type != unseen) (minigrid.py:613)
        if (cx == V / 2 && cy == V - 1) variant = 6;            // agent_pos=(V//2, V-1), agent_dir=3 (minigrid.py:1391-1396)
        copy_tile_row(atlas, tile, code * ATLAS_VARIANTS + variant, py, out + (size_t)i * tile * 3);
    }
}

// obs != NULL: MiniGridEnv.render(highlight=True) (minigrid.py:1415-1450): a cell is highlighted iff it lies in the
// agent's view and is visible there; visibility is read off the partial observation (type != unseen, as Grid.decode
// does, minigrid.py:613), view cell (vx,vy) of world cell p being  vx = (p-agent).r + V/2,  vy = V-1 - (p-agent).f.
template <typename IDX>
__global__ void k_render_full(DevCfg c, const uint32_t *__restrict__ state, const uint32_t *__restrict__ tmpl, const uint8_t *__restrict__ obs, int V,
                              const uint8_t *__restrict__ atlas, int tile, uint8_t *__restrict__ out, int64_t N) {
    const IDX rows = (IDX)(c.H * tile), uW = (IDX)c.W;
    const IDX total = (IDX)N * rows * uW;
    for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IDX)gridDim.x * blockDim.x) {
        const IDX r = i / uW;
        const int cx = (int)(i - r * uW);
        const IDX n = r / rows;
        const int row = (int)(r - n * rows);
        const int cy = row / tile, py = row - cy * tile;
        const uint32_t *base = state + (size_t)(n >> 5) * c.S * 32 + (n & 31);
        const int cidx = cx * c.HP + cy;
        const uint32_t cc = (grid_word(c, base, tmpl, cidx >> 2) >> ((cidx & 3) * 8)) & 0xFF;
        const uint32_t x = lut_entry((int)cc);                  // (type, colour, state) of the real object
        const int code = (int)(x & 0xFF) * 21 + (int)((x >> 8) & 0xFF) * 3 + (int)((x >> 16) & 0xFF);
        const uint32_t w0 = base[c.GW * 32];
        const int ax = (int)(w0 & 0xFF), ay = (int)((w0 >> 8) & 0xFF), dir = (int)((w0 >> 16) & 3);
        const bool agent_here = ax == cx && ay == cy;
        bool hl = false;
        if (obs) {
            const int fx = (dir & 1) ? 0 : 1 - dir, fy = (dir & 1) ? 2 - dir : 0;      // DIR_TO_VEC; right = (-fy, fx)
            const int rx = cx - ax, ry = cy - ay;
            const int vx = rx * -fy + ry * fx + V / 2, vy = V - 1 - (rx * fx + ry * fy);
            if ((unsigned)vx < (unsigned)V && (unsigned)vy < (unsigned)V) hl = obs[(((size_t)n * V + vx) * V + vy) * 3] != 0;
        }
        const int variant = agent_here ? (hl ? (dir == 3 ? 6 : 7 + dir) : 2 + dir) : (hl ? 1 : 0);
        copy_tile_row(atlas, tile, code * ATLAS_VARIANTS + variant, py, out + (size_t)i * tile * 3);
    }
}

// Uniform random policy (run_tests.py:43, benchmark.py:27-33: env.action_space.sample()) as a counter-based stream, so
// that a rollout needs no action input from the host: action of step t of the `epoch`-th random rollout of this handle
// for global env id g = mulhi32(Philox4x32-10(counter (t>>2, epoch, g lo, g hi), key (seed lo, seed hi ^ "ACT1"))[t&3],
// n_actions).  Filled by its own small kernel into the [T][N] action array the rollout kernel then reads: the hot kernel
// does not change.  One thread = one Philox block = 4 consecutive steps of one env.
constexpr uint32_t ACTION_KEY = 0x41435431u;
__global__ void k_policy_actions(uint8_t *__restrict__ actions, int64_t N, int T, uint64_t seed, int64_t env_id_base, uint32_t epoch,
                                 int n_actions) {
    const int64_t total = N * ((T + 3) / 4);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t blk = i / N, n = i - blk * N, gid = env_id_base + n;
        uint32_t o[4];
        philox4x32_10((uint32_t)blk, epoch, (uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), (uint32_t)seed, (uint32_t)(seed >> 32) ^ ACTION_KEY,
                      o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (blk * 4 + k < T) actions[(blk * 4 + k) * N + n] = (uint8_t)__umulhi(o[k], (uint32_t)n_actions);
    }
}

// ------------------------------------------------------------------------------------------
// bookkeeping wrappers: DACWrapper, ActionBonus / StateBonus, AppendActionWrapper,
// GoalPolicyWrapper (reference wrappers.py:35-154,418-526).  Small element-wise kernels on the outputs of a step.
// ------------------------------------------------------------------------------------------

// ActionBonus (wrappers.py:87-119, key (pos, dir, action)) and StateBonus (:121-154, key pos): the count table of
// env n is counts[n][table]; bonus = 1 / math.sqrt(new_count) in fp64 (sqrt and division correctly rounded, as CPython).
__global__ void k_visit_bonus(DevCfg c, const uint32_t *__restrict__ state, int by_action, const uint8_t *__restrict__ actions,
                              uint32_t *__restrict__ counts, int64_t table, double *__restrict__ reward, int64_t N, uint32_t *err) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t w0 = state[(size_t)(n >> 5) * c.S * 32 + (size_t)c.GW * 32 + (n & 31)];
        const int x = w0 & 0xFF, y = (w0 >> 8) & 0xFF, dir = (w0 >> 16) & 3;
        int64_t key = (int64_t)x * c.H + y;
        if (by_action) {
            const int a = actions[n];
            if (a >= c.n_actions) { atomicOr(err, ERR_ACTION); continue; }
            key = (key * 4 + dir) * c.n_actions + a;
        }
        if (key >= table) { atomicOr(err, ERR_BOUNDS); continue; }
        const uint32_t cnt = ++counts[n * table + key];
        reward[n] = __dadd_rn(reward[n], __ddiv_rn(1.0, __dsqrt_rn((double)cnt)));
    }
}

// DACWrapper.step (wrappers.py:56-77).  One thread per 4 bytes of the flat observation array blanks (image*0+1,
// :51-53) the envs whose episode is over; the first N threads also update the per-env scalars.  `env_done` is
// double-buffered (in -> out) so that the two roles do not race.
__global__ void k_dac(int64_t N, int ob, int count_ge_max, const uint8_t *__restrict__ done_in, const uint8_t *__restrict__ envdone_in,
                      uint8_t *__restrict__ envdone_out, const uint8_t *__restrict__ reset_dir, uint8_t *__restrict__ obs,
                      double *__restrict__ reward, uint8_t *__restrict__ done_out, uint8_t *__restrict__ dir) {
    const int64_t bytes = N * ob, words = (bytes + 3) >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < N) {
            const bool was = envdone_in[i] != 0, now = was || done_in[i] != 0;
            if (was) reward[i] = 0.0;                                  // `return self.last_obs, 0, ...` (:62-65)
            envdone_out[i] = now;
            done_out[i] = now && count_ge_max;                         // done only when the time is up (:62,74)
            if (now) dir[i] = reset_dir[i];                            // last_obs keeps the reset-time 'direction' (:48-51)
        }
        const int64_t b0 = i * 4;
        const int64_t e0 = b0 / ob, e1 = min((b0 + 3) / ob, N - 1);
        const bool f0 = envdone_in[e0] | done_in[e0], f1 = envdone_in[e1] | done_in[e1];
        if (!f0 && !f1) continue;
        if (b0 + 4 <= bytes && f0 && f1 && (reinterpret_cast<uintptr_t>(obs) & 3) == 0) {
            reinterpret_cast<uint32_t *>(obs)[i] = 0x01010101u;
        } else {
            for (int k = 0; k < 4 && b0 + k < bytes; ++k) {
                const int64_t e = (b0 + k) / ob;
                if (envdone_in[e] | done_in[e]) obs[b0 + k] = 1;
            }
        }
    }
}

// AppendActionWrapper (wrappers.py:418-458): hist[n][K] holds the last K action indices (255 = the all-zero vector
// a fresh deque holds, :429,438); a finished env starts its next episode with an empty history (reset(), :436-444).
__global__ void k_action_history(int64_t N, int K, const uint8_t *__restrict__ actions, const uint8_t *__restrict__ done,
                                 uint8_t *__restrict__ hist) {
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        uint8_t *h = hist + n * K;
        if (!actions || (done && done[n])) { for (int k = 0; k < K; ++k) h[k] = 255; continue; }
        for (int k = 0; k + 1 < K; ++k) h[k] = h[k + 1];               // popleft(); append(act) (:452-453)
        h[K - 1] = actions[n];
    }
}

// out[n] = obs[n] ++ one-hot(hist[n][0]) ++ ... ++ one-hot(hist[n][K-1]), uint8 (:454-457).  One thread per 4 output bytes.
__global__ void k_append_action(int64_t N, int D, int A, int K, const uint8_t *__restrict__ obs, const uint8_t *__restrict__ hist,
                                uint8_t *__restrict__ out) {
    const int L = D + A * K;
    const int64_t bytes = N * L, words = (bytes + 3) >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b0 = i * 4;
        int64_t n = b0 / L;
        int j = (int)(b0 - n * L);
        uint32_t w = 0;
        const int nb = (int)min((int64_t)4, bytes - b0);
        for (int k = 0; k < nb; ++k) {
            uint32_t v;
            if (j < D) v = obs[n * D + j];
            else { const int q = (j - D) / A; v = hist[n * K + q] == (j - D) - q * A; }
            w |= v << (8 * k);
            if (++j == L) { j = 0; ++n; }
        }
        if (nb == 4 && (reinterpret_cast<uintptr_t>(out) & 3) == 0) reinterpret_cast<uint32_t *>(out)[i] = w;
        else for (int k = 0; k < nb; ++k) out[b0 + k] = (uint8_t)(w >> (8 * k));
    }
}

// GoalPolicyWrapper._get_goals (wrappers.py:476-497) on FullyObsOneHotWrapper rows: [cells][planes] per env.
// achieved: goal plane cleared.  desired: agent cell -> empty, then goal cell -> agent.  One thread per cell.
__global__ void k_goal_policy(int64_t n_cells, int planes, int agent_idx, int empty_idx, int goal_idx,
                              const uint8_t *__restrict__ obs, uint8_t *__restrict__ achieved, uint8_t *__restrict__ desired) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *src = obs + i * planes;
        uint8_t *a = achieved + i * planes, *d = desired + i * planes;
        const bool is_agent = src[agent_idx] > 0, is_goal = src[goal_idx] > 0;
        for (int k = 0; k < planes; ++k) {
            const uint8_t v = src[k];
            a[k] = k == goal_idx ? 0 : v;
            uint8_t dv = v;
            if (is_agent) { if (k == agent_idx) dv = 0; if (k == empty_idx) dv = 1; }
            if (is_goal) { if (k == goal_idx) dv = 0; if (k == agent_idx) dv = 1; }
            d[k] = dv;
        }
    }
}

// Episode bookkeeping (SURVEY §5 "metrics": the reference only prints): running return and length per env; at a done
// step the finished episode's totals go to out_ret / out_len (elsewhere 0) and the running values restart.
__global__ void k_episode_stats(int64_t N, const double *__restrict__ reward, const uint8_t *__restrict__ done, double *__restrict__ run_ret,
                                int32_t *__restrict__ run_len, double *__restrict__ out_ret, int32_t *__restrict__ out_len,
                                unsigned long long *__restrict__ totals) {
    unsigned long long episodes = 0, steps = 0;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const double r = run_ret[n] + reward[n];
        const int32_t l = run_len[n] + 1;
        const bool d = done[n] != 0;
        out_ret[n] = d ? r : 0.0;
        out_len[n] = d ? l : 0;
        run_ret[n] = d ? 0.0 : r;
        run_len[n] = d ? 0 : l;
        if (d) { episodes++; steps += (unsigned long long)l; }
    }
    if (totals) {                                               // [0] finished episodes, [1] their summed length
        episodes = __reduce_add_sync(0xFFFFFFFFu, (unsigned)episodes);
        steps = __reduce_add_sync(0xFFFFFFFFu, (unsigned)steps);
        if ((threadIdx.x & 31) == 0 && episodes) { atomicAdd(&totals[0], episodes); atomicAdd(&totals[1], steps); }
    }
}

}  // namespace mgb

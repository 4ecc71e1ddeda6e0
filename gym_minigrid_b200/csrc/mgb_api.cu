// mgb200 C-ABI (include/mgb200.h): handle management, launch plumbing, host<->device pipeline.
// No torch types, no CPU fallback: every compute entry point launches sm_100a kernels.
#include "../../include/mgb200.h"
#include "mgb_kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

using namespace mgb;

static thread_local char g_err[512] = "";
static int fail(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return -1;
}
#define CUDA_OK(call)                                                                        \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int HOST_PIPE_STREAMS = 3;

// Every handle entry point runs on the handle's device and leaves the calling thread's current device as it found it
// (a process driving several GPUs, or torch's current device being another one).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
        if (cur != dev) {
            if (cudaSetDevice(dev) != cudaSuccess) { ok = false; return; }
            prev = cur;
        }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define ON_DEVICE(h)                                                                         \
    DeviceGuard guard_((h)->device);                                                         \
    if (!guard_.ok) return fail("cannot switch to device %d: %s", (h)->device, cudaGetErrorString(cudaGetLastError()))

static int elementwise_grid(int64_t total) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (total + 255) / 256;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 16));
}


struct mgb_handle {
    mgb_config cfg;
    DevCfg dc;
    int device = 0;
    int64_t n_envs = 0;
    int32_t n_groups = 0;
    uint64_t seed = 0;
    int64_t env_id_base = 0;
    int autoreset = 1;
    uint32_t *state = nullptr;
    uint32_t *spare = nullptr;               // pre-generated next layouts (spare_gen kernels)
    uint32_t *tmpl = nullptr;
    uint32_t *err = nullptr;
    uint32_t *ticket = nullptr;              // {next group ticket, warps done}: zeroed by the kernel itself when its last warp leaves
    const int32_t *tape = nullptr;
    const int64_t *tape_off = nullptr;
    uint32_t *pool = nullptr;
    int32_t pool_n = 0;
    int view = 7;
    int obs_bytes = OBS_BYTES;
    int sm_count = 0;
    // CTA shape per launch kind: [0] persistent rollouts (T > 1), [1] single steps / resets (state round-trips HBM)
    struct Shape { int blocks_per_sm = 0, warps_per_block = 4, warps = 0; size_t smem_bytes = 0; } shape[2];
    int64_t launches = 0;
    // host pipeline (mgb_step_host)
    cudaStream_t pipe[HOST_PIPE_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ev[HOST_PIPE_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t order_ev = nullptr;          // recorded after every state-changing launch on the caller's stream (mgb_step_host waits on it)
    uint32_t policy_epoch = 0;               // number of random-policy rollouts so far (counter word of their action stream)
    uint8_t *policy_scratch = nullptr;
    size_t policy_scratch_bytes = 0;
    uint8_t *d_actions = nullptr, *d_obs = nullptr, *d_done = nullptr, *d_dir = nullptr;
    double *d_reward = nullptr;
    // kernel timing
    int timing = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
};

typedef void (*rollout_fn)(const RolloutParams);
template <int V>
static rollout_fn pick_kernel_v(const mgb_config &c) {
    switch (c.gen) {
    case MGB_GEN_EMPTY: case MGB_GEN_DISTSHIFT: return c.see_through ? k_rollout<GEN_EMPTY, true, V> : k_rollout<GEN_EMPTY, false, V>;
    case MGB_GEN_DOORKEY: return c.see_through ? nullptr : k_rollout<GEN_DOORKEY, false, V>;
    case MGB_GEN_FOURROOMS: return c.see_through ? nullptr : k_rollout<GEN_FOURROOMS, false, V>;
    case MGB_GEN_DYNOBS: return c.see_through ? k_rollout<GEN_DYNOBS, true, V> : nullptr;
    case MGB_GEN_KEYCORRIDOR: return c.see_through ? nullptr : k_rollout<GEN_KEYCORRIDOR, false, V>;
    case MGB_GEN_POOL: return c.see_through ? k_rollout<GEN_POOL, true, V> : k_rollout<GEN_POOL, false, V>;
    case MGB_GEN_CROSSING: case MGB_GEN_LAVAGAP: case MGB_GEN_MULTIROOM:
        return c.see_through ? nullptr : k_rollout<GEN_PROC, false, V>;
    }
    return nullptr;
}
static int view_of(const mgb_config &c) { return c.agent_view_size == 0 ? 7 : c.agent_view_size; }
static rollout_fn pick_kernel(const mgb_config &c) {
    switch (view_of(c)) {
    case 3: return pick_kernel_v<3>(c);
    case 5: return pick_kernel_v<5>(c);
    case 7: return pick_kernel_v<7>(c);
    case 9: return pick_kernel_v<9>(c);
    case 11: return pick_kernel_v<11>(c);
    }
    return nullptr;
}

// static part of each layout (walls + fixed goal) as packed cell codes, x-major like Grid.encode
static std::vector<uint32_t> build_template(const mgb_config &c, int GW, int HP) {
    const int W = c.width, H = c.height;
    std::vector<uint8_t> g((size_t)GW * 4, (uint8_t)CODE_EMPTY);
    auto set = [&](int x, int y, int code) { g[(size_t)x * HP + y] = (uint8_t)code; };
    auto wall_rect = [&](int x0, int y0, int w, int h) {   // Grid.wall_rect minigrid.py:433-437
        for (int i = 0; i < w; ++i) { set(x0 + i, y0, CODE_WALL); set(x0 + i, y0 + h - 1, CODE_WALL); }
        for (int j = 0; j < h; ++j) { set(x0, y0 + j, CODE_WALL); set(x0 + w - 1, y0 + j, CODE_WALL); }
    };
    switch (c.gen) {
    case MGB_GEN_EMPTY: case MGB_GEN_DOORKEY: case MGB_GEN_DYNOBS: case MGB_GEN_CROSSING: case MGB_GEN_LAVAGAP:   // crossing.py:31-38, lavagap.py:28-37
        wall_rect(0, 0, W, H);
        set(W - 2, H - 2, CODE_GOAL);                      // empty.py:48, doorkey.py:23, dynamicobstacles.py:43
        break;
    case MGB_GEN_DISTSHIFT:                                // distshift.py:30-43
        wall_rect(0, 0, W, H);
        set(W - 2, 1, CODE_GOAL);
        for (int i = 0; i < W - 6; ++i) { set(3 + i, 1, code_of(T_LAVA, C_RED, 0)); set(3 + i, c.gen_param0, code_of(T_LAVA, C_RED, 0)); }
        break;
    case MGB_GEN_FOURROOMS: {                              // fourrooms.py:24-53 (walls only; gaps are drawn on device)
        wall_rect(0, 0, W, H);
        const int rw = W / 2, rh = H / 2;
        for (int y = 0; y < 2 * rh; ++y) set(rw, y, CODE_WALL);
        for (int x = 0; x < 2 * rw; ++x) set(x, rh, CODE_WALL);
        break;
    }
    case MGB_GEN_KEYCORRIDOR:                              // roomgrid.py:125-137
        for (int j = 0; j < c.num_rows; ++j)
            for (int i = 0; i < 3; ++i) wall_rect(i * (c.room_size - 1), j * (c.room_size - 1), c.room_size, c.room_size);
        break;
    }
    std::vector<uint32_t> w(GW);
    memcpy(w.data(), g.data(), (size_t)GW * 4);
    return w;
}

extern "C" {

#ifndef MGB_BUILD_SHA
#define MGB_BUILD_SHA "unknown"
#endif
#ifndef MGB_BUILD_DEFINES
#define MGB_BUILD_DEFINES "none"
#endif
// carries the source revision and the -D switches of the build: every file under profiles/ quotes it
const char *mgb_version(void) { return "mgb200 0.2 (sm_100a) src=" MGB_BUILD_SHA " defines=" MGB_BUILD_DEFINES; }
const char *mgb_last_error(void) { return g_err; }

int mgb_create(const mgb_config *cfg, int64_t num_envs, int device, uint64_t seed, int64_t env_id_base, mgb_handle **out) {
    if (!cfg || !out) return fail("mgb_create: null argument");
    *out = nullptr;
    const mgb_config &c = *cfg;
    if (c.width < 3 || c.height < 3 || c.width > 64 || c.height > 64) return fail("mgb_create: grid size %dx%d unsupported (3..64)", c.width, c.height);
    if (c.max_steps < 1 || c.max_steps > 65535) return fail("mgb_create: max_steps %d out of range", c.max_steps);
    if (num_envs < 1 || num_envs > ((int64_t)1 << 31) - 64) return fail("mgb_create: num_envs %lld out of range", (long long)num_envs);
    if (c.n_obstacles < 0 || c.n_obstacles > MGB_MAX_OBSTACLES) return fail("mgb_create: n_obstacles %d out of range", c.n_obstacles);
    if (c.gen == MGB_GEN_KEYCORRIDOR) {
        if (c.num_rows < 1 || c.num_rows > 3 || c.room_size < 3) return fail("mgb_create: bad RoomGrid shape");
        if (c.width != (c.room_size - 1) * 3 + 1 || c.height != (c.room_size - 1) * c.num_rows + 1) return fail("mgb_create: RoomGrid size mismatch (roomgrid.py:85-86)");
    }
    if (c.gen == MGB_GEN_CROSSING) {
        if ((c.width & 1) == 0 || (c.height & 1) == 0 || c.width < 5 || c.height < 5) return fail("mgb_create: Crossing needs an odd grid size (crossing.py:25)");
        if (c.gen_param0 < 0 || (c.gen_param1 & 3) > 2 || (c.gen_param1 & ~7)) return fail("mgb_create: bad Crossing parameters");
    }
    if (c.gen == MGB_GEN_DISTSHIFT && (c.width < 7 || c.gen_param0 < 1 || c.gen_param0 > c.height - 2 || c.random_start))
        return fail("mgb_create: bad DistShift parameters (width >= 7, strip2_row inside the grid, fixed start)");
    if (c.gen == MGB_GEN_LAVAGAP && (c.width < 5 || c.height < 5)) return fail("mgb_create: LavaGap needs at least 5x5 (lavagap.py:22)");
    if (c.gen == MGB_GEN_MULTIROOM && (c.gen_param0 < 1 || c.gen_param0 > 8 || c.gen_param1 < 4 || c.gen_param1 > 32 || c.width != c.height))
        return fail("mgb_create: bad MultiRoom parameters (1..8 rooms, maxRoomSize 4..32, square grid)");
    if (c.hook < 0 || c.hook > MGB_HOOK_MEMORY || (c.hook != 0 && c.gen != MGB_GEN_POOL)) return fail("mgb_create: hook %d needs a level-pool handle (MGB_GEN_POOL)", c.hook);
    if (c.gen == MGB_GEN_DYNOBS && c.n_actions != 3) return fail("mgb_create: Dynamic-Obstacles has 3 actions");
    if (c.gen != MGB_GEN_DYNOBS && c.n_actions != 7) return fail("mgb_create: n_actions must be 7");
    { const int v = view_of(c); if (v != 3 && v != 5 && v != 7 && v != 9 && v != 11) return fail("mgb_create: agent_view_size %d not built (3, 5, 7, 9, 11)", v); }
    rollout_fn fn = pick_kernel(c);
    if (!fn) return fail("mgb_create: unsupported (gen=%d, see_through=%d) combination", c.gen, c.see_through);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("mgb_create: no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("mgb_create: device %d out of range (%d devices)", device, ndev);
    DeviceGuard guard_(device);
    if (!guard_.ok) return fail("mgb_create: cannot switch to device %d", device);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("mgb_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

    mgb_handle *h = new (std::nothrow) mgb_handle();
    if (!h) return fail("mgb_create: out of host memory");
    h->cfg = c; h->device = device; h->n_envs = num_envs; h->seed = seed; h->env_id_base = env_id_base;
    h->n_groups = (int32_t)((num_envs + 31) / 32);
    h->view = view_of(c); h->obs_bytes = obs_bytes(h->view);
    DevCfg &d = h->dc;
    d.gen = c.gen == MGB_GEN_DISTSHIFT ? MGB_GEN_EMPTY : c.gen;      // DistShift = the Empty kernels on another template
    d.W = c.width; d.H = c.height; d.max_steps = c.max_steps; d.see_through = c.see_through;
    d.n_actions = c.n_actions; d.n_obst = c.n_obstacles; d.room_size = c.room_size; d.num_rows = c.num_rows;
    d.random_start = c.random_start; d.lava_v1 = c.lava_v1; d.hook = c.hook; d.gp0 = c.gen_param0; d.gp1 = c.gen_param1;
    d.HP = (c.height + 3) / 4 * 4;
    d.GW = c.width * d.HP / 4;
    d.goal_idx = c.gen == MGB_GEN_EMPTY ? (c.width - 2) * d.HP + c.height - 2 : (c.gen == MGB_GEN_DISTSHIFT ? (c.width - 2) * d.HP + 1 : -1);   // build_template
    d.S = d.GW + XWORDS + (c.n_obstacles > 0 ? OBST_WORDS : 0) + (c.gen == MGB_GEN_POOL ? 1 : 0);   // pool: + level word
    h->sm_count = prop.multiProcessorCount;
    auto cleanup = [&](int rc) { mgb_destroy(h); return rc; };
    // warps per CTA: whatever keeps the most warps resident per SM (shared memory is the limiter for the
    // larger grids: each warp needs its 32-env state block + a 4704-byte staging block)
    const size_t per_warp = (size_t)stage_bytes(h->view) + (size_t)(d.S + 1) * 32 * 4;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin) != cudaSuccess)
        return cleanup(fail("mgb_create: cannot opt in to %zu bytes of shared memory", (size_t)prop.sharedMemPerBlockOptin));
    // Measured on B200 (profiles/README.md §4): a single-step launch is bound by the latency of the state round
    // trip and wants every warp it can get; the persistent rollout is bound by the shared-memory pipe, which 24 warps
    // saturate -- beyond that more resident warps only cost (28 warps: -4 %), and at equal warps more, smaller CTAs
    // did slightly better (6 warps x 4 CTAs: +3 % over 8 x 3).
    int ROLLOUT_WARP_CAP = 24;
    const char *force = nullptr;
#ifdef MGB_EXPERIMENT       // experiment builds only (profiles/tools/*.sh): override the shape heuristics from the environment
    if (const char *cap = getenv("MGB_ROLLOUT_WARP_CAP")) ROLLOUT_WARP_CAP = std::max(2, atoi(cap));
    force = getenv("MGB_WARPS_PER_BLOCK");
#endif
    for (int wpb = MAX_WARPS_PER_BLOCK; wpb >= 2; --wpb) {
        if (force && atoi(force) != wpb) continue;
        const size_t smem = (size_t)table_bytes(d.gen) + (size_t)tmpl_smem_bytes(d.gen, d.GW) + (size_t)wpb * per_warp;
        if (smem > prop.sharedMemPerBlockOptin) continue;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, wpb * 32, smem) != cudaSuccess || nb < 1) continue;
        const int warps = nb * wpb;
        mgb_handle::Shape cand;
        cand.blocks_per_sm = nb; cand.warps_per_block = wpb; cand.warps = warps; cand.smem_bytes = smem;
        if (warps > h->shape[1].warps) h->shape[1] = cand;
        const mgb_handle::Shape &cur = h->shape[0];
        const int capped = std::min(warps, ROLLOUT_WARP_CAP), best = std::min(cur.warps, ROLLOUT_WARP_CAP);
        const int over = warps - capped, cur_over = cur.warps - best;
        if (c.see_through) {
            // see-through kernels (64 registers): at equal warps more, smaller CTAs, but not below 6 warps
            // (Empty-8x8: 6 x 4 beats 8 x 3 by 3 % and 4 x 6 by 1 %; profiles/r1_cta_shape_sweep.txt)
            if (capped > best || (capped == best && (over < cur_over || (over == cur_over && wpb >= 6 && nb > cur.blocks_per_sm))))
                h->shape[0] = cand;
        } else {
            // occluded kernels (128 registers = 4 warps per scheduler partition): CTAs of 4 or 8 warps fill the four
            // partitions evenly; a 6-warp CTA does not (FourRooms: 4 x 3 beats 6 x 2 and 3 x 4 by 4 %; DoorKey-8x8: 8 x 2
            // beats 4 x 4 by 3 % and 6 x 2 by 11 %).  Ties go to the larger CTA (fewer copies of the tables).
            const int score = capped - ((wpb & 3) ? 4 : 0), cur_score = cur.warps < 1 ? -1 : best - ((cur.warps_per_block & 3) ? 4 : 0);
            if (score > cur_score) h->shape[0] = cand;
        }
    }
#ifdef MGB_EXPERIMENT
    if (const char *nbs = getenv("MGB_BLOCKS_PER_SM"))
        h->shape[0].blocks_per_sm = std::max(1, std::min(h->shape[0].blocks_per_sm, atoi(nbs)));
    if (getenv("MGB_PRINT_SHAPE"))
        fprintf(stderr, "mgb_create: rollout shape %d CTAs x %d warps (%zu B smem), single-step shape %d x %d\n", h->shape[0].blocks_per_sm,
                h->shape[0].warps_per_block, h->shape[0].smem_bytes, h->shape[1].blocks_per_sm, h->shape[1].warps_per_block);
#endif
    if (h->shape[0].warps < 1) return cleanup(fail("mgb_create: kernel does not fit on an SM (%zu bytes of shared memory per warp)", per_warp));
    const size_t state_bytes = (size_t)h->n_groups * d.S * 32 * 4;
    if (cudaMalloc(&h->state, state_bytes) != cudaSuccess) return cleanup(fail("mgb_create: cudaMalloc(%zu) for env state failed", state_bytes));
    if (cudaMemset(h->state, 0, state_bytes) != cudaSuccess) return cleanup(fail("mgb_create: memset failed"));
    // Spare layouts (reset_with_spares) for the generators that cost more than a trip to HBM.  Measured with episode ends
    // spread over time (profiles/r2_desync_spares.txt): KeyCorridorS6R3 +50 %, S3R3 2.6x, SimpleCrossingS11N5 +36 %,
    // MultiRoom-N6 2.2x; LavaGap's generator is a handful of draws and loses 9 % to the round trip: it goes without.
    if (MGB_SPARES && (c.gen == MGB_GEN_KEYCORRIDOR || c.gen == MGB_GEN_CROSSING || c.gen == MGB_GEN_MULTIROOM)) {
        const size_t spare_bytes = (size_t)h->n_groups * spare_words(d.GW) * 32 * 4;
        if (cudaMalloc(&h->spare, spare_bytes) != cudaSuccess) return cleanup(fail("mgb_create: cudaMalloc(%zu) for the spare layouts failed", spare_bytes));
    }
    std::vector<uint32_t> t = build_template(c, d.GW, d.HP);
    if (cudaMalloc(&h->tmpl, t.size() * 4) != cudaSuccess || cudaMemcpy(h->tmpl, t.data(), t.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
        return cleanup(fail("mgb_create: template upload failed"));
    if (cudaMalloc(&h->err, 4) != cudaSuccess || cudaMemset(h->err, 0, 4) != cudaSuccess) return cleanup(fail("mgb_create: err flag alloc failed"));
    if (cudaMalloc(&h->ticket, 8) != cudaSuccess || cudaMemset(h->ticket, 0, 8) != cudaSuccess) return cleanup(fail("mgb_create: ticket counter alloc failed"));
    if (cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming) != cudaSuccess) return cleanup(fail("mgb_create: event creation failed"));
    if (cudaDeviceSynchronize() != cudaSuccess) return cleanup(fail("mgb_create: device synchronisation failed"));   // the memsets above ran on the legacy stream
    *out = h;
    return 0;
}

int mgb_destroy(mgb_handle *h) {
    if (!h) return 0;
    DeviceGuard guard_(h->device);
    cudaFree(h->state); cudaFree(h->spare); cudaFree(h->tmpl); cudaFree(h->err); cudaFree(h->ticket); cudaFree(h->pool);
    cudaFree(h->policy_scratch);
    cudaFree(h->d_actions); cudaFree(h->d_obs); cudaFree(h->d_done); cudaFree(h->d_dir); cudaFree(h->d_reward);
    for (auto &s : h->pipe) if (s) cudaStreamDestroy(s);
    for (auto &e : h->pipe_ev) if (e) cudaEventDestroy(e);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return 0;
}

int64_t mgb_num_envs(const mgb_handle *h) { return h ? h->n_envs : -1; }
int64_t mgb_kernel_launches(const mgb_handle *h) { return h ? h->launches : -1; }

int mgb_set_autoreset(mgb_handle *h, int on) {
    if (!h) return fail("null handle");
    h->autoreset = on ? 1 : 0;
    return 0;
}

int mgb_set_rng_tape(mgb_handle *h, const int32_t *draws, const int64_t *offsets) {
    if (!h) return fail("null handle");
    if ((draws == nullptr) != (offsets == nullptr)) return fail("mgb_set_rng_tape: draws and offsets must both be set or both be NULL");
    h->tape = draws; h->tape_off = offsets;
    return 0;
}

int mgb_set_kernel_timing(mgb_handle *h, int on) {
    if (!h) return fail("null handle");
    ON_DEVICE(h);
    h->timing = on ? 1 : 0;
    if (on && !h->ev0) { CUDA_OK(cudaEventCreate(&h->ev0)); CUDA_OK(cudaEventCreate(&h->ev1)); }
    h->ev_valid = false;
    return 0;
}

double mgb_last_kernel_ms(mgb_handle *h) {
    if (!h || !h->timing || !h->ev_valid) return -1.0;
    float ms = 0.f;
    if (cudaEventSynchronize(h->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

// launch over groups [g0, g0+ng)
static int launch(mgb_handle *h, int32_t g0, int32_t ng, int32_t T, int do_reset, const uint8_t *mask,
                  const uint8_t *actions, uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir,
                  int64_t stride, cudaStream_t stream, bool timed, bool order = true) {
    if (ng <= 0) return 0;
    RolloutParams p;
    p.cfg = h->dc; p.state = h->state; p.spare = h->spare; p.tmpl = h->tmpl; p.n_envs = h->n_envs;
    p.group0 = g0; p.n_groups = ng; p.T = T; p.do_reset = do_reset; p.autoreset = h->autoreset;
    p.reset_mask = mask; p.actions = actions; p.obs = obs; p.reward = reward; p.done = done; p.dir = dir;
    p.stride = stride; p.seed = h->seed; p.env_id_base = h->env_id_base;
    p.tape = h->tape; p.tape_off = h->tape_off; p.err = h->err; p.pool = h->pool; p.pool_n = h->pool_n;
    // group tickets: for the launches a handle runs one at a time (not the host pipeline's concurrent chunks, order == false)
    p.ticket = (MGB_DYNAMIC_GROUPS && order) ? h->ticket : nullptr;
    rollout_fn fn = pick_kernel(h->cfg);
    const mgb_handle::Shape &sh = h->shape[T > 1 ? 0 : 1];
    const int want = (ng + sh.warps_per_block - 1) / sh.warps_per_block;
    const int grid = std::max(1, std::min(want, h->sm_count * sh.blocks_per_sm));
    if (timed && h->timing) CUDA_OK(cudaEventRecord(h->ev0, stream));
#if MGB_PDL
    {
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)(sh.warps_per_block * 32)); lc.dynamicSmemBytes = sh.smem_bytes; lc.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        CUDA_OK(cudaLaunchKernelEx(&lc, fn, p));
    }
#else
    fn<<<grid, sh.warps_per_block * 32, sh.smem_bytes, stream>>>(p);
#endif
    CUDA_OK(cudaGetLastError());
    if (timed && h->timing) { CUDA_OK(cudaEventRecord(h->ev1, stream)); h->ev_valid = true; }
    if (order) CUDA_OK(cudaEventRecord(h->order_ev, stream));
    h->launches++;
    return 0;
}

int mgb_seed(mgb_handle *h, uint64_t seed) {
    if (!h) return fail("null handle");
    ON_DEVICE(h);
    h->seed = seed;
    // rewind the episode counters: word GW+2 of every env
    const DevCfg &d = h->dc;
    // the memset runs on the legacy default stream, which does not order against non-blocking streams: everything
    // enqueued before (any stream) is drained first, and the memset is complete before a later reset can be enqueued
    CUDA_OK(cudaDeviceSynchronize());
    CUDA_OK(cudaMemset2D(h->state + (size_t)(d.GW + 2) * 32, (size_t)d.S * 32 * 4, 0, 2 * 32 * 4, h->n_groups));
    if (h->spare) {                                             // layouts pre-generated under the old seed are void
        k_clear_flags<<<(unsigned)((h->n_envs + 255) / 256), 256>>>(h->state, d.S, d.GW, h->n_envs, (uint32_t)FLAG_SPARE);
        CUDA_OK(cudaGetLastError());
    }
    CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

int mgb_set_level_pool(mgb_handle *h, int32_t n_levels, const uint8_t *grid, const uint8_t *aux, const int32_t *agent,
                       const int32_t *hook_params, void *stream) {
    if (!h) return fail("null handle");
    if (h->cfg.gen != MGB_GEN_POOL) return fail("mgb_set_level_pool: handle was not created with MGB_GEN_POOL");
    if (n_levels < 1 || !grid || !agent) return fail("mgb_set_level_pool: need at least one level, grid and agent");
    ON_DEVICE(h);
    if (h->cfg.hook != 0 && !hook_params) return fail("mgb_set_level_pool: this handle has hook %d and needs hook_params", h->cfg.hook);
    const int PW = h->dc.GW + POOL_XW;
    uint32_t *np = nullptr;
    CUDA_OK(cudaMalloc(&np, (size_t)n_levels * PW * 4));
    const int64_t threads = (int64_t)n_levels * PW;
    k_pack_levels<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->dc, n_levels, grid, aux, agent, hook_params, np, h->err);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(np); return fail("k_pack_levels: %s", cudaGetErrorString(e)); }
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));      // the old pool may still be in use by earlier launches
    cudaFree(h->pool);
    h->pool = np; h->pool_n = n_levels;
    h->launches++;
    return 0;
}

static int levels_io(mgb_handle *h, int32_t *out, const int32_t *in, void *stream) {
    if (!h) return fail("null handle");
    if (h->cfg.gen != MGB_GEN_POOL) return fail("levels: handle was not created with MGB_GEN_POOL");
    ON_DEVICE(h);
    k_levels<<<(unsigned)((h->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(h->state, h->dc.S, h->dc.GW + XWORDS, h->n_envs, out, in);
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}
int mgb_get_levels(mgb_handle *h, int32_t *levels, void *stream) { return levels ? levels_io(h, levels, nullptr, stream) : fail("null buffer"); }
int mgb_set_levels(mgb_handle *h, const int32_t *levels, void *stream) { return levels ? levels_io(h, nullptr, levels, stream) : fail("null buffer"); }

int mgb_reset(mgb_handle *h, const uint8_t *mask, uint8_t *obs, uint8_t *dir, void *stream) {
    if (!h) return fail("null handle");
    if (h->cfg.gen == MGB_GEN_POOL && h->pool_n < 1) return fail("mgb_reset: no level pool (call mgb_set_level_pool first)");
    ON_DEVICE(h);
    return launch(h, 0, h->n_groups, 0, 1, mask, nullptr, obs, nullptr, nullptr, dir, h->n_envs, (cudaStream_t)stream, false);
}

int mgb_step(mgb_handle *h, const uint8_t *actions, uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir, void *stream) {
    if (!h) return fail("null handle");
    if (!actions) return fail("mgb_step: actions is NULL");
    ON_DEVICE(h);
    return launch(h, 0, h->n_groups, 1, 0, nullptr, actions, obs, reward, done, dir, h->n_envs, (cudaStream_t)stream, true);
}

int mgb_rollout(mgb_handle *h, int32_t T, const uint8_t *actions, uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir, void *stream) {
    if (!h) return fail("null handle");
    if (T < 1) return fail("mgb_rollout: T must be >= 1");
    if (!actions) return fail("mgb_rollout: actions is NULL");
    ON_DEVICE(h);
    return launch(h, 0, h->n_groups, T, 0, nullptr, actions, obs, reward, done, dir, h->n_envs, (cudaStream_t)stream, true);
}

int mgb_rollout_random(mgb_handle *h, int32_t T, uint8_t *actions_out, uint8_t *obs, double *reward, uint8_t *done, uint8_t *dir, void *stream) {
    if (!h) return fail("null handle");
    if (T < 1) return fail("mgb_rollout_random: T must be >= 1");
    ON_DEVICE(h);
    uint8_t *acts = actions_out;
    if (!acts) {                                             // nobody wants to see the actions: private scratch
        const size_t need = (size_t)T * h->n_envs;
        if (need > h->policy_scratch_bytes) {
            CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
            cudaFree(h->policy_scratch); h->policy_scratch = nullptr; h->policy_scratch_bytes = 0;
            CUDA_OK(cudaMalloc(&h->policy_scratch, need));
            h->policy_scratch_bytes = need;
        }
        acts = h->policy_scratch;
    }
    k_policy_actions<<<elementwise_grid(h->n_envs * ((T + 3) / 4)), 256, 0, (cudaStream_t)stream>>>(acts, h->n_envs, T, h->seed, h->env_id_base,
                                                                                           h->policy_epoch++, h->dc.n_actions);
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return launch(h, 0, h->n_groups, T, 0, nullptr, acts, obs, reward, done, dir, h->n_envs, (cudaStream_t)stream, true);
}

int mgb_step_host(mgb_handle *h, const uint8_t *actions_host, uint8_t *obs_host, double *reward_host, uint8_t *done_host, uint8_t *dir_host) {
    if (!h) return fail("null handle");
    if (!actions_host) return fail("mgb_step_host: actions is NULL");
    ON_DEVICE(h);
    const int64_t N = h->n_envs;
    if (!h->pipe[0]) {
        for (auto &s : h->pipe) CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        for (auto &e : h->pipe_ev) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CUDA_OK(cudaMalloc(&h->d_actions, N));
        CUDA_OK(cudaMalloc(&h->d_obs, (size_t)N * h->obs_bytes));
        CUDA_OK(cudaMalloc(&h->d_reward, (size_t)N * 8));
        CUDA_OK(cudaMalloc(&h->d_done, N));
        CUDA_OK(cudaMalloc(&h->d_dir, N));
    }
    // Chunk the env range so that H2D, kernel and the D2H of the observations of neighbouring chunks overlap.  The
    // copy engine pays a few microseconds per copy, so the chunks are few and the three small outputs (reward, done,
    // direction: 10 bytes per env) leave in one copy each after the last kernel instead of one per chunk.
    const int32_t G = h->n_groups;
#ifdef MGB_EXPERIMENT
    static const int forced = getenv("MGB_HOST_CHUNKS") ? atoi(getenv("MGB_HOST_CHUNKS")) : 0;
#else
    constexpr int forced = 0;
#endif
    int32_t nchunks = forced > 0 ? forced : (G >= 4096 ? 4 : (G >= 1024 ? 2 : 1));
    const int32_t per = (G + nchunks - 1) / nchunks;
    // Ordering against the caller's earlier work on this handle: every launch records `order_ev` on its stream and the
    // pipeline streams wait for it -- no device-wide synchronisation, other streams of the process keep running.
    // (State uploads / level pools synchronise their stream themselves.)
    int used = 0;
    for (int32_t c = 0, g0 = 0; g0 < G; ++c, g0 += per) {
        cudaStream_t s = h->pipe[c % HOST_PIPE_STREAMS];
        if (c < HOST_PIPE_STREAMS) CUDA_OK(cudaStreamWaitEvent(s, h->order_ev, 0));
        used = std::max(used, c % HOST_PIPE_STREAMS + 1);
        const int32_t ng = std::min(per, G - g0);
        const int64_t e0 = (int64_t)g0 * 32, ne = std::min((int64_t)ng * 32, N - e0);
        CUDA_OK(cudaMemcpyAsync(h->d_actions + e0, actions_host + e0, ne, cudaMemcpyHostToDevice, s));
        if (launch(h, g0, ng, 1, 0, nullptr, h->d_actions, obs_host ? h->d_obs : nullptr, reward_host ? h->d_reward : nullptr,
                   done_host ? h->d_done : nullptr, dir_host ? h->d_dir : nullptr, N, s, false, false)) return -1;
        CUDA_OK(cudaEventRecord(h->pipe_ev[c % HOST_PIPE_STREAMS], s));      // the last record per stream is the one waited on
        if (obs_host) CUDA_OK(cudaMemcpyAsync(obs_host + e0 * h->obs_bytes, h->d_obs + e0 * h->obs_bytes, (size_t)ne * h->obs_bytes, cudaMemcpyDeviceToHost, s));
    }
    cudaStream_t tail = h->pipe[0];
    for (int i = 1; i < used; ++i) CUDA_OK(cudaStreamWaitEvent(tail, h->pipe_ev[i], 0));
    if (reward_host) CUDA_OK(cudaMemcpyAsync(reward_host, h->d_reward, (size_t)N * 8, cudaMemcpyDeviceToHost, tail));
    if (done_host) CUDA_OK(cudaMemcpyAsync(done_host, h->d_done, N, cudaMemcpyDeviceToHost, tail));
    if (dir_host) CUDA_OK(cudaMemcpyAsync(dir_host, h->d_dir, N, cudaMemcpyDeviceToHost, tail));
    CUDA_OK(cudaEventRecord(h->order_ev, tail));                 // later launches on any stream are ordered by the host: all pipe streams are drained below
    for (auto &s : h->pipe) CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

static int state_io(mgb_handle *h, bool set, int full_obs, int64_t first, int64_t count, uint8_t *grid, uint8_t *aux, int32_t *agent,
                    uint8_t *carrying, int16_t *obstacles, uint8_t *target, uint32_t *rng, void *stream) {
    if (!h) return fail("null handle");
    if (first < 0 || count < 0 || first + count > h->n_envs) return fail("state range [%lld,+%lld) outside [0,%lld)", (long long)first, (long long)count, (long long)h->n_envs);
    if (count == 0) return 0;
    ON_DEVICE(h);
    StateIO io;
    io.cfg = h->dc; io.state = h->state; io.tmpl = h->tmpl; io.first = first; io.count = count;
    io.grid = grid; io.aux = aux; io.agent = agent; io.carrying = carrying;
    io.obstacles = h->dc.n_obst > 0 ? obstacles : nullptr; io.target = target; io.rng = rng; io.err = h->err;
    const int64_t threads = count * h->dc.S;
    const int block = 256;
    const int64_t grid_dim = (threads + block - 1) / block;
    if (grid_dim > 0x7FFFFFFF) return fail("state_io: too many envs for one launch");
    if (set) {
        // a partial upload into envs whose grid is implied (template + balls, DESIGN.md "pristine"): make it explicit first
        if (template_gen(h->dc.gen) && !(grid && (obstacles || h->dc.n_obst == 0)))
            k_materialize<<<(unsigned)((count + block - 1) / block), block, 0, (cudaStream_t)stream>>>(h->dc, h->state, h->tmpl, first, count);
        k_set_state<<<(unsigned)grid_dim, block, 0, (cudaStream_t)stream>>>(io);
    }
    else k_get_state<<<(unsigned)grid_dim, block, 0, (cudaStream_t)stream>>>(io, full_obs);
    CUDA_OK(cudaGetLastError());
    if (set) CUDA_OK(cudaEventRecord(h->order_ev, (cudaStream_t)stream));
    h->launches++;
    return 0;
}

int mgb_set_state(mgb_handle *h, int64_t first, int64_t count, const uint8_t *grid, const uint8_t *aux, const int32_t *agent,
                  const uint8_t *carrying, const int16_t *obstacles, const uint8_t *target, const uint32_t *rng, void *stream) {
    return state_io(h, true, 0, first, count, const_cast<uint8_t *>(grid), const_cast<uint8_t *>(aux), const_cast<int32_t *>(agent),
                    const_cast<uint8_t *>(carrying), const_cast<int16_t *>(obstacles), const_cast<uint8_t *>(target),
                    const_cast<uint32_t *>(rng), stream);
}

int mgb_get_state(mgb_handle *h, int64_t first, int64_t count, uint8_t *grid, uint8_t *aux, int32_t *agent, uint8_t *carrying,
                  int16_t *obstacles, uint8_t *target, uint32_t *rng, void *stream) {
    return state_io(h, false, 0, first, count, grid, aux, agent, carrying, obstacles, target, rng, stream);
}

int mgb_full_obs(mgb_handle *h, uint8_t *out, void *stream) {
    if (!out) return fail("mgb_full_obs: out is NULL");
    return state_io(h, false, 1, 0, h ? h->n_envs : 0, out, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}


int mgb_onehot(const uint8_t *cells, uint8_t *out, int64_t n_cells, const uint8_t *class_map, int32_t n_classes,
               int32_t n_colors, int32_t n_states, void *stream) {
    if (!cells || !out) return fail("mgb_onehot: null buffer");
    if (n_cells < 0 || n_classes < 1 || n_classes > 16 || n_colors < 0 || n_colors > 8 || n_states < 1 || n_states > 8)
        return fail("mgb_onehot: bad sizes");
    if (n_cells == 0) return 0;
    ClassMap cm;
    for (int i = 0; i < 16; ++i) cm.m[i] = class_map && i < 11 ? class_map[i] : (uint8_t)i;
    const int64_t total = n_cells * (n_classes + n_colors + n_states);
    k_onehot<<<elementwise_grid(total), 256, 0, (cudaStream_t)stream>>>(cells, out, n_cells, cm, n_classes, n_colors, n_states);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mgb_flat_obs(const uint8_t *img, int32_t img_bytes, const float *mission_table, int32_t mission_len,
                 const uint8_t *mission_idx, float *out, int64_t N, void *stream) {
    if (!img || !mission_table || !out) return fail("mgb_flat_obs: null buffer");
    if (img_bytes < 1 || mission_len < 0 || N < 0) return fail("mgb_flat_obs: bad sizes");
    if (N == 0) return 0;
    const int64_t total = N * ((int64_t)img_bytes + mission_len);
    k_flat_obs<<<elementwise_grid(total), 256, 0, (cudaStream_t)stream>>>(img, img_bytes, mission_table, mission_len, mission_idx, out, N);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mgb_render_partial(const uint8_t *obs, int32_t view, const uint8_t *atlas, int32_t tile, uint8_t *out, int64_t N, void *stream) {
    if (!obs || !atlas || !out) return fail("mgb_render_partial: null buffer");
    if (view < 1 || view > 11 || tile < 8 || tile % 8 != 0 || N < 0) return fail("mgb_render_partial: bad sizes (tile must be a multiple of 8)");
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(atlas)) & 7) return fail("mgb_render_partial: out and atlas must be 8-byte aligned");
    if (N == 0) return 0;
    const int64_t segs = N * view * tile * view;
    if (segs < ((int64_t)1 << 31) - (1 << 24)) k_render_partial<uint32_t><<<elementwise_grid(segs), 256, 0, (cudaStream_t)stream>>>(obs, view, atlas, tile, out, N);
    else k_render_partial<uint64_t><<<elementwise_grid(segs), 256, 0, (cudaStream_t)stream>>>(obs, view, atlas, tile, out, N);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mgb_render_full(mgb_handle *h, const uint8_t *obs, const uint8_t *atlas, int32_t tile, uint8_t *out, void *stream) {
    if (!h || !atlas || !out) return fail("mgb_render_full: null argument");
    if (tile < 8 || tile % 8 != 0) return fail("mgb_render_full: tile must be a multiple of 8");
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(atlas)) & 7) return fail("mgb_render_full: out and atlas must be 8-byte aligned");
    ON_DEVICE(h);
    const int64_t segs = h->n_envs * h->dc.H * tile * h->dc.W;
    if (segs < ((int64_t)1 << 31) - (1 << 24)) k_render_full<uint32_t><<<elementwise_grid(segs), 256, 0, (cudaStream_t)stream>>>(h->dc, h->state, h->tmpl, obs, h->view, atlas, tile, out, h->n_envs);
    else k_render_full<uint64_t><<<elementwise_grid(segs), 256, 0, (cudaStream_t)stream>>>(h->dc, h->state, h->tmpl, obs, h->view, atlas, tile, out, h->n_envs);
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int mgb_visit_bonus(mgb_handle *h, int32_t by_action, const uint8_t *actions, uint32_t *counts, int64_t table,
                    double *reward, void *stream) {
    if (!h || !counts || !reward) return fail("mgb_visit_bonus: null argument");
    if (by_action && !actions) return fail("mgb_visit_bonus: actions is NULL");
    const int64_t need = (int64_t)h->dc.W * h->dc.H * (by_action ? 4 * h->dc.n_actions : 1);
    if (table < need) return fail("mgb_visit_bonus: table must hold W*H%s entries per env", by_action ? "*4*n_actions" : "");
    ON_DEVICE(h);
    if (h->n_envs == 0) return 0;
    k_visit_bonus<<<elementwise_grid(h->n_envs), 256, 0, (cudaStream_t)stream>>>(h->dc, h->state, by_action, actions, counts, table,
                                                                              reward, h->n_envs, h->err);
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int mgb_dac(mgb_handle *h, int32_t count_ge_max, const uint8_t *done_in, const uint8_t *envdone_in, uint8_t *envdone_out,
            const uint8_t *reset_dir, uint8_t *obs, double *reward, uint8_t *done_out, uint8_t *dir, void *stream) {
    if (!h || !done_in || !envdone_in || !envdone_out || !reset_dir || !obs || !reward || !done_out || !dir)
        return fail("mgb_dac: null argument");
    if (envdone_in == envdone_out || done_in == done_out) return fail("mgb_dac: in and out buffers must differ");
    ON_DEVICE(h);
    if (h->n_envs == 0) return 0;
    const int64_t words = (h->n_envs * h->obs_bytes + 3) / 4;
    k_dac<<<elementwise_grid(words), 256, 0, (cudaStream_t)stream>>>(h->n_envs, h->obs_bytes, count_ge_max != 0, done_in, envdone_in,
                                                                    envdone_out, reset_dir, obs, reward, done_out, dir);
    CUDA_OK(cudaGetLastError());
    h->launches++;
    return 0;
}

int mgb_append_action(int64_t N, int32_t D, int32_t A, int32_t K, const uint8_t *obs, const uint8_t *actions,
                      const uint8_t *done, uint8_t *hist, uint8_t *out, void *stream) {
    if (!obs || !hist || !out) return fail("mgb_append_action: null buffer");
    if (N < 0 || D < 0 || A < 1 || A > 254 || K < 1) return fail("mgb_append_action: bad sizes");
    if (N == 0) return 0;
    k_action_history<<<elementwise_grid(N), 256, 0, (cudaStream_t)stream>>>(N, K, actions, done, hist);
    CUDA_OK(cudaGetLastError());
    const int64_t words = (N * ((int64_t)D + (int64_t)A * K) + 3) / 4;
    k_append_action<<<elementwise_grid(words), 256, 0, (cudaStream_t)stream>>>(N, D, A, K, obs, hist, out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mgb_goal_policy(int64_t n_cells, int32_t planes, int32_t agent_idx, int32_t empty_idx, int32_t goal_idx,
                    const uint8_t *obs, uint8_t *achieved, uint8_t *desired, void *stream) {
    if (!obs || !achieved || !desired) return fail("mgb_goal_policy: null buffer");
    if (n_cells < 0 || planes < 1 || agent_idx < 0 || agent_idx >= planes || empty_idx < 0 || empty_idx >= planes ||
        goal_idx < 0 || goal_idx >= planes) return fail("mgb_goal_policy: bad sizes");
    if (n_cells == 0) return 0;
    k_goal_policy<<<elementwise_grid(n_cells), 256, 0, (cudaStream_t)stream>>>(n_cells, planes, agent_idx, empty_idx, goal_idx, obs,
                                                                             achieved, desired);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mgb_episode_stats(int64_t N, const double *reward, const uint8_t *done, double *run_ret, int32_t *run_len, double *out_ret,
                      int32_t *out_len, uint64_t *totals, void *stream) {
    if (!reward || !done || !run_ret || !run_len || !out_ret || !out_len) return fail("mgb_episode_stats: null buffer");
    if (N < 0) return fail("mgb_episode_stats: bad size");
    if (N == 0) return 0;
    k_episode_stats<<<elementwise_grid(N), 256, 0, (cudaStream_t)stream>>>(N, reward, done, run_ret, run_len, out_ret, out_len,
                                                                        reinterpret_cast<unsigned long long *>(totals));
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mgb_error_flags(mgb_handle *h, void *stream, uint32_t *flags_host) {
    if (!h || !flags_host) return fail("null argument");
    ON_DEVICE(h);
    CUDA_OK(cudaMemcpyAsync(flags_host, h->err, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_OK(cudaMemsetAsync(h->err, 0, 4, (cudaStream_t)stream));
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

}  // extern "C"

#ifdef MGB_DEBUG_SPARES
extern "C" int mgb_debug_spares(unsigned long long *out, int clear) {
    if (cudaMemcpyFromSymbol(out, mgb::g_spare_dbg, sizeof(unsigned long long) * 8) != cudaSuccess) return 1;
    if (clear) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(mgb::g_spare_dbg, z, sizeof(z)); }
    return 0;
}
#endif
